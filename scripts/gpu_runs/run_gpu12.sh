set -x
python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu_r12.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_r12.log
