"""Experiment builds: the headline patterns' translation unit (kernel set 0, pattern group 0) recompiled with extra
definitions and linked with the other objects of the last `python -m gcm_b200.build` into exp_libs/<name>/ (git-ignored,
travels to the GPU box).  scripts/gpu_runs/r2_variants.py loads such a library when GCMB_EXP_LIB_DIR is set.

  python scripts/exp_build.py <name> [-DFLAG ...]
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gcm_b200 import build as B  # noqa: E402


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    out = os.path.join(ROOT, "exp_libs", name)
    os.makedirs(out, exist_ok=True)
    objs = []
    for src, obj, unit_flags in B.cuda_units():
        path = os.path.join(B.BUILD, obj)
        if obj == "stage_s0_g0.o":
            path = os.path.join(out, obj)
            B._run(["nvcc"] + unit_flags + flags + ["-c", os.path.join(B.HERE, src), "-o", path])
        objs.append(path)
    lib = os.path.join(out, "libgcm_b200.so")
    B._run(["nvcc", "-shared", "-Xlinker", "-Bsymbolic", "-o", lib] + objs + ["-ldl"])
    shutil.copy(os.path.join(B.HERE, "libgcm_b200_host.so"), out)
    os.remove(os.path.join(out, "stage_s0_g0.o"))
    print(lib)


if __name__ == "__main__":
    main()
