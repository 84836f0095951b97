"""Builds the native libraries of gcm_b200 in-tree (nvcc cross-compiles sm_100a without a GPU).

  gcm_b200/libgcm_b200.so       CUDA kernels + C ABI (include/gcm_b200.h)
  gcm_b200/libgcm_b200_host.so  C++ host layer (Task / Engine mirror of the reference) on top of the C ABI

python -m gcm_b200.build [--force]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BUILD = os.path.join(ROOT, "build")

NVCC_FLAGS = ["-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              # the reference is built without FMA contraction (CMakeLists.txt:6-7); bit-exact parity needs the same
              "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off"]
CXX_FLAGS = ["-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-Wall"]

CUDA_SOURCES = ["csrc/gcmb_capi.cu", "csrc/stage_dispatch.cu"]
CUDA_HEADERS = ["csrc/internal.cuh", "csrc/thread_fns.h", "csrc/march_async.h", "csrc/ztile.h", "csrc/patterns.inc",
                "csrc/simplex_fns.h", "csrc/simplex_capi.inc", "../include/gcm_b200.h"]
HOST_SOURCES = ["host/models.cpp", "host/engine.cpp", "host/task_file.cpp", "host/host_capi.cpp", "host/simplex_mesh.cpp", "host/simplex_engine.cpp", "host/vtk_writer.cpp"]
HOST_HEADERS = ["host/gcmb_host.hpp", "../include/gcm_b200.h"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    lib = os.path.join(HERE, "libgcm_b200.so")
    cuda_src = [os.path.join(HERE, s) for s in CUDA_SOURCES]
    cuda_dep = cuda_src + [os.path.join(HERE, h) for h in CUDA_HEADERS]
    if force or _newer(lib, cuda_dep):
        objs = [os.path.join(BUILD, os.path.basename(s) + ".o") for s in cuda_src]

        def compile_one(pair):
            src, obj = pair
            return _run(["nvcc"] + NVCC_FLAGS + ["-c", src, "-o", obj])

        with ThreadPoolExecutor(max_workers=len(cuda_src)) as ex:
            for out in ex.map(compile_one, zip(cuda_src, objs)):
                if verbose and out:
                    print(out)
        # -Bsymbolic: internal references bind inside the library whatever else the process has loaded
        _run(["nvcc", "-shared", "-Xlinker", "-Bsymbolic", "-o", lib] + objs + ["-ldl"])
    host = os.path.join(HERE, "libgcm_b200_host.so")
    host_src = [os.path.join(HERE, s) for s in HOST_SOURCES]
    host_dep = host_src + [os.path.join(HERE, h) for h in HOST_HEADERS] + [os.path.abspath(__file__)]
    if force or _newer(host, host_dep):
        _run(["g++"] + CXX_FLAGS + ["-shared", "-o", host] + host_src +
             ["-L" + HERE, "-lgcm_b200", "-Wl,-rpath,$ORIGIN"])
    # gcmb_exe: the reference launcher's command line (src/launcher/main.cpp) on top of the host layer
    exe = os.path.join(HERE, "gcmb_exe")
    exe_src = os.path.join(HERE, "host", "launcher.cpp")
    if force or _newer(exe, [exe_src, host]):
        _run(["g++"] + CXX_FLAGS + ["-o", exe, exe_src, "-L" + HERE, "-lgcm_b200_host", "-lgcm_b200", "-Wl,-rpath,$ORIGIN"])
    return lib, host


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
