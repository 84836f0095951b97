"""Builds the native libraries of gcm_b200 in-tree (nvcc cross-compiles sm_100a without a GPU).

  gcm_b200/libgcm_b200.so       CUDA kernels + C ABI (include/gcm_b200.h)
  gcm_b200/libgcm_b200_host.so  C++ host layer (Task / Engine mirror of the reference) on top of the C ABI

The stage kernels (csrc/stage_inst.cu) are compiled once per (kernel set, pattern group), in parallel:
  set 0  fp64, -fmad=false   bit-identical to the reference CPU engine (which is built without FMA contraction)
  set 1  fp64, FMA allowed   gcmb_set_fma / GCMB_FMA=1: within 1e-12 of the reference, not bit-identical
  set 2  fp32, FMA allowed   gcmb_create(device, 4)

python -m gcm_b200.build [--force]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BUILD = os.path.join(ROOT, "build")

# GCMB_NVCC_EXTRA: extra definitions for experiments, e.g. -DGCMB_TMA_ALL_MODES (all bulk-copy pipeline layouts)
NVCC_BASE = ["-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3"] + os.environ.get("GCMB_NVCC_EXTRA", "").split()
# the reference is built without FMA contraction (CMakeLists.txt:6-7); bit-exact parity needs the same
NVCC_EXACT = NVCC_BASE + ["-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off"]
NVCC_FMA = NVCC_BASE + ["-fmad=true", "-Xcompiler", "-fPIC,-ffp-contract=off"]
CXX_FLAGS = ["-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-Wall"]

STAGE_SETS = (0, 1, 2)
STAGE_GROUPS = (0, 1, 2, 3, 100)   # pattern groups of csrc/patterns.inc; 100 = dense kernels
CUDA_HEADERS = ["csrc/internal.cuh", "csrc/capi_internal.cuh", "csrc/thread_fns.h", "csrc/march_async.h", "csrc/ztile.h",
                "csrc/tma_pipe.h", "csrc/triangle_fns.h", "csrc/patterns.inc", "csrc/simplex_fns.h", "csrc/simplex_capi.inc", "../include/gcm_b200.h"]
HOST_SOURCES = ["host/models.cpp", "host/engine.cpp", "host/task_file.cpp", "host/host_capi.cpp", "host/simplex_mesh.cpp", "host/simplex_engine.cpp", "host/vtk_writer.cpp"]
HOST_HEADERS = ["host/gcmb_host.hpp", "../include/gcm_b200.h"]


def cuda_units():
    """(source, object name, flags) of every translation unit of libgcm_b200.so"""
    units = [("csrc/gcmb_capi.cu", "gcmb_capi.o", NVCC_EXACT), ("csrc/simplex_capi.cu", "simplex_capi.o", NVCC_EXACT),
             ("csrc/stage_dispatch.cu", "stage_dispatch.o", NVCC_EXACT)]
    for s in STAGE_SETS:
        for g in STAGE_GROUPS:
            flags = (NVCC_EXACT if s == 0 else NVCC_FMA) + ["-DGCMB_SET=%d" % s, "-DGCMB_GROUP=%d" % g]
            units.append(("csrc/stage_inst.cu", "stage_s%d_g%d.o" % (s, g), flags))
    return units


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
    headers = [os.path.join(HERE, h) for h in CUDA_HEADERS] + [os.path.abspath(__file__)]
    todo, objs = [], []
    for src, obj, flags in cuda_units():
        src, obj = os.path.join(HERE, src), os.path.join(BUILD, obj)
        objs.append(obj)
        if force or _newer(obj, [src] + headers):
            todo.append((src, obj, flags))

    def compile_one(unit):
        src, obj, flags = unit
        return _run(["nvcc"] + flags + ["-c", src, "-o", obj])

    if todo:
        # the heaviest units (3-D elastic patterns) first
        todo.sort(key=lambda u: ("_g0" in u[1] or "_g1" in u[1], "stage_" in u[1]), reverse=True)
        with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4)) as ex:
            for out in ex.map(compile_one, todo):
                if verbose and out:
                    print(out)
    if todo or force or _newer(lib, objs):
        # -Bsymbolic: internal references bind inside the library whatever else the process has loaded
        _run(["nvcc", "-shared", "-Xlinker", "-Bsymbolic", "-o", lib] + objs + ["-ldl"])
    host = os.path.join(HERE, "libgcm_b200_host.so")
    host_src = [os.path.join(HERE, s) for s in HOST_SOURCES]
    host_dep = host_src + [os.path.join(HERE, h) for h in HOST_HEADERS] + [os.path.abspath(__file__), lib]
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
