"""Builds tests/emul/libgcm_b200_emul.so and libgcm_b200_host_emul.so (see README.md: test infrastructure,
not a product path)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "gcm_b200", "csrc")


def _fresh(out, deps):
    return os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps)


def build_emul():
    """The product's CUDA sources compiled for the host; kernels are stepped thread by thread.  Every translation
    unit of gcm_b200/csrc is compiled by g++ with emul_prelude.h force-included; the stage kernels of the bit-exact
    fp64 set (0) and of the fp32 set (2) are built (set 1 differs from set 0 only by nvcc's FMA contraction)."""
    out = os.path.join(HERE, "libgcm_b200_emul.so")
    prelude = os.path.join(HERE, "emul_prelude.h")
    deps = [prelude, os.path.abspath(__file__)] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", ".inc"))]
    if _fresh(out, deps):
        return out
    from concurrent.futures import ThreadPoolExecutor
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    objdir = os.path.join(ROOT, "build", "emul")
    os.makedirs(objdir, exist_ok=True)
    units = [("gcmb_capi.cu", "gcmb_capi.o", []), ("simplex_capi.cu", "simplex_capi.o", []), ("stage_dispatch.cu", "stage_dispatch.o", [])]
    for s in (0, 2):
        for g in (0, 1, 2, 3, 100):
            units.append(("stage_inst.cu", "stage_s%d_g%d.o" % (s, g), ["-DGCMB_SET=%d" % s, "-DGCMB_GROUP=%d" % g]))

    def compile_one(u):
        src, obj, flags = u
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-x", "c++", "-include", prelude,
                        "-I" + cuda_inc, "-c", os.path.join(CSRC, src), "-o", os.path.join(objdir, obj)] + flags, check=True)
        return os.path.join(objdir, obj)

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(compile_one, units))
    subprocess.run(["g++", "-shared", "-Wl,-Bsymbolic", "-o", out] + objs + ["-ldl"], check=True)
    return out


def build_host_emul():
    """The product's host layer (gcm_b200/host/*.cpp, unchanged) linked against the stepping harness."""
    emul = build_emul()
    out = os.path.join(HERE, "libgcm_b200_host_emul.so")
    hdir = os.path.join(ROOT, "gcm_b200", "host")
    srcs = [os.path.join(hdir, f) for f in ("models.cpp", "engine.cpp", "task_file.cpp", "host_capi.cpp", "simplex_mesh.cpp", "simplex_engine.cpp", "vtk_writer.cpp")]
    deps = srcs + [os.path.join(hdir, "gcmb_host.hpp"), emul]
    if not _fresh(out, deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", out] + srcs +
                       ["-L" + HERE, "-lgcm_b200_emul", "-Wl,-rpath,$ORIGIN"], check=True)
    return out


def build_launcher_emul():
    """gcm_b200/host/launcher.cpp linked against the stepping harness (command-line logic on the build machine)."""
    host = build_host_emul()
    out = os.path.join(HERE, "gcmb_exe_emul")
    src = os.path.join(ROOT, "gcm_b200", "host", "launcher.cpp")
    if not _fresh(out, [src, host]):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-o", out, src, "-L" + HERE,
                        "-lgcm_b200_host_emul", "-lgcm_b200_emul", "-Wl,-rpath,$ORIGIN"], check=True)
    tasks = os.path.join(HERE, "tasks")
    if not os.path.islink(tasks):
        os.symlink(os.path.join(ROOT, "gcm_b200", "tasks"), tasks)
    return out
