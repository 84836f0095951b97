"""Builds tests/emul/libgcm_b200_emul.so (see README.md: test infrastructure, not a product path)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def build_emul():
    out = os.path.join(HERE, "libgcm_b200_emul.so")
    src = os.path.join(HERE, "emul_all.cpp")
    deps = [src] + [os.path.join(ROOT, "gcm_b200", "csrc", f) for f in
                    ("gcmb_capi.cu", "stage_dispatch.cu", "thread_fns.h", "internal.cuh", "patterns.inc")]
    if os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++",
                    "-Wl,-Bsymbolic",
                    "-I" + cuda_inc, src, "-o", out, "-ldl"], check=True)
    return out
