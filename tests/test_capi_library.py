"""The product library on a machine without GPU: it loads, exports every symbol include/gcm_b200.h
declares, and refuses to work (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import gcm_b200
from gcm_b200 import build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return gcm_b200.library()


def test_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "gcm_b200.h")).read()
    declared = set(re.findall(r"\b(gcmb_[a-z0-9_]+)\s*\(", header))
    declared -= {"gcmb_ctx", "gcmb_body"}
    assert declared == set(capi.C_ABI), declared ^ set(capi.C_ABI)
    for name in declared:
        assert getattr(lib.c, name) is not None


def test_is_sm100a_cuda_code(lib):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", lib.cuda_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    handle = ctypes.c_void_p()
    rc = lib.c.gcmb_create(0, 8, ctypes.byref(handle))
    assert rc == 101  # GCMB_E_NO_DEVICE
    assert b"no CUDA device" in lib.c.gcmb_last_error()
    with pytest.raises(capi.GcmError):
        capi.HostEngine(lib, "dimensionality 1\ncourant 0.5\nborder_size 1\nh 0.1\nsteps 1\n"
                             "body 0 elastic isotropic sizes 8 start 0\nmaterial default isotropic 1 1 1\n")


def test_product_never_touches_oracle_or_emul():
    for base, _, files in os.walk(os.path.join(ROOT, "gcm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".inc")):
                text = open(os.path.join(base, f)).read()
                assert "oracle" not in text, os.path.join(base, f)
                assert "libgcm_b200_emul" not in text, os.path.join(base, f)
