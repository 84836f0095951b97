"""Round-2 kernel experiments on the headline task (1024^3 layered isotropic elastic, or --size N): one fresh process per
variant (the kernel choice is read from the environment once), device-timed per stage class, one JSON line each.

  python scripts/gpu_runs/r2_variants.py [--size 1024] [--steps 5] [--only name,name]

The exp_* variants load an experiment build from exp_libs/<name>/ (scripts/exp_build.py <name> -DFLAG ...; the flags of every
build that was measured are listed with its numbers in profiles/r2_variants.md, experimental code paths that are not in the
tree any more as patches under scripts/exp_patches/); without that directory they report an error line and go on.
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

VARIANTS = {
    # name: (environment, real_bytes, fma, courant)
    "default": ({}, 8, False, 0.9),
    "separate_z_border_kernel": ({"GCMB_ZTILE_BORDER": "0"}, 8, False, 0.9),
    "fp32_default": ({}, 4, False, 0.9),
    # experiment builds (scripts/exp_build.py): bulk-copy marching kernels, blocks per SM x window layout
    "exp_tma6_slot_plane": ({"GCMB_EXP_LIB_DIR": "b6", "GCMB_STAGE_IMPL": "3", "GCMB_TMA_WINDOW": "1"}, 8, False, 0.9),
    "exp_tma6_register_window": ({"GCMB_EXP_LIB_DIR": "b6", "GCMB_STAGE_IMPL": "3", "GCMB_TMA_WINDOW": "0"}, 8, False, 0.9),
    "exp_ldgsts6_register_window": ({"GCMB_EXP_LIB_DIR": "b6", "GCMB_STAGE_IMPL": "2"}, 8, False, 0.9),
    "exp_tma7_slot_plane": ({"GCMB_EXP_LIB_DIR": "b7", "GCMB_STAGE_IMPL": "3", "GCMB_TMA_WINDOW": "1"}, 8, False, 0.9),
    "exp_tma7_register_window": ({"GCMB_EXP_LIB_DIR": "b7", "GCMB_STAGE_IMPL": "3", "GCMB_TMA_WINDOW": "0"}, 8, False, 0.9),
    "exp_ldgsts7_slot_plane": ({"GCMB_EXP_LIB_DIR": "b7s", "GCMB_STAGE_IMPL": "2"}, 8, False, 0.9),
    "exp_tma6_zt256": ({"GCMB_EXP_LIB_DIR": "zt256", "GCMB_STAGE_IMPL": "3"}, 8, False, 0.9),
    "exp_tma6_order1": ({"GCMB_EXP_LIB_DIR": "ord1", "GCMB_STAGE_IMPL": "3"}, 8, False, 0.9),
    "exp_tma6_order1_seg128": ({"GCMB_EXP_LIB_DIR": "ord1", "GCMB_STAGE_IMPL": "3", "GCMB_MARCH_SEG": "128"}, 8, False, 0.9),
    "exp_tma6_order1_seg1024": ({"GCMB_EXP_LIB_DIR": "ord1", "GCMB_STAGE_IMPL": "3", "GCMB_MARCH_SEG": "1024"}, 8, False, 0.9),
    "exp_tma6_order1_zt256": ({"GCMB_EXP_LIB_DIR": "ord1zt256", "GCMB_STAGE_IMPL": "3"}, 8, False, 0.9),
    "exp_tma6_order1_zt256_seg1024": ({"GCMB_EXP_LIB_DIR": "ord1zt256", "GCMB_STAGE_IMPL": "3", "GCMB_MARCH_SEG": "1024"}, 8, False, 0.9),
    "exp_ldgsts_base": ({"GCMB_EXP_LIB_DIR": "base"}, 8, False, 0.9),
    "exp_ldgsts_regwin_idring": ({"GCMB_EXP_LIB_DIR": "regwin_idring"}, 8, False, 0.9),
    "exp_ldgsts_slot_idring": ({"GCMB_EXP_LIB_DIR": "slot_idring"}, 8, False, 0.9),
    "exp_ldgsts_slot_idring7": ({"GCMB_EXP_LIB_DIR": "slot_idring7"}, 8, False, 0.9),
    "exp_ldgsts_cp16": ({"GCMB_EXP_LIB_DIR": "cp16"}, 8, False, 0.9),
    "exp_tma_base": ({"GCMB_STAGE_IMPL": "3"}, 8, False, 0.9),
    "exp_tma_nst4": ({"GCMB_EXP_LIB_DIR": "nst4", "GCMB_STAGE_IMPL": "3"}, 8, False, 0.9),
    "exp_tma_nst5_4blocks": ({"GCMB_EXP_LIB_DIR": "nst5b4", "GCMB_STAGE_IMPL": "3"}, 8, False, 0.9),
    "exp_tma_nst6_3blocks": ({"GCMB_EXP_LIB_DIR": "nst6b3", "GCMB_STAGE_IMPL": "3"}, 8, False, 0.9),
    "exp_tma_nst7_3blocks": ({"GCMB_EXP_LIB_DIR": "nst7b3", "GCMB_STAGE_IMPL": "3"}, 8, False, 0.9),
    "exp_tma_nst8_2blocks": ({"GCMB_EXP_LIB_DIR": "nst8b2", "GCMB_STAGE_IMPL": "3"}, 8, False, 0.9),
    "exp_tma_nst10_2blocks": ({"GCMB_EXP_LIB_DIR": "nst10b2", "GCMB_STAGE_IMPL": "3"}, 8, False, 0.9),
    "exp_tma_nst12_2blocks": ({"GCMB_EXP_LIB_DIR": "nst12b2", "GCMB_STAGE_IMPL": "3"}, 8, False, 0.9),
    "exp_deep_zt160": ({"GCMB_EXP_LIB_DIR": "zt160"}, 8, False, 0.9),
    "exp_deep_zt96_5blocks": ({"GCMB_EXP_LIB_DIR": "zt96b4"}, 8, False, 0.9),
    "exp_deep_nst4_3blocks": ({"GCMB_EXP_LIB_DIR": "nst4b3"}, 8, False, 0.9),
    "exp_deep_nst5_3blocks": ({"GCMB_EXP_LIB_DIR": "nst5b3"}, 8, False, 0.9),
    "exp_deep_nst6_2blocks": ({"GCMB_EXP_LIB_DIR": "nst6b2"}, 8, False, 0.9),
    "exp_tma5_slot_plane": ({"GCMB_EXP_LIB_DIR": "b5", "GCMB_STAGE_IMPL": "3", "GCMB_TMA_WINDOW": "1"}, 8, False, 0.9),
    "fp32_separate_z_border_kernel": ({"GCMB_ZTILE_BORDER": "0"}, 4, False, 0.9),
    "ldgsts_separate_border": ({"GCMB_STAGE_IMPL": "2"}, 8, False, 0.9),
    "ldgsts_fused_border": ({"GCMB_STAGE_IMPL": "2", "GCMB_FUSED_BORDER": "1"}, 8, False, 0.9),
    "tma_warp_pipes": ({"GCMB_STAGE_IMPL": "3", "GCMB_FUSED_BORDER": "1"}, 8, False, 0.9),
    "tma_warp_pipes_separate_border": ({"GCMB_STAGE_IMPL": "3"}, 8, False, 0.9),
    "tma_block_ring_inline": ({"GCMB_STAGE_IMPL": "3", "GCMB_TMA_MARCH": "1", "GCMB_TMA_ZTILE": "1"}, 8, False, 0.9),
    "tma_block_ring_producer_warp": ({"GCMB_STAGE_IMPL": "3", "GCMB_TMA_MARCH": "2", "GCMB_TMA_ZTILE": "0"}, 8, False, 0.9),
    "tma_warp_pipes_rows64": ({"GCMB_STAGE_IMPL": "3", "GCMB_ZTILE_ROWS": "64"}, 8, False, 0.9),
    "tma_warp_pipes_rows16": ({"GCMB_STAGE_IMPL": "3", "GCMB_ZTILE_ROWS": "16"}, 8, False, 0.9),
    "tma_warp_pipes_seg128": ({"GCMB_STAGE_IMPL": "3", "GCMB_MARCH_SEG": "128"}, 8, False, 0.9),
    "tma_warp_pipes_seg512": ({"GCMB_STAGE_IMPL": "3", "GCMB_MARCH_SEG": "512"}, 8, False, 0.9),
    "ldgsts_fma": ({"GCMB_STAGE_IMPL": "2"}, 8, True, 0.9),
    "tma_fma": ({"GCMB_STAGE_IMPL": "3"}, 8, True, 0.9),
    "ldgsts_courant1": ({"GCMB_STAGE_IMPL": "2"}, 8, False, 1.0),
    "tma_courant1": ({"GCMB_STAGE_IMPL": "3"}, 8, False, 1.0),
    "ldgsts_fp32": ({"GCMB_STAGE_IMPL": "2"}, 4, False, 0.9),
    "tma_fp32": ({"GCMB_STAGE_IMPL": "3"}, 4, False, 0.9),
}


def child(args):
    sys.path.insert(0, ROOT)
    import numpy as np
    import gcm_b200
    from gcm_b200 import capi
    import bench
    if os.environ.get("GCMB_EXP_LIB_DIR"):   # an experiment build of scripts/exp_build.py (loaded INSTEAD of the product's: same soname)
        d = os.path.join(ROOT, "exp_libs", os.environ["GCMB_EXP_LIB_DIR"])
        lib = capi.Library(os.path.join(d, "libgcm_b200.so"), os.path.join(d, "libgcm_b200_host.so"))
        gcm_b200._LIB = lib
    else:
        lib = gcm_b200.library()
    n = args.size
    text = bench.task_text(n, n, n, steps=10 ** 6, detector=False).replace("courant 0.9", "courant %r" % args.courant)
    os.chdir("/tmp")
    eng = capi.HostEngine(lib, text, real_bytes=args.real_bytes, fma=bool(args.fma))
    ctxh = eng.context_handle()
    eng.advance(3)
    lib.check(lib.c.gcmb_sync(ctxh))
    lib.check(lib.c.gcmb_profile_enable(ctxh, 1))
    launches0 = lib.c.gcmb_launch_count(ctxh)
    lib.check(lib.c.gcmb_timer_start(ctxh))
    eng.advance(args.steps)
    ms = capi.ctypes.c_float()
    lib.check(lib.c.gcmb_timer_stop(ctxh, capi.ctypes.byref(ms)))
    prof_ms = np.zeros(8)
    prof_n = np.zeros(8, dtype=np.int64)
    lib.check(lib.c.gcmb_profile_get(ctxh, 8, capi.dp(prof_ms), prof_n.ctypes.data_as(capi.c_ll_p)))
    chk = capi.ctypes.c_double()
    lib.check(lib.c.gcmb_cubic_checksum(eng.body_handle(0), capi.ctypes.byref(chk)))
    per_step = ms.value / args.steps
    bytes_stage = 2 * 9 * args.real_bytes * n ** 3
    peak = bench.measured_hbm_peak()[0]
    out = {"variant": args.name, "n": n, "ms_per_step": per_step, "node_updates_per_s": n ** 3 / (per_step * 1e-3),
           "whole_step_frac": 3 * bytes_stage / (per_step * 1e-3) / 1e9 / peak,
           "stage_ms": [prof_ms[a] / args.steps for a in range(3)],
           "stage_frac": [bytes_stage / (prof_ms[a] / args.steps * 1e-3) / 1e9 / peak if prof_ms[a] > 0 else None for a in range(3)],
           "border_ms": prof_ms[3] / args.steps, "launches_per_step": (lib.c.gcmb_launch_count(ctxh) - launches0) / args.steps,
           "kernels": [eng.kernel_name(0, d) for d in range(3)], "checksum": chk.value}
    print("VARIANT " + json.dumps(out))
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--name", default="")
    ap.add_argument("--real-bytes", type=int, default=8)
    ap.add_argument("--fma", type=int, default=0)
    ap.add_argument("--courant", type=float, default=0.9)
    args = ap.parse_args()
    if args.child:
        return child(args)
    names = [x for x in args.only.split(",") if x] or list(VARIANTS)
    for name in names:
        env, rb, fma, courant = VARIANTS[name]
        cmd = [sys.executable, os.path.abspath(__file__), "--child", "--name", name, "--size", str(args.size), "--steps", str(args.steps),
               "--real-bytes", str(rb), "--fma", str(int(fma)), "--courant", repr(courant)]
        r = subprocess.run(cmd, env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
        lines = [line for line in r.stdout.splitlines() if line.startswith("VARIANT ")]
        print(lines[0] if lines else "VARIANT " + json.dumps({"variant": name, "error": (r.stdout + r.stderr)[-600:]}), flush=True)


if __name__ == "__main__":
    main()
