"""A small pass over every stage-kernel family (for a run under a checking tool where one is available; compute-sanitizer is
closed on the round-2 pool): border inside the tile kernels,
random single stages against the oracle (cp.async and bulk-copy pipelines, fp64 and fp32), two engine fixtures, a simplex run."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import gcm_b200
from helpers import compare_with_golden, random_stage_check, tile_border_check
from scenarios import SCENARIOS

lib = gcm_b200.library()
os.chdir("/tmp")
print("tile border", tile_border_check(lib, 8), tile_border_check(lib, 4), flush=True)
random_stage_check(lib, ((3, (19, 13, 37), "elastic", 2), (3, (5, 9, 300), "acoustic", 2), (2, (23, 131), "elastic", 2), (3, (6, 11, 130), "elastic", 3)))
print("random stages ok", flush=True)
for name in ("elastic3d_layers", "elastic3d_contact_z", "elastic3d_layers_courant1", "ortho3d_rotated_plies"):
    compare_with_golden(lib, name, SCENARIOS[name])[0].close()
print("fixtures ok", flush=True)
print("SANITIZE_SMALL_OK", flush=True)
