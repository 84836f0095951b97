"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/gcm_ref, built from
/root/reference by oracle/Makefile) on every task in tests/scenarios.py.  Run in the build container
(where /root/reference exists):   python tests/golden/make_golden.py [scenario ...]
The fixtures travel to the GPU box; /root/reference does not."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_host as oh  # noqa: E402
from scenarios import SCENARIOS  # noqa: E402


def main():
    oh.build_oracle()
    here = os.path.dirname(os.path.abspath(__file__))
    only = sys.argv[1:]   # optional: names of the scenarios to (re)generate
    for name, text in SCENARIOS.items():
        if only and name not in only:
            continue
        with tempfile.TemporaryDirectory() as tmp:
            out = oh.run_reference(text, tmp, matrices=True)
        arrays = {"task": np.array(text)}
        for key, val in out.items():
            if key == "meta":
                arrays["tau"] = np.float64(val["tau"])
                arrays["time"] = np.float64(val["time"])
                arrays["steps"] = np.int64(val["steps"])
            elif key == "detector":
                arrays["detector"] = val
            elif isinstance(key, tuple):
                arrays["mat%d" % key[1]] = val
            else:
                arrays["body%d" % key] = val
        np.savez_compressed(os.path.join(here, name + ".npz"), **arrays)
        print(name, {k: getattr(v, "shape", None) for k, v in arrays.items() if k != "task"})


if __name__ == "__main__":
    main()
