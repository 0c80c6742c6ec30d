#!/usr/bin/env python3
"""Generate tests/golden/oracle_fixture_<case>.npz: inputs and oracle outputs of one SIAFD::update for small
cases, frozen so that (a) the GPU path is checked against committed vectors without running the oracle, and
(b) a change in the oracle itself shows up as a diff.  Run from the repo root:  python tools/make_golden.py
The reference itself cannot be built in this image, so these are outputs of the oracle
(the CPU restatement pinned by tests/test_oracle_known_answers.py), not of PISM."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402

CASES = ["dome_33_13", "C2t", "C4s"]  # gpbld dome (ragged sizes); Test G state, arr, cold converter; Greenland-like + smoother


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    for name in CASES:
        grid, cfg, inputs, gb = cases.case(name)
        run = cases.oracle_run(grid, cfg, inputs, gb, full=True)
        assert run.status == 0, (name, run.status)
        data = {"in_" + k: v for k, v in inputs.items()}
        for k in ("h_x", "h_y", "D", "Q", "u", "v", "work2d_0", "work2d_1"):
            data["out_" + k] = run.a[k]
        data["out_D_max"] = np.array(run.D_max)
        if gb is not None:
            data["global_bed"] = gb
        path = os.path.join(out, "oracle_fixture_%s.npz" % name)
        np.savez_compressed(path, **data)
        print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
