"""Pipelined vs plain host update: identical bits.  usage: e2e_check.py <case>"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import cases, gpu_util as U
name = sys.argv[1]
res = []
for pipe, band, rows in (("0", "4", "64"), ("1", "1", "16"), ("1", "3", "8"), ("1", "100", "64")):
    os.environ["SIAFD_B200_PIPELINE"], os.environ["SIAFD_B200_BAND"], os.environ["SIAFD_B200_ROWS"] = pipe, band, rows
    grid, cfg, inputs, gb = cases.case(name)
    sia = U.make_sia(grid, cfg, gb)
    U.gpu_update(sia, inputs, True)
    res.append({k: np.array(v, copy=True) for k, v in (("u", sia.velocity_u()), ("v", sia.velocity_v()), ("D", sia.diffusivity()), ("Q", sia.diffusive_flux()))})
for r in res[1:]:
    for k in r:
        assert np.array_equal(r[k], res[0][k]), k
print(name, "pipelined == plain, bitwise")
