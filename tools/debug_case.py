import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import cases, gpu_util as U
name, full, rows, bulk, skip = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
grid, cfg, inputs, gb = cases.case(name)
run = cases.oracle_run(grid, cfg, inputs, gb, full=bool(full))
sia = U.make_sia(grid, cfg, gb)
sia.set_tuning(rows, bulk, skip)
try:
    U.gpu_update(sia, inputs, bool(full))
    print(name, full, rows, bulk, skip, "OK", U.compare_with_oracle(sia, run, cfg, bool(full)))
except Exception as e:
    print(name, full, rows, bulk, skip, "FAIL", str(e)[:200])
