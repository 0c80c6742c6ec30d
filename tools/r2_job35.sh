#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pismv.py tests/test_gpu_exact_columns.py tests/test_golden_fixtures.py tests/test_gpu_decomposition.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2_t9.json 2> gpurun_out/r2_t9.err
python - <<P
import json
d=json.loads(open("gpurun_out/r2_t9.json").read().strip().split('\n')[-1])
print(d["vertical_velocity"]["ms"], d["vertical_velocity"]["frac"], d["consumers"]["vertical_velocity_plus_cfl_ms"], d["consumers"]["strain_heating"]["ms"])
P
