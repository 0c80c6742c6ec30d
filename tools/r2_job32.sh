#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_comm.py tests/test_gpu_pismv.py tests/test_gpu_exact_columns.py tests/test_golden_fixtures.py -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2_t8.json 2> gpurun_out/r2_t8.err
python - <<P
import json
d=json.loads(open("gpurun_out/r2_t8.json").read().strip().split('\n')[-1])
print(d["vertical_velocity"]["ms"], json.dumps(d["consumers"]["strain_heating"])[:300])
P
