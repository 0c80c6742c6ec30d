#!/bin/bash
# round 2: the whole GPU suite, smoke(), then the default bench command (N = 1)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_full_gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_full_gpu_tests.log
tail -4 gpurun_out/r2_full_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_final_n1.json 2> gpurun_out/r2_final_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_final_n1.err
python - <<P
import json
d=json.loads(open("gpurun_out/r2_final_n1.json").read().strip().split('\n')[-1]); r=d["roofline"]
print("step %.3f ms kernel %.3f ms frac %.3f launches %d e2e %.1f ms"%(d["ms_per_step"], r["kernel_ms"], r["frac"], d["gpu_launches"], d["e2e"]["ms_per_step"]), r["step_breakdown_ms"])
print(json.dumps(d["extras"])[:900]); print(d["cpu_baseline"]); print(d["vertical_velocity"]); print(json.dumps(d["consumers"])[:600])
P
