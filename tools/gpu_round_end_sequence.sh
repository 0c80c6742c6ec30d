#!/bin/bash
# what the driver runs at round end, in its order: GPU tests, smoke(), the reference arm, the default bench
mkdir -p gpurun_out
(lscpu | grep -i 'model name\|^CPU(s)\|numa\|socket'; free -g | head -2) > gpurun_out/r2_host_info.txt 2>&1
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r2_final_gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_final_gpu_tests.log; tail -3 gpurun_out/r2_final_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r2_final_ref.json
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2_final4_n1.json 2> gpurun_out/r2_final4_n1.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads(open("gpurun_out/r2_final4_n1.json").read().strip().split('\n')[-1]); r=d["roofline"]
print("step %.3f ms value %.4g kernel %.3f ms frac %.3f traffic %s dram_frac %s launches %d e2e %.1f ms"%(d["ms_per_step"], d["value"], r["kernel_ms"], r["frac"], r["traffic"], r["dram_frac"], d["gpu_launches"], d["e2e"]["ms_per_step"]))
print(d["clocks"]); print(d["cpu_baseline"]["value"], d["vertical_velocity"]["ms"], d["consumers"]["strain_heating"]["ms"])
P
