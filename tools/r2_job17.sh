#!/bin/bash
# round 2: 4 ranks, whole-step + e2e after the patch-edge clipping of the host pipeline
N=${1:-4}
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --no-extras > gpurun_out/r2b_bench_${N}gpu.json 2> gpurun_out/r2b_bench_${N}gpu.err; echo "bench rc=$?"
tail -5 gpurun_out/r2b_bench_${N}gpu.err
python - <<P
import json
d=json.loads(open("gpurun_out/r2b_bench_${N}gpu.json").read().strip().split('\n')[-1])
print("N=%d step %.3f ms kernel %.3f frac %.3f launches %d value %.3f G"%(d["n_gpus"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["gpu_launches"], d["value"]/1e9))
print("e2e", d["e2e"] and (d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d["e2e"]["d2h_bytes_per_step"], d["e2e"].get("verified")))
P
nproc; free -g | head -2; lscpu | grep -i "model name\|socket\|numa" ; nvidia-smi topo -m | head -20
