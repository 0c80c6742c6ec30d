#!/bin/bash
# round 2: N-GPU bench (N = $1) + pismv test C on the same ranks
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench_${N}gpu.err
python - <<P
import json
d=json.load(open("gpurun_out/r2_bench_${N}gpu.json"))
print("N=%d step %.3f ms kernel %.3f frac %.3f launches %d value %.3f G"%(d["n_gpus"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["gpu_launches"], d["value"]/1e9))
print("e2e", d["e2e"] and (d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d["e2e"]["d2h_bytes_per_step"], d["e2e"].get("verified")))
print(json.dumps(d["extras"])[:1200]); print(d["config"]["decomposition"]); print(d["checksum"], d["D_max"])
P
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/pismv_multi_gpu.py > gpurun_out/r2_pismv_${N}gpu.txt 2>&1; echo "pismv rc=$?"; tail -12 gpurun_out/r2_pismv_${N}gpu.txt
