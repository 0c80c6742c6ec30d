#!/bin/bash
# round 2: N ranks, the whole bench line (step, kernel by rank, breakdown, extras, e2e)
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2c_bench_${N}gpu.json 2> gpurun_out/r2c_bench_${N}gpu.err; echo "bench rc=$?"
tail -3 gpurun_out/r2c_bench_${N}gpu.err
python - <<P
import json
d=json.loads(open("gpurun_out/r2c_bench_${N}gpu.json").read().strip().split('\n')[-1])
r=d["roofline"]
print("N=%d step %.3f ms kernel %.3f frac %.3f launches %d value %.3f G"%(d["n_gpus"], d["ms_per_step"], r["kernel_ms"], r["frac"], d["gpu_launches"], d["value"]/1e9))
print("by rank", ["%.3f"%x for x in (r["kernel_ms_by_rank"] or [])]); print("breakdown", r["step_breakdown_ms"]); print(d["config"]["decomposition"])
print("e2e", d["e2e"] and (d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d["e2e"]["d2h_bytes_per_step"], d["e2e"].get("verified")))
print(json.dumps(d["extras"])[:1500])
P
