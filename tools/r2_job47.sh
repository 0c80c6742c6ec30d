#!/bin/bash
# 8 ranks: the driver's command (default extras, e2e) -- after the time-based wait bound
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2d_bench_8gpu.json 2> gpurun_out/r2d_bench_8gpu.err; echo "rc=$?"
tail -2 gpurun_out/r2d_bench_8gpu.err
python - <<P
import json
d=json.loads(open("gpurun_out/r2d_bench_8gpu.json").read().strip().split('\n')[-1]); r=d["roofline"]
print("N=%d step %.3f ms kernel %.3f value %.3f G e2e %.1f ms %s"%(d["n_gpus"], d["ms_per_step"], r["kernel_ms"], d["value"]/1e9, d["e2e"]["ms_per_step"], d["e2e"].get("verified")))
print("by rank", ["%.3f"%x for x in (r["kernel_ms_by_rank"] or [])]); print(r["step_breakdown_ms"]); print(json.dumps(d["extras"])[:500])
P
