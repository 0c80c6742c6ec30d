#!/bin/bash
# round 2: scaling sweep on one 8-GPU box: N = 8, 4, 1
mkdir -p gpurun_out
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_scale_${N}gpu.json 2> gpurun_out/r2_scale_${N}gpu.err; echo "N=$N rc=$?"
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-consumers > gpurun_out/r2_scale_1gpu.json 2> gpurun_out/r2_scale_1gpu.err; echo "N=1 rc=$?"
python - <<P
import json
for N in (1,4,8):
    try:
        d=json.loads(open("gpurun_out/r2_scale_%dgpu.json"%N).read().strip().split('\n')[-1])
    except Exception as e:
        print(N, "failed", e); continue
    print("N=%d step %.3f ms kernel %.3f step-kernel %.3f value %.3f G e2e %.1f ms"%(d["n_gpus"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["step_minus_kernel_ms"], d["value"]/1e9, d["e2e"]["ms_per_step"]))
    print("   ", json.dumps(d["extras"])[:700])
P
tail -3 gpurun_out/r2_scale_8gpu.err
