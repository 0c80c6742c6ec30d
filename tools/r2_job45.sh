#!/bin/bash
# 8 ranks: the host-buffer call with 2 / 3 / 4 fill threads per rank (32 host cores)
mkdir -p gpurun_out
for FT in 2 4 3; do
SIAFD_B200_FILL_THREADS=$FT timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2957$FT bench.py --gpus 8 --steps 5 --warmup 3 --no-extras --e2e-steps 4 > gpurun_out/r2_ft${FT}_8gpu.json 2> gpurun_out/r2_ft${FT}_8gpu.err; echo "rc=$?"
python - <<P
import json
d=json.loads(open("gpurun_out/r2_ft${FT}_8gpu.json").read().strip().split('\n')[-1])
print("fill threads $FT: step %.3f ms e2e %.1f ms"%(d["ms_per_step"], d["e2e"]["ms_per_step"]))
P
done
