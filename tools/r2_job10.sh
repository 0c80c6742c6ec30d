#!/bin/bash
# 8 ranks: variants of the step (segment order, rows per CTA), short runs
mkdir -p gpurun_out
run() { # tag, env..., -- extra args
  TAG=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e --no-extras $EXTRA > gpurun_out/r2_v8_${TAG}.json 2> gpurun_out/r2_v8_${TAG}.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r2_v8_${TAG}.json").read().strip().split('\n')[-1])
    print("${TAG}: step %.3f ms kernel %.3f step-kernel %.3f by rank %s"%(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["step_minus_kernel_ms"], ["%.3f"%x for x in d["roofline"]["kernel_ms_by_rank"]]))
except Exception as e:
    print("${TAG} failed", e)
P
}
EXTRA="" run order1 SIAFD_B200_ORDER=1
EXTRA="" run order0 SIAFD_B200_ORDER=0
EXTRA="" run rows16 SIAFD_B200_ORDER=1 SIAFD_B200_ROWS=16
EXTRA="--procs-y 1222,826,826,1222" run y1222 SIAFD_B200_ORDER=1
