#!/bin/bash
# round 2: N = 4 and N = 2 on one 4-GPU box, full bench lines
mkdir -p gpurun_out
for N in 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2c_bench_${N}gpu.json 2> gpurun_out/r2c_bench_${N}gpu.err; echo "N=$N rc=$?"
python - <<P
import json
d=json.loads(open("gpurun_out/r2c_bench_${N}gpu.json").read().strip().split('\n')[-1]); r=d["roofline"]
print("N=%d step %.3f ms kernel %.3f frac %.3f launches %d value %.3f G"%(d["n_gpus"], d["ms_per_step"], r["kernel_ms"], r["frac"], d["gpu_launches"], d["value"]/1e9))
print("by rank", ["%.3f"%x for x in (r["kernel_ms_by_rank"] or [])]); print("breakdown", r["step_breakdown_ms"]); print(d["config"]["decomposition"])
print("e2e", d["e2e"] and (d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d["e2e"]["d2h_bytes_per_step"], d["e2e"].get("verified")))
print(json.dumps(d["extras"])[:600])
P
done
# host memory bandwidth of this box (what bounds the host-buffer call with several ranks): numpy copy / fill on T threads
python - <<P
import numpy as np, threading, time, os
n=1<<27  # 1 GiB of doubles per thread
for T in (1,4,8,16,32):
    if T>os.cpu_count(): break
    a=[np.ones(n) for _ in range(T)]; b=[np.empty(n) for _ in range(T)]
    def work(i):
        np.copyto(b[i],a[i])
    for rep in range(2):
        th=[threading.Thread(target=work,args=(i,)) for i in range(T)]
        t0=time.perf_counter(); [t.start() for t in th]; [t.join() for t in th]; dt=time.perf_counter()-t0
    print("host copy %d threads: %.1f GB/s (read + write)"%(T, 2*8*n*T/dt/1e9))
    del a,b
P
