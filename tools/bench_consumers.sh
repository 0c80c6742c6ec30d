#!/bin/bash
# Times the consumers of SIAFD's outputs (vertical velocity + fused CFL, mass-continuity step) next to the update.
# Usage: tools/bench_consumers.sh "<kind> <wz> <rows>" ...   (kind 0 = k_vvel_slab, 1 = k_vvel_march)
mkdir -p gpurun_out
for cfg in "${@:-0 8 32}"; do
  set -- $cfg
  tag=$1_$2_$3
  SIAFD_B200_VVEL_KIND=$1 SIAFD_B200_VVEL_WZ=$2 SIAFD_B200_VVEL_ROWS=$3 python bench.py --size ${SIZE:-4096} --steps 5 --warmup 3 --with-w --no-e2e --no-cpu-baseline \
    2>gpurun_out/bench_consumers_$tag.err | grep '^{' > gpurun_out/bench_consumers_$tag.json
  python - "$tag" <<'PY'
import json, sys
r = sys.argv[1]
d = json.loads(open("gpurun_out/bench_consumers_%s.json" % r).read())
v = d["vertical_velocity"]
print("kind_wz_rows", r, "step ms", round(d["ms_per_step"], 3), "slab frac", round(d["roofline"]["frac"], 3), "| w ms", round(v["ms"], 3), "frac", round(v["frac"], 3), "| w+cfl ms", round(d["consumers"]["vertical_velocity_plus_cfl_ms"], 3), "mass step ms", round(d["consumers"]["mass_continuity_step_ms"], 3), "| heat ms", round(d["consumers"]["strain_heating"]["ms"], 3), "frac", round(d["consumers"]["strain_heating"]["frac"], 3))
PY
done
