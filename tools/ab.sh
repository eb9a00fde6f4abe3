#!/bin/bash
# A/B of two builds of the library on the default bench line: tools/ab.sh <old.so> [rounds]
# (the product build in-tree against DMIP_LIB=<old.so>; alternating runs so that clock drift hits both)
OLD=$1; R=${2:-2}
for i in $(seq 1 $R); do
  python bench.py --no-also --no-cpu-baseline --steps 5 > gpurun_out/ab_new_$i.json 2>/dev/null
  DMIP_LIB=$OLD python bench.py --no-also --no-cpu-baseline --steps 5 > gpurun_out/ab_old_$i.json 2>/dev/null
done
python - <<P
import json, glob
for f in sorted(glob.glob("gpurun_out/ab_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, "%.4e" % d["value"], "%.1f ms" % d["ms_per_step"], d["clocks"]["sm_mhz"])
P
