#!/bin/bash
# usage: tools/ab.sh tag "ENV=.. ENV=.." ...   -> gpurun_out/ab_<tag>.json for every (tag, env) pair
while [ $# -gt 1 ]; do
  tag=$1; envs=$2; shift 2
  env $envs python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err || tail -5 gpurun_out/ab_$tag.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/ab_$tag.json"))
    r=d["roofline"]
    print("$tag", "ms/step %.2f"%d["ms_per_step"], "e2e %.1f"%d["e2e"]["value"], {k:(round(v["ms_per_step"],2),round(v["tflops"])) for k,v in r["by_class"].items()}, "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("$tag failed", e)
P
done
