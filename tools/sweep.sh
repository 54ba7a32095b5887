#!/bin/bash
# usage: tools/sweep.sh tag "ENV=.. ENV=.." ...   (each argument = one bench configuration; profiling aid, not a bench line)
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-k1-probe > gpurun_out/sweep_$i.json 2> gpurun_out/sweep_$i.err
  echo "cfg[$i] '$cfg' rc=$?"
  python - gpurun_out/sweep_$i.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]
    print("  ms/step", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["ms_per_step"],1), {k:round(v,1) for k,v in r["phase_ms_per_step"].items() if v}, "frac", r["frac"], "batch", d["config"].get("users_per_batch"))
except Exception as e:
    print("  failed", e)
PY
done
