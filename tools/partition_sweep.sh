#!/bin/bash
# Profiling aid (not a bench line): what ONE GPU of an N-GPU song-partitioned job computes, for several head sizes.
mkdir -p gpurun_out
IFS=";" read -ra CFGS <<< "${PS_CFGS:-0/8 0;0/8 125;0/8 250;0/4 0;0/4 150;0/2 0;0/2 125}"
for cfg in "${CFGS[@]}"; do
  set -- $cfg
  tag=$(echo $1 | tr / _)_md$2
  MRSCORE_DEBUG_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-k1-probe --emulate-song-partition $1 --head-min-deg $2 > gpurun_out/ps_$tag.json 2> gpurun_out/ps_$tag.err
  python - $tag <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/ps_{sys.argv[1]}.json")); p=d["roofline"]["phase_ms_per_step"]
    print(sys.argv[1], "step", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["ms_per_step"],1), "steady", round(d["steady_state"]["ms_per_step"],1), "pre", round(p["precompute"],1), "head", round(p["head_rowsum"],1), "tail", round(p["tail_scatter"],1), "topk", round(p["topk"],1),
          "n_head", d["config"]["head_songs"], "batch", d["config"]["users_per_batch"], {k: round(v,1) for k,v in d["e2e"]["host_wall_ms_per_call_rank0"].items()})
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
grep "set_test_users" gpurun_out/ps_0_8_md125.err | tail -12
