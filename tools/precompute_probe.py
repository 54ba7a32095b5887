#!/usr/bin/env python
"""Profiling aid (not a bench line): head-row precompute time on the MSD shape as a function of the direct-path chunk size and of the
head size.   python tools/precompute_probe.py [chunk_mb ...]     (env MRSCORE_* tunables apply)"""
import json, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from musicrecommendation_b200.dataset import synth_config
from musicrecommendation_b200.recommender import MusicRecommender

ds = synth_config("c4").shard_test_users(0, 2048)
chunks = sys.argv[1:] or ["48"]
out = []
for min_deg in (0, 400):
    mr = MusicRecommender(ds, head_min_deg=min_deg)
    mr.prepare()
    for mb in chunks:
        os.environ["MRSCORE_DIRECT_CHUNK_MB"] = mb
        best = 1e9
        for _ in range(3):
            mr.invalidate_prepared()
            torch.cuda.synchronize(); t0 = time.perf_counter(); mr.prepare(); torch.cuda.synchronize()
            best = min(best, 1e3 * (time.perf_counter() - t0))
        s = mr.getTopK(0, k=500)[0]
        out.append({"min_deg": min_deg, "n_head": mr.info()["n_head"], "direct_chunk_mb": mb, "precompute_ms": round(best, 1),
                    "checksum": int(s[:, 0].astype("int64").sum())})
        print(json.dumps(out[-1]), flush=True)
    mr.close()
    if min_deg == 0:   # the all-staged construction of round 1 for comparison (needs a fresh handle: the split is fixed at load time)
        os.environ["MRSCORE_PRECOMPUTE_STAGE_ALL"] = "1"
        mr = MusicRecommender(ds)
        mr.prepare()
        best = 1e9
        for _ in range(3):
            mr.invalidate_prepared()
            torch.cuda.synchronize(); t0 = time.perf_counter(); mr.prepare(); torch.cuda.synchronize()
            best = min(best, 1e3 * (time.perf_counter() - t0))
        s = mr.getTopK(0, k=500)[0]
        out.append({"min_deg": 0, "n_head": mr.info()["n_head"], "direct_chunk_mb": "all rows staged (round-1 path)", "precompute_ms": round(best, 1),
                    "checksum": int(s[:, 0].astype("int64").sum())})
        print(json.dumps(out[-1]), flush=True)
        mr.close()
        os.environ.pop("MRSCORE_PRECOMPUTE_STAGE_ALL", None)
