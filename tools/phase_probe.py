#!/usr/bin/env python
"""Profiling aid (not a bench line): per-model, per-phase CUDA-event times of one MSD-shaped step on one GPU.
   python tools/phase_probe.py [--users N]     env MRSCORE_* tunables apply."""
import argparse, json, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from musicrecommendation_b200 import _lib
from musicrecommendation_b200.dataset import synth_config
from musicrecommendation_b200.recommender import MusicRecommender

ap = argparse.ArgumentParser()
ap.add_argument("--users", type=int, default=13750)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
ds = synth_config("c4").shard_test_users(0, a.users)
with MusicRecommender(ds, device=0) as mr:
    mr.prepare()
    lib, h = mr._lib, mr._h
    for m in (_lib.MR_UBM, _lib.MR_IBM):
        mr._check(lib.mr_topk_device(h, m, 0.0, 0, 500))
    lib.mr_set_profile(h, 1)
    out = {"info": mr.info()}
    for name, m in (("ubm", _lib.MR_UBM), ("ibm", _lib.MR_IBM), ("lc", _lib.MR_LC)):
        best = None
        for _ in range(a.reps):
            lib.mr_reset_timing(h)
            mr._check(lib.mr_topk_device(h, m, 0.5, 0, 500))
            t = {k: round(v, 2) for k, v in mr.timing().items() if v}
            if best is None or sum(t.values()) < sum(best.values()):
                best = t
        out[name] = best
    print(json.dumps(out))
