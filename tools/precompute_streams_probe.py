#!/usr/bin/env python
"""Profiling aid (not a bench line): the in-place head-row build (start_head_rows in mrscore.cu) with its L2-sized chunks on one stream or
alternating between two (MRSCORE_PRE_STREAMS) at several chunk sizes (MRSCORE_DIRECT_CHUNK_MB), MSD-shaped train set, one process.
Wall-clock of mr_invalidate_prepared + mr_prepare + mr_sync, best of --reps; after every variant the dense UBM and IBM rows of 128 test
users (~4 000 distinct head rows, every column) must equal the first variant's bit for bit.
   python tools/precompute_streams_probe.py > gpurun_out/precompute_streams_probe.json"""
import argparse, json, os, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from musicrecommendation_b200 import _lib
from musicrecommendation_b200.dataset import synth_config
from musicrecommendation_b200.recommender import MusicRecommender

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--variants", default="1:48,2:48,2:32,2:24,2:16,1:24,1:48", help="streams:chunk MB, comma separated; the first is the reference")
a = ap.parse_args()
VARIANTS = [tuple(int(x) for x in v.split(":")) for v in a.variants.split(",")]
ds = synth_config("c4").shard_test_users(0, 2048)
users = np.arange(0, 2048, 16, dtype=np.int32)
out = {}


def bits(x):
    return np.where(np.isnan(x), 0, x).view(np.int64)


with MusicRecommender(ds, device=0, head_min_deg=150) as mr:
    lib, h = mr._lib, mr._h
    mr.prepare()
    out["info"] = {k: mr.info()[k] for k in ("n_head", "n_cols", "head_exceptions")}
    ref = None
    for streams, mb in VARIANTS:
        os.environ["MRSCORE_PRE_STREAMS"] = str(streams)
        os.environ["MRSCORE_DIRECT_CHUNK_MB"] = str(mb)
        best = None
        for _ in range(a.reps):
            mr._check(lib.mr_invalidate_prepared(h))
            t0 = time.perf_counter()
            mr._check(lib.mr_prepare(h))
            mr._check(lib.mr_sync(h))
            dt = 1e3 * (time.perf_counter() - t0)
            best = dt if best is None else min(best, dt)
        got = [bits(mr.getRanks1(m, users)) for m in (_lib.MR_UBM, _lib.MR_IBM)]
        r = {"precompute_ms": round(best, 2)}
        if ref is None:
            ref = got
        else:
            r["dense_rows_equal_first"] = bool(all(np.array_equal(x, y) for x, y in zip(got, ref)))
        out[f"streams{streams}_chunk{mb}mb" + ("_again" if f"streams{streams}_chunk{mb}mb" in out else "")] = r
        print(streams, mb, r, file=sys.stderr, flush=True)
print(json.dumps(out))
