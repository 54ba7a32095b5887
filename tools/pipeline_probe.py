#!/usr/bin/env python
"""Profiling aid (not a bench line): the batch pipeline of the item-space top-k (run_batches in mrscore.cu: head pass of batch b + 1 on the
library stream beside the tail scatter + select of batch b on the slice stream) against the serial order (MRSCORE_NO_PIPELINE=1), in one
process.  A: the whole one-GPU job (110 000 users, 7 batches).  B: one GPU's share of an 8-GPU song-partitioned job (all users against 1/8
of the songs) as one batch and as 2 / 3 batches.  Wall-clock between mr_sync calls, best of --reps; the ranked lists of every variant must
equal the serial ones (all users) and the CPU oracle's (first --check users).
   python tools/pipeline_probe.py > gpurun_out/pipeline_probe.json"""
import argparse, json, os, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from musicrecommendation_b200 import _lib
from musicrecommendation_b200.dataset import synth_config
from musicrecommendation_b200.distributed import song_window
from musicrecommendation_b200.recommender import MusicRecommender

ap = argparse.ArgumentParser()
ap.add_argument("--users", type=int, default=110000)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--check", type=int, default=32)
ap.add_argument("--skip-a", action="store_true")
ap.add_argument("--skip-b", action="store_true")
a = ap.parse_args()
full = synth_config("c4")
ds = full.shard_test_users(0, min(a.users, full.U))
out = {}
MODELS = (("ubm", _lib.MR_UBM), ("ibm", _lib.MR_IBM))


KNOBS = ("MRSCORE_NO_PIPELINE",)


def env(**kw):
    for k in KNOBS:
        os.environ.pop(k, None)
    for k, v in kw.items():
        os.environ["MRSCORE_" + k] = str(v)


def timed_calls(mr, fn):
    lib, h = mr._lib, mr._h
    fn()
    mr._check(lib.mr_sync(h))
    best = None
    for _ in range(a.reps):
        t0 = time.perf_counter()
        fn()
        mr._check(lib.mr_sync(h))
        dt = 1e3 * (time.perf_counter() - t0)
        best = dt if best is None else min(best, dt)
    return round(best, 2)


def measure(mr, rebuild=False, per_model=False):
    lib, h = mr._lib, mr._h
    r = {}
    if per_model:
        for name, m in MODELS:
            r[name + "_ms"] = timed_calls(mr, lambda: mr._check(lib.mr_topk_device(h, m, 0.0, 0, 500)))

    def both():
        if rebuild:
            mr._check(lib.mr_invalidate_prepared(h))
            mr._check(lib.mr_prepare(h))
        for _, m in MODELS:
            mr._check(lib.mr_topk_device(h, m, 0.0, 0, 500))
    r["job_ms" if rebuild else "both_ms"] = timed_calls(mr, both)
    return r


def lists_of(mr):
    return {name: mr.getTopK(m, k=500) for name, m in MODELS}


def same(x, y):
    return {n: bool(all(np.array_equal(p, q) for p, q in zip(x[n], y[n]))) for n in x}


import oracle
oracle.build()
oracle.set_num_threads(len(os.sched_getaffinity(0)))
sub = full.shard_test_users(0, a.check)
want = {n: oracle.canon_scores(sub, m) for n, m in (("ubm", oracle.UBM), ("ibm", oracle.IBM))}
print("oracle done", file=sys.stderr, flush=True)


def oracle_equal(lists, lo=0, hi=None):
    res = {}
    for n in ("ubm", "ibm"):
        w = want[n] if hi is None else np.ascontiguousarray(want[n][:, lo:hi])
        ws, wv, wl = oracle.topk(w, 500)
        gs, gv, gl = lists[n]
        res[n] = bool(np.array_equal(gs[:a.check], np.where(ws >= 0, ws + lo, ws)) and np.array_equal(gv[:a.check].view(np.int64), wv.view(np.int64))
                      and np.array_equal(gl[:a.check], wl))
    return res


# ---- A: the whole one-GPU job
# (the head_hi / residency-cap / carve-out / chaining variants of profiles/r02_pipeline_probe_variants.json were measured with the knobs
#  of the branch wip/pipeline-variants; the shipped library has the plain pipeline and MRSCORE_NO_PIPELINE only)
#  profiles/r02_slice_streams_probe.json (two slice streams, matched carve-outs) with those of wip/slice-streams
VARIANTS = (
    ("serial", dict(NO_PIPELINE=1)),
    ("pipelined", dict()),
)
sec = {}
with (MusicRecommender(ds, device=0, head_min_deg=150) if not a.skip_a else __import__("contextlib").nullcontext()) as mr:
    if mr is not None:
        mr.prepare()
        sec["info"] = {k: mr.info()[k] for k in ("batch_rows", "n_head", "n_cols", "split_users", "device_bytes")}
    base = None
    for tag, e in (VARIANTS if mr is not None else ()):
        env(**e)
        r = measure(mr, per_model=tag in ("serial", "pipelined"))
        if tag in ("serial", "pipelined"):
            r.update(measure(mr, rebuild=True))
        got = lists_of(mr)
        if base is None:
            base = got
            r["oracle_equal"] = oracle_equal(got)
        else:
            r["lists_equal_serial"] = same(got, base)
            r["oracle_equal"] = oracle_equal(got)
        del got
        sec[tag] = r
        print(tag, r, file=sys.stderr, flush=True)
    env()
    del base
out["A_whole_job"] = sec

# ---- B: one GPU's share of the 8-GPU song-partitioned job (all users in one batch)
if not a.skip_b:
    lo, hi = song_window(full.S, 3, 8)
    sec = {}
    base = None
    with MusicRecommender(ds, device=0, head_min_deg=125, song_window=(lo, hi)) as mr:
        mr.prepare()
        for tag, e in (("one_batch_serial", dict(NO_PIPELINE=1)), ("one_batch_alternating_slice_streams", dict()), ("one_batch_serial_again", dict(NO_PIPELINE=1))):
            env(**e)
            r = measure(mr)
            r["batch_rows"] = mr.info()["batch_rows"]
            got = lists_of(mr)
            if base is None:
                base = got
                r["oracle_equal"] = oracle_equal(got, lo, hi)
            else:
                r["lists_equal_one_batch"] = same(got, base)
            sec[tag] = r
            print(tag, r, file=sys.stderr, flush=True)
        env()
    out["B_partition_3_of_8"] = sec
print(json.dumps(out))
