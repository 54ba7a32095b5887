#!/usr/bin/env python
"""Profiling aid (not a bench line): native GPU ingest (mr_ingest_tsv) of an MSD-shaped synthetic TSV set, stage times in ms, checked
against the generator's own CSR.   python tools/ingest_bench.py [--config c4] [--host-sample 200000]"""
import argparse, io, json, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from musicrecommendation_b200.dataset import synth_config
from musicrecommendation_b200.recommender import dataset_from_streams, dataset_from_streams_native


def tsv_bytes(ptr, col, users, songs):
    """Fixed-width records `user \\t song \\t 1 \\n` built with numpy (users / songs are equal-length ASCII ids)."""
    u = np.array(users, dtype="S"); s = np.array(songs, dtype="S")
    lu, ls = u.dtype.itemsize, s.dtype.itemsize
    rows = np.repeat(np.arange(len(ptr) - 1), np.diff(ptr))
    n = len(rows)
    rec = np.empty((n, lu + ls + 4), np.uint8)
    rec[:, :lu] = u.view(np.uint8).reshape(-1, lu)[rows]
    rec[:, lu] = 9
    rec[:, lu + 1:lu + 1 + ls] = s.view(np.uint8).reshape(-1, ls)[col]
    rec[:, lu + 1 + ls] = 9
    rec[:, lu + 2 + ls] = ord("1")
    rec[:, lu + 3 + ls] = 10
    return rec.tobytes()


ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c4")
ap.add_argument("--host-sample", type=int, default=200000, help="train lines given to the Python host mirror for a per-line rate")
a = ap.parse_args()
t0 = time.time()
ds = synth_config(a.config, with_strings=True)
tr = tsv_bytes(ds.tr_ptr, ds.tr_col, ds.train_users, ds.songs)
te = tsv_bytes(ds.te_ptr, ds.te_col, ds.test_users, ds.songs)
lb = tsv_bytes(ds.lab_ptr, ds.lab_col, ds.test_users, ds.songs)
gen_s = time.time() - t0
out = {"config": a.config, "T": ds.T, "U": ds.U, "S": ds.S, "lines": int(ds.tr_ptr[-1] + ds.te_ptr[-1] + ds.lab_ptr[-1]),
       "bytes": len(tr) + len(te) + len(lb), "generate_s": round(gen_s, 1)}
best = None
for _ in range(2):
    t0 = time.perf_counter()
    got = dataset_from_streams_native(tr, te, lb, with_strings=False)
    wall = time.perf_counter() - t0
    if best is None or wall < best[0]:
        best = (wall, got.meta["timing_ms"])
out["native_wall_s"] = round(best[0], 3)
out["native_stage_ms"] = {k: round(v, 1) for k, v in best[1].items()}
out["lines_per_s"] = out["lines"] / best[0]
ok = all(np.array_equal(getattr(got, f), getattr(ds, f)) for f in ("tr_ptr", "tr_col", "te_ptr", "te_col", "lab_ptr", "deg_tr", "deg_te", "deg_song"))   # label-only song ids are numbered differently by the generator
out["equals_generator_csr"] = bool(ok and (got.T, got.U, got.S) == (ds.T, ds.U, ds.S))
# host mirror (Python, as the JVM-less stand-in for MR:26-91) on a sample, for a per-line rate
n = min(a.host_sample, int(ds.tr_ptr[-1]))
reclen = len(tr) // int(ds.tr_ptr[-1])
t0 = time.perf_counter()
dataset_from_streams(io.StringIO(tr[:n * reclen].decode()), io.StringIO(te[:1000 * reclen].decode()), io.StringIO(""))
out["host_mirror_lines_per_s"] = n / (time.perf_counter() - t0)
print(json.dumps(out))
