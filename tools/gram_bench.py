#!/usr/bin/env python
"""Measure kernel K1 (tcgen05 int8 count GEMM) where it is the dominant cost: the item-space head-row precompute on a
dense-friendly shape (all operands resident as dense 0/1 u8 matrices).  Prints one JSON line with dense-equivalent int8 TOP/s.

  python tools/gram_bench.py [--train 131072] [--songs 32768]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from musicrecommendation_b200 import _lib
from musicrecommendation_b200.dataset import synth
from musicrecommendation_b200.recommender import MusicRecommender

ap = argparse.ArgumentParser()
ap.add_argument("--train", type=int, default=131072)
ap.add_argument("--songs", type=int, default=32768)
ap.add_argument("--users", type=int, default=1024)
ap.add_argument("--verify", action="store_true", help="compare top-k against the sparse-engine item-space path")
args = ap.parse_args()

t0 = time.time()
ds = synth(T=args.train, U=args.users, S=args.songs, seed=20230005)
print(f"generated T={ds.T} S={ds.S} nnz={ds.nnz_tr} in {time.time() - t0:.1f}s", file=sys.stderr)
mr = MusicRecommender(ds, engine=_lib.MR_ENGINE_TENSOR, space=_lib.MR_SPACE_ITEM, profile=True)
mr._lib.mr_reset_timing(mr._h)
t0 = time.time()
mr.prepare()
wall = time.time() - t0
t = mr.timing()
info = mr.info()
H, S, T = info["n_head"], ds.S, ds.T
Tpad = (T + 127) // 128 * 128
ops = 2.0 * H * S * Tpad * 5          # 1 count GEMM + 4 byte-plane GEMMs
line = {"kernel": "count_gemm_kernel (tcgen05.mma kind::i8, M=128 N=256 K=32, cta_group::1)", "M": H, "N": S, "K": Tpad, "gemms": 5,
        "gemm_ms": t["count"], "expand_ms": t["expand"], "wall_ms": wall * 1e3, "dense_int8_tops": ops / (t["count"] * 1e-3) / 1e12,
        "note": "bf16 cuBLAS peak on this pool 1640.9 TFLOP/s (MEASURED_PEAKS.json); nominal dense int8 4.5 POP/s"}
if args.verify:
    mr._lib.mr_set_profile(mr._h, 0)
    a = mr.getTopK(_lib.MR_LC, k=100, param=0.5)
    with MusicRecommender(ds, engine=_lib.MR_ENGINE_SPARSE, space=_lib.MR_SPACE_ITEM) as ref:
        b = ref.getTopK(_lib.MR_LC, k=100, param=0.5)
    line["matches_sparse_engine"] = bool(all(np.array_equal(x, y) for x, y in zip(a, b)))
print(json.dumps(line))
