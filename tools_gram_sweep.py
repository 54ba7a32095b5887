#!/usr/bin/env python
"""BASELINE configs[4]: item-item similarity sweep with a K-split across GPUs and one NCCL reduce-scatter per panel.

Each rank holds a contiguous range of TRAIN USERS (the contraction dimension K of G = A^T A), computes its partial int32 panel
G_r[p0:p1, :] with kernel K1 (tcgen05 count GEMM, or the inverted-index kernel when the dense operands do not fit), the panels are
summed with `reduce_scatter` (integers: exact in any order) so rank r owns rows r of every panel, and the owner applies the cosine
normalisation g / (sqrt(d_i) sqrt(d_j)) (MusicRecommender.scala:237-238).  Verified against the oracle on the smallest shape.

  torchrun --nproc-per-node N tools_gram_sweep.py --songs 10000 20000 50000
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent))
from musicrecommendation_b200 import _lib
from musicrecommendation_b200.dataset import synth
from musicrecommendation_b200.distributed import split_train_users, reduce_scatter_rows
from musicrecommendation_b200.recommender import MusicRecommender

ap = argparse.ArgumentParser()
ap.add_argument("--songs", type=int, nargs="+", default=[10000, 20000, 50000])
ap.add_argument("--panel", type=int, default=4096)
ap.add_argument("--engine", default="tensor", choices=["tensor", "sparse"])
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
engine = _lib.MR_ENGINE_TENSOR if args.engine == "tensor" else _lib.MR_ENGINE_SPARSE

for n_songs in args.songs:
    T = int(round(n_songs * 909318 / 384546))            # train users scaled with S from the MSD shape
    ds = synth(T=T, U=64, S=n_songs, seed=20230005)
    shard = split_train_users(ds, rank, world)
    mr = MusicRecommender(shard, device=local, engine=engine, space=_lib.MR_SPACE_USER)
    rs = torch.tensor(1.0 / np.sqrt(np.maximum(ds.deg_song, 1)), dtype=torch.float32, device="cuda")
    panel = args.panel // world * world
    checksum = torch.zeros((), dtype=torch.float64, device="cuda")
    verify_rows = None
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for p0 in range(0, n_songs, panel):
        p1 = min(n_songs, p0 + panel)
        part = mr.gram_rows_device(p0, p1)                       # partial panel of this rank's train users, int32 [p1-p0, ld]
        if (p1 - p0) % world:
            pad = torch.zeros((world - (p1 - p0) % world, part.shape[1]), dtype=part.dtype, device=part.device)
            part = torch.cat([part, pad])
        mine = reduce_scatter_rows(part, world, rank)            # rows [p0 + rank*n/world, ...) of the full G
        n_mine = mine.shape[0]
        r0 = p0 + rank * n_mine
        rows_valid = max(0, min(n_mine, p1 - r0))
        sim = mine[:rows_valid, :n_songs].to(torch.float32) * rs[r0:r0 + rows_valid, None] * rs[None, :n_songs]   # MR:237-238
        checksum += sim.sum(dtype=torch.float64)
        if p0 == 0 and rank == 0:
            verify_rows = (r0, mine[:min(rows_valid, 64), :n_songs].cpu().numpy())
    ev1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(checksum)
    if rank == 0:
        Tpad = (shard.T + 127) // 128 * 128
        line = {"workload": f"item-item sweep S={n_songs} T={T} nnz={ds.nnz_tr}", "n_gpus": world, "engine": args.engine, "ms": float(ms.item()),
                "dense_int8_tops_all_gpus": 2.0 * n_songs * n_songs * Tpad * world / (float(ms.item()) * 1e-3) / 1e12,
                "reduce_scatter_bytes_per_gpu": int(n_songs) * int(part.shape[1]) * 4, "checksum": float(checksum.item())}
        if n_songs <= 20000:
            import oracle
            r0, got = verify_rows
            want = oracle.gram_rows(ds, np.arange(r0, r0 + got.shape[0]))
            line["first_rows_equal_oracle"] = bool(np.array_equal(got, want))
        print(json.dumps(line), flush=True)
    mr.close()
if world > 1:
    dist.destroy_process_group()
