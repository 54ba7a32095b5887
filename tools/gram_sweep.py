#!/usr/bin/env python
"""BASELINE configs[4]: item-item similarity sweep with a K-split across GPUs and one NCCL reduce-scatter per panel.

Each rank holds a contiguous range of TRAIN USERS (the contraction dimension K of G = A^T A), computes its partial int32 panel
G_r[p0:p1, :] with kernel K1 (tcgen05 count GEMM, or the inverted-index kernel when the dense operands do not fit), the panels are
summed with `reduce_scatter` (integers: exact in any order) so rank r owns rows r of every panel, and the owner applies the cosine
normalisation g / (sqrt(d_i) sqrt(d_j)) (MusicRecommender.scala:237-238).  Verified against the oracle on the smallest shape.

  torchrun --nproc-per-node N tools/gram_sweep.py --songs 10000 20000 50000
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

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from musicrecommendation_b200 import _lib
from musicrecommendation_b200.dataset import synth
from musicrecommendation_b200.distributed import split_train_users, reduce_scatter_rows
from musicrecommendation_b200.recommender import MusicRecommender

ap = argparse.ArgumentParser()
ap.add_argument("--songs", type=int, nargs="+", default=[10000, 20000, 50000])
ap.add_argument("--panel", type=int, default=4096)
ap.add_argument("--engine", default="tensor", choices=["tensor", "sparse"])
ap.add_argument("--fused", action="store_true", help="reduce-scatter fused into the GEMM epilogue: peer stores over NVLink (CUDA IPC)")
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
    ld = (n_songs + 31) // 32 * 32
    rpo = panel // world                                   # rows of a panel each rank owns
    if args.fused:
        # receive buffer of this rank: 2 (double buffer) x world (one slot per sender) x rpo rows x ld int32, IPC-mapped by every peer
        slot_bytes = rpo * ld * 4
        my_ptr, my_handle = mr.peer_alloc(2 * world * slot_bytes)
        handles = [None] * world
        if world > 1:
            t = torch.tensor(list(my_handle), dtype=torch.uint8, device="cuda")
            got = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(got, t)
            handles = [bytes(g.cpu().tolist()) for g in got]
        bases = [my_ptr if r == rank else mr.peer_open(handles[r]) for r in range(world)]

        class _Arr:
            def __init__(self, ptr, shape):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": "<i4", "data": (ptr, False), "version": 2}
        recv = torch.as_tensor(_Arr(my_ptr, (2, world, rpo, ld)), device="cuda")
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for it, p0 in enumerate(range(0, n_songs, panel)):
        p1 = min(n_songs, p0 + panel)
        torch.cuda.synchronize()      # torch-side consumers of library-owned / peer-visible buffers are done before they are rewritten
        if args.fused:
            b = it & 1
            # the GEMM epilogue stores row m of the partial panel into owner (m // rpo)'s slot for sender `rank`
            slots = [bases[o] + ((b * world + rank) * slot_bytes) for o in range(world)]
            mr.gram_rows_scatter(p0, p1, slots, rpo, ld)
            if world > 1:
                dist.barrier()                                   # every sender's tiles have landed (the GEMM call returns after its stream drained)
            mine = recv[b].sum(dim=0, dtype=torch.int32)         # sum of the `world` partial slots: rows [p0 + rank*rpo, ...) of G
            n_mine = rpo
        else:
            part = mr.gram_rows_device(p0, p1)                   # partial panel of this rank's train users, int32 [p1-p0, ld]
            if part.shape[0] < panel:
                pad = torch.zeros((panel - part.shape[0], part.shape[1]), dtype=part.dtype, device=part.device)
                part = torch.cat([part, pad])
            mine = reduce_scatter_rows(part, world, rank)        # rows [p0 + rank*rpo, ...) of the full G
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
        line = {"workload": f"item-item sweep S={n_songs} T={T} nnz={ds.nnz_tr}", "n_gpus": world, "engine": args.engine, "fused_epilogue_scatter": bool(args.fused), "ms": float(ms.item()),
                "dense_int8_tops_all_gpus": 2.0 * n_songs * n_songs * Tpad * world / (float(ms.item()) * 1e-3) / 1e12,
                "reduce_scatter_bytes_per_gpu": int(n_songs) * int(ld) * 4, "checksum": float(checksum.item())}
        # closed form of the checksum: sum_ij G_ij rs_i rs_j = sum_v (sum_{s in I_v} rs_s)^2  — checks EVERY row of every panel
        rs64 = 1.0 / np.sqrt(np.maximum(ds.deg_song, 1).astype(np.float64))
        per_user = np.add.reduceat(rs64[ds.tr_col], ds.tr_ptr[:-1].astype(np.int64))
        want_sum = float(np.sum(per_user ** 2))
        line["checksum_expected"] = want_sum
        line["checksum_rel_err"] = abs(line["checksum"] - want_sum) / want_sum
        if n_songs <= 20000:
            import oracle
            r0, got = verify_rows
            want = oracle.gram_rows(ds, np.arange(r0, r0 + got.shape[0]))
            line["first_rows_equal_oracle"] = bool(np.array_equal(got, want))
        print(json.dumps(line), flush=True)
    if args.fused and world > 1:
        torch.cuda.synchronize(); dist.barrier()
        for r in range(world):
            if r != rank:
                mr.peer_close(bases[r])
        dist.barrier()
    mr.close()
if world > 1:
    dist.destroy_process_group()
