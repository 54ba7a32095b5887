"""N > 1 host logic on CPU: world_size-2 gloo processes shard the test users, each computes its shard (with the oracle standing
in for the GPU, which this container does not have), and the all-gathered top-k equals the single-process result — including
the index-dependent Aggregation / Stochastic blends that need each shard's global pair-index base."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from musicrecommendation_b200.dataset import synth
from musicrecommendation_b200.distributed import (shard_range, pair_index_bases, gather_topk, split_train_users, reduce_scatter_rows, song_window,
                                                  exchange_song_partitions)


def merge_lists_host(parts_song, parts_score, parts_len, k):
    """Checker for the join of the song partitions' ranked lists (on the GPU: mr_topk_merge): score descending, song id ascending."""
    P, n, _ = parts_song.shape
    song = np.full((n, k), -1, np.int32); score = np.zeros((n, k)); ln = np.zeros(n, np.int32)
    for u in range(n):
        cand = [(-float(parts_score[p, u, i]), int(parts_song[p, u, i])) for p in range(P) for i in range(int(parts_len[p, u]))]
        cand.sort()
        ln[u] = min(k, len(cand))
        for i, (neg, sg) in enumerate(cand[:k]):
            song[u, i], score[u, i] = sg, -neg
    return song, score, ln


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    ds = synth(T=120, U=11, S=700, seed=33)
    u0, u1 = shard_range(ds.U, rank, world)
    shard = ds.shard_test_users(u0, u1)
    starts, n_total = pair_index_bases(ds, world)
    ubm = oracle.canon_scores(shard, oracle.UBM)
    ibm = oracle.canon_scores(shard, oracle.IBM)
    mask = ~np.isnan(ubm)
    results = {}
    for name, kind in (("ubm", None), ("agg", oracle.AGG), ("stoch", oracle.STOCH)):
        if kind is None:
            model = ubm
        else:
            model = np.full(ubm.shape, np.nan)
            model[mask] = oracle.blend(kind, 0.5, ubm[mask], ibm[mask], seed=5, first_index=int(starts[rank]), n_total=n_total)
        song, score, ln = oracle.topk(model, 50)
        g = gather_topk(song, score, ln, ds.U, world, rank)
        results[name] = [t.numpy() for t in g]
        if kind is None:   # equal shards (first 5 users of each rank): one all_gather_into_tensor per array into a reused buffer
            for _ in range(2):
                g = gather_topk(song[:5], score[:5], ln[:5], 10, world, rank, reuse_buffers=True)
            results["eq"] = [t.numpy().copy() for t in g]
    # K-split of the item-item Gram: partial panels of the ranks' train-user shards, summed and row-scattered
    part = torch.from_numpy(oracle.gram_rows(split_train_users(ds, rank, world), np.arange(0, 64)))
    mine = reduce_scatter_rows(part, world, rank)
    np.save(os.path.join(out_dir, f"gram_rows_rank{rank}.npy"), mine.numpy())
    # song partitioning (DIST:459-461): every rank ranks ALL users inside its song window, one all-to-all hands rank r every window's
    # lists of ITS users, the join gives their global top-k
    lo, hi = song_window(ds.S, rank, world)
    full_ubm = oracle.canon_scores(ds, oracle.UBM)
    wsong, wscore, wlen = oracle.topk(np.ascontiguousarray(full_ubm[:, lo:hi]), 50)
    wsong = np.where(wsong >= 0, wsong + lo, wsong).astype(np.int32)
    for _ in range(2):          # the second call reuses the receive buffers
        ps, pv, pl = exchange_song_partitions(wsong, wscore, wlen, ds.U, world, rank)
    np.savez(os.path.join(out_dir, f"partition_join_rank{rank}.npz"), **dict(zip(("song", "score", "len"), merge_lists_host(ps.numpy(), pv.numpy(), pl.numpy(), 50))))
    if rank == 0:
        np.savez(os.path.join(out_dir, "gathered.npz"), **{f"{k}_{i}": v for k, vs in results.items() for i, v in enumerate(vs)})
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    for n in (1, 7, 13750, 110000):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, i, w) for i in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_two_process_gloo_gather_matches_single_process(tmp_path, oracle_lib):
    world = 2
    mp.start_processes(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    got = np.load(tmp_path / "gathered.npz")
    ds = synth(T=120, U=11, S=700, seed=33)
    ubm = oracle_lib.canon_scores(ds, oracle_lib.UBM)
    ibm = oracle_lib.canon_scores(ds, oracle_lib.IBM)
    want = {"ubm": ubm, "agg": oracle_lib.blend_dense(oracle_lib.AGG, 0.5, ubm, ibm), "stoch": oracle_lib.blend_dense(oracle_lib.STOCH, 0.5, ubm, ibm, seed=5)}
    full = oracle_lib.gram_rows(ds, np.arange(0, 64))
    for r in range(world):
        np.testing.assert_array_equal(np.load(tmp_path / f"gram_rows_rank{r}.npy"), full[r * 32:(r + 1) * 32])
    ws, wv, wl = oracle_lib.topk(ubm, 50)
    for r in range(world):
        u0, u1 = shard_range(ds.U, r, world)
        j = np.load(tmp_path / f"partition_join_rank{r}.npz")
        np.testing.assert_array_equal(j["song"], ws[u0:u1])
        np.testing.assert_array_equal(j["score"], wv[u0:u1])
        np.testing.assert_array_equal(j["len"], wl[u0:u1])
    pick = np.r_[0:5, 6:11]                      # shard_range(11, r, 2) = [0,6), [6,11)
    np.testing.assert_array_equal(got["eq_0"], ws[pick])
    np.testing.assert_array_equal(got["eq_1"], wv[pick])
    np.testing.assert_array_equal(got["eq_2"], wl[pick])
    for name, model in want.items():
        ws, wv, wl = oracle_lib.topk(model, 50)
        np.testing.assert_array_equal(got[f"{name}_0"], ws)
        np.testing.assert_array_equal(got[f"{name}_1"], wv)
        np.testing.assert_array_equal(got[f"{name}_2"], wl)
