"""BASELINE configs[4]: item-item similarity sweep with the train users split across GPUs (K-split) and a reduce-scatter per panel.

The reference's second partitioning distributes over songs (distributed.scala:459-461, 477-479: `ctx.parallelize(songs, 4).map(getRanks2)`);
the B200 form of it: each rank holds a contiguous range of TRAIN USERS — the contraction dimension of G = A^T A (MR:232-235) — computes
its partial int32 panel G_r[p0:p1, :] with kernel K1 (tcgen05 int8 count GEMM), the partial panels are summed (integers: exact in any
order) so that rank r ends up owning rows r of every panel, and the owner applies the cosine normalisation g / (sqrt d_i * sqrt d_j)
(MR:237-238, train + test-visible listener counts).  Two exchanges:

  fused   the reduce-scatter IS the GEMM epilogue: every tile is stored straight into the owner's receive slot — local HBM or a peer's
          HBM over NVLink (CUDA IPC mappings) — while the tensor cores work on the next tile.  Completion is signalled on the device:
          after its GEMM a sender raises a per-sender counter in every owner's flag block (st.release.sys), the owner's stream spins on
          its own flags (ld.acquire.sys) before it sums its `world` slots, and acknowledges consumption the same way so that the sender
          may reuse the slot two panels later (double buffer).  No host synchronisation, no barrier per panel.
  nccl    mr_gram_rows_device + one ncclReduceScatter per panel (the baseline the fused path has to beat).
  auto    fused from 40 000 songs up, nccl below (measured crossover: the fused epilogue stores 128-byte row segments per tile, which
          only pays once a panel is large enough to hide the peer-store latency behind the next tile's MMAs).
"""
from __future__ import annotations

import time

import numpy as np

from . import _lib
from .dataset import synth
from .distributed import split_train_users, reduce_scatter_rows
from .recommender import MusicRecommender

FLAG_BYTES = 4096          # flag block at the start of every rank's receive buffer: data[world] at 0, ack[world] at 512 (u64 each)
ACK_OFF = 512
FUSED_MIN_SONGS = 40000


def _dev_view(ptr, shape, typestr="<i4"):
    import torch

    class _Arr:
        __cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}
    return torch.as_tensor(_Arr(), device="cuda")


def int8_peak(device: int) -> float:
    """Measured dense int8 tensor peak of this GPU in TOP/s (cuBLASLt through torch._int_mm, 8192^3, best of 10): the roofline
    denominator of the count GEMM (SURVEY.md §8d asks for a measured int8 figure; MEASURED_PEAKS.json only holds bf16)."""
    import torch
    n = 8192
    a = torch.randint(-128, 127, (n, n), dtype=torch.int8, device=f"cuda:{device}")
    b = torch.randint(-128, 127, (n, n), dtype=torch.int8, device=f"cuda:{device}")
    for _ in range(3):
        torch._int_mm(a, b)
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch._int_mm(a, b)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


class KSplitSweep:
    """One rank of the sweep over a data set whose train users are split `world` ways."""

    def __init__(self, ds, rank: int, world: int, device: int, panel: int = 4096, mode: str = "auto"):
        import torch
        import torch.distributed as dist
        self.ds, self.rank, self.world, self.device = ds, rank, world, device
        self.mode = ("fused" if ds.S >= FUSED_MIN_SONGS else "nccl") if mode == "auto" else mode
        self.panel = max(world, panel // world * world)
        self.rpo = self.panel // world                      # rows of a panel each rank owns
        self.ld = (ds.S + 31) // 32 * 32
        shard = split_train_users(ds, rank, world)
        self.mr = MusicRecommender(shard, device=device, engine=_lib.MR_ENGINE_TENSOR, space=_lib.MR_SPACE_USER)
        self.stream = torch.cuda.ExternalStream(int(self.mr._lib.mr_stream(self.mr._h)), device=device)
        self.rs = torch.tensor(1.0 / np.sqrt(np.maximum(ds.deg_song, 1)), dtype=torch.float32, device="cuda")
        self.seq = 0                                        # panels exchanged so far (flag values keep counting across sweeps)
        self.bases = None
        if self.mode == "fused":
            self.slot_bytes = self.rpo * self.ld * 4
            self.my_ptr, my_handle = self.mr.peer_alloc(FLAG_BYTES + 2 * world * self.slot_bytes)
            handles = [None] * world
            if world > 1:
                t = torch.tensor(list(my_handle), dtype=torch.uint8, device="cuda")
                got = [torch.empty_like(t) for _ in range(world)]
                dist.all_gather(got, t)
                handles = [bytes(g.cpu().tolist()) for g in got]
            self.bases = [self.my_ptr if r == rank else self.mr.peer_open(handles[r]) for r in range(world)]
            self.recv = _dev_view(self.my_ptr + FLAG_BYTES, (2, world, self.rpo, self.ld))
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()

    def sweep(self, keep_rows: bool = False):
        """One pass over all panels.  Returns (checksum tensor, [(first row, int32 rows)] of the rows this rank owns if keep_rows)."""
        import torch
        ds, mr, world, rank, rpo, ld = self.ds, self.mr, self.world, self.rank, self.rpo, self.ld
        fused = self.mode == "fused"

        def on_stream():      # torch work of the fused mode is ordered on the library stream, between the flag kernels
            return torch.cuda.stream(self.stream if fused else torch.cuda.current_stream())
        with on_stream():
            checksum = torch.zeros((), dtype=torch.float64, device="cuda")
        kept = []
        for p0 in range(0, ds.S, self.panel):
            p1 = min(ds.S, p0 + self.panel)
            if self.mode == "fused":
                it = self.seq
                b = it & 1
                if it >= 2:       # every owner has consumed the panel that used slot b two panels ago
                    mr.peer_wait(self.my_ptr + ACK_OFF, world, it - 1)
                slots = [self.bases[o] + FLAG_BYTES + (b * world + rank) * self.slot_bytes for o in range(world)]
                mr.gram_rows_scatter(p0, p1, slots, rpo, ld, sync=False)
                mr.peer_signal([self.bases[o] + 8 * rank for o in range(world)], it + 1)          # my tiles of panel `it` have landed at owner o
                mr.peer_wait(self.my_ptr, world, it + 1)                                           # every sender's tiles have landed here
                with torch.cuda.stream(self.stream):
                    mine = self.recv[b].sum(dim=0, dtype=torch.int32)                              # rows [p0 + rank*rpo, ...) of G
                mr.peer_signal([self.bases[o] + ACK_OFF + 8 * rank for o in range(world)], it + 1)  # slot b of this rank is free again
                self.seq += 1
            else:
                torch.cuda.synchronize()      # torch-side consumers of the library-owned panel are done before it is rewritten
                part = mr.gram_rows_device(p0, p1)                   # partial panel of this rank's train users, int32 [p1-p0, ld]
                if part.shape[0] < self.panel:
                    pad = torch.zeros((self.panel - part.shape[0], part.shape[1]), dtype=part.dtype, device=part.device)
                    part = torch.cat([part, pad])
                mine = reduce_scatter_rows(part, world, rank)        # rows [p0 + rank*rpo, ...) of the full G
            r0 = p0 + rank * rpo
            rows_valid = max(0, min(rpo, p1 - r0))
            with on_stream():
                sim = mine[:rows_valid, :ds.S].to(torch.float32) * self.rs[r0:r0 + rows_valid, None] * self.rs[None, :ds.S]   # MR:237-238
                checksum += sim.sum(dtype=torch.float64)
                if keep_rows and rows_valid:
                    kept.append((r0, mine[:rows_valid, :ds.S].clone()))
        if self.mode == "fused":
            mr.sync()
        torch.cuda.synchronize()
        return checksum, kept

    def close(self):
        import torch
        import torch.distributed as dist
        if self.mode == "fused" and self.world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            for r in range(self.world):
                if r != self.rank:
                    self.mr.peer_close(self.bases[r])
            dist.barrier()
        self.mr.close()


def run_ksplit_bench(args, rank: int, world: int, device: int, log, checker=None):
    """bench.py --workload ksplit: one step = one whole sweep; returns the JSON line (rank 0) or None.  `checker(ds, first_row, n)` is the
    caller's reference for rows [first_row, first_row + n) of G (bench.py hands in the CPU oracle under --ksplit-verify)."""
    import torch
    import torch.distributed as dist
    sweeps = []
    for n_songs in args.ksplit_songs:
        T = int(round(n_songs * 909318 / 384546))            # train users scaled with S from the MSD shape
        t0 = time.time()
        ds = synth(T=T, U=64, S=n_songs, seed=20230005)
        log(f"[rank {rank}] ksplit data S={n_songs} T={T} nnz={ds.nnz_tr} generated in {time.time() - t0:.1f}s")
        sw = KSplitSweep(ds, rank, world, device, args.ksplit_panel, args.ksplit_mode)
        for _ in range(args.warmup):
            sw.sweep()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(sw.stream)
        checksum = None
        for _ in range(args.steps):
            checksum, _ = sw.sweep()
        ev1.record(sw.stream)
        ev1.synchronize()
        torch.cuda.synchronize()
        wall_ms = 1e3 * (time.perf_counter() - t0) / args.steps
        ms = torch.tensor([max(ev0.elapsed_time(ev1) / args.steps, wall_ms if sw.mode == "nccl" else 0.0)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(checksum)
        equal = None
        if checker is not None:
            _, kept = sw.sweep(keep_rows=True)
            ok = True
            for r0, rows in kept:
                got = rows.cpu().numpy()
                ok = ok and bool(np.array_equal(got, checker(ds, r0, got.shape[0])))
            flag = torch.tensor([1 if ok else 0], device="cuda")
            if world > 1:
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            equal = bool(flag.item())
        Tpad = (sw.mr.ds.T + 127) // 128 * 128
        ms_f = float(ms.item())
        # closed form of the checksum: sum_ij G_ij rs_i rs_j = sum_v (sum_{s in I_v} rs_s)^2  — checks EVERY row of every panel
        rs64 = 1.0 / np.sqrt(np.maximum(ds.deg_song, 1).astype(np.float64))
        per_user = np.add.reduceat(rs64[ds.tr_col], ds.tr_ptr[:-1].astype(np.int64))
        want_sum = float(np.sum(per_user ** 2))
        sweeps.append({"songs": n_songs, "train_users": T, "nnz": ds.nnz_tr, "mode": sw.mode, "panel_rows": sw.panel, "ms_per_sweep": ms_f,
                       "song_pairs_per_s": n_songs * n_songs / (ms_f * 1e-3),
                       "dense_int8_tops_all_gpus": 2.0 * n_songs * n_songs * Tpad * world / (ms_f * 1e-3) / 1e12,
                       "exchange_bytes_per_gpu": int(n_songs) * int(sw.ld) * 4 * (world - 1) // max(world, 1),
                       "checksum_rel_err": abs(float(checksum.item()) - want_sum) / want_sum, "every_row_equals_oracle": equal})
        launches = int(sw.mr.info()["launches"])
        log(f"[rank {rank}] ksplit {sweeps[-1]}")
        sw.close()
    if rank != 0:
        return None
    last = sweeps[-1]
    peak = None
    try:
        peak = int8_peak(device)              # measured in the same process (cuBLASLt int8)
    except Exception as e:  # noqa: BLE001
        log("int8 peak probe failed:", repr(e))
    return {"metric": "item-item similarity sweep (K-split over train users, reduce-scatter per panel): song pairs/s", "value": last["song_pairs_per_s"],
            "unit": "song pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": last["ms_per_sweep"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8 x u8 -> s32", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[4]: item-item sweep, S={last['songs']} songs, T={last['train_users']} train users split over {world} GPU(s)",
                       "mode": last["mode"], "every_row_equals_oracle": last["every_row_equals_oracle"], "checksum_rel_err": last["checksum_rel_err"]},
            "roofline": {"bound": "tensor", "achieved": last["dense_int8_tops_all_gpus"] / world, "peak": peak, "unit": "TOP/s",
                         "frac": (last["dense_int8_tops_all_gpus"] / world / peak) if peak else None, "traffic": None,
                         "peak_source": "torch._int_mm int8 8192^3 measured in this process", "kernel": "count_gemm_kernel<256, EPI_I32_SCATTER>"},
            "sweep": sweeps, "gpu_launches": launches}
