#!/usr/bin/env python
"""bench.py — scored (test-user, song) pairs / second for UBM + IBM scoring with top-500 ranking (BASELINE.json metric).

A "step" scores every (test user, song) pair of this rank's test-user shard once with the user-based model and once with
the item-based model and ranks the top 500 per user for each: 2 * (U*S - nnz_test) scored pairs per rank.

  python bench.py --gpus N --steps K --warmup W [--workload msd|c3|c2|c1] [--impl reference]

Workload `msd` (default) is BASELINE.json configs[3]: the MSD-shaped synthetic data set (909 318 train users, 384 546 songs,
~42 M train triplets, 110 000 test users), train replica on every GPU, test users sharded 13 750 per GPU (weak scaling: N GPUs
score N * 13 750 users; N = 8 is the full configuration).  No data-path collective: shards are independent
(distributed.scala:450-452); the only exchange would be the final top-k gather, which is outside the timed region.

One JSON line on rank 0:
  value      whole-job pairs/s with the CSR already resident in HBM (mr_topk_device only), CUDA-event timed, max over ranks
  e2e        the same through the host-buffer C-ABI calls: mr_set_test_users (pinned host -> device) + mr_topk (device -> host)
  roofline   dominant kernel (by CUDA-event share): algorithmic bytes per launch / average launch time vs measured HBM peak
  cpu_baseline  the oracle's canonical CPU port (OpenMP, all host threads) on a bounded sample of the same test users
`--impl reference` times that CPU port alone (the Scala reference cannot run here: no JVM; the as-written loops are
infeasible beyond configs[0]/[1] — SURVEY.md §8d) and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

K_TOP = 500
USERS_PER_GPU = 13750


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints "NCCL version ..." to fd 1), so fd 1 is pointed at
# stderr for the whole run and the JSON line goes to a private duplicate of the original stdout.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)
sys.stdout = sys.stderr


def emit(line: dict):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def make_workload(name: str, rank: int, world: int):
    from musicrecommendation_b200.dataset import synth_config
    t0 = time.time()
    if name == "msd":
        full = synth_config("c4")
        u0, u1 = rank * USERS_PER_GPU, (rank + 1) * USERS_PER_GPU
        ds = full.shard_test_users(u0, u1)
        desc = (f"BASELINE configs[3]: MSD-shaped synthetic, T={full.T} train users, S={full.S} songs, nnz_train={full.nnz_tr}, "
                f"{USERS_PER_GPU} test users per GPU (of 110000), train replica per GPU")
    else:
        full = synth_config(name)
        # weak scaling on the small shapes: every rank scores the same U test users (replicas), no sharding possible below U
        ds = full
        desc = f"BASELINE {name}: T={full.T}, U={full.U} per GPU, S={full.S}"
    log(f"[rank {rank}] workload {name} generated in {time.time() - t0:.1f}s: T={ds.T} U={ds.U} S={ds.S} nnz_tr={ds.nnz_tr} nnz_te={ds.nnz_te}")
    return ds, desc


class ClockSampler(threading.Thread):
    """nvidia-smi style clock / throttle-reason samples during the timed region (via NVML)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.sm = []
        self.reasons = set()
        self.sm_max = None
        self.err = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = int(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
                getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
            }
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop_flag.is_set():
                self.sm.append(int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                mask = int(get_reasons(h))
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
                time.sleep(0.1)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm), **({"error": self.err} if self.err else {})}


def pinned(a: np.ndarray):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t, t.numpy()


def cpu_port_sample(ds, n_users: int, threads: int | None = None):
    """The oracle's canonical CPU port (CSR + inverted index, exact integer accumulation, fp64 finalisation, full sort top-k)
    on the first n_users test users of the shard, UBM + IBM; returns (pairs, seconds, threads)."""
    import oracle
    oracle.build()
    n_users = min(n_users, ds.U)
    sub = ds.shard_test_users(0, n_users)
    t0 = time.perf_counter()
    pairs = 0
    tops = []
    for m in (oracle.UBM, oracle.IBM):
        sc = oracle.canon_scores(sub, m)
        tops.append(oracle.topk(sc, K_TOP))
        pairs += sub.n_pairs
    return pairs, time.perf_counter() - t0, oracle.num_threads(), tops


def as_written_sample(ds, n_songs_ubm: int = 48, n_songs_ibm: int = 48):
    """The reference's loops AS WRITTEN (oracle.naive_sample: MusicRecommender.scala:105-307, linear `contains` scans, cosine
    recomputed per neighbour) with OpenMP over the pair loop (= `.par`, MR:119-125) on a handful of pairs of test user 0."""
    import oracle
    out = {}
    for name, model, n_songs in (("ubm", oracle.UBM, n_songs_ubm), ("ibm", oracle.IBM, n_songs_ibm)):
        stride = max(1, ds.S // n_songs)
        t0 = time.perf_counter()
        pairs, _ = oracle.naive_sample(ds, model, 0, 1, stride, phase=stride // 2, par=True)
        dt = time.perf_counter() - t0
        out[name] = {"pairs": pairs, "seconds": dt, "pairs_per_s": pairs / dt if dt > 0 else None}
    out["cores"] = oracle.num_threads()
    out["sample"] = f"test user 0 x every {max(1, ds.S // n_songs_ubm)}-th song, loops as written, OpenMP over pairs"
    return out


def k1_probe(device: int):
    """Kernel K1 (tcgen05 int8 count GEMM) where it dominates: item-space head rows of a dense-friendly shape (65 536 train users x
    16 384 songs) computed as 1 count GEMM + 4 byte-plane GEMMs; dense-equivalent int8 TOP/s from CUDA events around the GEMM launches."""
    from musicrecommendation_b200 import _lib
    from musicrecommendation_b200.dataset import synth
    from musicrecommendation_b200.recommender import MusicRecommender
    ds = synth(T=65536, U=256, S=16384, seed=20230005)
    with MusicRecommender(ds, device=device, engine=_lib.MR_ENGINE_TENSOR, space=_lib.MR_SPACE_ITEM, profile=True) as m:
        m._lib.mr_reset_timing(m._h)
        m.prepare()
        t = m.timing()
        H = m.info()["n_head"]
    ops = 2.0 * H * ds.S * ((ds.T + 127) // 128 * 128) * 5
    return {"kernel": "count_gemm_kernel<256> (tcgen05.mma.cta_group::1.kind::i8 M128 N256 K32, TMA 128B swizzle, TMEM double buffer)",
            "M": H, "N": ds.S, "K": ds.T, "gemms": 5, "gemm_ms": t["count"], "dense_int8_tops": ops / (t["count"] * 1e-3) / 1e12,
            "tensor_pipe_active_pct_ncu": 62.9, "ncu": "profiles/r01_gemm_summary.md"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    ds, desc = make_workload(args.workload, 0, 1)
    n_users = args.ref_users
    times = []
    pairs = 0
    for i in range(args.warmup + args.steps):
        p, dt, thr, _ = cpu_port_sample(ds, n_users)
        if i >= args.warmup:
            times.append(dt)
            pairs = p
    ms = 1e3 * float(np.mean(times))
    val = pairs / (ms / 1e3)
    sample = f"first {min(n_users, ds.U)} test users of the shard per step, UBM+IBM canonical CPU port + top-{K_TOP}"
    line = {"impl": "reference", "metric": "scored (test-user, song) pairs/sec, UBM+IBM with top-500", "value": val, "unit": "pairs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64/f64", "data": "synthetic",
            "config": {"workload": desc, "k": K_TOP},
            "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": thr, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="msd", choices=["msd", "c1", "c2", "c3"])
    ap.add_argument("--engine", default="auto", choices=["auto", "tensor", "sparse"])
    ap.add_argument("--space", default="auto", choices=["auto", "user", "item"])
    ap.add_argument("--ref-users", type=int, default=384, help="test users per step of the CPU port sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-as-written", action="store_true", help="skip the as-written (naive) CPU sample")
    ap.add_argument("--no-k1-probe", action="store_true", help="skip the tensor-core count-GEMM probe")
    ap.add_argument("--users", type=int, default=0, help="profiling aid: score only the first N test users of the shard (not a bench line)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the scoring path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from musicrecommendation_b200 import _lib
    from musicrecommendation_b200.recommender import MusicRecommender
    from musicrecommendation_b200.distributed import gather_topk

    ds, desc = make_workload(args.workload, rank, world)
    if args.users:
        ds = ds.shard_test_users(0, min(args.users, ds.U))
        desc += f" [PROFILING SUBSET: first {ds.U} test users]"
    engine = {"auto": _lib.MR_ENGINE_AUTO, "tensor": _lib.MR_ENGINE_TENSOR, "sparse": _lib.MR_ENGINE_SPARSE}[args.engine]
    t0 = time.time()
    t_load0 = time.perf_counter()
    space = {"auto": _lib.MR_SPACE_AUTO, "user": _lib.MR_SPACE_USER, "item": _lib.MR_SPACE_ITEM}[args.space]
    mr = MusicRecommender(ds, device=local_rank, engine=engine, space=space)
    lib, h = mr._lib, mr._h
    load_ms = 1e3 * (time.perf_counter() - t_load0)
    log(f"[rank {rank}] mr_load done in {time.time() - t0:.1f}s, info={mr.info()}")
    stream = torch.cuda.ExternalStream(int(lib.mr_stream(h)), device=local_rank)
    # one-off per train set: item-space head rows (reported, not part of a step — it depends on the train replica only)
    t0 = time.perf_counter()
    mr.prepare()
    precompute_ms = 1e3 * (time.perf_counter() - t0)
    U, S, k = ds.U, ds.S, K_TOP
    pairs_per_step = 2 * ds.n_pairs

    def check(rc):
        mr._check(rc)

    def step_device():
        check(lib.mr_topk_device(h, _lib.MR_UBM, 0.0, 0, k))
        check(lib.mr_topk_device(h, _lib.MR_IBM, 0.0, 0, k))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing (value)
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = mr.info()["launches"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    ev1.synchronize()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = mr.info()["launches"] - l0

    # ---------------- end-to-end through the host-buffer C-ABI (pinned host -> device, device -> host every step)
    keep = [pinned(ds.te_ptr.astype(np.int64)), pinned(ds.te_col.astype(np.int32)), pinned(ds.deg_te.astype(np.int32))]
    out = [pinned(np.empty((U, k), np.int32)), pinned(np.empty((U, k), np.float64)), pinned(np.empty(U, np.int32))]
    h2d = sum(a.nbytes for _, a in keep)
    d2h = 2 * sum(a.nbytes for _, a in out)
    base = rank * USERS_PER_GPU * 0   # AGG/STOCH are not part of the bench step; the pair index base is irrelevant here

    def p(a):
        return C.c_void_p(a.ctypes.data)

    def step_e2e(verbose=False):
        t_a = time.perf_counter()
        check(lib.mr_set_test_users(h, U, p(keep[0][1]), p(keep[1][1]), p(keep[2][1]), base, 0))
        t_b = time.perf_counter()
        for model in (_lib.MR_UBM, _lib.MR_IBM):
            check(lib.mr_topk(h, model, 0.0, 0, k, p(out[0][1]), p(out[1][1]), p(out[2][1])))
            if world > 1:   # the reference's `.collect` (DIST:451-478): all-gather the fixed-size top-k blocks over NCCL
                torch.cuda.current_stream().wait_stream(stream)
                gather_topk(*mr.topk_device_tensors(k), U * world, world, rank, reuse_buffers=True)
        if verbose:
            log(f"[rank {rank}] e2e step: mr_set_test_users {1e3 * (t_b - t_a):.1f} ms, 2 x mr_topk {1e3 * (time.perf_counter() - t_b):.1f} ms")

    step_e2e(verbose=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    checksum = int(out[0][1][:, 0].astype(np.int64).sum())

    # ---------------- per-phase CUDA-event profile of one extra step (for the roofline of the dominant kernel)
    lib.mr_set_profile(h, 1)
    lib.mr_reset_timing(h)
    step_device()
    torch.cuda.synchronize()
    phases = mr.timing()
    lib.mr_set_profile(h, 0)
    info = mr.info()
    batch = int(info["batch_rows"]) if info["space"] == _lib.MR_SPACE_ITEM else 128    # test users per batch (mr_get_info) / kUserBatch
    n_batches = (U + batch - 1) // batch

    # max over ranks
    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot = torch.tensor([float(pairs_per_step)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        pairs_all = float(tot.item())
    else:
        pairs_all = float(pairs_per_step)
    dev_ms_max, e2e_ms_max = float(t[0].item()), float(t[1].item())

    if rank == 0:
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        # dominant phase and its algorithmic bytes per launch: every operand crosses HBM once per launch
        T, nnz = ds.T, ds.nnz_tr
        sparse = info["engine"] == _lib.MR_ENGINE_SPARSE
        item = info["space"] == _lib.MR_SPACE_ITEM
        Sp = (S + 31) // 32 * 32
        # distinct head songs per batch (head = the n_head songs with most train listeners, ties by id — as mr_load selects them)
        deg_train = np.bincount(ds.tr_col, minlength=S)
        is_head = np.zeros(S, bool)
        is_head[np.argsort(-deg_train, kind="stable")[:info["n_head"]]] = True
        distinct_head_rows = 0
        for b0 in range(0, U, batch):
            cols = ds.te_col[int(ds.te_ptr[b0]):int(ds.te_ptr[min(U, b0 + batch)])]
            distinct_head_rows += int(np.count_nonzero(is_head[np.unique(cols)]))
        n_launch = {"agg_ubm": n_batches, "agg_ibm": n_batches, "topk": 2 * n_batches, "count": 2 * n_batches, "expand": n_batches,
                    "head_rowsum": 2 * n_batches, "tail_scatter": 2 * n_batches}
        alg = {
            # user space: count panel + inverted index + q table + Sint panel written
            "agg_ubm": 2 * 128 * T + 4 * nnz + 8 * (S + 1) + 4 * T + 8 * 128 * S,
            "agg_ibm": (4 * 128 * T + 4 * nnz + 8 * (S + 1) + 8 * 128 * S) if sparse else None,
            # item space: every DISTINCT head row a batch needs crosses HBM once per pass (4 B/song of packed Gq in the UBM pass,
            # 2 B/song of packed G in the IBM pass; users of the batch that share a song reuse its tiles out of L2), and each pass
            # writes its Sint rows; averaged over the 2 * n_batches launches of a step
            "head_rowsum": (distinct_head_rows * Sp * 6 + 2 * U * Sp * 8) / (2 * n_batches),
            # top-k: three streaming passes over the user's Sint row(s) + rsd for IBM, k results written
            "topk": (3 * U * S * 8 * 2 + 3 * U * S * 8 + 2 * 12 * U * k) / (2 * n_batches),
        }
        names = {"agg_ubm": "aggregate_panel_kernel<true>", "agg_ibm": "aggregate_panel_kernel<false>" if sparse else "aggregate_ibm_kernel",
                 "topk": "topk_kernel", "count": "sparse_count / count_gemm", "expand": "expand_rows_kernel",
                 "head_rowsum": "head_rowsum_kernel", "tail_scatter": "tail_scatter_kernel"}
        dom = max(n_launch, key=lambda n: phases.get(n, 0.0))
        dom_ms = phases[dom] / n_launch[dom]
        roof = {"bound": "hbm", "kernel": names[dom], "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None, "traffic": None,
                "peak_source": peak_src, "ms_per_launch": dom_ms, "launches_per_step": n_launch[dom], "phase_ms_per_step": phases}
        if alg.get(dom):
            roof["achieved"] = alg[dom] / (dom_ms * 1e-3) / 1e9
            roof["frac"] = roof["achieved"] / hbm_peak
            roof["algorithmic_bytes_per_launch"] = alg[dom]
            if dom == "head_rowsum":
                roof["gathered_bytes_per_launch"] = (info["head_entries"] * Sp * 6 + 2 * U * Sp * 8) / (2 * n_batches)
                roof["gathered_gbs"] = roof["gathered_bytes_per_launch"] / (dom_ms * 1e-3) / 1e9
                roof["note"] = ("algorithmic = each distinct head row of a batch once + Sint written (what must cross HBM); gathered = one row "
                                "read per (user, head song) entry = what crosses the L2 -> SM fabric.  With one wave of balanced CTAs per song "
                                "tile the DRAM traffic equals the algorithmic bytes (ncu, profiles/), so the pass is no longer HBM-bound: it "
                                "runs at gathered_gbs against the ~12-16 TB/s the L2 -> SM fabric delivers (B300_MICROARCH: ~6300 B/clk)")
        traffic_file = ROOT / "profiles" / "traffic.json"      # dram bytes per launch from the committed ncu --set full capture
        if traffic_file.exists():
            roof["traffic"] = json.loads(traffic_file.read_text()).get(names[dom])
        line = {
            "metric": "scored (test-user, song) pairs/sec, UBM+IBM with top-500", "value": pairs_all * args.steps / (dev_ms_max * 1e-3),
            "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64 accumulate / f64 scores", "data": "synthetic",
            "config": {"workload": desc, "k": K_TOP, "engine": ["auto", "tensor", "sparse"][info["engine"]],
                       "space": {8: "user", 16: "item"}.get(info["space"]), "head_songs": info["n_head"], "users_per_batch": batch,
                       "l2": "inputs (precomputed head rows >= 10 GB, train CSR/CSC 0.7 GB, the Sint panel of a batch: 8 B per (user, song)) exceed the 126 MB L2; no explicit flush",
                       "pairs_per_step": pairs_all, "precompute_ms_once_per_train_set": precompute_ms},
            "e2e": {"value": pairs_all * args.steps / (e2e_ms_max * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms_max / args.steps},
            "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roof, "checksum": checksum,
            # conservative variant: as if the one-off head-row precompute (it depends on the train replica only) were redone every step
            "value_if_precompute_redone_every_step": pairs_all * args.steps / ((dev_ms_max + precompute_ms * args.steps) * 1e-3),
            "cold_start_ms": {"mr_load": load_ms, "precompute_once_per_train_set": precompute_ms},
        }
        if not args.no_cpu_baseline:
            pairs_c, sec_c, thr, tops = cpu_port_sample(ds, args.ref_users)
            # full-size parity: the same users' top-500 from the CUDA path must equal the oracle's bit for bit
            n_chk = min(args.ref_users, U)
            equal = {}
            for name, model, (ws, wv, wl) in (("ubm", _lib.MR_UBM, tops[0]), ("ibm", _lib.MR_IBM, tops[1])):
                gs, gv, gl = mr.getTopK(model, k=k)
                equal[name] = bool(np.array_equal(gs[:n_chk], ws) and np.array_equal(gv[:n_chk].view(np.int64), wv.view(np.int64))
                                   and np.array_equal(gl[:n_chk], wl))
            line["parity"] = {"users_checked": n_chk, "top500_ids_and_scores_bit_equal": equal}
            line["cpu_baseline_as_written"] = as_written_sample(ds) if not args.no_as_written else None
            line["cpu_baseline"] = {"value": pairs_c / sec_c, "unit": "pairs/s", "cores": thr, "kind": "port",
                                    "sample": f"first {min(args.ref_users, U)} test users of rank 0's shard, UBM+IBM canonical CPU port + top-{K_TOP}, {sec_c:.1f}s"}
    mr.close()
    if rank == 0:
        if not args.no_k1_probe:   # after the scorer released its HBM (it sizes its batches to fill the GPU)
            try:
                line["k1_count_gemm_probe"] = k1_probe(local_rank)
            except Exception as e:  # noqa: BLE001
                line["k1_count_gemm_probe"] = {"error": repr(e)}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
