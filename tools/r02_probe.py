#!/usr/bin/env python
"""Profiling aid (not a bench line): facts the round-2 plan needs, measured on one B200.
   1. int8 tensor peak: torch._int_mm (cuBLASLt) on random int8 8192^3 — the denominator SURVEY 8(d) asks for
   2. head-row precompute time as a function of the head size (MRSCORE_HEAD_MIN_DEG), wall and per-phase CUDA events
   3. cudaMalloc cost of the packed head rows (first vs second prepare on the same process)
   4. per-model, per-phase step times (UBM / IBM / LC = both panels in one call)
   python tools/r02_probe.py > gpurun_out/r02_probe.json"""
import json, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from musicrecommendation_b200 import _lib
from musicrecommendation_b200.dataset import synth_config
from musicrecommendation_b200.recommender import MusicRecommender

out = {}
# ---- 1. int8 peak (library GEMM, used only as the roofline denominator)
try:
    n = 8192
    a = torch.randint(-128, 127, (n, n), dtype=torch.int8, device="cuda")
    b = torch.randint(-128, 127, (n, n), dtype=torch.int8, device="cuda")
    for _ in range(3):
        torch._int_mm(a, b)
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch._int_mm(a, b); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 200
    for _ in range(reps):
        torch._int_mm(a, b)
    e1.record(); e1.synchronize()
    out["int8_peak"] = {"how": "torch._int_mm int8 8192^3 (2*N^3 ops)", "burst_tops": 2 * n ** 3 / (best * 1e-3) / 1e12,
                        "sustained_tops": 2 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12}
    del a, b
    torch.cuda.empty_cache()
except Exception as e:  # noqa: BLE001
    out["int8_peak"] = {"error": repr(e)}
print(json.dumps(out), file=sys.stderr, flush=True)

# ---- 2./3. precompute vs head size
full = synth_config("c4")
ds = full.shard_test_users(0, 13750)
out["precompute"] = []
for min_deg in (0, 128, 400, 1500, 6000):
    if min_deg:
        os.environ["MRSCORE_HEAD_MIN_DEG"] = str(min_deg)
    else:
        os.environ.pop("MRSCORE_HEAD_MIN_DEG", None)
    mr = MusicRecommender(ds, device=0, profile=False)
    torch.cuda.synchronize(); t0 = time.perf_counter(); mr.prepare(); torch.cuda.synchronize(); wall = 1e3 * (time.perf_counter() - t0)
    info = mr.info()
    row = {"min_deg": min_deg or "default", "n_head": info["n_head"], "wall_ms": round(wall, 1),
           "head_entries": info["head_entries"], "tail_entries": info["tail_entries"]}
    if not min_deg:
        # step phases for the default head
        lib, h = mr._lib, mr._h
        lib.mr_set_profile(h, 1)
        for m in (_lib.MR_UBM, _lib.MR_IBM):
            mr._check(lib.mr_topk_device(h, m, 0.0, 0, 500))
        steps = {}
        for name, m in (("ubm", _lib.MR_UBM), ("ibm", _lib.MR_IBM), ("lc", _lib.MR_LC)):
            best = None
            for _ in range(2):
                lib.mr_reset_timing(h)
                mr._check(lib.mr_topk_device(h, m, 0.5, 0, 500))
                tt = {k: round(v, 2) for k, v in mr.timing().items() if v}
                if best is None or sum(tt.values()) < sum(best.values()):
                    best = tt
            steps[name] = best
        row["step_phase_ms"] = steps
        mr.close()
        mr = MusicRecommender(ds, device=0, profile=True)      # per-phase sums of the precompute kernels (every phase synchronised)
        mr.prepare()
        row["precompute_phase_ms"] = {k: round(v, 2) for k, v in mr.timing().items() if v}
    mr.close()
    out["precompute"].append(row)
    print(json.dumps(row), file=sys.stderr, flush=True)
os.environ.pop("MRSCORE_HEAD_MIN_DEG", None)
print(json.dumps(out))
