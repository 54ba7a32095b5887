"""GPU parity tests (-m gpu), second set: the MSD-shaped configuration, the K-split Gram scatter, mAP@500, the per-partition
granularities of distributed.scala, the two head-row construction paths and the parameter / degree guards — all through the C-ABI."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from musicrecommendation_b200 import _lib
from musicrecommendation_b200.dataset import synth, synth_config
from musicrecommendation_b200.recommender import MusicRecommender, ParameterRange

ROOT = Path(__file__).resolve().parent.parent
ITEM = dict(engine=_lib.MR_ENGINE_SPARSE, space=_lib.MR_SPACE_ITEM)


def assert_bits_equal(a, b):
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    np.testing.assert_array_equal(np.isnan(a), np.isnan(b))
    np.testing.assert_array_equal(np.where(np.isnan(a), 0, a).view(np.int64), np.where(np.isnan(b), 0, b).view(np.int64))


def assert_topk_equal(got, want):
    np.testing.assert_array_equal(got[2], want[2])
    np.testing.assert_array_equal(got[0], want[0])
    assert_bits_equal(got[1], want[1])


BLENDS = ((_lib.MR_LC, "LC", 0.5, 0), (_lib.MR_AGG, "AGG", 0.5, 0), (_lib.MR_STOCH, "STOCH", 0.5, 42))


def test_msd_shape_shard(mrlib, oracle_lib):
    """BASELINE configs[3] shape (909 318 train users, 384 546 songs, 42 M triplets): a 2 560-user shard on the default engine (item
    space, sparse counts, default head of ~38 k songs built by both construction paths, users split over work groups), top-500 of
    UBM, IBM and the three blends bit-equal to the oracle on the first and the last 40 users of the shard; then the same users as a
    second, differently sized shard of the same handle."""
    full = synth_config("c4")
    ds = full.shard_test_users(0, 2560)
    n_total = ds.n_pairs
    per_user = ds.S - np.diff(ds.te_ptr)
    base = np.concatenate([[0], np.cumsum(per_user)])
    subsets = ((0, 40), (ds.U - 40, ds.U))
    want = {}
    for u0, u1 in subsets:
        sub = ds.shard_test_users(u0, u1)
        cu, ci = oracle_lib.canon_scores(sub, oracle_lib.UBM), oracle_lib.canon_scores(sub, oracle_lib.IBM)
        want[(u0, "ubm")] = oracle_lib.topk(cu, 500)
        want[(u0, "ibm")] = oracle_lib.topk(ci, 500)
        mask = ~np.isnan(cu)
        for kind, name, param, seed in BLENDS:
            okind = getattr(oracle_lib, name)
            b = np.full(cu.shape, np.nan)
            b[mask] = oracle_lib.blend(okind, param, cu[mask], ci[mask], seed, first_index=int(base[u0]), n_total=n_total)
            want[(u0, name)] = oracle_lib.topk(b, 500)
    with MusicRecommender(ds) as mr:
        info = mr.info()
        assert info["space"] == _lib.MR_SPACE_ITEM and info["engine"] == _lib.MR_ENGINE_SPARSE
        assert info["n_head"] > 30000 and info["split_users"] > 0
        models = [(_lib.MR_UBM, "ubm", 0.0, 0), (_lib.MR_IBM, "ibm", 0.0, 0)] + list(BLENDS)
        for kind, name, param, seed in models:
            got = mr.getTopK(kind, k=500, param=param, seed=seed)
            for u0, u1 in subsets:
                assert_topk_equal(tuple(a[u0:u1] for a in got), want[(u0, name)])
        # the last 40 users as their own shard (pair index base carried over): same rows again
        u0, u1 = subsets[1]
        mr.set_test_users(ds.shard_test_users(u0, u1), int(base[u0]), n_total)
        for kind, name, param, seed in models:
            assert_topk_equal(mr.getTopK(kind, k=500, param=param, seed=seed), want[(u0, name)])


def test_head_rows_direct_equals_staged(mrlib, oracle_lib, monkeypatch):
    """The in-place construction of the head rows (32-bit atomics into the final u16 / u32 arrays) and the staged one (64-bit
    accumulators + pack) give the same rows: both score to the oracle's bits, also after mr_invalidate_prepared + a rebuild."""
    ds = synth(T=6000, U=1100, S=9000, seed=31)
    want = {"ubm": oracle_lib.canon_scores(ds, oracle_lib.UBM), "ibm": oracle_lib.canon_scores(ds, oracle_lib.IBM)}
    wtop = {k: oracle_lib.topk(v, 300) for k, v in want.items()}
    # chunk_mb = "1": ~50 chunks of in-place rows, rotating over the four build streams (default) or all on one (MRSCORE_PRE_STREAMS=1)
    for stage_all, chunk_mb, streams in ((False, None, None), (False, "1", None), (False, "1", "1"), (True, None, None)):
        if stage_all:
            monkeypatch.setenv("MRSCORE_PRECOMPUTE_STAGE_ALL", "1")
        if chunk_mb:
            monkeypatch.setenv("MRSCORE_DIRECT_CHUNK_MB", chunk_mb)
        if streams:
            monkeypatch.setenv("MRSCORE_PRE_STREAMS", streams)
        with MusicRecommender(ds, head_min_deg=2, **ITEM) as mr:
            assert mr.info()["n_head"] > 1000
            for rebuild in (False, True):
                if rebuild:
                    mr.invalidate_prepared()
                    mr.prepare()
                assert_bits_equal(mr.getUserBasedModel().scores, want["ubm"])
                assert_bits_equal(mr.getItemBasedModel().scores, want["ibm"])
                assert_topk_equal(mr.getTopK(_lib.MR_UBM, k=300), wtop["ubm"])
                assert_topk_equal(mr.getTopK(_lib.MR_IBM, k=300), wtop["ibm"])
        monkeypatch.delenv("MRSCORE_DIRECT_CHUNK_MB", raising=False)
        monkeypatch.delenv("MRSCORE_PRE_STREAMS", raising=False)


def test_head_min_deg_option_changes_nothing_but_the_split(mrlib, oracle_lib):
    ds = synth(T=3000, U=1050, S=5000, seed=32)
    want = oracle_lib.topk(oracle_lib.canon_scores(ds, oracle_lib.IBM), 200)
    heads = []
    for min_deg in (0, 3, 40, 100000):
        with MusicRecommender(ds, head_min_deg=min_deg, **ITEM) as mr:
            heads.append(mr.info()["n_head"])
            assert_topk_equal(mr.getTopK(_lib.MR_IBM, k=200), want)
    assert heads[1] > heads[2] > heads[3] == 0


@pytest.mark.parametrize("cfg", [dict(engine=_lib.MR_ENGINE_TENSOR, space=_lib.MR_SPACE_USER), ITEM], ids=["tensor-userspace", "sparse-itemspace"])
def test_map_at_k(mrlib, oracle_lib, cfg):
    """mAP@k of the ranked lists: per-user AP and the mean, bit-equal to the CPU restatement, from host lists and from the
    device-resident result of the last top-k call."""
    ds = synth(T=800, U=140, S=4000, seed=33)
    with MusicRecommender(ds, **cfg) as mr:
        for kind, m in ((_lib.MR_UBM, oracle_lib.UBM), (_lib.MR_IBM, oracle_lib.IBM)):
            for k in (500, 37):
                song, score, ln = mr.getTopK(kind, k=k)
                want, want_ap = oracle_lib.map_at_k(song, ln, ds, per_user=True)
                got, got_ap = mr.mapAtK(k, top=(song, ln), per_user=True)
                assert got == want and want > 0
                np.testing.assert_array_equal(got_ap.view(np.int64), want_ap.view(np.int64))
                assert mr.mapAtK(k) == want                       # device-resident lists of the call above
        # users without label rows are skipped, not counted as zeros
        ds2 = synth(T=800, U=140, S=4000, seed=33)
        keep = np.ones(ds2.U, bool); keep[::3] = False
        cnt = np.where(keep, np.diff(ds2.lab_ptr), 0)
        ds2.lab_col = np.concatenate([ds2.lab_col[ds2.lab_ptr[u]:ds2.lab_ptr[u + 1]] for u in range(ds2.U) if keep[u]])
        ds2.lab_ptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
        song, score, ln = mr.getTopK(_lib.MR_UBM, k=100)
        mr.ds = ds2
        assert mr.mapAtK(100, top=(song, ln)) == oracle_lib.map_at_k(song, ln, ds2)


def test_map_at_500_msd_shard_ranking_quality(mrlib, oracle_lib):
    """mAP@500 on an MSD-shaped shard is the same number from the CUDA lists and from the oracle's lists (north_star: identical)."""
    ds = synth_config("c4").shard_test_users(0, 1100)
    sub = ds.shard_test_users(0, 48)
    with MusicRecommender(ds) as mr:
        for kind, m in ((_lib.MR_UBM, oracle_lib.UBM), (_lib.MR_IBM, oracle_lib.IBM)):
            song, score, ln = mr.getTopK(kind, k=500)
            ws, wv, wl = oracle_lib.topk(oracle_lib.canon_scores(sub, m), 500)
            assert oracle_lib.map_at_k(song[:48], ln[:48], sub) == oracle_lib.map_at_k(ws, wl, sub)
            assert mr.mapAtK(500) == oracle_lib.map_at_k(song, ln, ds)


@pytest.mark.parametrize("cfg", [dict(engine=_lib.MR_ENGINE_TENSOR, space=_lib.MR_SPACE_USER), dict(engine=_lib.MR_ENGINE_SPARSE, space=_lib.MR_SPACE_USER), ITEM],
                         ids=["tensor-userspace", "sparse-userspace", "sparse-itemspace"])
def test_score_users_and_songs(mrlib, oracle_lib, cfg):
    """DIST getRanks1(user) / getRanks2(song) granularities (DIST:198-221, 269-292) = rows / columns of the dense model."""
    ds = synth(T=500, U=150, S=2500, seed=34)
    want = {_lib.MR_UBM: oracle_lib.canon_scores(ds, oracle_lib.UBM), _lib.MR_IBM: oracle_lib.canon_scores(ds, oracle_lib.IBM)}
    users = [149, 0, 128, 127, 5, 5]
    songs = [0, 2499, 1234, 77, 77]
    with MusicRecommender(ds, **cfg) as mr:
        for kind, w in want.items():
            assert_bits_equal(mr.getRanks1(kind, users), w[users])
            assert_bits_equal(mr.getRanks2(kind, songs), w[:, songs].T)
        assert mr.getRanks1(_lib.MR_UBM, []).shape == (0, ds.S)
        with pytest.raises(_lib.MrError, match="out of range"):
            mr.getRanks1(_lib.MR_UBM, [150])
        with pytest.raises(_lib.MrError, match="out of range"):
            mr.getRanks2(_lib.MR_IBM, [-1])


def test_parameter_and_degree_guards(mrlib, oracle_lib):
    ds = synth(T=64, U=3, S=130, seed=2)
    with MusicRecommender(ds) as mr:
        ubm, ibm = mr.getUserBasedModel(), mr.getItemBasedModel()
        # the materialised linear combination follows the reference (no range check on alpha, MR:317-330) ...
        assert_bits_equal(mr.getLinearCombinationModel(ubm, ibm, 1.5).scores,
                          oracle_lib.blend_dense(oracle_lib.LC, 1.5, ubm.scores, ibm.scores))
        # ... but a ranking of possibly negative blends is refused, and NaN never passes a range check
        for bad in (1.5, -0.25, float("nan")):
            with pytest.raises(ParameterRange, match="alpha"):
                mr.getTopK(_lib.MR_LC, k=10, param=bad)
        for kind in (_lib.MR_AGG, _lib.MR_STOCH):
            with pytest.raises(ParameterRange, match="between 0 and 1"):
                mr.getTopK(kind, k=10, param=float("nan"))
        with pytest.raises(ParameterRange):
            mr.getAggregationModel(ubm, ibm, float("nan"))
        for ok in (0.0, 1.0):
            want = oracle_lib.topk(oracle_lib.blend_dense(oracle_lib.LC, ok, ubm.scores, ibm.scores), 10)
            assert_topk_equal(mr.getTopK(_lib.MR_LC, k=10, param=ok), want)
    # degrees beyond the fixed-point tolerance are refused at load time
    big = synth(T=64, U=3, S=130, seed=2)
    big.deg_tr = big.deg_tr.copy(); big.deg_tr[5] = 112590
    with pytest.raises(_lib.MrError, match="1e-5"):
        MusicRecommender(big)
    ok = synth(T=64, U=3, S=130, seed=2)
    ok.deg_tr = ok.deg_tr.copy(); ok.deg_tr[5] = 112589
    with MusicRecommender(ok) as mr:
        assert_bits_equal(mr.getUserBasedModel().scores, oracle_lib.canon_scores(ok, oracle_lib.UBM))
    big = synth(T=64, U=3, S=130, seed=2)
    big.deg_song = big.deg_song.copy(); big.deg_song[7] = 1801440
    with pytest.raises(_lib.MrError, match="1e-5"):
        MusicRecommender(big)


def test_many_head_entries_per_group_split_batches(mrlib, oracle_lib, monkeypatch):
    """A shard whose work groups would not fit the 48 KB staging area of the head pass is split into more batches instead of failing."""
    monkeypatch.setenv("MRSCORE_HEAD_GROUPS", "2")
    ds = synth(T=1500, U=1300, S=1200, seed=35)
    with MusicRecommender(ds, head_min_deg=1, **ITEM) as mr:
        info = mr.info()
        assert info["batch_rows"] < ds.U                      # the 1300-user shard did not stay one batch
        assert_topk_equal(mr.getTopK(_lib.MR_UBM, k=100), oracle_lib.topk(oracle_lib.canon_scores(ds, oracle_lib.UBM), 100))
        assert_topk_equal(mr.getTopK(_lib.MR_IBM, k=100), oracle_lib.topk(oracle_lib.canon_scores(ds, oracle_lib.IBM), 100))


# ---------------------------------------------------------------------------------------------------------------- configs[4]
def _dev_view(ptr, shape, typestr="<i4"):
    import torch

    class _Arr:
        __cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}
    return torch.as_tensor(_Arr(), device="cuda")


@pytest.mark.parametrize("n_owners", [1, 2, 4])
def test_gram_rows_scatter_self_slots(mrlib, oracle_lib, n_owners):
    """BASELINE configs[4], one GPU standing in for every owner: mr_gram_rows_scatter (reduce-scatter fused into the count-GEMM
    epilogue) stores row m of the panel into owner (m // rows_per_owner)'s slot; with local slots from mr_peer_alloc every row of every
    panel must equal the oracle's Gram row, pad rows and pad columns stay zero, and the plain mr_gram_rows_device agrees."""
    import torch
    ds = synth(T=3000, U=8, S=2200, seed=36)
    ld = (ds.S + 31) // 32 * 32
    panel = 1024
    rpo = panel // n_owners
    want = oracle_lib.gram_rows(ds, np.arange(ds.S))
    with MusicRecommender(ds, engine=_lib.MR_ENGINE_TENSOR, space=_lib.MR_SPACE_USER) as mr:
        ptrs = [mr.peer_alloc(rpo * ld * 4)[0] for _ in range(n_owners)]
        for p0 in range(0, ds.S, panel):
            p1 = min(ds.S, p0 + panel)
            for p in ptrs:
                _dev_view(p, (rpo, ld)).zero_()
            torch.cuda.synchronize()
            mr.gram_rows_scatter(p0, p1, ptrs, rpo, ld)
            got = torch.cat([_dev_view(p, (rpo, ld)) for p in ptrs]).cpu().numpy()
            np.testing.assert_array_equal(got[:p1 - p0, :ds.S], want[p0:p1])
            assert not got[p1 - p0:].any() and not got[:, ds.S:].any()
            plain = mr.gram_rows_device(p0, p1).cpu().numpy()
            np.testing.assert_array_equal(plain[:p1 - p0, :ds.S], want[p0:p1])
        with pytest.raises(_lib.MrError):
            mr.gram_rows_scatter(0, panel, ptrs, rpo // 2 if rpo > 1 else 0, ld)     # slots too small for the panel


@pytest.mark.parametrize("mode", ["fused", "nccl"])
def test_ksplit_sweep_single_rank(mrlib, oracle_lib, mode):
    """The sweep driver of bench.py --workload ksplit with one rank: the device-flag protocol of the fused exchange (signal / wait / ack,
    double-buffered slots, sequence numbers running across two sweeps) must leave every row of G equal to the oracle's."""
    from musicrecommendation_b200.ksplit import KSplitSweep
    ds = synth(T=4000, U=8, S=3000, seed=37)
    want = oracle_lib.gram_rows(ds, np.arange(ds.S))
    sw = KSplitSweep(ds, 0, 1, 0, panel=512, mode=mode)
    try:
        for _ in range(2):
            checksum, kept = sw.sweep(keep_rows=True)
            got = np.concatenate([rows.cpu().numpy() for _, rows in kept])
            assert [r0 for r0, _ in kept] == list(range(0, ds.S, 512))
            np.testing.assert_array_equal(got, want)
            rs = 1.0 / np.sqrt(np.maximum(ds.deg_song, 1).astype(np.float64))
            per_user = np.add.reduceat(rs[ds.tr_col], ds.tr_ptr[:-1].astype(np.int64))
            assert abs(float(checksum.item()) - float(np.sum(per_user ** 2))) < 1e-5 * float(np.sum(per_user ** 2))
    finally:
        sw.close()


def test_ksplit_two_ranks_under_torchrun(mrlib):
    """configs[4] with a real exchange: two ranks, each holding half of the train users, K-split Gram panels summed by the fused
    epilogue scatter over CUDA IPC peer memory and by NCCL reduce-scatter; every row is checked against the oracle inside the job."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    for mode in ("fused", "nccl"):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29533",
               str(ROOT / "bench.py"), "--workload", "ksplit", "--ksplit-songs", "6000", "--ksplit-mode", mode, "--ksplit-verify", "--steps", "1", "--warmup", "1"]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="4"))
        assert out.returncode == 0, out.stderr[-2000:]
        import json
        line = json.loads(out.stdout.strip().splitlines()[-1])
        assert line["config"]["every_row_equals_oracle"] is True


def test_batch_pipeline_equals_the_serial_order(mrlib, oracle_lib, monkeypatch):
    """Item-space top-k of a pure model over several batches runs as a two-stream pipeline (head pass of batch b + 1 beside the slices of
    batch b, consecutive batches in alternating Sint panels; run_batches in mrscore.cu).  Device-resident calls back to back, in both
    orders, with a blend (both panels), a dense probe and a host-buffer call in between: always the oracle's lists, and the same lists as
    the serial order (MRSCORE_NO_PIPELINE)."""
    import ctypes as C
    monkeypatch.setenv("MRSCORE_ITEM_BATCH", "600")
    ds = synth(T=3000, U=2500, S=20000, seed=31)     # 5 batches of 500 users; rows long enough for the sampled select
    k = 200
    ubm, ibm = oracle_lib.canon_scores(ds, oracle_lib.UBM), oracle_lib.canon_scores(ds, oracle_lib.IBM)
    want = {_lib.MR_UBM: oracle_lib.topk(ubm, k), _lib.MR_IBM: oracle_lib.topk(ibm, k)}
    want_lc = oracle_lib.topk(oracle_lib.blend_dense(oracle_lib.LC, 0.5, ubm, ibm, 0), k)
    with MusicRecommender(ds, **ITEM) as mr:
        assert mr.info()["batch_rows"] == 500
        lib, h = mr._lib, mr._h

        def device_topk(model):
            mr._check(lib.mr_topk_device(h, model, 0.0, 0, k))
            song, score, ln = np.empty((ds.U, k), np.int32), np.empty((ds.U, k), np.float64), np.empty(ds.U, np.int32)
            mr._check(lib.mr_topk_fetch(h, k, song.ctypes.data_as(C.c_void_p), score.ctypes.data_as(C.c_void_p), ln.ctypes.data_as(C.c_void_p)))
            return song, score, ln

        for order in ((_lib.MR_UBM, _lib.MR_IBM, _lib.MR_IBM, _lib.MR_UBM), (_lib.MR_IBM, _lib.MR_UBM)):
            for model in order:
                mr._check(lib.mr_topk_device(h, model, 0.0, 0, k))      # back to back: the next call starts while nothing has been fetched
            assert_topk_equal(device_topk(order[-1]), want[order[-1]])
            for model in order:
                assert_topk_equal(device_topk(model), want[model])
            assert_topk_equal(mr.getTopK(_lib.MR_LC, k=k, param=0.5), want_lc)          # both panels, serial path
            assert_bits_equal(mr.getRanks1(_lib.MR_IBM, [3, 1700]), ibm[[3, 1700]])     # dense probe reuses the panels
            assert_topk_equal(mr.getTopK(_lib.MR_UBM, k=k), want[_lib.MR_UBM])          # host buffers: slices copied out on the copy stream
        monkeypatch.setenv("MRSCORE_NO_PIPELINE", "1")
        for model in (_lib.MR_UBM, _lib.MR_IBM):
            assert_topk_equal(device_topk(model), want[model])
