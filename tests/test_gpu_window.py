"""GPU parity tests (-m gpu) of the song partitioning (distributed.scala:459-461, 477-479: `ctx.parallelize(songs, n).map(getRanks2)`):
a handle with a song window scores all test users against its songs only — same bits as the matching columns of the unpartitioned
model — and mr_topk_merge joins the partitions' ranked lists into the exact global top-k.  Everything through the C-ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from musicrecommendation_b200 import _lib
from musicrecommendation_b200.dataset import synth, synth_config
from musicrecommendation_b200.distributed import song_window
from musicrecommendation_b200.recommender import MusicRecommender

ITEM = dict(engine=_lib.MR_ENGINE_SPARSE, space=_lib.MR_SPACE_ITEM)
BLENDS = ((_lib.MR_LC, "LC", 0.3, 0), (_lib.MR_AGG, "AGG", 0.5, 0), (_lib.MR_STOCH, "STOCH", 0.5, 42))


def assert_bits_equal(a, b):
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    np.testing.assert_array_equal(np.isnan(a), np.isnan(b))
    np.testing.assert_array_equal(np.where(np.isnan(a), 0, a).view(np.int64), np.where(np.isnan(b), 0, b).view(np.int64))


def assert_topk_equal(got, want):
    np.testing.assert_array_equal(got[2], want[2])
    np.testing.assert_array_equal(got[0], want[0])
    assert_bits_equal(got[1], want[1])


def window_topk(oracle_lib, model, lo, hi, k):
    """The oracle's ranking of the columns [lo, hi) of a dense model, with global song ids."""
    song, score, ln = oracle_lib.topk(np.ascontiguousarray(model[:, lo:hi]), k)
    return np.where(song >= 0, song + lo, song).astype(np.int32), score, ln


def merge_on_device(mr, parts, k):
    import torch
    ps = torch.from_numpy(np.stack([p[0] for p in parts])).cuda()
    pv = torch.from_numpy(np.stack([p[1] for p in parts])).cuda()
    pl = torch.from_numpy(np.stack([p[2] for p in parts])).cuda()
    n = ps.shape[1]
    os_, ov, ol = (torch.empty((n, k), dtype=torch.int32, device="cuda"), torch.empty((n, k), dtype=torch.float64, device="cuda"),
                   torch.empty(n, dtype=torch.int32, device="cuda"))
    mr.mergeTopK(ps, pv, pl, os_, ov, ol)
    torch.cuda.synchronize()
    return os_.cpu().numpy(), ov.cpu().numpy(), ol.cpu().numpy()


@pytest.mark.parametrize("shape", [dict(T=800, U=60, S=3001, seed=5, k=50, parts=3, min_deg=0),
                                   dict(T=12000, U=1040, S=30000, seed=6, k=500, parts=4, min_deg=8),
                                   dict(T=400, U=9, S=700, seed=9, k=500, parts=2, min_deg=2)])
def test_song_partitions_equal_the_unpartitioned_model(mrlib, oracle_lib, shape):
    """Every partition: dense UBM / IBM rows == the oracle's columns [lo, hi) bit for bit; top-k of UBM, IBM and the three blends (the
    Aggregation / Stochastic pair index is the global one of main.scala:57-59) == the oracle's ranking of those columns; the device join
    of the partitions' lists == the oracle's global top-k.  Shapes: windows that are no multiple of 32, head rows with exception
    entries (T = 12 000), k larger than a window holds unlistened songs (S = 700 in two parts)."""
    ds = synth(T=shape["T"], U=shape["U"], S=shape["S"], seed=shape["seed"])
    k, n_parts = shape["k"], shape["parts"]
    ubm, ibm = oracle_lib.canon_scores(ds, oracle_lib.UBM), oracle_lib.canon_scores(ds, oracle_lib.IBM)
    models = {"ubm": ubm, "ibm": ibm}
    for _, name, param, seed in BLENDS:
        models[name] = oracle_lib.blend_dense(getattr(oracle_lib, name), param, ubm, ibm, seed)
    kinds = [(_lib.MR_UBM, "ubm", 0.0, 0), (_lib.MR_IBM, "ibm", 0.0, 0)] + list(BLENDS)
    parts = {name: [] for _, name, _, _ in kinds}
    saw_exceptions = False
    for r in range(n_parts):
        lo, hi = song_window(ds.S, r, n_parts)
        with MusicRecommender(ds, head_min_deg=shape["min_deg"], song_window=(lo, hi), **ITEM) as mr:
            info = mr.info()
            assert info["n_cols"] == hi - lo and info["win_lo"] == lo and info["space"] == _lib.MR_SPACE_ITEM
            assert_bits_equal(mr.getUserBasedModel().scores, ubm[:, lo:hi])
            assert_bits_equal(mr.getItemBasedModel().scores, ibm[:, lo:hi])
            saw_exceptions |= mr.info()["head_exceptions"] > 0
            for kind, name, param, seed in kinds:
                got = mr.getTopK(kind, k=k, param=param, seed=seed)
                assert_topk_equal(got, window_topk(oracle_lib, models[name], lo, hi, k))
                parts[name].append(got)
            # per-partition granularities (DIST getRanks1 / getRanks2) inside the window
            users = np.array([0, ds.U - 1], np.int32)
            assert_bits_equal(mr.getRanks1(_lib.MR_IBM, users), ibm[users][:, lo:hi])
            songs = np.array([lo, (lo + hi) // 2, hi - 1], np.int32)
            assert_bits_equal(mr.getRanks2(_lib.MR_UBM, songs), ubm[:, songs].T)
            with pytest.raises(_lib.MrError):
                mr.getRanks2(_lib.MR_UBM, np.array([hi if hi < ds.S else lo - 1], np.int32))
            if r == n_parts - 1:
                for _, name, _, _ in kinds:
                    assert_topk_equal(merge_on_device(mr, parts[name], k), oracle_lib.topk(models[name], k))
    if shape["T"] == 12000:
        assert saw_exceptions


def test_song_window_argument_checks(mrlib):
    ds = synth(T=300, U=20, S=2000, seed=1)
    with pytest.raises(_lib.MrError):
        MusicRecommender(ds, song_window=(100, 50), **ITEM)
    with pytest.raises(_lib.MrError):
        MusicRecommender(ds, song_window=(0, 2001), **ITEM)
    with pytest.raises(_lib.MrError):   # the partitioned path exists for inverted-index counts in item space only
        MusicRecommender(ds, song_window=(0, 1000), engine=_lib.MR_ENGINE_TENSOR)
    with MusicRecommender(ds, song_window=(0, 2000)) as mr:     # the whole range is no window
        assert mr.info()["n_cols"] == 2000
    with MusicRecommender(ds, song_window=(500, 1500)) as mr:   # Gram-row probes address songs by their global id: refused under a window
        with pytest.raises(_lib.MrError):
            mr.counts_ibm(0, 10)


def test_msd_shape_song_partition(mrlib, oracle_lib):
    """BASELINE configs[3] shape, partition 3 of 8 of the songs (48 068 columns), a 2 560-user shard: top-500 of UBM, IBM and the
    Aggregation blend inside the partition bit-equal to the oracle's columns for 48 users; joined with the other columns' oracle
    lists it gives the oracle's global top-500."""
    ds = synth_config("c4").shard_test_users(0, 2560)
    lo, hi = song_window(ds.S, 3, 8)
    sub = ds.shard_test_users(0, 48)
    ubm, ibm = oracle_lib.canon_scores(sub, oracle_lib.UBM), oracle_lib.canon_scores(sub, oracle_lib.IBM)
    mask = ~np.isnan(ubm)
    agg = np.full(ubm.shape, np.nan)
    agg[mask] = oracle_lib.blend(oracle_lib.AGG, 0.5, ubm[mask], ibm[mask], 0, first_index=0, n_total=ds.n_pairs)
    with MusicRecommender(ds, song_window=(lo, hi)) as mr:
        assert mr.info()["n_cols"] == hi - lo and mr.info()["n_head"] > 30000
        for kind, model, param in ((_lib.MR_UBM, ubm, 0.0), (_lib.MR_IBM, ibm, 0.0), (_lib.MR_AGG, agg, 0.5)):
            got = tuple(a[:48] for a in mr.getTopK(kind, k=500, param=param))
            assert_topk_equal(got, window_topk(oracle_lib, model, lo, hi, 500))
            if kind != _lib.MR_AGG:
                rest = [window_topk(oracle_lib, model, a, b, 500) for a, b in ((0, lo), (hi, ds.S))]
                assert_topk_equal(merge_on_device(mr, [rest[0], got, rest[1]], 500), oracle_lib.topk(model, 500))


def test_prepare_async_overlaps_set_test_users(mrlib, oracle_lib):
    """mr_prepare_async starts the head-row build on its own stream; mr_set_test_users of the next shard runs meanwhile and the first
    scoring call completes the build — same bits as the synchronous order, also when a rebuild is abandoned half-way."""
    ds = synth(T=12000, U=1040, S=30000, seed=6)
    want = {m: oracle_lib.topk(oracle_lib.canon_scores(ds, m), 200) for m in (oracle_lib.UBM, oracle_lib.IBM)}
    with MusicRecommender(ds, head_min_deg=8, **ITEM) as mr:
        for _ in range(2):
            mr.invalidate_prepared()
            mr.prepare_async()
            mr.set_test_users(ds)
            assert_topk_equal(mr.getTopK(_lib.MR_UBM, k=200), want[oracle_lib.UBM])
            assert_topk_equal(mr.getTopK(_lib.MR_IBM, k=200), want[oracle_lib.IBM])
        assert mr.info()["head_exceptions"] > 0
        mr.invalidate_prepared()
        mr.prepare_async()
        mr.invalidate_prepared()          # abandons the build in flight
        mr.prepare_async()
        mr.prepare()                      # waits for it
        half = ds.shard_test_users(0, 520)
        mr.set_test_users(half)
        assert_topk_equal(mr.getTopK(_lib.MR_IBM, k=200), tuple(a[:520] for a in want[oracle_lib.IBM]))


def test_many_users_take_the_threaded_host_path(mrlib, oracle_lib):
    """17 000 test users: mr_set_test_users validates and rotates the rows on several host threads (>= 16 384 rows); the second of
    two song partitions still ranks to the oracle's columns, and a malformed row is still reported by its index."""
    ds = synth(T=3000, U=17000, S=5000, seed=15)
    lo, hi = song_window(ds.S, 1, 2)
    ubm = oracle_lib.canon_scores(ds, oracle_lib.UBM)
    with MusicRecommender(ds, song_window=(lo, hi), **ITEM) as mr:
        assert_topk_equal(mr.getTopK(_lib.MR_UBM, k=100), window_topk(oracle_lib, ubm, lo, hi, 100))
        bad = ds.shard_test_users(0, ds.U)
        bad.te_col = bad.te_col.copy()
        e = int(bad.te_ptr[16999])
        bad.te_col[e + 1] = bad.te_col[e]                    # row 16 999 no longer strictly ascending
        with pytest.raises(_lib.MrError, match="row 16999 not ascending"):
            mr.set_test_users(bad)
