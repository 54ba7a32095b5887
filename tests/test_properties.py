"""Property tests (hypothesis) of the oracle's canonical restatement — size-independent invariants the domain offers."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st, HealthCheck

from musicrecommendation_b200.dataset import from_triplets


@st.composite
def datasets(draw):
    T = draw(st.integers(2, 12))
    U = draw(st.integers(1, 5))
    S = draw(st.integers(3, 25))
    rng = np.random.default_rng(draw(st.integers(0, 2**31 - 1)))
    n_tr = draw(st.integers(T, 4 * T))
    n_te = draw(st.integers(U, 3 * U))
    tr = (np.concatenate([np.arange(T), rng.integers(0, T, n_tr)]), rng.integers(0, S, T + n_tr))
    te = (np.concatenate([np.arange(U), rng.integers(0, U, n_te)]), rng.integers(0, S, U + n_te))
    lab = (np.arange(U), rng.integers(0, S + 2, U))
    return from_triplets(tr, te, lab, T, U, S)


SETTINGS = dict(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])


@settings(**SETTINGS)
@given(datasets())
def test_canonical_equals_as_written_within_tolerance(oracle_lib, ds):
    for m in (oracle_lib.UBM, oracle_lib.IBM):
        n = oracle_lib.naive_scores(ds, m)
        c = oracle_lib.canon_scores(ds, m)
        mask = ~np.isnan(n)
        assert np.array_equal(mask, ~np.isnan(c)) and np.array_equal(mask, ~ds.listened_mask())
        np.testing.assert_allclose(c[mask], n[mask], rtol=1e-5, atol=0)          # north_star tolerance
        assert np.array_equal(c[mask] == 0, n[mask] == 0)


@settings(**SETTINGS)
@given(datasets(), st.integers(0, 2**31 - 1))
def test_relabelling_train_users_changes_nothing(oracle_lib, ds, seed):
    """Train users are only a summation index: any permutation of their ids leaves the exact integer numerators unchanged."""
    perm = np.random.default_rng(seed).permutation(ds.T)
    rows = perm[np.repeat(np.arange(ds.T), np.diff(ds.tr_ptr))]
    ds2 = from_triplets((rows, ds.tr_col), (np.repeat(np.arange(ds.U), np.diff(ds.te_ptr)), ds.te_col),
                        (np.repeat(np.arange(ds.U), np.diff(ds.lab_ptr)), ds.lab_col), ds.T, ds.U, ds.S)
    for m in (oracle_lib.UBM, oracle_lib.IBM):
        np.testing.assert_array_equal(oracle_lib.canon_sint(ds, m), oracle_lib.canon_sint(ds2, m))


@settings(**SETTINGS)
@given(datasets(), st.integers(0, 2**31 - 1))
def test_relabelling_songs_permutes_scores(oracle_lib, ds, seed):
    """Permuting song ids permutes the score columns (and nothing else): scores do not depend on the id order, only ties do."""
    perm = np.random.default_rng(seed).permutation(ds.S)
    mp = np.concatenate([perm, np.arange(ds.S, ds.S + 8)])
    ds2 = from_triplets((np.repeat(np.arange(ds.T), np.diff(ds.tr_ptr)), mp[ds.tr_col]), (np.repeat(np.arange(ds.U), np.diff(ds.te_ptr)), mp[ds.te_col]),
                        (np.repeat(np.arange(ds.U), np.diff(ds.lab_ptr)), mp[ds.lab_col]), ds.T, ds.U, ds.S)
    for m in (oracle_lib.UBM, oracle_lib.IBM):
        a = oracle_lib.canon_scores(ds, m)
        b = oracle_lib.canon_scores(ds2, m)
        np.testing.assert_array_equal(np.nan_to_num(a, nan=-1.0), np.nan_to_num(b[:, perm], nan=-1.0))
        assert oracle_lib.evaluate(a, ds) == pytest.approx(oracle_lib.evaluate(b, ds2), abs=1e-12)


@settings(**SETTINGS)
@given(datasets(), st.floats(0, 1), st.integers(0, 2**40))
def test_blend_identities(oracle_lib, ds, p, seed):
    u = oracle_lib.canon_scores(ds, oracle_lib.UBM)
    i = oracle_lib.canon_scores(ds, oracle_lib.IBM)
    same = lambda a, b: np.testing.assert_array_equal(np.nan_to_num(a, nan=-1.0), np.nan_to_num(b, nan=-1.0))
    same(oracle_lib.blend_dense(oracle_lib.LC, 1.0, u, i), u)
    same(oracle_lib.blend_dense(oracle_lib.LC, 0.0, u, i), i)
    same(oracle_lib.blend_dense(oracle_lib.AGG, 0.0, u, i), u)
    same(oracle_lib.blend_dense(oracle_lib.AGG, 1.0, u, i), i)
    same(oracle_lib.blend_dense(oracle_lib.STOCH, 0.0, u, i, seed), u)        # nextFloat() < 0 never holds
    same(oracle_lib.blend_dense(oracle_lib.STOCH, p, u, u, seed), u)          # blending a model with itself is the identity
    agg = oracle_lib.blend_dense(oracle_lib.AGG, p, u, i)
    m = ~np.isnan(u)
    thr = int(p * np.count_nonzero(m))
    assert np.array_equal(agg[m][:thr], i[m][:thr]) and np.array_equal(agg[m][thr:], u[m][thr:])


@settings(**SETTINGS)
@given(datasets(), st.integers(1, 30))
def test_topk_is_a_sorted_prefix_of_the_full_ranking(oracle_lib, ds, k):
    sc = oracle_lib.canon_scores(ds, oracle_lib.IBM)
    song, val, ln = oracle_lib.topk(sc, k)
    full_song, full_val, full_ln = oracle_lib.topk(sc, ds.S)
    for u in range(ds.U):
        n = ln[u]
        assert n == min(k, ds.S - (ds.te_ptr[u + 1] - ds.te_ptr[u]))
        assert np.array_equal(song[u, :n], full_song[u, :n]) and np.all(song[u, n:] == -1)
        keys = list(zip((-val[u, :n]).tolist(), song[u, :n].tolist()))
        assert keys == sorted(keys)                                         # score descending, then song id ascending
        assert not np.isin(song[u, :n], ds.te_col[ds.te_ptr[u]:ds.te_ptr[u + 1]]).any()
