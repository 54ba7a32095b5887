"""GPU parity tests (-m gpu) of the native TSV ingest (mr_ingest_tsv) against the host mirror of MR:26-91."""
import io

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from musicrecommendation_b200 import _lib
from musicrecommendation_b200.dataset import synth
from musicrecommendation_b200.recommender import MusicRecommender, dataset_from_streams, dataset_from_streams_native

FIELDS = ("tr_ptr", "tr_col", "te_ptr", "te_col", "lab_ptr", "lab_col", "deg_tr", "deg_te", "deg_song")


def tsv_of(ds, rng, shuffle=True, crlf=False):
    """The three TSV texts of a synthetic data set with string ids, rows in random order (the reference does not need them sorted)."""
    def lines(ptr, col, users):
        rows = np.repeat(np.arange(len(ptr) - 1), np.diff(ptr))
        out = [f"{users[r]}\t{ds.songs[c]}\t{1 + (i % 7)}" for i, (r, c) in enumerate(zip(rows, col))]
        if shuffle:
            rng.shuffle(out)
        return ("\r\n" if crlf else "\n").join(out) + ("\r\n" if crlf else "\n")
    return lines(ds.tr_ptr, ds.tr_col, ds.train_users), lines(ds.te_ptr, ds.te_col, ds.test_users), lines(ds.lab_ptr, ds.lab_col, ds.test_users)


def assert_same(a, b):
    assert (a.T, a.U, a.S) == (b.T, b.U, b.S)
    for f in FIELDS:
        np.testing.assert_array_equal(getattr(a, f), getattr(b, f), err_msg=f)
    assert a.train_users == b.train_users and a.test_users == b.test_users and a.songs == b.songs


@pytest.mark.parametrize("T,U,S,seed,crlf", [(300, 20, 2000, 1, False), (64, 3, 130, 2, True), (3000, 150, 9000, 5, False)])
def test_native_ingest_matches_host_mirror(mrlib, T, U, S, seed, crlf):
    ds = synth(T=T, U=U, S=S, seed=seed, with_strings=True)
    tr, te, lb = tsv_of(ds, np.random.default_rng(seed), crlf=crlf)
    want = dataset_from_streams(io.StringIO(tr, newline=""), io.StringIO(te, newline=""), io.StringIO(lb, newline=""))
    got = dataset_from_streams_native(tr.encode(), te.encode(), lb.encode())
    assert_same(got, want)
    # and the synthetic generator's own ids (sorted strings) survive the round trip through text
    for f in ("tr_ptr", "tr_col", "te_ptr", "te_col", "deg_tr", "deg_te"):
        np.testing.assert_array_equal(getattr(got, f), getattr(ds, f), err_msg=f)
    assert got.meta["timing_ms"]["total"] > 0


def test_native_ingest_edge_cases(mrlib):
    """Duplicate rows inflate the `.length` degrees but not the CSR (MR:40-41); trailing tabs are dropped by Java's split; label rows of
    unknown users are dropped; label-only songs get ids >= S; ids of different lengths order like String.compareTo; no trailing newline."""
    train = "B\ts2\t1\nA\ts10\t1\nA\ts2\t5\nA\ts2\t7\nB\ts3\t1\t\t\nC\ts3\t1\nAB\ts1\t2\n\tsX\t1"
    test = "X\ts1\t1\r\nX\ts4\t1\r\nY\ts2\t1\r\nX\ts1\t9\r\n"
    labels = "X\ts2\t1\nY\ts3\t1\nY\tzz_only_in_labels\t1\nQ\ts1\t1\nY\ts3\t1\n"
    want = dataset_from_streams(io.StringIO(train, newline=""), io.StringIO(test, newline=""), io.StringIO(labels, newline=""))
    got = dataset_from_streams_native(train.encode(), test.encode(), labels.encode())
    assert_same(got, want)
    assert got.train_users == ["", "A", "AB", "B", "C"] and got.songs[-1] == "zz_only_in_labels" and got.meta["label_only_songs"] == 1
    assert got.deg_tr.tolist() == [1, 3, 1, 2, 1] and got.deg_te.tolist() == [3, 1]          # duplicates counted
    assert np.diff(got.tr_ptr).tolist() == [1, 2, 1, 2, 1]                                     # but not stored twice
    # empty label file, single line without newline
    got = dataset_from_streams_native(b"A\ts1\t1", b"X\ts1\t1", b"")
    assert (got.T, got.U, got.S) == (1, 1, 1) and got.lab_ptr.tolist() == [0, 0]


def test_native_ingest_reports_the_malformed_line(mrlib):
    with pytest.raises(ValueError, match=r"MatchError: line 3 of the test file"):
        dataset_from_streams_native(b"A\ts1\t1\n", b"X\ts1\t1\nX\ts2\t1\nX\ts3\n", b"X\ts2\t1\n")
    with pytest.raises(ValueError, match=r"MatchError: line 2 of the train file"):
        dataset_from_streams_native(b"A\ts1\t1\n\nA\ts2\t1\n", b"X\ts1\t1\n", b"")
    with pytest.raises(ValueError, match="MatchError"):
        dataset_from_streams(io.StringIO("A\ts1\t1\n\nA\ts2\t1\n"), io.StringIO("X\ts1\t1\n"), io.StringIO(""))


def test_scoring_from_native_ingest(mrlib, oracle_lib):
    """End to end: TSV bytes -> GPU ingest -> mr_load -> scores, bit-identical to the oracle on the host-ingested data set."""
    ds = synth(T=300, U=20, S=2000, seed=1, with_strings=True)
    tr, te, lb = tsv_of(ds, np.random.default_rng(3))
    with MusicRecommender(tr.encode(), te.encode(), lb.encode(), ingest="native") as mr:
        got = mr.getUserBasedModel().scores
    want = oracle_lib.canon_scores(ds, oracle_lib.UBM)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    np.testing.assert_array_equal(np.nan_to_num(got).view(np.int64), np.nan_to_num(want).view(np.int64))
