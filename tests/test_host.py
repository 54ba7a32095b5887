"""CPU tests of the host-side mirror and of the C-ABI library as a loadable artefact (no compute calls without a GPU)."""
import ctypes as C
import io
import re
from pathlib import Path

import numpy as np
import pytest

from musicrecommendation_b200 import _lib, recommender
from musicrecommendation_b200.dataset import fixture_4_3, synth
from musicrecommendation_b200.javafmt import double_to_string

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(mrlib):
    header = (ROOT / "include" / "mrscore.h").read_text()
    declared = set(re.findall(r"\b(mr_[a-z_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS)
    for sym in declared:
        assert hasattr(mrlib, sym), sym


def test_no_cpu_fallback_without_gpu(mrlib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    dev = (C.c_int * 1)(0)
    rc = mrlib.mr_create(C.byref(h), dev, 1, 0)
    assert rc == _lib.MR_ERR_CUDA
    assert b"no CPU fallback" in mrlib.mr_last_error(h)
    mrlib.mr_destroy(h)
    with pytest.raises(_lib.MrError):
        recommender.MusicRecommender(fixture_4_3())


def test_product_path_never_imports_oracle():
    for p in list((ROOT / "musicrecommendation_b200").rglob("*.py")) + list((ROOT / "musicrecommendation_b200").rglob("*.cu*")) \
            + list((ROOT / "musicrecommendation_b200").rglob("*.h")):
        txt = p.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt and "mr_oracle" not in txt, p


def test_tsv_ingest_matches_fixture():
    train = io.StringIO("A\ts1\t1\nA\ts2\t1\nB\ts2\t1\nB\ts3\t1\nC\ts3\t1\n")
    test = io.StringIO("X\ts1\t1\nX\ts4\t1\nY\ts2\t1\n")
    labels = io.StringIO("X\ts2\t1\nY\ts3\t1\nY\ts5\t1\n")
    ds = recommender.dataset_from_streams(train, test, labels)
    fx = fixture_4_3()
    assert (ds.T, ds.U, ds.S) == (3, 2, 4)
    for f in ("tr_ptr", "tr_col", "te_ptr", "te_col", "lab_ptr", "lab_col", "deg_tr", "deg_te", "deg_song"):
        np.testing.assert_array_equal(getattr(ds, f), getattr(fx, f))
    assert ds.songs == ["s1", "s2", "s3", "s4", "s5"] and ds.test_users == ["X", "Y"]
    with pytest.raises(ValueError, match="MatchError"):
        recommender.parse_triplets(io.StringIO("A\ts1\n"))
    # Java split drops trailing empties: "A\ts1\t1\t" still has 3 fields
    assert recommender.parse_triplets(io.StringIO("A\ts1\t1\t\n")) == (["A"], ["s1"])


def test_duplicate_rows_inflate_degrees_only():
    """MR:40-41: the per-user / per-song lists are not de-duplicated, so `.length` counts duplicates."""
    train = io.StringIO("A\ts1\t1\nA\ts1\t3\nA\ts2\t1\n")
    test = io.StringIO("X\ts1\t1\n")
    labels = io.StringIO("X\ts2\t1\n")
    ds = recommender.dataset_from_streams(train, test, labels)
    assert ds.deg_tr.tolist() == [3] and ds.tr_col.tolist() == [0, 1] and ds.deg_song.tolist() == [3, 1]


def test_java_double_to_string():
    cases = {0.4999999999999999: "0.4999999999999999", 0.0: "0.0", 1.0: "1.0", 1e7: "1.0E7", 1.234e-5: "1.234E-5",
             0.001: "0.001", 9999999.999: "9999999.999", 12345678.9: "1.23456789E7", 0.7071067811865475: "0.7071067811865475",
             123.0: "123.0", 1e-4: "1.0E-4"}
    for x, want in cases.items():
        assert double_to_string(x) == want
    for x in np.random.default_rng(0).random(200) * 10:
        assert float(double_to_string(float(x))) == float(x)


def test_host_evaluate_matches_oracle(oracle_lib):
    ds = synth(T=60, U=8, S=700, seed=21)
    for m in (oracle_lib.UBM, oracle_lib.IBM):
        sc = oracle_lib.canon_scores(ds, m)
        for nt in (10, 11):
            assert recommender.evaluate_map(sc, ds, nt) == pytest.approx(oracle_lib.evaluate(sc, ds, nt), rel=0, abs=1e-15)
    ds = fixture_4_3()
    sc = oracle_lib.naive_scores(ds, oracle_lib.UBM)
    assert round(recommender.evaluate_map(sc, ds), 10) == 0.6666666667


def test_model_tuples_order():
    ds = fixture_4_3()
    scores = np.where(ds.listened_mask(), np.nan, np.arange(8, dtype=float).reshape(2, 4))
    m = recommender.Model(scores, ds.test_users, ds.songs)
    assert [(u, s) for u, (s, _) in m.tuples()] == [("X", "s2"), ("X", "s3"), ("Y", "s1"), ("Y", "s3"), ("Y", "s4")]
    assert len(m) == 5


def test_native_ingest_has_no_cpu_fallback(mrlib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.MrError) as ei:
        recommender.dataset_from_streams_native(b"A\ts1\t1\n", b"X\ts1\t1\n", b"X\ts2\t1\n")
    assert ei.value.code == _lib.MR_ERR_CUDA and "no CPU fallback" in ei.value.msg


def test_ingest_selectors_match_the_header():
    """The MR_ING_* selectors of mr_ingest_get are positional: the Python constants must follow the header's enum order."""
    header = (ROOT / "include" / "mrscore.h").read_text()
    body = header[header.index("enum { MR_ING_TR_PTR"):]
    body = body[:body.index("};")]
    names = re.findall(r"\bMR_ING_[A-Z_]+", re.sub(r"/\*.*?\*/", "", body, flags=re.S))
    assert len(names) == 16 and len(set(names)) == 16
    for i, name in enumerate(names):
        assert getattr(_lib, name) == i, name


def test_header_is_plain_c_and_the_library_links_from_c(mrlib, tmp_path):
    """include/mrscore.h must be consumable from C99 (JNI / JNA / Panama bind a C ABI): compile and link examples/c_abi_demo.c with gcc
    against the built library and run it; without a GPU it must stop at mr_create with MR_ERR_CUDA."""
    import subprocess
    import torch
    exe = tmp_path / "c_abi_demo"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", str(ROOT / "include"), str(ROOT / "examples" / "c_abi_demo.c"),
                    "-L", str(ROOT / "musicrecommendation_b200"), "-lmrscore", "-Wl,-rpath," + str(ROOT / "musicrecommendation_b200"), "-o", str(exe)],
                   check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr
        assert "user 0  #1" in r.stdout
    else:
        assert r.returncode == _lib.MR_ERR_CUDA and "no CPU fallback" in r.stderr


def test_native_double_formatter_is_java_double_to_string(mrlib):
    """mr_format_double (csrc/modelio.cu, std::to_chars digits + Java's layout rules) against the independent Python formatter."""
    buf = C.create_string_buffer(40)
    rng = np.random.default_rng(7)
    vals = list(rng.random(3000) * 10) + list(10.0 ** rng.uniform(-15, 15, 3000)) + [0.0, 1.0, 1e7, 1e-3, 9999999.999, 0.00099999999, 123.0,
                                                                                      5e-324, 1.7976931348623157e308, 0.4999999999999999, 2.0 ** 53, 1e22, 1e23]
    for x in vals:
        n = mrlib.mr_format_double(float(x), buf)
        assert buf.value.decode() == double_to_string(float(x)) and n == len(buf.value)
        assert float(buf.value) == float(x)                      # `.toDouble` in importModelFromFile gets the same bits back (MR:509)


def test_model_file_round_trip(mrlib, oracle_lib, tmp_path):
    """writeModelOnFile -> importModelFromFile (MR:489-512) on real score arrays: every emitted pair comes back with the same bits, in
    (user, song) order, listened pairs absent; the native writer's bytes equal the pure-Python writer's."""
    ds = synth(T=120, U=9, S=900, seed=13)
    users = [f"{i:040x}" for i in range(ds.U)]
    songs = ["SO" + f"{i:016X}" for i in range(ds.S)]
    for m in (oracle_lib.UBM, oracle_lib.IBM):
        scores = oracle_lib.canon_scores(ds, m)
        model = recommender.Model(scores, users, songs, "m")
        p_native, p_py = tmp_path / f"native_{m}.txt", tmp_path / f"py_{m}.txt"
        rows = recommender.write_model_file(model, str(p_native), mrlib)
        recommender.write_model_file_py(model, str(p_py))
        assert rows == ds.n_pairs == len(model)
        assert p_native.read_bytes() == p_py.read_bytes()
        back = recommender.import_model_file(str(p_native))
        assert len(back) == ds.n_pairs
        want = list(model.tuples())
        assert [(u, s) for u, s, _ in back] == [(u, s) for u, (s, _) in want]         # ids are zero-padded: string order == id order
        np.testing.assert_array_equal(np.array([v for _, _, v in back]).view(np.int64), np.array([v for _, (_, v) in want]).view(np.int64))
    # the hand-derived fixture: the score that is NOT 0.5 keeps all its digits
    fx = fixture_4_3()
    model = recommender.Model(oracle_lib.naive_scores(fx, oracle_lib.UBM), fx.test_users, fx.songs, "ubm")
    p = tmp_path / "fx.txt"
    recommender.write_model_file(model, str(p), mrlib)
    assert p.read_text().splitlines()[0] == "X\ts2\t0.4999999999999999"
    assert p.read_text().splitlines()[1] == "X\ts3\t0.0"
