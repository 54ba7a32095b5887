"""Host-side mirror of the reference's `MusicRecommender` class (MusicRecommender.scala:12-640) over libmrscore.so.

Same method names, argument meaning and error behaviour as the Scala class, so a caller of the reference finds
`getUserBasedModel`, `getItemBasedModel`, the three blends, `writeModelOnFile`, `importModelFromFile` and `evaluateModel`
where they expect them; every scoring method is one call through the C-ABI (include/mrscore.h) into the sm_100a kernels.
The JVM is absent from this image, so this mirror is Python; INTEGRATION.md shows the Scala/JNI binding of the same ABI.

Differences a caller can observe (all documented in DESIGN.md):
  * models come back already in the order of the alignment sort main.scala:57-59 (user asc, song asc), as a `Model`
    holding the dense U x S score matrix (NaN = pair not emitted, MR:109); `Model.tuples()` yields the reference's
    `(user, (song, score))` elements;
  * where the reference calls System.exit(2) / System.exit(-1) this raises `KeyMismatch` / `ParameterRange` (the Scala
    wrapper maps the error codes back to the exits);
  * `getStochasticCombinationModel` takes an explicit `seed` (the reference uses an unseeded `new Random`, MR:439);
  * `getTopK` is new (north_star): the only output that exists at MSD scale.
"""
from __future__ import annotations

import ctypes as C
import io
import sys
from dataclasses import dataclass

import numpy as np

from . import _lib
from .dataset import Dataset, _csr_from_pairs
from .javafmt import double_to_string


class KeyMismatch(RuntimeError):
    """ubm / ibm not aligned — the reference calls System.exit(2) (MR:326, 347, 379, 413, 445, 476)."""
    exit_code = 2


class ParameterRange(ValueError):
    """blend parameter outside [0,1] — the reference prints to stderr and calls System.exit(-1) (MR:366-369, 434-437)."""
    exit_code = -1


@dataclass
class Model:
    """A scored model in main.scala:57-59 order.  scores[u, s] is NaN where the reference emits no element (MR:109)."""
    scores: np.ndarray            # [U, S] float64
    test_users: list | None = None
    songs: list | None = None
    name: str = ""

    def __len__(self) -> int:
        return int(np.count_nonzero(~np.isnan(self.scores)))

    def compact(self) -> np.ndarray:
        """The score column of the sorted `Array[(String, (String, Double))]`."""
        return self.scores[~np.isnan(self.scores)]

    def keys(self):
        u, s = np.nonzero(~np.isnan(self.scores))
        return u.astype(np.int32), s.astype(np.int32)

    def tuples(self):
        u, s = self.keys()
        v = self.scores[u, s]
        un = self.test_users if self.test_users is not None else list(range(self.scores.shape[0]))
        sn = self.songs if self.songs is not None else list(range(self.scores.shape[1]))
        for a, b, c in zip(u.tolist(), s.tolist(), v.tolist()):
            yield un[a], (sn[b], c)


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


def parse_triplets(stream):
    """One TSV stream -> (users list, songs list) exactly as MR:32-42 accepts it: 3 tab-separated fields, third ignored.
    Java's String.split drops trailing empty strings; a line without exactly 3 fields raises (scala.MatchError, MR:34-35)."""
    users, songs = [], []
    for line in stream:
        line = line.rstrip("\n").rstrip("\r")
        parts = line.split("\t")
        while parts and parts[-1] == "":
            parts.pop()
        if len(parts) != 3:
            raise ValueError(f"scala.MatchError: line does not have 3 tab-separated fields: {line!r}")
        users.append(parts[0])
        songs.append(parts[1])
    return users, songs


def dataset_from_streams(train, test, labels) -> Dataset:
    """The constructor's ingest (MR:26-91) with ids assigned in ascending string order.  Degrees are the `.length` of the
    reference's non-deduplicated lists (MR:40-41), so duplicate rows inflate denominators exactly as they do there."""
    tr_u, tr_s = parse_triplets(train)
    te_u, te_s = parse_triplets(test)
    lb_u, lb_s = parse_triplets(labels)
    train_users = sorted(set(tr_u))
    test_users = sorted(set(te_u))
    songs = sorted(set(tr_s) | set(te_s))                    # mutSongs: train ∪ test-visible (MR:38, 51, 58)
    label_only = sorted(set(lb_s) - set(songs))              # label songs that occur nowhere else get ids >= S
    tix = {u: i for i, u in enumerate(train_users)}
    uix = {u: i for i, u in enumerate(test_users)}
    six = {s: i for i, s in enumerate(songs + label_only)}
    T, U, S = len(train_users), len(test_users), len(songs)
    tr_rows = np.fromiter((tix[u] for u in tr_u), np.int64, len(tr_u))
    tr_cols = np.fromiter((six[s] for s in tr_s), np.int64, len(tr_s))
    te_rows = np.fromiter((uix[u] for u in te_u), np.int64, len(te_u))
    te_cols = np.fromiter((six[s] for s in te_s), np.int64, len(te_s))
    lb_keep = [(uix[u], six[s]) for u, s in zip(lb_u, lb_s) if u in uix]
    lb_rows = np.array([a for a, _ in lb_keep], np.int64)
    lb_cols = np.array([b for _, b in lb_keep], np.int64)
    tr_ptr, tr_col = _csr_from_pairs(tr_rows, tr_cols, T)
    te_ptr, te_col = _csr_from_pairs(te_rows, te_cols, U)
    lab_ptr, lab_col = _csr_from_pairs(lb_rows, lb_cols, U)
    deg_tr = np.bincount(tr_rows, minlength=T).astype(np.int32)           # .length incl. duplicates (MR:147)
    deg_te = np.bincount(te_rows, minlength=U).astype(np.int32)
    deg_song = (np.bincount(tr_cols, minlength=S) + np.bincount(te_cols, minlength=S)).astype(np.int32)   # MR:237
    return Dataset(T, U, S, tr_ptr, tr_col, te_ptr, te_col, lab_ptr, lab_col, deg_tr, deg_te, deg_song,
                   train_users, test_users, songs + label_only, {"source": "tsv"})


def _read_bytes(f) -> bytes:
    if isinstance(f, (bytes, bytearray, memoryview)):
        return bytes(f)
    if isinstance(f, str) or hasattr(f, "__fspath__"):
        with open(f, "rb") as fh:
            return fh.read()
    data = f.read()
    return data.encode("utf-8") if isinstance(data, str) else bytes(data)


def dataset_from_streams_native(train, test, labels, device: int = 0, with_strings: bool = True) -> Dataset:
    """The same ingest on the GPU (`mr_ingest_tsv`, csrc/k6_ingest.cu): the three TSV files as bytes -> int-id CSR, ids in ascending
    string order, `.length` degrees.  Same result as dataset_from_streams (tests/test_gpu_parity.py compares them field by field);
    a malformed line raises ValueError like the host path.  `timing_ms` of the stages is kept in Dataset.meta."""
    lib = _lib.load()
    bufs = [_read_bytes(f) for f in (train, test, labels)]
    g = C.c_void_p()
    rc = lib.mr_ingest_tsv(device, bufs[0], len(bufs[0]), bufs[1], len(bufs[1]), bufs[2], len(bufs[2]), C.byref(g))
    try:
        if rc != _lib.MR_OK:
            msg = (lib.mr_ingest_error(g) or b"").decode()
            if rc == _lib.MR_ERR_BAD_ARG and msg.startswith("scala.MatchError"):
                raise ValueError(msg)
            raise _lib.MrError(rc, msg)
        dims = [C.c_int32() for _ in range(4)]
        lib.mr_ingest_dims(g, *[C.byref(d) for d in dims])
        T, U, S, n_extra = (d.value for d in dims)

        def get(which, dtype):
            ptr, n = C.c_void_p(), C.c_int64()
            lib.mr_ingest_get(g, which, C.byref(ptr), C.byref(n))
            if n.value == 0:
                return np.zeros(0, dtype)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(n.value,)).copy()

        def table(chars_id, off_id):
            if not with_strings:
                return None
            chars = get(chars_id, np.uint8).tobytes()
            off = get(off_id, np.int64)
            return [chars[off[i]:off[i + 1]].decode("utf-8") for i in range(len(off) - 1)]

        return Dataset(T, U, S, get(_lib.MR_ING_TR_PTR, np.int64), get(_lib.MR_ING_TR_COL, np.int32), get(_lib.MR_ING_TE_PTR, np.int64),
                       get(_lib.MR_ING_TE_COL, np.int32), get(_lib.MR_ING_LAB_PTR, np.int64), get(_lib.MR_ING_LAB_COL, np.int32),
                       get(_lib.MR_ING_DEG_TRAIN, np.int32), get(_lib.MR_ING_DEG_TEST, np.int32), get(_lib.MR_ING_DEG_SONG, np.int32),
                       table(_lib.MR_ING_TRAIN_USER_CHARS, _lib.MR_ING_TRAIN_USER_OFF), table(_lib.MR_ING_TEST_USER_CHARS, _lib.MR_ING_TEST_USER_OFF),
                       table(_lib.MR_ING_SONG_CHARS, _lib.MR_ING_SONG_OFF),
                       {"source": "tsv-native", "label_only_songs": n_extra, "timing_ms": dict(zip(
                           ["h2d", "lines_parse", "unique_users", "unique_songs", "host_sort", "csr", "unused", "total"], get(_lib.MR_ING_TIMING_MS, np.float64).tolist()))})
    finally:
        lib.mr_ingest_free(g)


class MusicRecommender:
    """`new MusicRecommender(trainFile, testFile, testLabelsFile)` (MR:12).  Also accepts a ready `Dataset`."""

    def __init__(self, trainFile, testFile=None, testLabelsFile=None, device: int = 0, engine: int = _lib.MR_ENGINE_AUTO,
                 profile: bool = False, space: int = _lib.MR_SPACE_AUTO, ingest: str = "host", head_min_deg: int = 0,
                 item_batch: int = 0, song_window=None):
        """song_window = (lo, hi): score only the songs [lo, hi) of every test user — one song partition of distributed.scala:459-461
        (MR_OPT_SONG_WINDOW_*): dense models become [U, hi - lo], getTopK ranks inside the window (global song ids)."""
        if isinstance(trainFile, Dataset):
            ds = trainFile
        elif ingest == "native":
            ds = dataset_from_streams_native(trainFile, testFile, testLabelsFile, device=device)
        else:
            def op(f):
                return open(f, "r", encoding="utf-8") if isinstance(f, (str, bytes)) or hasattr(f, "__fspath__") else f
            ds = dataset_from_streams(op(trainFile), op(testFile), op(testLabelsFile))
        self.ds = ds
        self._lib = _lib.load()
        self._h = C.c_void_p()
        dev = (C.c_int * 1)(device)
        rc = self._lib.mr_create(C.byref(self._h), dev, 1, engine | space | (_lib.MR_PROFILE if profile else 0))
        self._check(rc)
        if head_min_deg:
            self._check(self._lib.mr_set_option(self._h, _lib.MR_OPT_HEAD_MIN_DEG, int(head_min_deg)))
        if item_batch:
            self._check(self._lib.mr_set_option(self._h, _lib.MR_OPT_ITEM_BATCH, int(item_batch)))
        self.song_window = (0, ds.S) if song_window is None else (int(song_window[0]), int(song_window[1]))
        if song_window is not None:
            self._check(self._lib.mr_set_option(self._h, _lib.MR_OPT_SONG_WINDOW_LO, self.song_window[0]))
            self._check(self._lib.mr_set_option(self._h, _lib.MR_OPT_SONG_WINDOW_HI, self.song_window[1]))
        a = self._arrs = dict(
            tr_ptr=np.ascontiguousarray(ds.tr_ptr, np.int64), tr_col=np.ascontiguousarray(ds.tr_col, np.int32),
            te_ptr=np.ascontiguousarray(ds.te_ptr, np.int64), te_col=np.ascontiguousarray(ds.te_col, np.int32),
            deg_tr=np.ascontiguousarray(ds.deg_tr, np.int32), deg_te=np.ascontiguousarray(ds.deg_te, np.int32),
            deg_song=np.ascontiguousarray(ds.deg_song, np.int32))
        rc = self._lib.mr_load(self._h, ds.T, ds.U, ds.S, _p(a["tr_ptr"]), _p(a["tr_col"]), _p(a["te_ptr"]), _p(a["te_col"]),
                               _p(a["deg_tr"]), _p(a["deg_te"]), _p(a["deg_song"]))
        self._check(rc)

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int):
        if rc == _lib.MR_OK:
            return
        msg = self._lib.mr_last_error(self._h).decode(errors="replace") if self._h else "mr_create failed"
        if rc == _lib.MR_ERR_PARAM_RANGE:
            sys.stderr.write(msg + "\n\n")            # System.err.println("... must be between 0 and 1\n"), MR:367, 435
            raise ParameterRange(msg)
        if rc == _lib.MR_ERR_KEY_MISMATCH:
            raise KeyMismatch(msg)
        raise _lib.MrError(rc, msg)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_test_users(self, ds_shard: Dataset, pair_index_base: int = 0, n_pairs_total: int = 0):
        """Score another shard of test users against the resident train replica (DIST:450-452 partitioning)."""
        a = self._arrs
        a["te_ptr"] = np.ascontiguousarray(ds_shard.te_ptr, np.int64)
        a["te_col"] = np.ascontiguousarray(ds_shard.te_col, np.int32)
        a["deg_te"] = np.ascontiguousarray(ds_shard.deg_te, np.int32)
        self._check(self._lib.mr_set_test_users(self._h, ds_shard.U, _p(a["te_ptr"]), _p(a["te_col"]), _p(a["deg_te"]),
                                                pair_index_base, n_pairs_total))
        self.ds = ds_shard

    def _model(self, kind: int, name: str) -> Model:
        lo, hi = self.song_window
        out = np.empty((self.ds.U, hi - lo), np.float64)
        self._check(self._lib.mr_score_dense(self._h, kind, _p(out)))
        songs = self.ds.songs if (lo, hi) == (0, self.ds.S) or self.ds.songs is None else self.ds.songs[lo:hi]
        return Model(out, self.ds.test_users, songs, name)

    # ------------------------------------------------------------------ reference method surface
    def getUserBasedModel(self) -> Model:                    # MR:132
        return self._model(_lib.MR_UBM, "userBasedModel")

    getUserBasedModelP = getUserBasedModel                   # MR:177 (.par twin: same values)

    def getItemBasedModel(self) -> Model:                    # MR:222
        return self._model(_lib.MR_IBM, "itemBasedModel")

    getItemBasedModelP = getItemBasedModel                   # MR:268

    def _blend(self, kind: int, ubm: Model, ibm: Model, param: float, seed: int, name: str) -> Model:
        mu, mi = ~np.isnan(ubm.scores), ~np.isnan(ibm.scores)
        cu, ci = ubm.scores[mu], ibm.scores[mi]
        n = min(len(cu), len(ci))                            # zip truncates (MR:322)
        ku, ki = np.flatnonzero(mu.ravel())[:n], np.flatnonzero(mi.ravel())[:n]
        # parameter range first (MR:366-369 precedes the zip), then the per-element key check (MR:326)
        out = np.empty(n, np.float64)
        rc = self._lib.mr_blend_dense(self._h, kind, float(param), C.c_uint64(seed & (2**64 - 1)), _p(cu), _p(ci), _p(out), n, 0,
                                      len(cu))
        self._check(rc)
        if not np.array_equal(ku, ki):
            raise KeyMismatch("(user, song) keys of ubm and ibm differ")
        res = np.full(ubm.scores.shape, np.nan)
        res.ravel()[ku] = out
        return Model(res, ubm.test_users, ubm.songs, name)

    def getLinearCombinationModel(self, ubm: Model, ibm: Model, alpha: float) -> Model:                    # MR:317
        return self._blend(_lib.MR_LC, ubm, ibm, alpha, 0, "linearCombinationModel")

    getLinearCombinationModelP = getLinearCombinationModel                                                  # MR:340

    def getAggregationModel(self, ubm: Model, ibm: Model, itemBasedPercentage: float = 0.5) -> Model:      # MR:361
        return self._blend(_lib.MR_AGG, ubm, ibm, itemBasedPercentage, 0, "aggregationModel")

    getAggregationModelP = getAggregationModel                                                              # MR:396

    def getStochasticCombinationModel(self, ubm: Model, ibm: Model, itemBasedProbability: float = 0.5,
                                      seed: int = 0) -> Model:                                              # MR:429
        return self._blend(_lib.MR_STOCH, ubm, ibm, itemBasedProbability, seed, "stochasticCombinationModel")

    getStochasticCombinationModelP = getStochasticCombinationModel                                          # MR:461

    def getTopK(self, model: int = _lib.MR_UBM, k: int = 500, param: float = 0.5, seed: int = 0):
        """Per test user the k best unlistened songs (score desc, song id asc): (song [U,k] int32, score [U,k] f64, len [U])."""
        U = self.ds.U
        song = np.empty((U, k), np.int32)
        score = np.empty((U, k), np.float64)
        ln = np.empty(U, np.int32)
        self._check(self._lib.mr_topk(self._h, model, float(param), C.c_uint64(seed & (2**64 - 1)), k, _p(song), _p(score), _p(ln)))
        return song, score, ln

    def prepare(self):
        """Build the item-space head rows now (one-off per train set) instead of lazily."""
        self._check(self._lib.mr_prepare(self._h))

    def prepare_async(self):
        """Start the head-row build on its own stream and return; the next scoring call completes it (overlaps set_test_users)."""
        self._check(self._lib.mr_prepare_async(self._h))

    def invalidate_prepared(self):
        """Forget the head rows so that the next prepare() / scoring call rebuilds them (bench.py: the whole model build inside a step)."""
        self._check(self._lib.mr_invalidate_prepared(self._h))

    def getRanks1(self, model: int, users) -> np.ndarray:
        """DIST UserBasedModel/ItemBasedModel.getRanks1(user) (DIST:198-205, 269-276) for the listed test users of the current shard:
        [len(users), S] fp64, NaN at listened pairs."""
        ids = np.ascontiguousarray(users, np.int32)
        out = np.empty((len(ids), self.song_window[1] - self.song_window[0]), np.float64)
        self._check(self._lib.mr_score_users(self._h, model, _p(ids), len(ids), _p(out)))
        return out

    def getRanks2(self, model: int, songs) -> np.ndarray:
        """DIST getRanks2(song) (DIST:214-221, 285-292) for the listed songs: [len(songs), U] fp64, NaN where the user listened."""
        ids = np.ascontiguousarray(songs, np.int32)
        out = np.empty((len(ids), self.ds.U), np.float64)
        self._check(self._lib.mr_score_songs(self._h, model, _p(ids), len(ids), _p(out)))
        return out

    def mapAtK(self, k: int = 500, top=None, per_user: bool = False, labels: Dataset | None = None):
        """mAP@k (north_star's mAP@500; MSD-challenge definition, include/mrscore.h) of ranked lists against this data set's label rows.
        top = (song [U,k], len [U]) as getTopK returns them, or None for the device-resident result of the last getTopK / mr_topk_device."""
        lab = labels if labels is not None else self.ds      # labels: the data set whose label rows match explicit lists of other users
        lp = np.ascontiguousarray(lab.lab_ptr, np.int64)
        lc = np.ascontiguousarray(lab.lab_col, np.int32)
        out = C.c_double(0.0)
        ap = np.zeros(lab.U, np.float64)
        if top is None:
            self._check(self._lib.mr_map_at_k(self._h, k, None, None, self.ds.U, _p(lp), _p(lc), C.byref(out), _p(ap)))
        else:
            song = np.ascontiguousarray(top[0], np.int32)
            ln = np.ascontiguousarray(top[1], np.int32)
            self._check(self._lib.mr_map_at_k(self._h, song.shape[1], _p(song), _p(ln), song.shape[0], _p(lp), _p(lc), C.byref(out), _p(ap)))
        return (float(out.value), ap) if per_user else float(out.value)

    def topk_device_tensors(self, k: int):
        """torch views (no copy) of the last mr_topk_device result in this GPU's HBM: (song int32 [U,k], score f64 [U,k], len int32 [U])."""
        import torch
        ps, pv, pl = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self._lib.mr_topk_device_ptrs(self._h, k, C.byref(ps), C.byref(pv), C.byref(pl)))
        U = self.ds.U

        class _Arr:
            def __init__(self, ptr, shape, typestr):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}
        return (torch.as_tensor(_Arr(ps.value, (U, k), "<i4"), device="cuda"), torch.as_tensor(_Arr(pv.value, (U, k), "<f8"), device="cuda"),
                torch.as_tensor(_Arr(pl.value, (U,), "<i4"), device="cuda"))

    def mergeTopK(self, parts_song, parts_score, parts_len, out_song, out_score, out_len):
        """Join the ranked lists of several song partitions for the same users (mr_topk_merge; the `.collect` + ranking of
        distributed.scala:459-461).  All arguments are CUDA torch tensors: parts_* [P, n, k] / [P, n], out_* [n, k] / [n]; the join runs
        on the library stream."""
        P, n, k = parts_song.shape
        tab = lambda t: (C.c_void_p * P)(*[t[i].data_ptr() for i in range(P)])      # noqa: E731
        self._check(self._lib.mr_topk_merge(self._h, k, P, n, tab(parts_song), tab(parts_score), tab(parts_len), C.c_void_p(out_song.data_ptr()),
                                            C.c_void_p(out_score.data_ptr()), C.c_void_p(out_len.data_ptr())))

    def topk_packed_tensor(self, k: int):
        """The last top-k result as ONE uint8 torch view of device memory plus the byte offsets of its parts:
        (block, score_offset, len_offset); song int32 [U,k] starts at 0."""
        import torch
        base, nbytes, so, lo = C.c_void_p(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self._lib.mr_topk_packed(self._h, k, C.byref(base), C.byref(nbytes), C.byref(so), C.byref(lo)))

        class _Arr:
            __cuda_array_interface__ = {"shape": (int(nbytes.value),), "typestr": "|u1", "data": (base.value, False), "version": 2}
        return torch.as_tensor(_Arr(), device="cuda"), int(so.value), int(lo.value)

    # ------------------------------------------------------------------ parity probes / similarity products
    def counts_ubm(self) -> np.ndarray:
        out = np.empty((self.ds.U, self.ds.T), np.int32)
        self._check(self._lib.mr_counts_ubm(self._h, _p(out)))
        return out

    def counts_ibm(self, s0: int, s1: int) -> np.ndarray:
        out = np.empty((s1 - s0, self.ds.S), np.int32)
        self._check(self._lib.mr_counts_ibm(self._h, s0, s1, _p(out)))
        return out

    def gram_rows_device(self, s0: int, s1: int):
        """Rows [s0,s1) of this handle's (partial) train co-occurrence matrix as a torch int32 view [s1-s0, ld] of device memory."""
        import torch
        ptr, ld = C.c_void_p(), C.c_int64()
        self._check(self._lib.mr_gram_rows_device(self._h, s0, s1, C.byref(ptr), C.byref(ld)))

        class _Arr:
            __cuda_array_interface__ = {"shape": (s1 - s0, ld.value), "typestr": "<i4", "data": (ptr.value, False), "version": 2}
        return torch.as_tensor(_Arr(), device="cuda")

    # ---- fused K-split reduce-scatter (GEMM epilogue stores into the owners' receive slots, local or NVLink peer memory)
    def peer_alloc(self, nbytes: int):
        """(device pointer, 64-byte CUDA IPC handle) of a zeroed buffer other ranks can map with peer_open."""
        ptr = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        self._check(self._lib.mr_peer_alloc(self._h, nbytes, C.byref(ptr), handle))
        return int(ptr.value), bytes(handle)

    def peer_open(self, handle: bytes) -> int:
        ptr = C.c_void_p()
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        self._check(self._lib.mr_peer_open(self._h, buf, C.byref(ptr)))
        return int(ptr.value)

    def peer_close(self, ptr: int):
        self._check(self._lib.mr_peer_close(self._h, C.c_void_p(ptr)))

    def gram_rows_scatter(self, s0: int, s1: int, slot_ptrs, rows_per_owner: int, ld: int, sync: bool = True):
        arr = (C.c_void_p * len(slot_ptrs))(*slot_ptrs)
        fn = self._lib.mr_gram_rows_scatter if sync else self._lib.mr_gram_rows_scatter_async
        self._check(fn(self._h, s0, s1, arr, len(slot_ptrs), rows_per_owner, ld))

    def peer_signal(self, flag_ptrs, value: int):
        arr = (C.c_void_p * len(flag_ptrs))(*flag_ptrs)
        self._check(self._lib.mr_peer_signal(self._h, arr, len(flag_ptrs), value))

    def peer_wait(self, flags_ptr: int, n: int, value: int):
        self._check(self._lib.mr_peer_wait(self._h, C.c_void_p(flags_ptr), n, value))

    def sync(self):
        self._check(self._lib.mr_sync(self._h))

    def similarity_ubm(self) -> np.ndarray:
        out = np.empty((self.ds.U, self.ds.T), np.float32)
        self._check(self._lib.mr_similarity_ubm(self._h, _p(out)))
        return out

    def similarity_ibm(self, s0: int, s1: int) -> np.ndarray:
        out = np.empty((s1 - s0, self.ds.S), np.float32)
        self._check(self._lib.mr_similarity_ibm(self._h, s0, s1, _p(out)))
        return out

    def timing(self, reset: bool = False) -> dict:
        t = (C.c_double * 9)()
        self._lib.mr_get_timing(self._h, t, 9)
        if reset:
            self._lib.mr_reset_timing(self._h)
        return dict(zip(_lib.TIMING_NAMES, list(t)))

    def info(self) -> dict:
        v = (C.c_int64 * 19)()
        self._lib.mr_get_info(self._h, v, 19)
        return dict(zip(["engine", "launches", "dense_bytes", "n_items", "num_sms", "device_bytes", "space", "n_head", "head_entries",
                         "tail_entries", "head_exceptions", "batch_rows", "head_groups", "split_users", "n_cols", "win_lo", "topk_fast_rejects", "topk_short_rows", "topk_degenerate_rows"], list(v)))

    # ------------------------------------------------------------------ model file I/O (MR:489-512)
    def writeModelOnFile(self, model: Model, outputFileName: str = "") -> int:
        """`user \\t song \\t Double.toString(score) \\n` per element, in model order (MR:492-494), through the native writer
        `mr_write_model` (csrc/modelio.cu).  Returns the number of lines written."""
        return write_model_file(model, outputFileName, self._lib)

    def importModelFromFile(self, pathToModel: str):
        """Array[(String, String, Double)] sorted by (user, song, score desc) (MR:505-512)."""
        return import_model_file(pathToModel)

    # ------------------------------------------------------------------ evaluation (MR:521-639), host-side like the reference
    def evaluateModel(self, model: Model, parallel: bool = False, n_thresholds: int = 10) -> float:
        """The reference's threshold-sweep mAP (MR:521-639; 10 thresholds MR:590, 11 in distributed.scala:395) on the GPU
        (`mr_evaluate_dense`).  `evaluate_map` below is the same computation in numpy for hosts without a GPU handle."""
        sc = np.ascontiguousarray(model.scores, np.float64)
        lp = np.ascontiguousarray(self.ds.lab_ptr, np.int64)
        lc = np.ascontiguousarray(self.ds.lab_col, np.int32)
        out = C.c_double(0.0)
        self._check(self._lib.mr_evaluate_dense(self._h, _p(sc), sc.shape[0], sc.shape[1], _p(lp), _p(lc), n_thresholds, C.byref(out)))
        return float(out.value)


def _string_table(names, n):
    """id -> string list as (uint8 chars, int64 offsets [n+1]); ints stand in for missing tables."""
    enc = [(str(i) if names is None else names[i]).encode("utf-8") for i in range(n)]
    off = np.zeros(n + 1, np.int64)
    np.cumsum([len(b) for b in enc], out=off[1:])
    chars = np.frombuffer(b"".join(enc) or b"\0", np.uint8)
    return np.ascontiguousarray(chars), off


def write_model_file(model: Model, path: str, lib=None) -> int:
    """writeModelOnFile (MR:489-496) natively: one `user\\tsong\\tDouble.toString(score)\\n` line per emitted pair."""
    lib = lib or _lib.load()
    sc = np.ascontiguousarray(model.scores, np.float64)
    U, S = sc.shape
    uc, uo = _string_table(model.test_users, U)
    scs, so = _string_table(model.songs, S)
    rows = C.c_int64(0)
    rc = lib.mr_write_model(str(path).encode(), _p(sc), U, S, _p(uc), _p(uo), _p(scs), _p(so), 0, C.byref(rows))
    if rc != _lib.MR_OK:
        raise OSError(f"mr_write_model({path!r}) failed with code {rc}")
    return int(rows.value)


def write_model_file_py(model: Model, path: str) -> None:
    """The same file from pure Python (javafmt.double_to_string): the independent check of the native formatter in tests."""
    with open(path, "w", encoding="utf-8", newline="") as out:
        buf = io.StringIO()
        for user, (song, score) in model.tuples():
            buf.write(f"{user}\t{song}\t{double_to_string(score)}\n")
        out.write(buf.getvalue())


def import_model_file(path: str):
    """importModelFromFile (MR:505-512): Array[(String, String, Double)] sorted by (user, song, score desc)."""
    rows = []
    with open(path, "r", encoding="utf-8") as f:
        for line in f:
            parts = line.rstrip("\n").split("\t")
            if len(parts) != 3:
                raise ValueError(f"scala.MatchError: {line!r}")
            rows.append((parts[0], parts[1], float(parts[2])))
    rows.sort(key=lambda r: (r[0], r[1], -r[2]))
    return rows


def evaluate_map(scores: np.ndarray, ds: Dataset, n_thresholds: int = 10) -> float:
    """MR:521-639 vectorised over label songs: global min/max normalisation (MR:524-529), per-song confusion counts over
    test users (MR:541-553), AP = sum_i (R_i - R_{i+1}) P_i with the last two terms as in MR:601-609, mean over the
    distinct label songs (MR:626)."""
    U, S = scores.shape
    valid = ~np.isnan(scores)
    mn, mx = scores[valid].min(), scores[valid].max()
    with np.errstate(invalid="ignore", divide="ignore"):
        norm = (scores - mn) / (mx - mn)
    new_songs = np.unique(ds.lab_col)
    labelled = np.zeros((U, int(max(S, new_songs.max() + 1))), bool)
    rows = np.repeat(np.arange(U), np.diff(ds.lab_ptr))
    labelled[rows, ds.lab_col] = True
    th = [0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0][:n_thresholds]
    in_model = new_songs[new_songs < S]
    prec = np.zeros((len(th), len(new_songs)))
    rec = np.zeros((len(th), len(new_songs)))
    lab = labelled[:, new_songs]                       # [U, n_new]
    pos = np.flatnonzero(new_songs < S)
    for t, thr in enumerate(th):
        pred = np.zeros((U, len(new_songs)), bool)
        with np.errstate(invalid="ignore"):
            pred[:, pos] = norm[:, in_model] > thr     # NaN compares false (listened pairs, max == min)
        tp = (pred & lab).sum(0).astype(np.float64)
        fp = (pred & ~lab).sum(0)
        fn = (~pred & lab).sum(0)
        with np.errstate(invalid="ignore", divide="ignore"):
            prec[t] = np.where(tp + fp > 0, tp / (tp + fp), 0.0)
            rec[t] = np.where(tp + fn > 0, tp / (tp + fn), 0.0)
    n = len(th)
    ap = np.zeros(len(new_songs))
    for t in range(n):                                  # List.sum: left fold from 0.0 in threshold order
        if t == n - 1:
            term = 0.0
        elif t == n - 2:
            term = (rec[t] - 0.0) * prec[t]
        else:
            term = (rec[t] - rec[t + 1]) * prec[t]
        ap = ap + term
    total = 0.0
    for x in ap.tolist():                               # foldLeft(0.0)(_+_) (MR:626)
        total += x
    return total / len(new_songs)
