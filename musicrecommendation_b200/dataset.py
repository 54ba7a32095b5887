"""Int-id data model for the scoring path, and the synthetic MSD-shaped generator.

`Dataset` is what `MusicRecommender`'s constructor builds from the three TSV streams
(reference MusicRecommender.scala:26-91), with the strings replaced by dense int32 ids assigned in
ascending `String.compareTo` order, so that int order == `Ordering.String`
(main.scala:57) and "ties broken by song id" is well defined:

  train CSR   tr_ptr[T+1], tr_col     songs of each train user, ascending, unique     (MR:55)
  test  CSR   te_ptr[U+1], te_col     visible half of each test user's history         (MR:56)
  labels CSR  lab_ptr[U+1], lab_col   hidden half (MR:70-91); ids >= S are songs that occur nowhere else
  deg_tr / deg_te                     `.length` of the per-user arrays (MR:147)
  deg_song                            `.length` of songsToUsersMap(s): train AND test-visible listeners (MR:41,53,237)

`synth()` follows SURVEY.md §8(d): heavy-tailed user degrees (min 10), Zipf–Mandelbrot song
popularity, no duplicate (user, song) rows, every one of the S songs heard at least once in
train ∪ test-visible, ceil(n/2) / floor(n/2) visible / label split of test users
(dataExtraction.ipynb cell 11).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class Dataset:
    T: int
    U: int
    S: int
    tr_ptr: np.ndarray
    tr_col: np.ndarray
    te_ptr: np.ndarray
    te_col: np.ndarray
    lab_ptr: np.ndarray
    lab_col: np.ndarray
    deg_tr: np.ndarray
    deg_te: np.ndarray
    deg_song: np.ndarray
    train_users: list | None = None   # id -> string tables (only when built from TSV / requested)
    test_users: list | None = None
    songs: list | None = None
    meta: dict = field(default_factory=dict)

    @property
    def nnz_tr(self) -> int:
        return int(self.tr_ptr[-1])

    @property
    def nnz_te(self) -> int:
        return int(self.te_ptr[-1])

    @property
    def n_pairs(self) -> int:
        """Number of scored pairs = U*S - nnz_te (getModel's filter, MR:109)."""
        return self.U * self.S - self.nnz_te

    def listened_mask(self) -> np.ndarray:
        m = np.zeros((self.U, self.S), bool)
        rows = np.repeat(np.arange(self.U), np.diff(self.te_ptr))
        m[rows, self.te_col] = True
        return m

    def shard_test_users(self, u0: int, u1: int) -> "Dataset":
        """Test users [u0,u1) against the full train replica — the variant-1 partitioning of
        distributed.scala:450-452.  deg_song is kept GLOBAL (it counts every test-visible listener)."""
        a, b = int(self.te_ptr[u0]), int(self.te_ptr[u1])
        la, lb = int(self.lab_ptr[u0]), int(self.lab_ptr[u1])
        return Dataset(self.T, u1 - u0, self.S, self.tr_ptr, self.tr_col,
                       (self.te_ptr[u0:u1 + 1] - a).astype(np.int64), self.te_col[a:b],
                       (self.lab_ptr[u0:u1 + 1] - la).astype(np.int64), self.lab_col[la:lb],
                       self.deg_tr, self.deg_te[u0:u1], self.deg_song,
                       self.train_users, None if self.test_users is None else self.test_users[u0:u1], self.songs,
                       dict(self.meta, shard=(u0, u1)))


def _sorted_unique(key: np.ndarray) -> np.ndarray:
    """np.unique for int64 keys via sort + adjacent compare (numpy 2.3's hash-based unique is ~6x slower at 5e7 keys)."""
    key = np.sort(key)
    if len(key) == 0:
        return key
    keep = np.empty(len(key), bool)
    keep[0] = True
    np.not_equal(key[1:], key[:-1], out=keep[1:])
    return key[keep]


def _csr_from_pairs(rows: np.ndarray, cols: np.ndarray, n_rows: int):
    """Sort (row, col) pairs, drop duplicates, return (ptr int64, col int32)."""
    key = rows.astype(np.int64) * (1 << 32) | cols.astype(np.int64)
    key = _sorted_unique(key)
    r = (key >> 32).astype(np.int64)
    c = (key & 0xFFFFFFFF).astype(np.int32)
    ptr = np.zeros(n_rows + 1, np.int64)
    np.cumsum(np.bincount(r, minlength=n_rows), out=ptr[1:])
    return ptr, c


def from_triplets(train, test, labels, T: int, U: int, S: int, **kw) -> Dataset:
    """Build a Dataset from int-id (user, song) pair arrays; degrees = row lengths after de-duplication."""
    tr_ptr, tr_col = _csr_from_pairs(np.asarray(train[0]), np.asarray(train[1]), T)
    te_ptr, te_col = _csr_from_pairs(np.asarray(test[0]), np.asarray(test[1]), U)
    lab_ptr, lab_col = _csr_from_pairs(np.asarray(labels[0]), np.asarray(labels[1]), U)
    deg_song = (np.bincount(tr_col, minlength=S) + np.bincount(te_col, minlength=S)).astype(np.int32)
    return Dataset(T, U, S, tr_ptr, tr_col, te_ptr, te_col, lab_ptr, lab_col,
                   np.diff(tr_ptr).astype(np.int32), np.diff(te_ptr).astype(np.int32), deg_song, **kw)


def _degrees(rng, n, mean, cap):
    """10 + lognormal tail tuned so that the mean is `mean` (MSD: min 10, mean 47.5, max ~4.4k)."""
    sigma = 1.1
    mu = np.log(max(mean - 10.0, 1.0)) - sigma * sigma / 2
    d = 10 + np.floor(rng.lognormal(mu, sigma, n)).astype(np.int64)
    return np.minimum(d, cap)


def _popularity_cdf(n_songs: int, s0: float) -> np.ndarray:
    p = 1.0 / (np.arange(1, n_songs + 1, dtype=np.float64) + s0)
    c = np.cumsum(p)
    return c / c[-1]


def synth(T: int, U: int, S: int, seed: int, mean_deg: float = 47.5, with_strings: bool = False) -> Dataset:
    rng = np.random.Generator(np.random.PCG64(seed))
    cap = max(10, min(4400, S // 3))
    s0 = max(5.0, S / 7000.0)           # top song reaches ~10 % of users at MSD scale
    cdf = _popularity_cdf(S, s0)
    extra = max(1, S // 50)             # label-only ("new") songs get ids >= S
    cdf_lab = _popularity_cdf(S + extra, s0)
    perm = rng.permutation(S).astype(np.int64)   # popularity rank -> song id (string order is unrelated to rank)

    def draw(n_per_user, c):
        rows = np.repeat(np.arange(len(n_per_user), dtype=np.int64), n_per_user)
        ranks = np.searchsorted(c, rng.random(rows.shape[0]), side="right")
        return rows, np.minimum(ranks, len(c) - 1)

    # train
    deg_v = _degrees(rng, T, mean_deg, cap)
    need = int(np.ceil(1.10 * S)) - int(deg_v.sum())      # the MSD subsets have S ~ 0.95 * triplets: keep S coverable
    if need > 0:
        deg_v = deg_v + need // T + (np.arange(T) < need % T)
    tr_rows, tr_rank = draw(deg_v, cdf)
    # test: ceil(n/2) visible, floor(n/2) labels (dataExtraction.ipynb cell 11)
    deg_u = _degrees(rng, U, mean_deg * 1.1, cap)
    te_rows, te_rank = draw((deg_u + 1) // 2, cdf)
    lab_rows, lab_rank = draw(np.maximum(deg_u // 2, 1), cdf_lab)

    # de-duplicate (user, song) rows
    def dedupe(rows, ranks):
        key = _sorted_unique(rows * (1 << 32) | ranks)
        return key >> 32, key & 0xFFFFFFFF
    tr_rows, tr_rank = dedupe(tr_rows, tr_rank)
    te_rows, te_rank = dedupe(te_rows, te_rank)

    # coverage: every song in [0,S) must be heard in train ∪ test-visible (S == reference `songs.length`)
    cnt = np.bincount(tr_rank, minlength=S) + np.bincount(te_rank, minlength=S)
    missing = np.flatnonzero(cnt == 0)
    if len(missing):
        # spare train entries = every occurrence but the first of a song that train holds more than once
        order = np.argsort(tr_rank, kind="stable")
        sorted_rank = tr_rank[order]
        first = np.ones(len(order), bool)
        first[1:] = sorted_rank[1:] != sorted_rank[:-1]
        spare = order[~first]
        n_swap = min(len(spare), len(missing))
        pick = rng.choice(spare, size=n_swap, replace=False)
        tr_rank[pick] = missing[:n_swap]   # a missing song is new to every user, so no duplicate row appears
        if n_swap < len(missing):          # not enough repeats to recycle: hand the rest out as extra rows
            rest = missing[n_swap:]
            tr_rows = np.concatenate([tr_rows, rng.integers(0, T, size=len(rest))])
            tr_rank = np.concatenate([tr_rank, rest])
    # labels: drop the ones the user already has in the visible half; keep at least one per user
    lab_key = _sorted_unique(lab_rows * (1 << 32) | lab_rank)
    vis_key = te_rows * (1 << 32) | te_rank
    lab_key = lab_key[~np.isin(lab_key, vis_key)]
    lab_rows, lab_rank = lab_key >> 32, lab_key & 0xFFFFFFFF
    lacking = np.setdiff1d(np.arange(U), lab_rows)
    if len(lacking):                      # give them one unseen song each
        lab_rows = np.concatenate([lab_rows, lacking])
        lab_rank = np.concatenate([lab_rank, np.full(len(lacking), S + extra - 1)])

    to_id = np.concatenate([perm, np.arange(S, S + extra, dtype=np.int64)])
    ds = from_triplets((tr_rows, to_id[tr_rank]), (te_rows, to_id[te_rank]), (lab_rows, to_id[lab_rank]), T, U, S)
    ds.meta.update(seed=seed, generator="synth/zipf-mandelbrot", s0=s0, n_label_only=extra)
    if with_strings:
        ds.train_users, ds.test_users = _user_ids(rng, T, U)
        ds.songs = _song_ids(rng, S + extra)
    return ds


_HEX = np.frombuffer(b"0123456789abcdef", np.uint8)
_ALNUM = np.frombuffer(b"0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZ", np.uint8)


def _random_strings(rng, n, length, alphabet, prefix=""):
    out = set()
    while len(out) < n:
        m = n - len(out)
        arr = alphabet[rng.integers(0, len(alphabet), size=(m, length))]
        out.update(prefix + bytes(r).decode() for r in arr)
    return sorted(out)


def _user_ids(rng, T, U):
    """40-hex user ids; train and test sets disjoint (dataExtraction.ipynb cells 6, 8), each sorted."""
    allu = _random_strings(rng, T + U, 40, _HEX)
    pick = set(rng.choice(T + U, size=U, replace=False).tolist())
    return [s for i, s in enumerate(allu) if i not in pick], [s for i, s in enumerate(allu) if i in pick]


def _song_ids(rng, n):
    """'SO' + 16 [A-Z0-9] song ids, sorted, so id order == String.compareTo order."""
    return _random_strings(rng, n, 16, _ALNUM, "SO")


# Named shapes of BASELINE.json `configs`; seed = 20230000 + config number.
CONFIGS = {
    "c1": dict(T=100, U=10, S=4798, seed=20230001),
    "c2": dict(T=500, U=10, S=16785, seed=20230002),
    "c3": dict(T=2000, U=100, S=44451, seed=20230003),
    "c4": dict(T=909318, U=110000, S=384546, seed=20230004),
}


def synth_config(name: str, **over) -> Dataset:
    kw = dict(CONFIGS[name]); kw.update(over)
    ds = synth(**kw)
    ds.meta["config"] = name
    return ds


def fixture_4_3() -> Dataset:
    """The hand-derived known-answer fixture of SURVEY.md §4.3 (users A,B,C / X,Y; songs s1..s4, label-only s5)."""
    A, B, Cc = 0, 1, 2
    X, Y = 0, 1
    s1, s2, s3, s4, s5 = 0, 1, 2, 3, 4
    train = ([A, A, B, B, Cc], [s1, s2, s2, s3, s3])
    test = ([X, X, Y], [s1, s4, s2])
    labels = ([X, Y, Y], [s2, s3, s5])
    ds = from_triplets(train, test, labels, T=3, U=2, S=4,
                       train_users=["A", "B", "C"], test_users=["X", "Y"], songs=["s1", "s2", "s3", "s4", "s5"])
    ds.meta["config"] = "fixture_4_3"
    return ds
