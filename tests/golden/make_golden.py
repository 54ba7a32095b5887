"""Regenerate the golden vectors under tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

The reference ships no fixtures, so the goldens are:
  fixture_4_3.json   the hand-derived known-answer fixture of SURVEY.md §4.3 (typed in from the survey, NOT computed)
  small_seed11.npz   outputs of the as-written restatement (oracle.naive_scores) and of the canonical restatement on a
                     seeded synthetic data set small enough for the naive loops — pins both restatements against drift.
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

import oracle  # noqa: E402
from musicrecommendation_b200.dataset import synth  # noqa: E402

HERE = Path(__file__).resolve().parent

FIXTURE = {
    "source": "SURVEY.md §4.3 (hand-derived from MusicRecommender.scala semantics)",
    "train": [["A", "s1"], ["A", "s2"], ["B", "s2"], ["B", "s3"], ["C", "s3"]],
    "test": [["X", "s1"], ["X", "s4"], ["Y", "s2"]],
    "labels": [["X", "s2"], ["Y", "s3"], ["Y", "s5"]],
    "pairs": [["X", "s2"], ["X", "s3"], ["Y", "s1"], ["Y", "s3"], ["Y", "s4"]],
    "deg_song": {"s1": 2, "s2": 3, "s3": 2, "s4": 1},
    "ubm_counts": {"X": [1, 0, 0], "Y": [1, 1, 0]},
    "ubm": [0.4999999999999999, 0.0, 0.7071067811865475, 0.7071067811865475, 0.0],
    "ibm": [0.40824829046386296, 0.0, 0.40824829046386296, 0.40824829046386296, 0.0],
    "lc_0.5": [0.45412414523193145, 0.0, 0.5576775358252052, 0.5576775358252052, 0.0],
    "agg_0.5": [0.40824829046386296, 0.0, 0.7071067811865475, 0.7071067811865475, 0.0],
    "map_rounded": 0.6666666667,
    "top2": {"ubm": {"X": ["s2", "s3"], "Y": ["s1", "s3"]}, "ibm": {"X": ["s2", "s3"], "Y": ["s1", "s3"]}},
}


def main():
    (HERE / "fixture_4_3.json").write_text(json.dumps(FIXTURE, indent=1) + "\n")
    ds = synth(T=40, U=6, S=500, seed=11)
    out = {}
    for name, m in (("ubm", oracle.UBM), ("ibm", oracle.IBM)):
        out[f"naive_{name}"] = oracle.naive_scores(ds, m)
        out[f"canon_{name}"] = oracle.canon_scores(ds, m)
        out[f"sint_{name}"] = oracle.canon_sint(ds, m)
    out["counts_ubm"] = oracle.counts_ubm(ds)
    out["gram_0_64"] = oracle.gram_rows(ds, np.arange(64))
    ts, tv, tl = oracle.topk(out["canon_ubm"], 50)
    out["top50_ubm_song"], out["top50_ubm_score"], out["top50_ubm_len"] = ts, tv, tl
    out["lc"] = oracle.blend_dense(oracle.LC, 0.5, out["canon_ubm"], out["canon_ibm"])
    out["agg"] = oracle.blend_dense(oracle.AGG, 0.5, out["canon_ubm"], out["canon_ibm"])
    out["stoch_seed42"] = oracle.blend_dense(oracle.STOCH, 0.5, out["canon_ubm"], out["canon_ibm"], seed=42)
    out["map_ubm"] = np.float64(oracle.evaluate(out["canon_ubm"], ds))
    out["map_ibm"] = np.float64(oracle.evaluate(out["canon_ibm"], ds))
    np.savez_compressed(HERE / "small_seed11.npz", **out)
    print("wrote", sorted(out))


if __name__ == "__main__":
    main()
