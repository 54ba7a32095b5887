"""musicrecommendation_b200 — B200-native scoring path for MusicRecommendation (UBM / IBM cosine scoring,
blends, top-500).  The compute lives in csrc/ (hand-written sm_100a CUDA behind the C-ABI of include/mrscore.h);
this package holds the host-side mirror of the reference's `MusicRecommender` interface."""
from .dataset import Dataset, synth, synth_config, fixture_4_3, from_triplets, CONFIGS  # noqa: F401
