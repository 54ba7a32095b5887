import os, sys, time
sys.path.insert(0, '/root/repo')
from musicrecommendation_b200 import _lib
from musicrecommendation_b200.dataset import synth_config
from musicrecommendation_b200.recommender import MusicRecommender
import torch
ds = synth_config('c4').shard_test_users(0, 2368)
for chunk in sys.argv[1:]:
    if chunk != 'default': os.environ['MRSCORE_PRECOMPUTE_CHUNK'] = chunk
    mr = MusicRecommender(ds, engine=_lib.MR_ENGINE_SPARSE, space=_lib.MR_SPACE_ITEM)
    torch.cuda.synchronize(); t0 = time.perf_counter(); mr.prepare(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    s = mr.getTopK(_lib.MR_UBM, k=500)[0]
    print(chunk, 'precompute wall ms', round(dt * 1e3, 1), 'n_head', mr.info()['n_head'], 'checksum', int(s[:, 0].astype('int64').sum()), flush=True)
    mr.close()
