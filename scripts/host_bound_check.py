"""Is the REINFORCE step host-launch-bound?  Compares host enqueue time per step with device time per step."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import __graft_entry__ as ge
ge.build()
import s2vt_b200
vocab, by, order = bench.load_corpus()
w2i, bias = bench.peaked_bias(vocab, by)
B, K, Tv = 64, 5, 80
model = s2vt_b200.Video_Caption_Generator(batch_size=B, n_video_lstm_step=Tv, bias_init_vector=bias, max_videos=B, max_rows=K * B)
scorer = s2vt_b200.cider.CiderD([by[v] for v in order], w2i)
tr = s2vt_b200.trainer.ReinforceTrainer(model, scorer, n_samples=K)
feats = torch.from_numpy(bench.features(B, Tv, 1)).cuda(); vidx = torch.arange(B, dtype=torch.int32, device='cuda')
for _ in range(3):
    tr.step(feats, vidx)
torch.cuda.synchronize()
n = 5
t0 = time.perf_counter()
for _ in range(n):
    tr.step(feats, vidx)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print('host enqueue per step %.2f ms, total per step %.2f ms (launches per step %d)' % (1e3 * (t1 - t0) / n, 1e3 * (t2 - t0) / n, model.launch_count() // 8))
# rollout only
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(n):
    model.rollout(feats, K, 1)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print('rollout: host %.2f ms, total %.2f ms' % (1e3 * (t1 - t0) / n, 1e3 * (t2 - t0) / n))
