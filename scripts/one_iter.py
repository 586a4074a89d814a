"""A few full REINFORCE iterations at the bench configuration (target for ncu captures)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import s2vt_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
vocab, by, order = bench.load_corpus()
w2i, bias = bench.peaked_bias(vocab, by)
B, K, Tv = 64, 5, 80
model = s2vt_b200.Video_Caption_Generator(batch_size=B, n_video_lstm_step=Tv, bias_init_vector=bias, max_videos=B, max_rows=K * B)
scorer = s2vt_b200.cider.CiderD([by[v] for v in order], w2i)
tr = s2vt_b200.trainer.ReinforceTrainer(model, scorer, n_samples=K)
feats = torch.from_numpy(bench.features(B, Tv, 1)).cuda(); vidx = torch.arange(B, dtype=torch.int32, device='cuda')
for _ in range(n):
    tr.step(feats, vidx)
torch.cuda.synchronize()
print('done')
