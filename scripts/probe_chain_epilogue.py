import collections, ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, '/root/repo')
import __graft_entry__ as ge
import s2vt_b200
B, K, Tv = 64, 5, 80
m = s2vt_b200.Video_Caption_Generator(batch_size=B, n_video_lstm_step=Tv, max_videos=B, max_rows=K * B)
video = torch.rand(B, Tv, 1536, device='cuda')
samp, gr = m.rollout(video, K, 1)
mask, _ = m.caption_masks(samp)
r = torch.rand(K * B, device='cuda'); b = torch.rand(K * B, device='cuda')
m.rl_backward(video, samp, mask, r, b, drop_seed=3); torch.cuda.synchronize()
buf = torch.zeros(8 * 4001, dtype=torch.int64, device='cuda')
m.lib.s2vt_debug_probe(C.c_void_p(buf.data_ptr()))
m.set_reuse_frontend(False)
samp, gr = m.rollout(video, K, 2)
m.rl_backward(video, samp, mask, r, b, drop_seed=3)
torch.cuda.synchronize()
a = buf.cpu().numpy(); n = int(a[0]); rec = a[8:8 * (n + 1)].reshape(n, 8)
agg = collections.OrderedDict()
for i in range(1, n):
    t = rec[i]
    if t[6] == 0 or t[0] == 0: continue
    meta = int(t[6]); key = (meta >> 32, meta & 0xffffffff)
    if (key[0] % 100000) // 1000 != 4: continue
    agg.setdefault(key, []).append((t[7] - t[4], t[2] - t[7], t[5] - t[2], t[5] - t[4]))
print('split-K chains: tmem_ld + DSMEM send | cluster barrier | reduce + cell math + stores | epilogue total (ns, medians)')
for k, v in agg.items():
    print(k, len(v), np.round(np.median(np.array(v, dtype=np.float64), axis=0)).astype(int).tolist())
