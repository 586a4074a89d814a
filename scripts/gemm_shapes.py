"""Isolated timing of every batched-GEMM shape of one REINFORCE iteration (B=64, K=5, T_v=80) against the measured bf16 peak:
python scripts/gemm_shapes.py [backend ...]   (default: auto single_cta).  Each shape runs alone on the GPU (no side stream), 20
repetitions after 3 warm-ups, CUDA events; operands are zeros (tensor-core timing does not depend on the values)."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
if __name__ == '__main__':
    ge.build()
import s2vt_b200

# (name, M, N, K, mn_major, fp32_out) -- padded sizes as the engine launches them
SHAPES = [
    ('frame projection img = X.We', 5120, 512, 1536, 0, 0),
    ('G1x = img.W1x', 5120, 4096, 512, 0, 1),
    ('G2x rollout = h1.W2x', 7360, 4096, 1024, 0, 1),
    ('G2x training = out1.W2x', 36800, 4096, 1024, 0, 1),
    ('logits = out2.Wo', 11200, 9984, 1024, 0, 1),
    ('Etab = Wemb.W2e', 9984, 4096, 512, 0, 1),
    ('dout2 = dlogits.Wo^T', 11200, 1024, 9984, 0, 1),
    ('dout1 = dG2.W2x^T', 36800, 1024, 4096, 0, 1),
    ('dEmb = dG2.W2e^T', 11200, 512, 4096, 0, 1),
    ('dimg = dG1.W1x^T', 5120, 512, 4096, 0, 1),
    ('dWo = out2^T.dlogits', 1024, 9984, 11200, 1, 1),
    ('dW2[out1] = out1^T.dG2', 1024, 4096, 36800, 1, 1),
    ('dW2[emb] = emb^T.dG2', 512, 4096, 11200, 1, 1),
    ('dW1[h] = h1^T.dG1', 1024, 4096, 7360, 1, 1),
    ('dW1[x] = img^T.dG1', 512, 4096, 5120, 1, 1),
    ('dWe = X^T.dimg', 1536, 512, 5120, 1, 1),
]
def time_shapes(m, torch, reps=20):
    """{(name, M, N, K, mn): us} for every shape, each alone on the GPU through the handle's own dispatch."""
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = {}
    for name, M, N, K, mn, f32 in SHAPES:
        run = lambda: m._check(m.lib.s2vt_debug_gemm(m.h, M, N, K, mn, f32, st))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record(); torch.cuda.synchronize()
        out[(name, M, N, K, mn)] = e0.elapsed_time(e1) * 1e3 / reps
    return out


def main():
    backends = sys.argv[1:] or ['auto', 'single_cta']
    pk = os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')
    peaks = json.load(open(pk)) if os.path.exists(pk) else {}
    peak = peaks.get('bf16_tflops', 1681.8)
    rows = {}
    for be in backends:
        m = s2vt_b200.Video_Caption_Generator(batch_size=64, n_video_lstm_step=80, max_videos=64, max_rows=320, gemm_backend=be)
        for k, us in time_shapes(m, torch).items():
            rows.setdefault(k, {})[be] = us
        del m
        torch.cuda.empty_cache()
    print('| product | M | N | K | form | ' + ' | '.join('%s us | TF/s | of %.0f' % (b, peak) for b in backends) + ' |')
    print('|---|---:|---:|---:|---|' + '---:|---:|---:|' * len(backends))
    out = []
    for (name, M, N, K, mn), r in rows.items():
        cells = []
        for b in backends:
            tf = 2.0 * M * N * K / (r[b] * 1e-6) / 1e12
            cells.append('%.1f | %.0f | %.2f' % (r[b], tf, tf / peak))
        print('| %s | %d | %d | %d | %s | %s |' % (name, M, N, K, 'X^T.Y' if mn else 'A.B^T', ' | '.join(cells)))
        out.append(dict(name=name, M=M, N=N, K=K, mn_major=mn, us=r))
    json.dump(dict(peak_tflops=peak, shapes=out), open('gpurun_out/gemm_shapes.json', 'w'))


if __name__ == '__main__':
    main()
