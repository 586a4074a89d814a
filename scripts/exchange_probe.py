#!/usr/bin/env python
"""Where the data-parallel exchange spends its time, per variant (run under torchrun on N GPUs):
  * the exchange alone, back to back (warm): own peer kernel vs NCCL all-reduce;
  * inside the training step: CUDA events around [exchange + optimiser step + refresh] of every iteration, and the whole iteration.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/exchange_probe.py [steps]"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import s2vt_b200  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    vocab, by, order = bench.load_corpus()
    w2i, bias = bench.peaked_bias(vocab, by)
    B, K, Tv = 64, 5, 80
    D = bench.DIMS
    model = s2vt_b200.Video_Caption_Generator(dim_image=D['D'], n_words=D['V'], word_dim=D['E'], lstm_dim=D['H'], batch_size=B, n_video_lstm_step=Tv,
                                              n_caption_lstm_step=35, bias_init_vector=bias, dropout_rate=0.9, precision='bf16', max_videos=B, max_rows=K * B, seed=4)
    model.variable('embed_word_W').mul_(3.0)
    model.refresh()
    scorer = s2vt_b200.cider.CiderD([by[v] for v in order], w2i)
    tr = s2vt_b200.trainer.ReinforceTrainer(model, scorer, n_samples=K, start_learning_rate=1e-6, decay_steps=1000, clip_norm=5.0, seed=2024)
    assert tr.peer_exchange, 'peer exchange did not come up'
    feats = torch.from_numpy(bench.features(B, Tv, 1234 + rank)).cuda()
    vidx = torch.from_numpy(((rank * B + np.arange(B)) % len(order)).astype(np.int32)).cuda()
    out = {'world': world}

    def note(msg):
        torch.cuda.synchronize()
        print('[rank %d] %s' % (rank, msg), file=sys.stderr, flush=True)

    def timed(fn, n):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / n], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- alone, warm
    for name in ('peer', 'nccl'):
        model.peer_world = world if name == 'peer' else 0
        s2vt_b200.trainer.allreduce_gradients(model)
        out['alone_%s_ms' % name] = timed(lambda: s2vt_b200.trainer.allreduce_gradients(model), 20)
        note('alone %s done' % name)
    model.peer_world = world
    model.grads.zero_()            # the timing loops below must not walk the parameters away (no backward pass fills the block here)
    out['alone_peer_fused_step_ms'] = timed(lambda: model.peer_optimizer_step(0.0, 5.0, normalize=False), 20)
    note('fused alone done')
    model.gather_optimizer_state()
    note('gather done')
    out['alone_adam_full_ms'] = timed(lambda: model.optimizer_step(0.0, 5.0, normalize=False), 20)

    # ---- in the step: events around the exchange + optimiser region
    marks = []
    orig_fused, orig_step, orig_ar = model.peer_optimizer_step, model.optimizer_step, s2vt_b200.trainer.allreduce_gradients

    def fused(*a, **k):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); r = orig_fused(*a, **k); e1.record(); marks.append((e0, e1))
        return r

    state = {}

    def ar(m, *a, **k):
        state['e0'] = torch.cuda.Event(enable_timing=True); state['e0'].record()
        return orig_ar(m, *a, **k)

    def step(*a, **k):
        r = orig_step(*a, **k)
        e1 = torch.cuda.Event(enable_timing=True); e1.record(); marks.append((state['e0'], e1))
        return r

    model.peer_optimizer_step = fused; model.optimizer_step = step; s2vt_b200.trainer.allreduce_gradients = ar
    for variant in ('peer', 'peer_allreduce', 'nccl'):
        s2vt_b200.trainer.DP_EXCHANGE = variant
        model.gather_optimizer_state()
        model.peer_world = 0 if variant == 'nccl' else world
        tr.peer_exchange = variant != 'nccl'
        for _ in range(3):
            tr.step(feats, vidx)
        del marks[:]
        ms = timed(lambda: tr.step(feats, vidx), steps)
        torch.cuda.synchronize()
        region = float(np.mean([a.elapsed_time(b) for a, b in marks[-steps:]]))
        t = torch.tensor([region], device='cuda'); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out['step_%s_ms' % variant] = ms
        note('variant %s done' % variant)
        out['step_%s_exchange_region_ms' % variant] = float(t.item())
    model.gather_optimizer_state()
    if rank == 0:
        print(json.dumps(out, indent=1))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
