"""Data-parallel invariant on 2 GPUs (NCCL): the update computed by two ranks, each owning half of the videos and all
K samples of them, equals the single-process update on the concatenated batch.  Skipped with < 2 GPUs."""
import os

import numpy as np
import pytest
import torch

from oracle import s2vt_numpy as M

pytestmark = pytest.mark.gpu
DIMS = dict(D=200, E=60, H=72, V=301)
Tv, Tc, B, K = 3, 7, 4, 2


def _inputs():
    rng = np.random.RandomState(5)
    p = M.init_params(seed=4, dtype=np.float32, **DIMS)
    video = M.synthetic_features(2 * B, Tv, DIMS['D'])
    cap = rng.randint(2, DIMS['V'], size=(2, K * B, Tc)).astype(np.int32)        # [rank, K*B rows (sample-major), Tc]
    cap[:, ::3, 4:] = 0
    mask = np.ones_like(cap, dtype=np.float32); mask[:, ::3, 5:] = 0
    r = rng.uniform(0, 2, (2, K * B)).astype(np.float32); b = rng.uniform(0, 2, (2, K * B)).astype(np.float32)
    return p, video, cap, mask, r, b


def _model(p):
    import s2vt_b200
    m = s2vt_b200.Video_Caption_Generator(dim_image=DIMS['D'], n_words=DIMS['V'], word_dim=DIMS['E'], lstm_dim=DIMS['H'], batch_size=2 * B,
                                          n_video_lstm_step=Tv, n_caption_lstm_step=Tc, dropout_rate=0.9, precision='fp32', max_videos=2 * K * B,
                                          max_rows=2 * K * B)
    m.load_variables(p)
    return m


def _worker(rank, world, port, out, exchange):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    import s2vt_b200
    p, video, cap, mask, r, b = _inputs()
    m = _model(p)
    m.rl_backward(video[rank * B:(rank + 1) * B], cap[rank], mask[rank], r[rank], b[rank], norm=1.0, drop_seed=11, row_base=rank * K * B)
    if exchange != 'nccl':
        assert s2vt_b200.trainer.connect_peers(m), 'the peer exchange did not come up'
        g_before = m.grads.clone()
    if exchange == 'peer_fused':      # exchange + clip + Adam on the own slice + parameter gather in one kernel per rank
        out_t = m.peer_optimizer_step(1e-3, 0.05, wemb_slice_norm=True, normalize=True)
        with pytest.raises(RuntimeError):
            m.optimizer_step(1e-3, 0.05)          # the Adam slots are sharded: the plain step refuses
        m.adam_step -= 1
        m.gather_optimizer_state()
        torch.cuda.synchronize()
        if rank == 0:
            torch.save({'params': m.params.cpu(), 'out': out_t.cpu(), 'm': m.adam_m.cpu(), 'v': m.adam_v.cpu()}, out)
        dist.barrier()
        dist.destroy_process_group()
        return
    s2vt_b200.trainer.allreduce_gradients(m)
    if exchange == 'peer':      # the same sum through NCCL
        g_peer = m.grads.clone()
        m.grads.copy_(g_before)
        m.peer_world = 0
        s2vt_b200.trainer.allreduce_gradients(m)
        err = float((g_peer - m.grads).abs().max() / m.grads.abs().max())
        assert err < 1e-6, err
        m.grads.copy_(g_peer)
    out_t = m.optimizer_step(1e-3, 0.05, wemb_slice_norm=True, normalize=True)
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({'params': m.params.cpu(), 'out': out_t.cpu(), 'm': m.adam_m.cpu(), 'v': m.adam_v.cpu()}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('exchange', ['nccl', 'peer', 'peer_fused'])
def test_two_rank_update_equals_single_process_update(tmp_path, exchange):
    """exchange: the NCCL all-reduce, or the library's own kernel over NVLink peer memory (csrc/peer.cuh; also held equal to NCCL's sum)."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    out = str(tmp_path / 'dp.pt')
    mp.spawn(_worker, args=(2, 29633 + ['nccl', 'peer', 'peer_fused'].index(exchange), out, exchange), nprocs=2, join=True)
    got = torch.load(out)
    # single process: the same rows as one batch.  Rows are sample-major per rank, so the concatenated batch keeps each
    # rank's block with its own videos: use one video per row (the literal feed) and the matching Philox row ids.
    p, video, cap, mask, r, b = _inputs()
    m = _model(p)
    vid_rows = np.concatenate([np.concatenate([video[rk * B:(rk + 1) * B]] * K, 0) for rk in range(2)], 0)
    m2 = m
    m2.rl_backward(vid_rows, cap.reshape(-1, Tc), mask.reshape(-1, Tc), r.reshape(-1), b.reshape(-1), norm=1.0, drop_seed=11, row_base=0)
    ref = m2.optimizer_step(1e-3, 0.05, wemb_slice_norm=True, normalize=True)
    e = float((got['params'] - m2.params.cpu()).abs().max() / m2.params.abs().max().item())
    print('\n[dp] 2-rank vs 1-process parameters rel diff %.3e, grad norm %.5f vs %.5f, loss %.6f vs %.6f'
          % (e, got['out'][0], ref[0].item(), got['out'][1], ref[1].item()))
    assert e < 1e-5
    for slot, ref_slot in (('m', m2.adam_m), ('v', m2.adam_v)):       # the Adam slots (gathered from their slices in the fused form)
        es = float((got[slot] - ref_slot.cpu()).abs().max() / ref_slot.abs().max().item())
        assert es < 1e-4, (slot, es)
    assert abs(got['out'][0] - ref[0].item()) < 1e-4 * ref[0].item() and abs(got['out'][1] - ref[1].item()) < 1e-4 * abs(ref[1].item()) + 1e-7
