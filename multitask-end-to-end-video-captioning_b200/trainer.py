"""Training / decoding loops of the reference scripts on top of the B200 library.

`ReinforceTrainer.step` is one iteration of the hot loop of reinforcement_multisampling_tf_s2vt.py:734-829:
K sampled captions + greedy baseline per video, CIDEr-D rewards, REINFORCE gradient, global-norm clip, Adam -- with
every stage on the device and, under torch.distributed (one process per GPU), one NCCL all-reduce of the flat
gradient block per iteration.
"""
import numpy as np
import torch

try:
    import torch.distributed as dist
except Exception:  # pragma: no cover
    dist = None


def _world():
    if dist is not None and dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def exponential_decay(lr0, global_step, decay_steps, rate=0.5):
    """tf.train.exponential_decay(..., staircase=True) (:639-640; tf_s2vt.py:442-443)."""
    return lr0 * rate ** (global_step // decay_steps)


_AR_STREAMS = {}
# Which early-final gradient segments are all-reduced UNDER the backward kernels (S2VT_AR_SEGMENTS="0,1,2": embed_word_W/b, Wemb,
# LSTM2).  Default: none.  Measured on 8 B200 (round 2, 20 iterations each): no overlap 9.25 ms per iteration, segment 0 only 9.28 ms,
# segments 0,1,2 9.37 ms -- the NCCL kernels take SMs and L2 bandwidth from the persistent recurrent chains they run beside, which
# costs more than the 0.39 ms the collective takes on an idle GPU.
import os as _os
EARLY_SEGMENTS = tuple(int(x) for x in _os.environ.get('S2VT_AR_SEGMENTS', '').split(',') if x.strip() != '')


# The gradient exchange itself: 'peer' = the library's own kernels over NVLink peer memory (csrc/peer.cuh; every rank maps the other ranks'
# state blocks through CUDA IPC) with the exchange FUSED with the optimiser step where a trainer allows it (reduce-scatter, Adam on the own
# slice, all-gather of the parameters), 'peer_allreduce' = the peer kernel as a plain all-reduce + the full Adam pass, 'nccl' =
# torch.distributed all-reduce.  S2VT_DP_EXCHANGE overrides; the default tries 'peer' and keeps NCCL when the memory cannot be shared
# (another node, an allocator that does not hand out IPC-capable memory).
DP_EXCHANGE = _os.environ.get('S2VT_DP_EXCHANGE', 'peer')


def connect_peers(model):
    """Collective: map the gradient blocks of all ranks of the default process group into this rank's library handle.  Returns True when
    allreduce_gradients will use the peer kernel from now on; every rank gets the same answer."""
    rank, world = _world()
    if world == 1 or not model.grads.is_cuda or DP_EXCHANGE not in ('peer', 'peer_allreduce') or not hasattr(model, 'peer_export'):
        return False
    try:
        mine = model.peer_export()
    except Exception as e:          # not shareable: say so once, keep NCCL
        mine = None
        if rank == 0:
            print('s2vt: peer exchange unavailable (%s); using the NCCL all-reduce' % e)
    exports = [None] * world
    dist.all_gather_object(exports, mine)
    ok = all(e is not None for e in exports)
    if ok:
        try:
            model.peer_connect(rank, exports)
        except Exception as e:
            ok = False
            print('s2vt: rank %d cannot map its peers (%s); using the NCCL all-reduce' % (rank, e))
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok))
    if not all(flags):
        if getattr(model, 'peer_world', 0):
            model.peer_disconnect()
        return False
    return True


def allreduce_gradients(model, bucket_bytes=0, overlap=True):
    """Sum the flat fp32 gradient block (+ aux slots: slice norm, loss, sum(mask)) over the ranks.

    overlap: three ranges are final before the backward call's last kernel -- embed_word_W / embed_word_b before the BPTT chains
    start, Wemb and the LSTM2 weights while the LSTM1 chain still runs.  For the segments listed in EARLY_SEGMENTS the all-reduce is
    enqueued on a stream that waits for exactly that point (`s2vt_grad_segment_ready`), so it runs under the remaining kernels; the
    default list is empty because that overlap measured SLOWER than one collective after the backward call (see EARLY_SEGMENTS).
    What is left follows once the backward call has finished, one collective per contiguous range (bucket_bytes > 0 splits ranges
    into buckets; NVSwitch bandwidth does not ask for it)."""
    rank, world = _world()
    if world == 1:
        return
    if getattr(model, 'peer_world', 0) == world:
        model.peer_allreduce()       # one kernel per rank over NVLink peer memory, on the current stream
        return
    g = model.grads
    n = g.numel()
    works, done = [], []
    if overlap and g.is_cuda and hasattr(model, 'grad_segment_ready'):
        dev = g.device
        st = _AR_STREAMS.get(dev)
        if st is None:
            st = _AR_STREAMS[dev] = torch.cuda.Stream(device=dev)
        for segment in EARLY_SEGMENTS:
            seg = model.grad_segment_ready(st, segment)
            if seg is None:
                continue
            off, cnt = seg
            with torch.cuda.stream(st):
                works.append(dist.all_reduce(g[off:off + cnt], op=dist.ReduceOp.SUM, async_op=True))
            done.append((off, off + cnt))
    for lo, hi in complement_ranges(done, n):
        step = max(1, bucket_bytes // 4) if bucket_bytes > 0 else hi - lo
        for i in range(lo, hi, step):
            works.append(dist.all_reduce(g[i:min(hi, i + step)], op=dist.ReduceOp.SUM, async_op=True))
    for w in works:
        w.wait()


def complement_ranges(done, n):
    """[0, n) minus the (disjoint) half-open ranges in `done`, as a sorted list of half-open ranges."""
    out, pos = [], 0
    for lo, hi in sorted(done):
        if lo > pos:
            out.append((pos, lo))
        pos = max(pos, hi)
    if pos < n:
        out.append((pos, n))
    return out


class FeaturePipe(object):
    """Double-buffered staging of the per-step feed (features [B, T_v, D] fp32 and the corpus index of each video) from
    pinned host memory, on its own copy stream, so the copy of step i+1 runs under the kernels of step i.  The
    reference feeds through feed_dict, i.e. a blocking copy before every sess.run (:823-826).

        pipe.put(host_features, host_index)       # enqueue the copy of a future step (at most 2 outstanding)
        v, vi, slot = pipe.get()                  # device tensors of the oldest staged step, ordered after its copy
        ... launch the step ...
        pipe.release(slot)                        # the step's kernels have been enqueued; the slot may be refilled

    Host features may be float32 (the reference feed) or float16 (`ingest.FeatureFile.cache_fp16`): the tensor-core mode rounds the frames
    to fp16 before its first product anyway, so an fp16 host cache feeds bit-identical numbers with half the host -> device bytes (which
    matters when eight ranks share one host's memory and PCIe bandwidth).  The device-side fp16 -> fp32 widening runs on the copy stream.
    """

    def __init__(self, device, batch, n_frames, dim_image, depth=2):
        self.device = torch.device(device)
        self.bufs = [(torch.empty(batch, n_frames, dim_image, dtype=torch.float32, device=self.device),
                      torch.empty(batch, dtype=torch.int32, device=self.device)) for _ in range(depth)]
        self.half = [None] * depth                  # fp16 landing buffers, allocated on first use
        self.stream = torch.cuda.Stream(device=self.device)
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.free = [None] * depth
        self.rows = [0] * depth
        self.n_put = self.n_get = 0

    def put(self, features, video_index):
        i = self.n_put % len(self.bufs)
        if self.n_put - self.n_get >= len(self.bufs):
            raise RuntimeError('FeaturePipe: every slot holds a staged step; get() one first')
        f = features if torch.is_tensor(features) else torch.from_numpy(np.ascontiguousarray(features, dtype=(np.float16 if getattr(features, 'dtype', None) == np.float16 else np.float32)))
        vi = video_index if torch.is_tensor(video_index) else torch.from_numpy(np.ascontiguousarray(video_index, dtype=np.int32))
        v, x = self.bufs[i]
        if self.free[i] is not None:
            self.stream.wait_event(self.free[i])       # kernels of the step that last read this slot
        with torch.cuda.stream(self.stream):
            if f.dtype == torch.float16:
                if self.half[i] is None:
                    self.half[i] = torch.empty_like(v, dtype=torch.float16)
                self.half[i][:f.shape[0]].copy_(f, non_blocking=True)
                v[:f.shape[0]].copy_(self.half[i][:f.shape[0]])          # widen on the device
            else:
                v[:f.shape[0]].copy_(f, non_blocking=True)
            x[:vi.shape[0]].copy_(vi, non_blocking=True)
            self.ready[i].record(self.stream)
        self.rows[i] = f.shape[0]
        self.n_put += 1

    def get(self):
        if self.n_get >= self.n_put:
            raise RuntimeError('FeaturePipe: nothing staged')
        i = self.n_get % len(self.bufs)
        torch.cuda.current_stream(self.device).wait_event(self.ready[i])
        self.n_get += 1
        v, x = self.bufs[i]
        return v[:self.rows[i]], x[:self.rows[i]], i

    def release(self, slot):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.free[slot] = ev


class ReinforceTrainer(object):
    """Stage-2 trainer (K-sample REINFORCE with CIDEr-D reward and greedy baseline).

    Data parallel rule: each rank owns its videos and all K samples of them; gradients are accumulated
    with norm = 1, summed over ranks together with sum(mask), and divided by the global sum(mask) inside the
    optimiser kernel, so the update equals the single-process update on the concatenated batch.
    """

    def __init__(self, model, scorer, n_samples=5, start_learning_rate=1e-6, decay_steps=1000, clip_norm=5.0, seed=2024,
                 dropout=True, wemb_slice_norm=True):
        self.model, self.scorer, self.K = model, scorer, int(n_samples)
        self.lr0, self.decay_steps, self.clip = start_learning_rate, decay_steps, clip_norm
        self.seed, self.dropout, self.wemb_slice_norm = int(seed), dropout, wemb_slice_norm
        self.global_step = 0
        self.last = {}
        model.set_reuse_frontend(True)      # rollout and update run on the same feature tensor within step()
        self.peer_exchange = connect_peers(model)      # collective under torch.distributed; False -> NCCL all-reduce

    def step(self, video, video_index):
        """video: float32 [B, T_v, D] on the device or in (pinned) host memory; video_index: int32 [B] corpus video
        of each row (reference lists for the reward).  Returns a device tensor [global grad norm, loss]."""
        m, K = self.model, self.K
        rank, world = _world()
        v = video if torch.is_tensor(video) else torch.from_numpy(np.ascontiguousarray(video, dtype=np.float32))
        v = v.to(m.device, torch.float32, non_blocking=True)
        vi = torch.as_tensor(video_index).to(m.device, torch.int32, non_blocking=True)
        B = v.shape[0]
        it = self.global_step
        row_base = rank * K * B
        samp, greedy = m.rollout(v, K, seed=self.seed + it, row_base=row_base)                    # :743-753
        mask, _ = m.caption_masks(samp)                                                           # :784
        scores = self.scorer.score_ids(torch.cat([samp, greedy]), vi.repeat(K + 1)).to(torch.float32)    # one launch for both calls
        r = scores[:K * B]                                                                        # :806
        b = scores[K * B:].repeat(K)                                                              # :790-795
        drop_seed = (self.seed * 7919 + it + 1) if self.dropout else 0
        m.rl_backward(v, samp, mask, r, b, norm=1.0, drop_seed=drop_seed, row_base=row_base)      # :643-650, norm deferred
        lr = exponential_decay(self.lr0, it, self.decay_steps)
        if self.peer_exchange and DP_EXCHANGE == 'peer':      # exchange + clip + Adam in one kernel per rank
            out = m.peer_optimizer_step(lr, self.clip, wemb_slice_norm=self.wemb_slice_norm, normalize=True)
        else:
            allreduce_gradients(m)
            out = m.optimizer_step(lr, self.clip, wemb_slice_norm=self.wemb_slice_norm, normalize=True)   # :650-652
        self.global_step += 1
        self.last = dict(samples=samp, greedy=greedy, rewards=r, baseline=b, mask=mask)
        return out


class XETrainer(object):
    """Stage-1 trainer: tf_s2vt.py:442-448, 482-497 (label-smoothed XE + L2, Adam 1e-3 halved every 5000, clip 10)."""

    def __init__(self, model, start_learning_rate=1e-3, decay_steps=5000, clip_norm=10.0, seed=2024, dropout=True):
        self.model, self.lr0, self.decay_steps, self.clip = model, start_learning_rate, decay_steps, clip_norm
        self.seed, self.dropout, self.global_step = int(seed), dropout, 0
        self.peer_exchange = connect_peers(model)

    def step(self, video, captions, mask):
        m = self.model
        rank, world = _world()
        it = self.global_step
        drop_seed = (self.seed * 7919 + it + 1) if self.dropout else 0
        n = captions.shape[0] if hasattr(captions, 'shape') else len(captions)
        if world == 1:
            loss = m.xe_backward(video, captions, mask, drop_seed=drop_seed, row_base=0).clone()
        else:
            # Q3 couples the rows of a batch (step loss = mean_b(CE_b) * sum_b mask[b, i] / sum(mask)): exchange the per-step mask sums
            # and the row count first, then every rank back-propagates its share with the GLOBAL statistics and the shares add up --
            # the update equals the single-process one on the concatenated batch.  Weight decay is added once (rank 0).
            msk = m._f32(mask)
            stats = torch.cat([msk.sum(0), torch.tensor([float(n)], device=msk.device)])
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)
            colsum, n_global = stats[:-1].contiguous(), int(round(float(stats[-1].item())))
            loss = m.xe_backward_sharded(video, captions, msk, colsum, n_global, norm=float(colsum.sum().item()),
                                         decay=(None if rank == 0 else 0.0), drop_seed=drop_seed, row_base=rank * n).clone()
            allreduce_gradients(m, overlap=False)
            dist.all_reduce(loss, op=dist.ReduceOp.SUM)
        m.optimizer_step(exponential_decay(self.lr0, it, self.decay_steps), self.clip, wemb_slice_norm=False)
        self.global_step += 1
        return loss


def greedy_decode_all(model, features_by_video, batch_size):
    """evaluation() loop (reinforcement_multisampling_tf_s2vt.py:969-976): greedy ids for every video, in chunks."""
    vids = list(features_by_video)
    out = {}
    for i in range(0, len(vids), batch_size):
        chunk = vids[i:i + batch_size]
        feats = np.stack([features_by_video[v] for v in chunk]).astype(np.float32)
        ids = model.greedy(feats).cpu().numpy()
        for v, row in zip(chunk, ids):
            out[v] = row
    return out
