#!/usr/bin/env python
"""bench.py -- REINFORCE train videos/s (BASELINE.json metric) of the S2VT caption hot path on B200.

One step = one iteration of the reference's RL hot loop (reinforcement_multisampling_tf_s2vt.py:734-829) on a batch of
synthetic MSVD-shaped features: K sampled captions + greedy baseline per video, CIDEr-D rewards against the real
MSVD training references, teacher-forced forward with output dropout, BPTT, global-norm clip, Adam.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                     the reference algorithm on the host CPU (NumPy oracle)

Prints ONE JSON line (rank 0).
"""
import argparse
import gzip
import json
import os
import subprocess
import sys
import threading
import time

if '--impl' in sys.argv and 'reference' in sys.argv:
    # the CPU arm uses every host core: torchrun exports OMP_NUM_THREADS=1 to its workers, which would pin NumPy's BLAS to one thread
    # (the pools read the variable when the library loads, i.e. before `import numpy` below)
    for _v in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, 'tests', 'golden')
DIMS = dict(D=1536, E=500, H=1000, V=9972)
METRIC = 'reinforce_train_videos_per_s'
# ncu dram__bytes_read.sum + dram__bytes_write.sum per launch of the persistent chain kernels at the bench configuration, keyed by
# (rows, N, K) (profiles/r2_chain_ncu_full.md; the 64-row forward figure is the 80-step LSTM2 encoder chain)
STEP_KERNEL_DRAM_BYTES_PER_LAUNCH = {(320, 1024, 4096): 791533056, (320, 4096, 1024): 1147060992,
                                     (64, 4096, 1024): 98122496, (64, 1024, 4096): 166504960}

def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--videos', type=int, default=64, help='videos per GPU per step (BASELINE config 2: batch 64)')
    ap.add_argument('--samples', type=int, default=5, help='K sampled captions per video')
    ap.add_argument('--frames', type=int, default=80, help='T_v (BASELINE features [64, 80, 1536]; the reference literal is 5)')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--ref-videos', type=int, default=0, help='videos per step of the CPU reference arm / cpu_baseline sample (0: --videos, the same configuration)')
    ap.add_argument('--ref-seconds', type=float, default=150.0, help='CPU reference arm: stop timing further steps once this many seconds are spent (>= 2 steps are always timed)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--gemm-backend', default='auto')
    ap.add_argument('--overlap', type=int, default=-1, help='debug: side-stream overlap mask (s2vt_set_overlap)')
    ap.add_argument('--workload', default='train', choices=['train', 'beam'],
                    help="train: the REINFORCE iteration (headline metric); beam: BASELINE config 5, the beam-5 captioning sweep over batch 1..1024")
    ap.add_argument('--beam-batches', default='1,2,4,8,16,32,64,128,256,512,1024')
    ap.add_argument('--quick', action='store_true', help='timed region only (for ncu launch lists): no e2e / roofline / CPU passes')
    a = ap.parse_args()
    if a.ref_videos <= 0:
        a.ref_videos = a.videos
    return a


# ------------------------------------------------------------------------------------------------------------------
# shared synthetic workload
# ------------------------------------------------------------------------------------------------------------------
def load_corpus():
    def rd(name):
        with gzip.open(os.path.join(GOLD, name), 'rt') as f:
            return [l.rstrip('\n') for l in f]
    vocab = [l.rstrip() for l in rd('msvd_vocabulary1.txt.gz')]
    by, order = {}, []
    for line in rd('msvd_sents_train_noval_lc_nopunc.txt.gz'):
        vid, sent = line.strip().split('\t')[:2]
        if vid not in by:
            by[vid] = []; order.append(vid)
        by[vid].append(sent)
    return vocab, by, order


def peaked_bias(vocab, by):
    """'set B' of SURVEY 8(d): embed_word_b = ln(unigram count + 1) (the reference's bias_init_vector hook, :95-96)."""
    w2i = {'<eos>': 0, '<bos>': 1}
    w2i.update((w, i + 2) for i, w in enumerate(vocab))
    counts = np.zeros(len(w2i))
    for sents in by.values():
        for s in sents:
            for w in s.split(' '):
                counts[w2i.get(w, 2)] += 1
            counts[0] += 1
    return w2i, np.log(counts + 1.0)


def features(B, Tv, seed):
    rng = np.random.RandomState(seed)
    return np.maximum(0.0, rng.normal(0.25, 0.5, size=(B, Tv, DIMS['D']))).astype(np.float32)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(',')])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace('.', '').isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[4:8]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# CPU reference arm: the reference's algorithm (NumPy oracle restatement -- TF 1.1 / py2 cannot run here)
# ------------------------------------------------------------------------------------------------------------------
class OracleIteration(object):
    """reinforcement_multisampling_tf_s2vt.py:734-829 statement by statement on the host: K separate multinomial sampler
    runs + one greedy run (each re-encoding the frames), host CIDEr-D, build_loss forward at K*B rows with dropout,
    BPTT, clip 5, Adam.  NumPy (fp32 for the timing arms, fp64 for the parity test), BLAS-threaded.

    Random streams are those of trainer.ReinforceTrainer (Philox sampler seed `seed + it` over rows k*B+j, dropout seed
    `seed * 7919 + it + 1`), so one iteration here and one `ReinforceTrainer.step` see the same draws
    (tests/test_gpu_bench_config_parity.py)."""

    def __init__(self, samples, frames, videos, vocab, by, order, bias, seed=2024, lr0=1e-6, decay_steps=1000, clip=5.0, dtype=np.float32,
                 video=None, video_index=None):
        from oracle import s2vt_numpy as M, ciderd, philox
        self.M, self.philox = M, philox
        self.K, self.Tv, self.B = samples, frames, videos
        self.seed, self.lr0, self.decay_steps, self.clip, self.dtype = seed, lr0, decay_steps, clip, dtype
        self.p = M.init_params(seed=4, dtype=dtype, peaked_bias=bias, logit_scale=3.0, **DIMS)
        self.opt = M.TFAdam(self.p)
        self.scorer = ciderd.CiderD([by[v] for v in order])
        self.i2w = {0: '<eos>', 1: '<bos>'}
        self.i2w.update((i + 2, w) for i, w in enumerate(vocab))
        vidx = np.arange(self.B) % len(order) if video_index is None else np.asarray(video_index)
        self.refs = [by[order[int(j)]] for j in vidx]
        self.video = (features(self.B, self.Tv, 1234) if video is None else np.asarray(video)).astype(dtype)
        self.step_no = 0
        self.last = {}

    def step(self, use_samples=None, use_greedy=None):
        """One iteration.  use_samples / use_greedy (parity test only): continue with these ids instead of the oracle's own draws -- they
        are still drawn and kept in `last['own_samples']` / `last['own_greedy']` -- so that one categorical draw decided differently at
        an fp32 near-tie does not void the comparison of everything downstream."""
        from oracle import text, ciderd
        M, K, B = self.M, self.K, self.B
        it = self.step_no
        samples = [M.multinomial_sampler(self.p, self.video, self.seed + it, np.arange(k * B, (k + 1) * B)) for k in range(K)]   # :743-753
        greedy = M.greedy_sampler(self.p, self.video)
        samples = np.vstack(samples)                                                                  # :764-782 sample-major rows
        own = dict(own_samples=samples, own_greedy=greedy)
        if use_samples is not None:
            samples = np.asarray(use_samples)
        if use_greedy is not None:
            greedy = np.asarray(use_greedy)
        vid_rows = np.concatenate([self.video] * K, 0)
        mask, multi_decoded = text.decode_captions_masks(samples, self.i2w)                           # :784
        _, greedy_decoded = text.decode_captions_masks(greedy, self.i2w)
        ref = {i: self.refs[i % B] for i in range(K * B)}
        b = ciderd.evaluate_captions_cider(self.scorer, ref, greedy_decoded)                          # :790-795
        b = np.tile(b, K)
        r = ciderd.evaluate_captions_cider(self.scorer, ref, multi_decoded)                           # :806
        T = self.Tv + samples.shape[1]
        rows = np.arange(K * B)
        ds = self.seed * 7919 + it + 1
        d1 = np.stack([self.philox.dropout_mask(ds, self.philox.STREAM_DROP1, rows, t, DIMS['H'], 0.9) for t in range(T)]).astype(self.dtype)
        d2 = np.stack([self.philox.dropout_mask(ds, self.philox.STREAM_DROP2, rows, t, DIMS['H'], 0.9) for t in range(T)]).astype(self.dtype)
        # the rewards pass through float32 as in the feed (rewards / base_line are tf.float32 placeholders, :631-632)
        loss, grads, aux = M.rl_objective(self.p, vid_rows, samples, np.asarray(mask, np.float32), r.astype(np.float32), b.astype(np.float32), d1, d2)
        clipped, gn = M.clip_by_global_norm(grads, self.clip, emb_slice_sqnorm=aux['emb_slice_sqnorm'])     # :650
        self.p = self.opt.apply(self.p, clipped, M.exponential_decay(self.lr0, it, self.decay_steps))        # :639-652
        self.step_no += 1
        self.last = dict(samples=samples, greedy=greedy, rewards=r, baseline=b, mask=np.asarray(mask), loss=float(loss), grad_norm=float(gn), **own)
        return float(loss)


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [i.get('num_threads') for i in threadpool_info() if i.get('user_api') == 'blas']
        return int(max(n)) if n else os.cpu_count()
    except Exception:
        return os.cpu_count()


def oracle_beam_captions_per_s(args, bias, n_videos, warm=1):
    """final_beam_search.py:248-294 (= e2e_beam_search.py:301-344) as the reference runs it: one video at a time, beam 5, length
    normalisation 1, one beam_probability call per hypothesis and step; NumPy fp32 restatement (oracle/beam.py, oracle/s2vt_numpy.py)."""
    from oracle import beam as obeam
    from oracle import s2vt_numpy as M
    p = M.init_params(seed=4, dtype=np.float32, peaked_bias=bias, logit_scale=3.0, **DIMS)
    video = features(n_videos + warm, args.frames, 99)
    step = M.beam_step_fn(p, 5)

    def one(v):
        s1, s2 = M.beam_initial_states(p, video[v:v + 1])
        return obeam.beam_search(step, s1, s2, 5, 35, 1.0)

    for v in range(warm):
        one(v)
    t0 = time.perf_counter()
    for v in range(warm, warm + n_videos):
        one(v)
    return n_videos / (time.perf_counter() - t0)


def run_reference_beam(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    use_all_host_threads()
    vocab, by, order = load_corpus()
    _, bias = peaked_bias(vocab, by)
    n = max(1, min(args.ref_videos, 8)) * max(args.steps, 1)
    val = oracle_beam_captions_per_s(args, bias, n, warm=max(args.warmup, 1))
    sample = '%d videos, one at a time (the reference loop), beam 5, T_v=%d, fp32 NumPy/OpenBLAS' % (n, args.frames)
    out = {'impl': 'reference', 'metric': 'beam5_decode_captions_per_s', 'value': val, 'unit': 'captions/s', 'n_gpus': args.gpus, 'steps': args.steps,
           'warmup': args.warmup, 'ms_per_step': 1e3 / val, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32',
           'data': 'synthetic',
           'config': {'workload': 'beam-5 captioning (BASELINE config 5), batch 1 as in final_beam_search.py / e2e_beam_search.py, T_v=%d, '
                                  'length_normalization_factor 1, precomputed [1, T_v, 1536] features' % args.frames,
                      'beam_size': 5, 'T_v': args.frames, 'T_c': 35, 'n_words': DIMS['V'], 'lstm_dim': DIMS['H'], 'parallelism': 'dp1'},
           'cpu_baseline': {'value': val, 'unit': 'captions/s', 'cores': cpu_threads(), 'kind': 'port', 'sample': sample},
           'e2e': {'value': val, 'unit': 'captions/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(out), flush=True)


def shard_range(batch, rank, world):
    """Videos [lo, hi) of a batch that `rank` decodes: contiguous blocks of ceil(batch / world); ranks beyond the batch get nothing."""
    per = (batch + world - 1) // world
    return min(batch, rank * per), min(batch, (rank + 1) * per)


def use_all_host_threads():
    """BLAS pools of this process -> every host core (also when the environment asked for one thread: see the top of the file)."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:
        pass
    return n


def run_reference(args):
    """The reference's algorithm on the host cores, on the SAME configuration as the B200 arm (videos per step, K, T_v): W warm-up
    iterations, then up to K timed ones -- timing stops early once --ref-seconds are spent (an iteration on 64 videos costs tens of
    seconds), and `steps` reports what was timed."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    use_all_host_threads()
    vocab, by, order = load_corpus()
    w2i, bias = peaked_bias(vocab, by)
    o = OracleIteration(args.samples, args.frames, args.ref_videos, vocab, by, order, bias)
    for _ in range(min(args.warmup, 1)):          # one warm-up iteration pages in BLAS and the corpus; more would only burn minutes
        o.step()
    t0 = time.perf_counter()
    done, losses = 0, []
    while done < max(args.steps, 1):
        losses.append(o.step())
        done += 1
        if done >= 2 and time.perf_counter() - t0 > args.ref_seconds:
            break
    dt = time.perf_counter() - t0
    val = done * args.ref_videos / dt
    sample = ('%d timed iterations (of %d requested, %d warm-up) of the full iteration on %d videos x K=%d, T_v=%d, fp32 NumPy/OpenBLAS'
              % (done, args.steps, min(args.warmup, 1), args.ref_videos, args.samples, args.frames))
    out = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'videos/s', 'n_gpus': args.gpus, 'steps': done, 'warmup': min(args.warmup, 1),
           'ms_per_step': 1e3 * dt / done, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
           'config': workload_config(args, args.ref_videos, 1), 'loss': losses[-1], 'losses': losses,
           'cpu_baseline': {'value': val, 'unit': 'videos/s', 'cores': cpu_threads(), 'kind': 'port', 'sample': sample},
           'e2e': {'value': val, 'unit': 'videos/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(out), flush=True)


def workload_config(args, videos_per_gpu, world):
    return {'workload': 'REINFORCE iteration (rollout K+1, CIDEr-D reward, teacher-forced fwd+bwd with dropout 0.9, clip 5, Adam); '
                        'BASELINE config 2 on MSVD-shaped features',
            'videos_per_gpu': videos_per_gpu, 'global_videos': videos_per_gpu * world, 'K': args.samples, 'T_v': args.frames, 'T_c': 35,
            'dim_image': DIMS['D'], 'word_dim': DIMS['E'], 'lstm_dim': DIMS['H'], 'n_words': DIMS['V'], 'parallelism': 'dp%d' % world,
            'weights': 'random init seed 4, embed_word_b = ln(unigram count+1), embed_word_W x3 (SURVEY 8d set B)',
            'references': 'MSVD train sentences (1200 videos, 48 774 refs)',
            'l2': 'per-step working set (~3 GB of activations / gradients) is >> the 126 MB L2; no explicit flush'}


# ------------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------------
def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner from C), so the process's
    fd 1 is pointed at stderr for the whole run and the JSON line goes to a private duplicate of the original stdout."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(keep, 'w')


def beam_sweep(s2vt_b200, torch, dist, args, bias, rank, world, batches, reps=3):
    """BASELINE config 5 (e2e_beam_search.py after the CNN / final_beam_search.py): beam-5 captioning, length normalisation 1, of
    `batch` videos per call, sharded over the ranks with no collective (decode paths are independent videos).
    Device-resident features (`ms`) and host feeds (`e2e_ms`: pinned features -> device, sentences / lengths / scores -> host
    inside the timed region), CUDA events, max over ranks."""
    Tv = args.frames
    per_max = (max(batches) + world - 1) // world
    model = s2vt_b200.Video_Caption_Generator(dim_image=DIMS['D'], n_words=DIMS['V'], word_dim=DIMS['E'], lstm_dim=DIMS['H'], batch_size=per_max,
                                              n_video_lstm_step=Tv, n_caption_lstm_step=35, bias_init_vector=bias, dropout_rate=1.0, beam_size=5,
                                              precision=args.precision, max_videos=per_max, max_rows=min(per_max, 64), seed=4,
                                              gemm_backend=args.gemm_backend)
    if args.overlap >= 0:
        model.lib.s2vt_set_overlap(model.h, args.overlap)
    model.variable('embed_word_W').mul_(3.0)
    model.refresh()
    host = torch.from_numpy(features(per_max, Tv, 99 + rank)).pin_memory()
    dev = host.cuda()
    rows = []

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    pipe = s2vt_b200.trainer.FeaturePipe(model.device, per_max, Tv, DIMS['D'])
    idx_host = torch.zeros(per_max, dtype=torch.int32).pin_memory()
    res_host = [[torch.empty(per_max, 35, dtype=torch.int32).pin_memory(), torch.empty(per_max, dtype=torch.int32).pin_memory(),
                 torch.empty(per_max, dtype=torch.float32).pin_memory(), torch.empty(per_max, dtype=torch.float32).pin_memory()] for _ in range(2)]
    res_ev = [torch.cuda.Event(), torch.cuda.Event()]

    def timed(fn, whole=False):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if whole:
            fn()                          # fn runs the `reps` calls itself
        else:
            for _ in range(reps):
                fn()
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device='cuda', dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / reps

    for b in batches:
        lo, hi = shard_range(b, rank, world)
        n = hi - lo                      # this rank's videos of the batch (ranks beyond the batch idle)

        def resident():
            if n:
                model.beam_search(dev[:n], 5, 1.0)

        def fed():
            if n:
                out = model.beam_search(host[:n].cuda(non_blocking=True), 5, 1.0)
                return [x.cpu() for x in out]

        def piped():
            """a stream of calls: the features of call i+1 are staged on the copy stream under call i, results read one call behind"""
            if not n:
                return
            pipe.put(host[:n], idx_host[:n])
            pending = None
            for i in range(reps):
                if i + 1 < reps:
                    pipe.put(host[:n], idx_host[:n])
                v, _, slot = pipe.get()
                out = model.beam_search(v, 5, 1.0)
                pipe.release(slot)
                for dst, src in zip(res_host[i % 2], out):
                    dst[:n].copy_(src, non_blocking=True)
                res_ev[i % 2].record()
                if pending is not None:
                    res_ev[pending].synchronize()
                pending = i % 2
            res_ev[pending].synchronize()

        for _ in range(2):
            resident()
        ms = timed(resident)
        fed()
        ms_e2e = timed(fed)
        piped()
        ms_pipe = timed(piped, whole=True)
        rows.append({'batch': b, 'ms': ms, 'captions_per_s': b / (ms / 1e3), 'e2e_ms': ms_e2e, 'e2e_captions_per_s': b / (ms_e2e / 1e3),
                     'e2e_pipelined_ms': ms_pipe, 'e2e_pipelined_captions_per_s': b / (ms_pipe / 1e3)})
    del model
    return rows


def run_beam(args):
    """`--workload beam`: one JSON line for the secondary BASELINE metric (beam-5 decode captions/s), value = the largest batch."""
    out_stream = _claim_stdout()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1')); local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    import s2vt_b200
    vocab, by, order = load_corpus()
    _, bias = peaked_bias(vocab, by)
    batches = [int(x) for x in args.beam_batches.split(',')]
    sampler = ClockSampler(local)
    sampler.start()
    rows = beam_sweep(s2vt_b200, torch, dist, args, bias, rank, world, batches, reps=max(args.steps, 3))
    clocks = sampler.stop()
    if rank == 0:
        top = max(rows, key=lambda r: r['batch'])
        Tv = args.frames
        out = {'metric': 'beam5_decode_captions_per_s', 'value': top['captions_per_s'], 'unit': 'captions/s', 'n_gpus': world,
               'steps': max(args.steps, 3), 'warmup': 2, 'ms_per_step': top['ms'], 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
               'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
               'config': {'workload': 'beam-5 captioning sweep (BASELINE config 5): batch %d videos per call split over the GPUs, T_v=%d, '
                                      'length_normalization_factor 1, precomputed [B, T_v, 1536] features' % (top['batch'], Tv),
                          'beam_size': 5, 'T_v': Tv, 'T_c': 35, 'n_words': DIMS['V'], 'lstm_dim': DIMS['H'], 'parallelism': 'dp%d' % world,
                          'l2': 'no explicit flush: every call streams its own activations; weights (83 MB bf16) are meant to stay L2-resident'},
               'clocks': clocks,
               'e2e': {'value': top['e2e_pipelined_captions_per_s'], 'unit': 'captions/s', 'h2d_bytes_per_step': top['batch'] * Tv * DIMS['D'] * 4,
                       'd2h_bytes_per_step': top['batch'] * (35 + 3) * 4, 'ms_per_step': top['e2e_pipelined_ms'],
                       'mode': 'a stream of calls with HOST feeds: features of call i+1 staged from pinned memory on a copy stream under call i '
                               '(trainer.FeaturePipe), sentences / lengths / log-probs / scores read back one call behind; every copy inside the timed region',
                       'blocking_feed_value': top['e2e_captions_per_s']},
               'sweep': rows}
        if world == 1 and not args.no_cpu_baseline:
            n = 2
            out['cpu_baseline'] = {'value': oracle_beam_captions_per_s(args, bias, n), 'unit': 'captions/s', 'cores': cpu_threads(), 'kind': 'port',
                                   'sample': '%d videos, one at a time (the reference loop), beam 5, T_v=%d, fp32 NumPy/OpenBLAS' % (n, args.frames)}
        print(json.dumps(out), file=out_stream, flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_b200(args):
    out_stream = _claim_stdout()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1')); local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    import s2vt_b200
    vocab, by, order = load_corpus()
    w2i, bias = peaked_bias(vocab, by)
    B, K, Tv = args.videos, args.samples, args.frames
    model = s2vt_b200.Video_Caption_Generator(dim_image=DIMS['D'], n_words=DIMS['V'], word_dim=DIMS['E'], lstm_dim=DIMS['H'], batch_size=B,
                                              n_video_lstm_step=Tv, n_caption_lstm_step=35, bias_init_vector=bias, dropout_rate=0.9,
                                              precision=args.precision, max_videos=B, max_rows=K * B, seed=4, gemm_backend=args.gemm_backend)
    if args.overlap >= 0:
        model.lib.s2vt_set_overlap(model.h, args.overlap)
    model.variable('embed_word_W').mul_(3.0)
    model.refresh()
    scorer = s2vt_b200.cider.CiderD([by[v] for v in order], w2i)
    trainer = s2vt_b200.trainer.ReinforceTrainer(model, scorer, n_samples=K, start_learning_rate=1e-6, decay_steps=1000, clip_norm=5.0, seed=2024)
    feats_host = torch.from_numpy(features(B, Tv, 1234 + rank)).pin_memory()
    feats_dev = feats_host.cuda(non_blocking=True)
    vidx_host = torch.from_numpy(((rank * B + np.arange(B)) % len(order)).astype(np.int32)).pin_memory()
    vidx_dev = vidx_host.cuda(non_blocking=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(n, fn):
        """n steps bracketed by barrier + synchronize; device time via CUDA events; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device='cuda', dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    step_dev = lambda: trainer.step(feats_dev, vidx_dev)

    def step_e2e_blocking():
        out = trainer.step(feats_host, vidx_host)          # H2D of the step's features on the compute stream, inside the step
        return out.cpu()                                   # D2H of [grad norm, loss]

    pipe = s2vt_b200.trainer.FeaturePipe(model.device, B, Tv, DIMS['D'])
    res_host = torch.empty(2, 2, dtype=torch.float32).pin_memory()
    res_ev = [torch.cuda.Event(), torch.cuda.Event()]
    results = []

    def run_e2e(n):
        """n steps through the public API with HOST feeds: every step's features + indices are copied from pinned memory
        (staged one step ahead on the copy stream) and every step's [grad norm, loss] is read back (one step behind)."""
        pipe.put(feats_host, vidx_host)
        pending = None
        for i in range(n):
            if i + 1 < n:
                pipe.put(feats_host, vidx_host)
            v, vi, slot = pipe.get()
            out = trainer.step(v, vi)
            pipe.release(slot)
            res_host[i % 2].copy_(out, non_blocking=True)
            res_ev[i % 2].record()
            if pending is not None:
                res_ev[pending].synchronize()
                results.append(res_host[pending].tolist())
            pending = i % 2
        res_ev[pending].synchronize()
        results.append(res_host[pending].tolist())

    for _ in range(max(args.warmup, 3)):
        step_dev()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start(); time.sleep(0.3)
    l0 = model.launch_count()
    ms = timed(args.steps, step_dev)
    launches = model.launch_count() - l0 + args.steps              # + the CIDEr-D kernel of each step (launched outside the handle)
    clocks = sampler.stop() if sampler else None
    value = world * B * args.steps / (ms / 1e3)
    if args.quick:
        if rank == 0:
            print(json.dumps({'metric': METRIC, 'value': value, 'unit': 'videos/s', 'ms_per_step': ms / args.steps, 'gpu_launches': int(launches),
                              'quick': True}), file=out_stream, flush=True)
        return
    # ---- data-parallel extras (N > 1): strong scaling on BASELINE config 2's global batch, and the gradient all-reduce alone
    strong, collective = None, None
    if world > 1:
        nbytes = int(model.grads.numel() * 4)
        def time_exchange():
            s2vt_b200.trainer.allreduce_gradients(model, overlap=False)
            return timed(10, lambda: s2vt_b200.trainer.allreduce_gradients(model, overlap=False)) / 10
        peer_on = getattr(model, 'peer_world', 0) == world
        ar_ms = time_exchange()                      # the exchange the timed steps used
        nccl_ms = ar_ms
        if peer_on:                                  # ... and NCCL's all-reduce of the same block beside it
            model.peer_world = 0
            nccl_ms = time_exchange()
            model.peer_world = world
        collective = {'op': ('own kernel over NVLink peer memory (csrc/peer.cuh): reduce-scatter by cp.async.bulk loads from every rank, sums pushed into every rank by bulk stores, '
                             'two flag barriers, one launch per rank' if peer_on else 'NCCL all-reduce(sum)') + ' of the flat fp32 gradient block, nothing overlapped',
                      'bytes': nbytes, 'ms': ar_ms,
                      'algbw_GBps': nbytes / (ar_ms * 1e-3) / 1e9, 'busbw_GBps': nbytes / (ar_ms * 1e-3) / 1e9 * 2 * (world - 1) / world,
                      'nccl_allreduce_ms': nccl_ms, 'nvlink5_per_direction_GBps': 900.0,
                      'in_step': 'one exchange after the backward call; reducing the early-final segments under the backward kernels '
                                 '(s2vt_grad_segment_ready, S2VT_AR_SEGMENTS=0,1,2, NCCL) measured slower on 8 B200: 9.37 vs 9.25 ms per iteration',
                      'early_segments': list(s2vt_b200.trainer.EARLY_SEGMENTS),
                      'exchange_in_the_timed_steps': ('peer kernel fused with clip + Adam on the own slice (s2vt_peer_optimizer_step)'
                                                      if peer_on and s2vt_b200.trainer.DP_EXCHANGE == 'peer' else
                                                      'peer kernel all-reduce + full Adam' if peer_on else 'NCCL all-reduce + full Adam')}
        if args.videos % world == 0:
            Bs = args.videos // world
            m2 = s2vt_b200.Video_Caption_Generator(dim_image=DIMS['D'], n_words=DIMS['V'], word_dim=DIMS['E'], lstm_dim=DIMS['H'], batch_size=Bs,
                                                   n_video_lstm_step=Tv, n_caption_lstm_step=35, bias_init_vector=bias, dropout_rate=0.9,
                                                   precision=args.precision, max_videos=Bs, max_rows=K * Bs, seed=4, gemm_backend=args.gemm_backend)
            m2.variable('embed_word_W').mul_(3.0)
            m2.refresh()
            t2 = s2vt_b200.trainer.ReinforceTrainer(m2, scorer, n_samples=K, start_learning_rate=1e-6, decay_steps=1000, clip_norm=5.0, seed=2024)
            f2, v2 = feats_dev[:Bs].contiguous(), vidx_dev[:Bs].contiguous()
            for _ in range(3):
                t2.step(f2, v2)
            ms_s = timed(args.steps, lambda: t2.step(f2, v2))
            strong = {'scaling': 'strong', 'global_videos': args.videos, 'videos_per_gpu': Bs, 'rows_per_gpu': K * Bs, 'ms_per_step': ms_s / args.steps,
                      'value': args.videos * args.steps / (ms_s / 1e3), 'unit': 'videos/s',
                      'note': 'BASELINE config 2 literally: global batch %d split over %d GPUs; the per-GPU work shrinks to %d rows, the %d sequential '
                              'recurrent steps and the all-reduce do not' % (args.videos, world, K * Bs, 3 * (Tv + 35) + 2 * 35)}
            if getattr(m2, 'peer_world', 0):      # unmap the peers' blocks on every rank BEFORE any rank frees its own (CUDA IPC rule)
                m2.peer_disconnect()
            barrier()
            del t2, m2
            torch.cuda.empty_cache()
    run_e2e(2)
    ms_e2e = timed(1, lambda: run_e2e(args.steps))
    e2e = world * B * args.steps / (ms_e2e / 1e3)
    # the same loop fed from an fp16 host cache (ingest.FeatureFile.cache_fp16): identical numbers in the tensor-core mode, half the H2D bytes
    feats_host32, feats_host = feats_host, feats_host.to(torch.float16).pin_memory()
    run_e2e(2)
    ms_e2e_h = timed(1, lambda: run_e2e(args.steps))
    feats_host = feats_host32
    for _ in range(2):
        step_e2e_blocking()
    ms_e2e_blk = timed(args.steps, step_e2e_blocking)
    # roofline pass: the same steps with CUDA-event brackets around every GEMM launch (separate from the timed region
    # above so the brackets do not perturb `value`)
    model.profile(True)
    barrier()
    for _ in range(args.steps):
        step_dev()
    shapes = model.profile_shapes()
    prof = model.profile_read()
    model.profile(False)
    loss = float(trainer.step(feats_dev, vidx_dev)[1].item())
    # every batched-GEMM shape of the iteration timed ALONE (no side-stream kernels beside it) through the engine's own dispatch
    iso = None
    if world == 1:
        sys.path.insert(0, os.path.join(ROOT, 'scripts'))
        import gemm_shapes
        t_iso = gemm_shapes.time_shapes(model, torch, reps=10)
        iso_rows = [{'product': k[0], 'M': k[1], 'N': k[2], 'K': k[3], 'form': 'X^T.Y' if k[4] else 'A.B^T', 'us': us,
                     'tflops': 2.0 * k[1] * k[2] * k[3] / (us * 1e-6) / 1e12} for k, us in t_iso.items()]
        twice = lambda r: 2 if r['product'].startswith('dW2[out1]') else 1      # the h2 rows of dW2 have the same shape
        iso = {'shapes': iso_rows, 'us_total': sum(r['us'] * twice(r) for r in iso_rows),
               'tflop_total': sum(2.0 * r['M'] * r['N'] * r['K'] * twice(r) for r in iso_rows) / 1e12}
    # secondary BASELINE metric: beam-5 decode captions/s (e2e_beam_search.py semantics, length normalisation 1), batch = B videos
    beam = {}
    if world == 1:
        for _ in range(2):
            model.beam_search(feats_dev, 5, 1.0)
        bs_ms = timed(3, lambda: model.beam_search(feats_dev, 5, 1.0))
        gr_ms = timed(3, lambda: model.greedy(feats_dev))
        beam = {'beam5_captions_per_s': 3 * B / (bs_ms / 1e3), 'beam5_ms_per_batch': bs_ms / 3, 'greedy_captions_per_s': 3 * B / (gr_ms / 1e3),
                'batch': B, 'T_v': Tv, 'beam_size': 5, 'length_normalization_factor': 1.0}
        # BASELINE config 3: temporal-attention decoder (original_attention.py) greedy decode on MSR-VTT-shaped [B, 32, 1536] features
        att = s2vt_b200.attention.Video_Caption_Generator(dim_image=DIMS['D'], n_words=DIMS['V'], dim_hidden=DIMS['H'], batch_size=B, n_video_lstm_steps=32,
                                                          n_caption_lstm_steps=35, drop_out_rate=1.0, precision=args.precision)
        feats32 = torch.from_numpy(features(B, 32, 4321)).cuda()
        for _ in range(2):
            att.build_generator(feats32)
        at_ms = timed(3, lambda: att.build_generator(feats32))
        beam.update({'attention_greedy_captions_per_s': 3 * B / (at_ms / 1e3), 'attention_ms_per_batch': at_ms / 3, 'attention_frames': 32})
        del att
        # reward kernels on the rollout's [K*B, 35] id matrix (CIDEr-D is the one inside the timed step; BLEU-4 / ROUGE-L are the
        # rewards of bleu4_/rouge_reinforcement_multisampling_tf_s2vt.py)
        ids = trainer.last['samples']
        rows = vidx_dev.repeat(K)
        rw = {}
        for name, sc in (('ciderd', scorer), ('bleu4', s2vt_b200.rewards.Bleu4([by[v] for v in order], w2i)), ('rouge_l', s2vt_b200.rewards.RougeL([by[v] for v in order], w2i))):
            sc.score_ids(ids, rows)
            rw[name + '_us_per_%d_hyps' % ids.shape[0]] = 1e3 * timed(20, lambda: sc.score_ids(ids, rows)) / 20
        beam['reward_kernels'] = rw
        # BASELINE config 3 (frame count of the attention file, 32): S2VT greedy and beam-5 decode on [B, 32, 1536]
        m32 = s2vt_b200.Video_Caption_Generator(dim_image=DIMS['D'], n_words=DIMS['V'], word_dim=DIMS['E'], lstm_dim=DIMS['H'], batch_size=B,
                                                n_video_lstm_step=32, n_caption_lstm_step=35, bias_init_vector=bias, dropout_rate=1.0, beam_size=5,
                                                precision=args.precision, max_videos=B, max_rows=B, seed=4)
        m32.variable('embed_word_W').mul_(3.0); m32.refresh()
        for _ in range(2):
            m32.beam_search(feats32, 5, 1.0); m32.greedy(feats32)
        b32 = timed(3, lambda: m32.beam_search(feats32, 5, 1.0)); g32 = timed(3, lambda: m32.greedy(feats32))
        beam['config3_s2vt_32_frames'] = {'beam5_captions_per_s': 3 * B / (b32 / 1e3), 'greedy_captions_per_s': 3 * B / (g32 / 1e3), 'batch': B, 'T_v': 32}
        del m32
        # BASELINE config 4: multitask REINFORCE step with the 400-way attribute head (reinforce_multitask_e2e_attribute_loss.py:957), batch 128,
        # single sample, T_v = 5: rollout(1) + CIDEr-D + -(1-alpha) RL backward + alpha attribute-head backward + clip 10 + Adam
        B4, alpha = 128, 0.05
        m4 = s2vt_b200.Video_Caption_Generator(dim_image=DIMS['D'], n_words=DIMS['V'], word_dim=DIMS['E'], lstm_dim=DIMS['H'], batch_size=B4,
                                               n_video_lstm_step=5, n_caption_lstm_step=35, bias_init_vector=bias, dropout_rate=0.9, n_attributes=400,
                                               precision=args.precision, max_videos=B4, max_rows=B4, seed=4)
        m4.variable('embed_word_W').mul_(3.0); m4.refresh()
        f4 = torch.from_numpy(features(B4, 5, 777)).cuda()
        v4 = torch.from_numpy((np.arange(B4) % len(order)).astype(np.int32)).cuda()
        y4 = (torch.rand(B4, 400, device='cuda') < 0.05).float()
        it4 = [0]

        def step4():
            it4[0] += 1
            samp, greedy = m4.rollout(f4, 1, seed=100 + it4[0])
            mask, _ = m4.caption_masks(samp)
            sc = scorer.score_ids(torch.cat([samp, greedy]), v4.repeat(2)).to(torch.float32)
            m4.rl_backward(f4, samp, mask, sc[:B4], sc[B4:], grad_scale=1.0 - alpha, drop_seed=it4[0])
            m4.attribute_backward(f4, y4, grad_scale=alpha)
            m4.optimizer_step(1e-6, 10.0)

        for _ in range(3):
            step4()
        c4 = timed(5, step4)
        beam['config4_multitask_attribute'] = {'videos_per_s': 5 * B4 / (c4 / 1e3), 'ms_per_step': c4 / 5, 'batch': B4, 'K': 1, 'T_v': 5, 'n_attributes': 400, 'alpha': alpha}
        del m4, f4
        del feats32
        torch.cuda.empty_cache()
        # BASELINE config 1: tf_s2vt.py cross-entropy train step (teacher-forced forward with dropout, label-smoothed CE + L2, BPTT,
        # clip 10, Adam) on [B, T_v, 1536] features and the first reference sentence of each video; measured last, it moves the weights
        unk = dict(w2i); unk.setdefault('<en_unk>', 2)
        cap_ids, cap_mask = s2vt_b200.text.sentence_padding_toix([by[order[(rank * B + j) % len(order)]][0] for j in range(B)], unk, 35)
        cap_ids, cap_mask = torch.from_numpy(cap_ids).cuda(), torch.from_numpy(cap_mask).cuda()
        for _ in range(2):
            model.xe_step(feats_dev, cap_ids, cap_mask, 1e-4)
        xe_ms = timed(5, lambda: model.xe_step(feats_dev, cap_ids, cap_mask, 1e-4))
        beam.update({'xe_train_videos_per_s': 5 * B / (xe_ms / 1e3), 'xe_ms_per_step': xe_ms / 5})
        # same-precision reference point: the whole iteration in the fp32 mode (fp32 operands and storage, SIMT FMA GEMMs, one launch per recurrent
        # step -- the 1e-5 parity mode, not a tuned product path; the reference computes in fp32)
        m32f = s2vt_b200.Video_Caption_Generator(dim_image=DIMS['D'], n_words=DIMS['V'], word_dim=DIMS['E'], lstm_dim=DIMS['H'], batch_size=B,
                                                 n_video_lstm_step=Tv, n_caption_lstm_step=35, bias_init_vector=bias, dropout_rate=0.9,
                                                 precision='fp32', max_videos=B, max_rows=K * B, seed=4)
        m32f.variable('embed_word_W').mul_(3.0); m32f.refresh()
        t32 = s2vt_b200.trainer.ReinforceTrainer(m32f, scorer, n_samples=K, start_learning_rate=1e-6, decay_steps=1000, clip_norm=5.0, seed=2024)
        t32.step(feats_dev, vidx_dev)
        f32_ms = timed(2, lambda: t32.step(feats_dev, vidx_dev)) / 2
        beam['same_precision_fp32'] = {'videos_per_s': B / (f32_ms / 1e3), 'ms_per_step': f32_ms, 'note': 'precision=fp32: SIMT fp32 GEMMs + per-step launches (parity mode)'}
        del t32, m32f
        torch.cuda.empty_cache()
    # BASELINE config 5: latency vs throughput of beam-5 captioning over the batch size, sharded over the ranks with no collective (N > 1: fewer batch sizes)
    sweep_batches = [int(x) for x in args.beam_batches.split(',')] if world == 1 else [64, 1024, 1024 * world]
    beam['beam5_sweep'] = beam_sweep(s2vt_b200, torch, dist, args, bias, rank, world, sweep_batches)
    if rank != 0:
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak_tf = peaks.get('bf16_tflops_sustained', 1400.0)
    peak_src = 'MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)' if peaks else 'fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)'
    bms, bfl, bn, bby = prof['batched']
    sms, sfl, sn, sby = prof['step']
    ach_tf = bfl / (bms * 1e-3) / 1e12 if bms > 0 else 0.0
    ach_gb = sby / (sms * 1e-3) / 1e9 if sms > 0 else 0.0
    hbm = peaks.get('hbm_gbs', 6650.0)
    # dominant kernel by time: the recurrent-step shape with the largest total (the LSTM2 backward-through-time chain at the
    # metric's configuration): one persistent launch walks all its time steps
    dom = max([x for x in shapes if x[0] == 1 and (x[1], x[2], x[3]) in STEP_KERNEL_DRAM_BYTES_PER_LAUNCH] or [x for x in shapes if x[0] == 1], key=lambda x: x[4])
    _, dM, dN, dK, dms, dcnt, dby, dln = dom
    dname = 'EpiLstmBwd' if dK > dN else 'EpiLstmFwd'
    d_ach = dby / (dms * 1e-3) / 1e9 if dms > 0 else 0.0
    def chain_kernel_name(M_, N_, K_):
        if K_ > N_:
            return 'tc::gemm_tc_chain_kernel<%d, EpiLstmBwd<bf16, f16>, 4> (persistent BPTT chain, split-K cluster of 4)' % (128 if M_ > 128 else 32)
        if M_ > 128:
            return 'tc::gemm_tc_ws2_chain_kernel<EpiLstmFwd<f16>> (weights-stationary slabs, two pipelined M=64 halves per row group)'
        return 'tc::gemm_tc_chain_kernel<32, EpiLstmFwd<f16>, 1, weights-stationary> (persistent forward chain)'

    chains = []
    for c_, M_, N_, K_, ms_, n_, b_, l_ in sorted(shapes, key=lambda x: -x[4]):
        if c_ == 1 and l_ and n_ > l_:       # persistent chains only (one launch walks many steps)
            ach = b_ / (ms_ * 1e-3) / 1e9
            chains.append({'kernel': chain_kernel_name(M_, N_, K_), 'rows': M_, 'N': N_, 'K': K_, 'direction': 'backward' if K_ > N_ else 'forward',
                           'launches_per_step': l_ / args.steps, 'recurrent_steps_per_launch': n_ / l_, 'us_per_recurrent_step': 1e3 * ms_ / n_,
                           'ms_per_step': ms_ / args.steps, 'achieved': ach, 'unit': 'GB/s', 'frac': ach / hbm if hbm else None,
                           'traffic': STEP_KERNEL_DRAM_BYTES_PER_LAUNCH.get((M_, N_, K_))})
    roof = {'kernel': '%s rows=%d N=%d K=%d: per time step h.W_h on tcgen05 + fused BasicLSTMCell %s'
                      % (chain_kernel_name(dM, dN, dK), dM, dN, dK, 'backward' if dK > dN else 'forward'),
            'bound': 'hbm', 'achieved': d_ach, 'peak': hbm, 'unit': 'GB/s', 'frac': d_ach / hbm if hbm else None,
            'traffic': STEP_KERNEL_DRAM_BYTES_PER_LAUNCH.get((dM, dN, dK)),
            'traffic_source': 'ncu --set full dram__bytes_read+write per launch, cold caches (profiles/r2_chain_ncu_full.md)',
            'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback 6.65 TB/s (B200_PROFILING.md)',
            'launches_per_step': dln / args.steps, 'recurrent_steps_per_launch': dcnt / dln if dln else None,
            'us_per_launch': 1e3 * dms / dln if dln else None, 'ms_per_step': dms / args.steps,
            'algorithmic_bytes_per_launch': dby / dln if dln else None,
            'bytes_model': 'per recurrent step: bf16 W_h once + activation rows in + gate addends / saved gate activations / cell state / '
                           'outputs (DESIGN.md section 4), times the steps one launch walks; weights and state stay L2-resident, the '
                           'binding limits are the grid-barrier latency per step and the ~42 B/clk/SM L2->SM fill rate, not HBM',
            'pass': 'separate instrumented pass of the same %d steps: CUDA events around each chain launch' % args.steps}
    out = {'metric': METRIC, 'value': value, 'unit': 'videos/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
           'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
           'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic', 'config': workload_config(args, B, world),
           'clocks': clocks, 'e2e': {'value': e2e, 'unit': 'videos/s', 'h2d_bytes_per_step': int(feats_host.numel() * 4 + vidx_host.numel() * 4),
                                     'd2h_bytes_per_step': 8, 'ms_per_step': ms_e2e / args.steps,
                                     'mode': 'feed staged one step ahead on a copy stream (FeaturePipe), result read one step behind; every copy inside the timed region',
                                     'blocking_feed_value': world * B * args.steps / (ms_e2e_blk / 1e3),
                                     'fp16_host_cache_value': world * B * args.steps / (ms_e2e_h / 1e3), 'fp16_host_cache_h2d_bytes_per_step': int(feats_host.numel() * 2 + vidx_host.numel() * 4)},
           'gpu_launches': int(launches), 'loss': loss, 'decode': beam,
           'gemm_shapes': [{'class': 'batched' if c == 0 else 'step', 'M': M_, 'N': N_, 'K': K_, 'gemms_per_step': n_ / args.steps,
                            'launches_per_step': l_ / args.steps, 'us_per_gemm': 1e3 * ms_ / n_, 'ms_per_step': ms_ / args.steps}
                           for c, M_, N_, K_, ms_, n_, b_, l_ in sorted(shapes, key=lambda x: -x[4])],
           'roofline': roof, 'roofline_chains': chains, 'strong_scaling': strong, 'collective': collective,
           'roofline_recurrent_family': {'kernels': 'all recurrent-step kernels (persistent chains + the per-step launches of the sampling loops)',
                                         'bound': 'hbm', 'achieved': ach_gb, 'peak': hbm, 'unit': 'GB/s', 'frac': ach_gb / hbm if hbm else None,
                                         'recurrent_steps_per_iteration': sn / args.steps, 'ms_per_step': sms / args.steps,
                                         'us_per_recurrent_step': 1e3 * sms / sn if sn else None},
           'roofline_batched_gemm': {'kernel': 'tc::gemm_tc_kernel<256, EpiStore|EpiGradStore> (128x256 tcgen05 tiles: projections, vocab logits, weight gradients)',
                                     'bound': 'tensor', 'achieved': ach_tf, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': ach_tf / peak_tf if peak_tf else None,
                                     'peak_source': peak_src, 'launches_per_step': bn / args.steps, 'ms_per_step': bms / args.steps,
                                     'algorithmic_gflop_per_step': bfl / args.steps / 1e9,
                                     'note': 'measured INSIDE the step: side-stream kernels (persistent chains) share the GPU with several of these GEMMs; '
                                             'roofline_batched_gemm_isolated times every shape alone'}}
    if iso:
        ach_iso = iso['tflop_total'] / (iso['us_total'] * 1e-6)
        out['roofline_batched_gemm_isolated'] = {'bound': 'tensor', 'achieved': ach_iso, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': ach_iso / peak_tf if peak_tf else None,
                                                 'peak_source': peak_src, 'us_total': iso['us_total'], 'padded_tflop_total': iso['tflop_total'], 'shapes': iso['shapes'],
                                                 'how': 's2vt_debug_gemm: each padded shape of the iteration alone on the GPU, 10 repetitions, CUDA events'}
    if world == 1 and not args.no_cpu_baseline:
        use_all_host_threads()
        OracleIteration(args.samples, 5, 2, vocab, by, order, bias).step()        # page in BLAS / the corpus on a tiny case
        o = OracleIteration(args.samples, args.frames, args.ref_videos, vocab, by, order, bias)
        t0 = time.perf_counter(); cpu_loss = o.step(); dt = time.perf_counter() - t0
        out['cpu_baseline'] = {'value': args.ref_videos / dt, 'unit': 'videos/s', 'cores': cpu_threads(), 'kind': 'port', 'loss': cpu_loss,
                               'sample': '1 full iteration on %d videos x K=%d, T_v=%d (the same configuration), fp32 NumPy/OpenBLAS oracle, %.1f s' % (args.ref_videos, K, Tv, dt)}
    print(json.dumps(out), file=out_stream, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference' and a.workload == 'beam':
        run_reference_beam(a)
    elif a.impl == 'reference':
        run_reference(a)
    elif a.workload == 'beam':
        run_beam(a)
    else:
        run_b200(a)
