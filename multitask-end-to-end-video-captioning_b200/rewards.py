"""GPU BLEU-4 and ROUGE-L rewards behind the `evaluate_captions_cider(ref, cand)` (sic) of the reference's
bleu_evaluation.py:60-87 and rouge_evaluation.py:60-87 -- the reward helpers of
bleu4_reinforcement_multisampling_tf_s2vt.py and rouge_reinforcement_multisampling_tf_s2vt.py.  Like cider.CiderD, the
per-video reference tables are built once (csrc/rewards.cu) and scoring is one kernel launch on token ids; both classes
offer `score_ids` (what trainer.ReinforceTrainer calls), `score_strings` and the literal drop-in call."""
import ctypes as C

import numpy as np
import torch

from . import _lib

MAX_TOKENS = 64


class _ReferenceScorer(object):
    SPLIT = None           # None: str.split() (Bleu's precook); ' ': str.split(' ') (Rouge.calc_score)

    def __init__(self, ref_sets, wordtoix, device=None):
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
        self.wordtoix = dict(wordtoix)
        self._next_id = max(self.wordtoix.values()) + 1
        self.empty_token = self._next_id                      # the '' word that split(' ') can produce
        self.wordtoix[''] = self.empty_token
        self._next_id += 1
        self.n_videos = len(ref_sets)
        toks, ref_off, vid_off = [], [0], [0]
        for refs in ref_sets:
            for s in refs:
                toks.extend(self._ids(s, grow=True))
                ref_off.append(len(toks))
            vid_off.append(len(ref_off) - 1)
        if self._next_id > 65000:
            raise ValueError('n-gram keys hold 16-bit token ids; vocabulary + OOV words = %d' % self._next_id)
        toks = np.asarray(toks, dtype=np.int32)
        ref_off = np.asarray(ref_off, dtype=np.int64)
        vid_off = np.asarray(vid_off, dtype=np.int64)
        corpus = C.c_void_p()
        rc = self.lib.s2vt_reward_corpus_create(toks.ctypes.data_as(C.c_void_p), ref_off.ctypes.data_as(C.c_void_p), len(ref_off) - 1,
                                                vid_off.ctypes.data_as(C.c_void_p), self.n_videos, C.byref(corpus))
        if rc != 0:
            raise _lib.S2vtError(rc, 's2vt_reward_corpus_create failed')
        try:
            nbytes = self.lib.s2vt_reward_corpus_device_bytes(corpus)
            host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
            rc = self.lib.s2vt_reward_corpus_serialize(corpus, C.c_void_p(host.data_ptr()))
            if rc != 0:
                raise _lib.S2vtError(rc, 's2vt_reward_corpus_serialize failed')
            self.table = host.to(self.device)
        finally:
            self.lib.s2vt_reward_corpus_destroy(corpus)
        self.table_bytes = nbytes
        self._key_of_refs = {}
        for i, refs in enumerate(ref_sets):
            self._key_of_refs.setdefault(tuple(refs), i)

    def _ids(self, sentence, grow=False):
        out, private = [], {}
        for w in (sentence.split() if self.SPLIT is None else sentence.split(self.SPLIT)):
            i = self.wordtoix.get(w)
            if i is None:
                if grow:
                    i = self._next_id
                    self.wordtoix[w] = i
                    self._next_id += 1
                else:                    # hypothesis-only word: a private id that never equals a reference token
                    i = private.setdefault(w, 65534 - len(private))
            out.append(i)
        return out

    def _launch(self, hyp, vid, N, Tc, st):
        raise NotImplementedError

    def score_ids(self, hyp_ids, video_of_row):
        """hyp_ids int32 [N, T_c] device tensor (words before the first 0 count), video_of_row int32 [N] -> float64 [N]."""
        hyp = hyp_ids.to(self.device, torch.int32).contiguous()
        vid = torch.as_tensor(video_of_row).to(self.device, torch.int32).contiguous()
        return self._launch(hyp, vid, hyp.shape[0], hyp.shape[1], C.c_void_p(torch.cuda.current_stream().cuda_stream))

    def score_strings(self, cands, video_of_row):
        rows = [self._ids(c) for c in cands]
        rows = [[] if r == [self.empty_token] else r for r in rows]          # '' -> no words; the kernel re-creates the empty token
        L = max([len(r) for r in rows] + [1]) + 1
        if L > MAX_TOKENS:
            raise ValueError('hypothesis longer than %d tokens' % (MAX_TOKENS - 1))
        ids = np.zeros((len(rows), L), dtype=np.int32)
        for i, r in enumerate(rows):
            ids[i, :len(r)] = r
        return self.score_ids(torch.from_numpy(ids), video_of_row)

    def evaluate_captions_cider(self, ref, cand):
        """Literal drop-in: ref {i: [reference sentences]}, cand [str] -> float64 ndarray [N]; every ref[i] must be the
        reference list of one of the corpus videos (it is in the reference's train loop)."""
        vids = []
        for i in range(len(cand)):
            key = tuple(ref[i])
            if key not in self._key_of_refs:
                raise KeyError('reference set %d is not one of the corpus videos' % i)
            vids.append(self._key_of_refs[key])
        return self.score_strings(cand, np.asarray(vids, dtype=np.int32)).cpu().numpy()


class Bleu4(_ReferenceScorer):
    """Bleu(4).compute_score(refe, hypo)[1][3]: per-sentence BLEU-4 ('closest' reference length)."""
    SPLIT = None

    def score_all_orders(self, hyp_ids, video_of_row, want_comps=False):
        hyp = hyp_ids.to(self.device, torch.int32).contiguous()
        vid = torch.as_tensor(video_of_row).to(self.device, torch.int32).contiguous()
        N, Tc = hyp.shape
        out = torch.empty(N, 4, dtype=torch.float64, device=self.device)
        comps = torch.empty(N, 10, dtype=torch.int32, device=self.device) if want_comps else None
        rc = self.lib.s2vt_bleu_score(C.c_void_p(self.table.data_ptr()), C.c_void_p(hyp.data_ptr()), C.c_void_p(vid.data_ptr()), N, Tc,
                                      C.c_void_p(out.data_ptr()), C.c_void_p(comps.data_ptr()) if want_comps else None,
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise _lib.S2vtError(rc, 's2vt_bleu_score failed')
        return (out, comps) if want_comps else out

    def corpus_bleu(self, cands, video_of_row):
        """Bleu(4).compute_score(ref, hypo)[0]: corpus-level [Bleu_1 .. Bleu_4] (totals of the per-sentence components)."""
        rows = [self._ids(c) for c in cands]
        L = max([len(r) for r in rows] + [1]) + 1
        ids = np.zeros((len(rows), L), dtype=np.int32)
        for i, r in enumerate(rows):
            ids[i, :len(r)] = r
        _, comps = self.score_all_orders(torch.from_numpy(ids), video_of_row, want_comps=True)
        tot = comps.to(torch.int64).sum(0).cpu().numpy()
        tiny, small = 1e-15, 1e-9
        bleus, bleu = [], 1.0
        for k in range(4):
            bleu *= float(tot[k] + tiny) / (tot[4 + k] + small)
            bleus.append(bleu ** (1.0 / (k + 1)))
        ratio = (tot[8] + tiny) / (tot[9] + small)
        if ratio < 1:
            bleus = [b * float(np.exp(1 - 1 / ratio)) for b in bleus]
        return bleus

    def _launch(self, hyp, vid, N, Tc, st):
        return self.score_all_orders(hyp, vid)[:, 3].contiguous()


class RougeL(_ReferenceScorer):
    """Rouge().compute_score(refe, hypo)[1]: per-sentence ROUGE-L, beta = 1.2."""
    SPLIT = ' '

    def _launch(self, hyp, vid, N, Tc, st):
        out = torch.empty(N, dtype=torch.float64, device=self.device)
        rc = self.lib.s2vt_rouge_score(C.c_void_p(self.table.data_ptr()), C.c_void_p(hyp.data_ptr()), C.c_void_p(vid.data_ptr()), N, Tc,
                                       int(self.empty_token), C.c_void_p(out.data_ptr()), st)
        if rc != 0:
            raise _lib.S2vtError(rc, 's2vt_rouge_score failed')
        return out


def make_scorer(kind, ref_sets, wordtoix, device=None):
    """'cider' | 'bleu4' | 'rouge' -> reward scorer with score_ids(hyp_ids, video_of_row)."""
    from . import cider
    kind = kind.lower()
    if kind in ('cider', 'ciderd', 'cider-d'):
        return cider.CiderD(ref_sets, wordtoix, device=device)
    if kind in ('bleu', 'bleu4', 'bleu-4'):
        return Bleu4(ref_sets, wordtoix, device=device)
    if kind in ('rouge', 'rouge_l', 'rouge-l', 'rougel'):
        return RougeL(ref_sets, wordtoix, device=device)
    raise ValueError('unknown reward %r' % kind)


def evaluate_for_particular_captions(cand, ref_captions, wordtoix, device=None):
    """cider_evaluation.evaluate_for_particular_captions / score_all (cider_evaluation.py:14-58): cand {key: [caption]},
    ref_captions {key: [references]} -> {'Bleu_1'..'Bleu_4' (corpus level), 'ROUGE_L' (mean), 'CIDEr'}.  coco-caption's Cider
    is the clipped, length-penalised CIDEr-D with document frequencies taken from the scored reference sets, i.e. cider.CiderD
    built on these references.  METEOR needs the external Java scorer and is not computed."""
    from . import cider
    keys = [k for k in cand]
    ref_sets = [list(ref_captions[k]) for k in keys]
    hyps = [cand[k][0] if isinstance(cand[k], (list, tuple)) else cand[k] for k in keys]
    rows = np.arange(len(keys), dtype=np.int32)
    out = {}
    b = Bleu4(ref_sets, wordtoix, device=device)
    for m, s in zip(('Bleu_1', 'Bleu_2', 'Bleu_3', 'Bleu_4'), b.corpus_bleu(hyps, rows)):
        out[m] = float(s)
    out['ROUGE_L'] = float(RougeL(ref_sets, wordtoix, device=device).score_strings(hyps, rows).mean().item())
    out['CIDEr'] = float(cider.CiderD(ref_sets, wordtoix, device=device).score_strings(hyps, rows).mean().item())
    return out
