"""GPU CIDEr-D reward behind the reference's `evaluate_captions_cider(ref, cand)` call (cider_evaluation.py:60-87).

The reference wraps the third-party `CiderD(df='msvd')` scorer and runs it in host Python between the rollout and
the update; here the per-video reference tables live in HBM (built once by the C++ host code of
csrc/ciderd.cu) and scoring is one kernel launch on token ids -- no strings on the hot path.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

MAX_TOKENS = 64


class CiderD(object):
    """`CiderD(df=<corpus>)`: document frequencies from `df_ref_sets`, scoring against `ref_sets`.

    ref_sets    : list (one entry per video) of lists of reference sentences to score against
    df_mask     : optional bool per video -- which videos form the document-frequency corpus (default: all)
    wordtoix    : vocabulary of the captioning model; reference words outside it get private ids >= len(wordtoix)
    """

    def __init__(self, ref_sets, wordtoix, df_mask=None, device=None):
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
        self.wordtoix = dict(wordtoix)
        self.n_model_words = len(wordtoix)
        self._next_id = max(self.wordtoix.values()) + 1
        self.n_videos = len(ref_sets)
        toks, ref_off, vid_off = [], [0], [0]
        for refs in ref_sets:
            for s in refs:
                ids = self._ids(s, grow=True)
                toks.extend(ids)
                ref_off.append(len(toks))
            vid_off.append(len(ref_off) - 1)
        if self._next_id > 65534:
            raise ValueError('CIDEr-D n-gram keys hold 16-bit token ids; vocabulary + OOV words = %d' % self._next_id)
        toks = np.asarray(toks, dtype=np.int32)
        ref_off = np.asarray(ref_off, dtype=np.int64)
        vid_off = np.asarray(vid_off, dtype=np.int64)
        dfm = None if df_mask is None else np.ascontiguousarray(np.asarray(df_mask, dtype=np.uint8))
        corpus = C.c_void_p()
        rc = self.lib.ciderd_corpus_create(toks.ctypes.data_as(C.c_void_p), ref_off.ctypes.data_as(C.c_void_p), len(ref_off) - 1,
                                           vid_off.ctypes.data_as(C.c_void_p), self.n_videos,
                                           dfm.ctypes.data_as(C.c_void_p) if dfm is not None else None, C.byref(corpus))
        if rc != 0:
            raise _lib.S2vtError(rc, 'ciderd_corpus_create failed')
        try:
            nbytes = self.lib.ciderd_corpus_device_bytes(corpus)
            host = torch.empty(nbytes, dtype=torch.uint8).pin_memory() if torch.cuda.is_available() else torch.empty(nbytes, dtype=torch.uint8)
            rc = self.lib.ciderd_corpus_serialize(corpus, C.c_void_p(host.data_ptr()))
            if rc != 0:
                raise _lib.S2vtError(rc, 'ciderd_corpus_serialize failed')
            self.table = host.to(self.device)
        finally:
            self.lib.ciderd_corpus_destroy(corpus)
        self.table_bytes = nbytes
        self._key_of_refs = {}
        for i, refs in enumerate(ref_sets):
            self._key_of_refs.setdefault(tuple(refs), i)

    def _ids(self, sentence, grow=False):
        out, private = [], {}
        for w in sentence.split():
            i = self.wordtoix.get(w)
            if i is None:
                if grow:
                    i = self._next_id
                    self.wordtoix[w] = i
                    self._next_id += 1
                else:                    # hypothesis-only word: a private id that never equals a reference token
                    i = private.setdefault(w, 65534 - len(private))
                    if i < self._next_id:
                        raise ValueError('too many distinct out-of-vocabulary words')
            out.append(i)
        return out

    def score_ids(self, hyp_ids, video_of_row, want_counts=False):
        """hyp_ids int32 [N, T_c] device tensor (tokens before the first 0 count), video_of_row int32 [N].
        Returns float64 [N] device tensor (and the exact n-gram (key, count) tables when want_counts)."""
        hyp = hyp_ids.to(self.device, torch.int32).contiguous()
        vid = torch.as_tensor(video_of_row).to(self.device, torch.int32).contiguous()
        N, Tc = hyp.shape
        scores = torch.empty(N, dtype=torch.float64, device=self.device)
        counts = torch.zeros(N, 4 * Tc, 2, dtype=torch.int64, device=self.device) if want_counts else None
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        rc = self.lib.ciderd_score(C.c_void_p(self.table.data_ptr()), C.c_void_p(hyp.data_ptr()), C.c_void_p(vid.data_ptr()), N, Tc,
                                   C.c_void_p(scores.data_ptr()), C.c_void_p(counts.data_ptr()) if want_counts else None, st)
        if rc != 0:
            raise _lib.S2vtError(rc, 'ciderd_score failed')
        return (scores, counts) if want_counts else scores

    def score_strings(self, cands, video_of_row):
        rows = [self._ids(c) for c in cands]
        L = max([len(r) for r in rows] + [1]) + 1
        if L > MAX_TOKENS:
            raise ValueError('hypothesis longer than %d tokens' % (MAX_TOKENS - 1))
        ids = np.zeros((len(rows), L), dtype=np.int32)
        for i, r in enumerate(rows):
            ids[i, :len(r)] = r
        return self.score_ids(torch.from_numpy(ids), video_of_row)

    def evaluate_captions_cider(self, ref, cand):
        """Drop-in for cider_evaluation.evaluate_captions_cider(ref: {i: [refs]}, cand: [str]) -> float64 ndarray [N].
        Every ref[i] must be the reference list of one of the corpus videos (it is, in the reference's train loop,
        reinforcement_multisampling_tf_s2vt.py:788-806)."""
        vids = []
        for i in range(len(cand)):
            key = tuple(ref[i])
            if key not in self._key_of_refs:
                raise KeyError('reference set %d is not one of the corpus videos' % i)
            vids.append(self._key_of_refs[key])
        return self.score_strings(cand, np.asarray(vids, dtype=np.int32)).cpu().numpy()


def decode_ngram_key(key):
    """uint64 key -> tuple of token ids (inverse of the exact packing in csrc/ciderd.cu)."""
    out = []
    key = int(key) & 0xFFFFFFFFFFFFFFFF
    while key:
        out.append((key & 0xFFFF) - 1)
        key >>= 16
    return tuple(out)
