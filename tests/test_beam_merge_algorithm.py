"""The algorithm of the fused beam step, restated in NumPy and checked against a plain top-k on CPU: per-part top-8 lists with
tf.nn.top_k's tie rule (EpiLogitsTopK, csrc/gemm.cuh), the soft-max statistics merge  lse = log sum_p s_p e^{m_p}  and the k rounds of
arg-max that each take the best candidate strictly AFTER the previous winner in (value descending, index ascending) order
(beam_step_kernel, csrc/beam.cuh).  The CUDA code itself is held to the un-fused launches and to the oracle by tests/test_gpu_cider_beam.py;
this file pins the reasoning (ties, padded vocabulary tail, parts with fewer than 8 words, k = 1 .. 8) where no GPU is needed."""
import numpy as np
import pytest

TOPK_MAX = 8
INT_MAX = 0x7fffffff


def part_epilogue(logits_row, V, Vp, PW):
    """One tile row of EpiLogitsTopK: for every part of PW columns its 8 best (value, index) pairs in insertion order and (max, sum exp)."""
    nparts = Vp // PW
    val = np.full((nparts, TOPK_MAX), -np.inf, np.float32)
    idx = np.full((nparts, TOPK_MAX), INT_MAX, np.int64)
    stat = np.zeros((nparts, 2), np.float32)
    for p in range(nparts):
        c0 = p * PW
        cols = [c for c in range(c0, c0 + PW) if c < V]
        mx = np.float32(-np.inf)
        for c in cols:
            mx = max(mx, logits_row[c])
        tv = [np.float32(-np.inf)] * TOPK_MAX
        ti = [INT_MAX] * TOPK_MAX
        se = np.float32(0)
        for c in cols:                                     # ascending scan
            e = logits_row[c]
            se = np.float32(se + np.exp(np.float32(e - mx)))
            if e > tv[-1]:                                 # strict: an equal value never displaces an earlier (lower) index
                tv[-1], ti[-1] = e, c
                for q in range(TOPK_MAX - 1, 0, -1):       # one bubble pass, strict compares
                    if tv[q] > tv[q - 1]:
                        tv[q], tv[q - 1] = tv[q - 1], tv[q]
                        ti[q], ti[q - 1] = ti[q - 1], ti[q]
        val[p], idx[p], stat[p] = tv, ti, (mx, se)
    return val, idx, stat


def merge(val, idx, stat, V, k):
    """beam_step_kernel phase 1 for one row: (top-k indices, their log-probs)."""
    mx = stat[:, 0].max()
    with np.errstate(invalid='ignore'):
        se = np.sum(np.where(stat[:, 1] > 0, stat[:, 1] * np.exp(stat[:, 0] - mx), 0.0), dtype=np.float64)
    lse = mx + np.log(se)
    cv, ci = val.ravel(), idx.ravel()
    prev_v, prev_i = np.inf, -1
    out_i, out_lp = [], []
    for _ in range(k):
        best_v, best_i = -np.inf, INT_MAX
        for v, i in zip(cv, ci):
            after = v < prev_v or (v == prev_v and i > prev_i)
            if after and i < V and (v > best_v or (v == best_v and i < best_i)):
                best_v, best_i = v, i
        out_i.append(int(best_i)); out_lp.append(float(best_v - lse))
        prev_v, prev_i = best_v, best_i
    return out_i, out_lp


def reference_topk(logits_row, V, k):
    """tf.nn.top_k(softmax(logits)): descending value, lower index first on ties; log-probs from a plain log-softmax."""
    x = logits_row[:V].astype(np.float64)
    order = sorted(range(V), key=lambda i: (-x[i], i))[:k]
    lse = x.max() + np.log(np.exp(x - x.max()).sum())
    return order, [float(x[i] - lse) for i in order]


@pytest.mark.parametrize('V,Vp,PW', [(9972, 9984, 64), (9972, 9984, 32), (60, 128, 32), (200, 256, 8), (5, 128, 32)])
@pytest.mark.parametrize('ties', [False, True])
def test_part_lists_and_ordered_merge_equal_plain_topk(V, Vp, PW, ties):
    rng = np.random.RandomState(V + PW + ties)
    for trial in range(3):
        row = rng.normal(0, 3, Vp).astype(np.float32)
        if ties:                                           # few distinct values: every rank has many equal candidates across and inside parts
            row = rng.choice(np.asarray([-1.5, 0.0, 0.25, 2.0], np.float32), Vp)
        row[V:] = 1e9                                      # padded columns hold garbage: they must never be picked
        val, idx, stat = part_epilogue(row, V, Vp, PW)
        for k in (1, 3, 5, 8):
            if k > V:
                continue
            got_i, got_lp = merge(val, idx, stat, V, k)
            ref_i, ref_lp = reference_topk(row, V, k)
            assert got_i == ref_i, (V, PW, ties, k, got_i, ref_i)
            np.testing.assert_allclose(got_lp, ref_lp, atol=2e-5)


def test_every_valid_word_lives_in_exactly_one_part_and_empty_parts_are_neutral():
    V, Vp, PW = 70, 128, 32
    row = np.linspace(-1, 1, Vp).astype(np.float32)
    val, idx, stat = part_epilogue(row, V, Vp, PW)
    assert np.isneginf(stat[3, 0]) and stat[3, 1] == 0 and (idx[3] == INT_MAX).all()          # part 3 = columns 96..127: all padding
    assert (idx[2][:6] == np.arange(69, 63, -1)).all() and (idx[2][6:] == INT_MAX).all()        # part 2 holds only words 64..69
    seen = idx[idx != INT_MAX]
    assert len(set(seen.tolist())) == len(seen) and seen.max() < V
