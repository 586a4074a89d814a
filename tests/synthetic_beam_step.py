"""A deterministic stand-in for `beam_probability` (final_beam_search.py:202-223) used to pin the beam HOST LOOP: a pure function of
(language state, previous word) that returns the top-k words, their probabilities and the next states.  No model is involved, so the
reference's own loop (executed by scripts/make_reference_fixtures.py) and the restatement in oracle/beam.py can be driven by exactly
the same step function.  The eos probability grows with the depth of the hypothesis so that searches finish at different lengths,
finals arrive one by one (the `exclude_num` paths B2/B3) and some searches run out of steps without any final."""
import numpy as np

VOCAB = 40


def make_step(case_seed, beam_size, eos_ramp):
    def step(state1, state2, word):
        depth = int(state1[0, 0])
        key = (int(state2[0, 0]) * 1000003 + int(word[0]) * 7919 + case_seed * 104729 + depth) % (2 ** 31 - 1)
        rng = np.random.RandomState(key)
        logits = rng.normal(0.0, 1.5, VOCAB)
        logits[0] += eos_ramp * depth - 3.0                      # <eos> = 0 becomes likelier with depth
        logits[1] = -30.0                                        # <bos> is never produced
        p = np.exp(logits - logits.max())
        p /= p.sum()
        idx = np.argsort(-p, kind='stable')[:beam_size]
        new2 = np.array([[float(key % 99991)]])
        new1 = np.array([[float(depth + 1)]])
        return idx.astype(np.int32), p[idx].astype(np.float32), new2, new1
    return step


def initial_states():
    return np.zeros((1, 1)), np.array([[17.0]])


CASES = [(seed, k, lnf, ramp, tc) for seed in range(6) for k in (3, 4, 5) for lnf in (0.0, 1.0) for ramp, tc in ((0.6, 35), (0.15, 12))]
