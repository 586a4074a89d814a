"""Oracle beam search.  TEST INFRASTRUCTURE.

Restates beam_search.py:6-80 (Caption / TopN) and the host beam loop of
final_beam_search.py:248-294 (identical in e2e_beam_search.py:301-344) on top of a
`step_fn(state1, state2, word) -> (idx[k], p[k], state2', state1')` callable that
plays the role of the `beam_probability` graph (final_beam_search.py:202-223).
Semantics B1-B7 of SURVEY.md section 3.4 are kept, including the global, monotone
`exclude_num` and the `== beam_size` early-exit test (B2).
"""
import heapq
import math


class Caption(object):
    """beam_search.py:6-42: ordered by score only."""

    def __init__(self, sentence, img_state, language_state, logprob, score, metadata=None):
        self.sentence = sentence
        self.language_state = language_state
        self.img_state = img_state
        self.logprob = logprob
        self.score = score
        self.metadata = metadata

    def __lt__(self, other):
        return self.score < other.score

    def __eq__(self, other):
        return self.score == other.score


class TopN(object):
    """beam_search.py:44-80: bounded min-heap keeping the n largest."""

    def __init__(self, n):
        self._n = n
        self._data = []

    def size(self):
        return len(self._data)

    def push(self, x):
        if len(self._data) < self._n:
            heapq.heappush(self._data, x)
        else:
            heapq.heappushpop(self._data, x)

    def extract(self, sort=False):
        data = self._data
        self._data = None
        if sort:
            data.sort(reverse=True)
        return data

    def reset(self):
        self._data = []


def beam_search(step_fn, initial_state1, initial_state2, beam_size, n_caption_lstm_step=35,
                length_normalization_factor=0.0, trace=None):
    """final_beam_search.py:248-294.  Returns (sentence ids, logprob, score).

    `trace`, if a list, receives per step the tuple (step, [(sentence, logprob)] of the k
    surviving parents, exclude_num) -- used by the parity test to locate divergences.
    """
    captions = TopN(beam_size * beam_size)
    final_captions = TopN(beam_size)
    initial_word = [1]
    word_index, probs, state2, state1 = step_fn(initial_state1, initial_state2, initial_word)
    for beam in range(beam_size):
        captions.push(Caption(sentence=[int(word_index[beam])], img_state=state1, language_state=state2,
                              logprob=math.log(probs[beam]), score=math.log(probs[beam])))
    exclude_num = 0
    for i in range(1, n_caption_lstm_step):
        mid_captions = captions.extract(sort=True)[:beam_size]
        captions.reset()
        if trace is not None:
            trace.append((i, [(list(c.sentence), c.logprob) for c in mid_captions], exclude_num))
        for mid_caption in mid_captions:
            word_index, probs, state2, state1 = step_fn(mid_caption.img_state, mid_caption.language_state,
                                                        [mid_caption.sentence[-1]])
            for beam in range(beam_size - exclude_num):
                sentence = mid_caption.sentence + [int(word_index[beam])]
                logprob = mid_caption.logprob + math.log(probs[beam])
                score = logprob
                if word_index[beam] == 0:
                    if length_normalization_factor > 0:
                        score /= len(sentence) ** length_normalization_factor
                    final_captions.push(Caption(sentence, state1, state2, logprob, score))
                    exclude_num += 1
                else:
                    captions.push(Caption(sentence, state1, state2, logprob, score))
        if exclude_num == beam_size:
            break
    if not final_captions.size():
        final_captions = captions
    final_cap = final_captions.extract(sort=True)[0]
    return final_cap.sentence, final_cap.logprob, final_cap.score
