#!/usr/bin/env python
"""Pin the oracle with EXECUTED reference code (run in the build container, where /root/reference exists).

The reference scripts are Python-2 / TensorFlow-1.1 programs and cannot run as a whole, but the pure-Python pieces of the hot path
can: this script loads their SOURCE TEXT from /root/reference at run time (nothing is copied into the repo), executes it under
Python 3 -- `import tensorflow` stubbed, `xrange` / `iteritems` / `print` statements mapped -- and writes the OUTPUTS as fixtures:

  beam_search.py:6-80                      Caption / TopN (heap order incl. ties, extract(sort=True), reset)
  final_beam_search.py:248-294             the beam host loop of build_generator (B1-B7), driven by a synthetic step function
  e2e_beam_search.py:301-344               the same loop in the end-to-end script (must agree with the one above)
  cider_evaluation.py:122-172              decode_captions, decode_captions_masks
  tf_s2vt.py:345-401                       preProBuildWordVocab, sentence_padding_toix
  reinforcement_multisampling_tf_s2vt.py:600-601      get_captions
  reinforce_multitask_e2e_attribute_loss.py:874-893   get_multilabel

tests/test_oracle_pinned_by_reference.py holds oracle/ (and the product's host mirror) to these outputs.
"""
import gzip
import json
import math
import os
import re
import sys
import textwrap
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
REF = '/root/reference'
OUT = os.path.join(ROOT, 'tests', 'golden', 'reference_exec_golden.json.gz')


def source(name):
    with open(os.path.join(REF, name)) as f:
        return f.read().replace('\t', '        ')


def py3(src):
    src = src.replace('xrange(', 'range(').replace('.iteritems()', '.items()')
    return re.sub(r'^(\s*)print (.*)$', r'\1print(\2)', src, flags=re.M)


def top_level_def(src, name):
    """Text of the top-level `def name(` up to the next top-level statement."""
    lines = src.split('\n')
    start = next(i for i, l in enumerate(lines) if l.startswith('def %s(' % name))
    end = next((i for i in range(start + 1, len(lines)) if lines[i] and not lines[i][0].isspace() and not lines[i].startswith('#')), len(lines))
    return py3('\n'.join(lines[start:end]))


def load_beam_module():
    """beam_search.py as a module, with the unused `import tensorflow` satisfied by an empty stub."""
    sys.modules.setdefault('tensorflow', types.ModuleType('tensorflow'))
    mod = types.ModuleType('ref_beam_search')
    exec(compile(py3(source('beam_search.py')), os.path.join(REF, 'beam_search.py'), 'exec'), mod.__dict__)
    return mod


def host_loop(script, first_line, last_line, beam_mod):
    """The beam host loop inside Video_Caption_Generator.build_generator, lines [first_line, last_line] of `script` (1-based), turned
    into a function of (self, sess, state1, state2, length_normalization_factor): the TF graph construction above it is not executed;
    `sess.run` is served by FakeSession below."""
    lines = source(script).split('\n')[first_line - 1:last_line]
    assert 'captions = TopN(beam_size*beam_size)' in lines[0] and 'return tf_sentence' in lines[-1], (lines[0], lines[-1])
    body = textwrap.dedent('\n'.join(lines))
    # preamble = what the lines above the loop leave in scope: beam_size (:232), and in e2e_beam_search.py self.sess and the encoder
    # states already fetched (:296); final_beam_search.py fetches them itself (:251-252)
    pre = ('def host_loop(self, sess, state1, state2, length_normalization_factor):\n    beam_size = self.beam_size\n    self.sess = sess\n'
           '    initial_state1 = sess.run(state1)\n    initial_state2 = sess.run(state2)\n')
    fn = pre + textwrap.indent(py3(body), '    ')
    ns = {'TopN': beam_mod.TopN, 'Caption': beam_mod.Caption, 'math': math, 'np': np}
    exec(compile(fn, os.path.join(REF, script), 'exec'), ns)
    return ns['host_loop']


class Token(object):
    def __init__(self, name):
        self.name = name

    def __hash__(self):
        return hash(self.name)

    def __eq__(self, o):
        return isinstance(o, Token) and o.name == self.name


class FakeModel(object):
    def __init__(self, beam_size, Tc):
        self.beam_size, self.n_caption_lstm_step = beam_size, Tc

    def beam_probability(self):
        return [Token(n) for n in ('word_index', 'probs', 'state2', 'state1', 'state2_feed', 'state1_feed', 'input_feed')]


class FakeSession(object):
    """sess.run(state) -> the encoder state; sess.run([word_index, probs, state2, state1], feed_dict) -> one beam_probability call."""

    def __init__(self, step, s1, s2):
        self.step, self.init = step, {'init1': s1, 'init2': s2}
        self.calls = 0

    def run(self, fetch, feed_dict=None):
        if feed_dict is None:
            return self.init[fetch.name]
        feed = {k.name: v for k, v in feed_dict.items()}
        self.calls += 1
        return self.step(feed['state1_feed'], feed['state2_feed'], feed['input_feed'])


def topn_trace(beam_mod, seed):
    """Push a stream of scored items (ties included) through TopN / Caption; record every extract."""
    rng = np.random.RandomState(seed)
    out = []
    for n in (1, 3, 5, 9):
        t = beam_mod.TopN(n)
        scores = np.round(rng.normal(0, 1, 40), 1).tolist()          # rounded: many exact ties
        for i, s in enumerate(scores):
            t.push(beam_mod.Caption([i], None, None, s, s))
        size = t.size()
        got = t.extract(sort=True)
        t.reset()
        out.append({'n': n, 'scores': scores, 'size': size, 'sorted_scores': [c.score for c in got],
                    'sorted_ids_by_score': sorted(((c.score, c.sentence[0]) for c in got), reverse=True), 'size_after_reset': t.size()})
    return out


def main():
    import synthetic_beam_step as S
    beam_mod = load_beam_module()
    g = {'source_files': ['beam_search.py', 'final_beam_search.py', 'e2e_beam_search.py', 'cider_evaluation.py', 'tf_s2vt.py',
                          'reinforcement_multisampling_tf_s2vt.py', 'reinforce_multitask_e2e_attribute_loss.py']}
    g['topn'] = [topn_trace(beam_mod, s) for s in range(3)]
    c = beam_mod.Caption([1], None, None, -1.0, -2.0)
    d = beam_mod.Caption([2], None, None, -5.0, -2.0)
    g['caption_cmp'] = {'lt': c < d, 'eq': c == d, 'lt_lower': beam_mod.Caption([3], None, None, 0, -3.0) < c}

    loop_final = host_loop('final_beam_search.py', 248, 294, beam_mod)
    loop_e2e = host_loop('e2e_beam_search.py', 301, 344, beam_mod)
    cases = []
    for seed, k, lnf, ramp, tc in S.CASES:
        s1, s2 = S.initial_states()
        res = []
        for loop in (loop_final, loop_e2e):
            sess = FakeSession(S.make_step(seed, k, ramp), s1, s2)
            sent, lp, sc = loop(FakeModel(k, tc), sess, Token('init1'), Token('init2'), lnf)
            res.append(([int(w) for w in sent], float(lp), float(sc), sess.calls))
        assert res[0] == res[1], 'final_beam_search.py and e2e_beam_search.py loops disagree'
        cases.append({'seed': seed, 'beam_size': k, 'lnf': lnf, 'eos_ramp': ramp, 'Tc': tc, 'sentence': res[0][0], 'logprob': res[0][1],
                      'score': res[0][2], 'step_calls': res[0][3]})
    g['beam_loop'] = cases
    finished = sum(1 for x in cases if x['sentence'][-1] == 0)
    print('beam host loop: %d cases, %d finished with <eos>, %d ran out of steps' % (len(cases), finished, len(cases) - finished))

    # ---- text glue ----------------------------------------------------------------------------------------------
    ce = {'np': np}
    exec(top_level_def(source('cider_evaluation.py'), 'decode_captions'), ce)
    exec(top_level_def(source('cider_evaluation.py'), 'decode_captions_masks'), ce)
    tfs = {'np': np, 'n_caption_lstm_step': 35}
    exec(top_level_def(source('tf_s2vt.py'), 'preProBuildWordVocab'), tfs)
    exec(top_level_def(source('tf_s2vt.py'), 'sentence_padding_toix'), tfs)
    rl = {}
    exec(top_level_def(source('reinforcement_multisampling_tf_s2vt.py'), 'get_captions'), rl)
    al = {'np': np, 'defaultdict': __import__('collections').defaultdict}
    exec(top_level_def(source('reinforce_multitask_e2e_attribute_loss.py'), 'get_multilabel'), al)

    with open(os.path.join(REF, 'msvd_vocabulary1.txt')) as f:
        vocab = [l.strip() for l in f]           # tf_s2vt.py:411-415 reads the file this way
    w2i, i2w = tfs['preProBuildWordVocab'](vocab, word_count_threshold=0)
    g['vocab'] = {'n_words': len(w2i), 'probe': {w: w2i[w] for w in ('<eos>', '<bos>', '<en_unk>', 'a', 'man', 'zucchini') if w in w2i},
                  'ixtoword_probe': {str(i): i2w[i] for i in (0, 1, 2, 3, 100, 9971)}}
    sents = []
    with open(os.path.join(REF, 'msvd_sents_train_noval_lc_nopunc.txt')) as f:
        for line in f:
            vid, s = line.strip().split('\t')[:2]
            sents.append((vid, s))
    batch = [s for _, s in sents[:200]] + [s for _, s in sents[20000:20100]] + \
        [' '.join(['a'] * 34), ' '.join(['man'] * 35), ' '.join(['is'] * 50), 'a  double space', 'Unknownword here', '', 'A MAN Is Running']
    ids, mask = tfs['sentence_padding_toix'](list(batch), w2i)
    g['sentence_padding_toix'] = {'captions': batch, 'ids': [[int(x) for x in r] for r in ids], 'mask': np.asarray(mask).astype(int).tolist()}

    rng = np.random.RandomState(0)
    caps = rng.randint(0, 60, size=(64, 35))
    caps[::7, 0] = 0
    caps[1::7] = np.where(caps[1::7] == 0, 5, caps[1::7])      # rows without any <eos>
    masks, dec = ce['decode_captions_masks'](caps, i2w)
    g['decode'] = {'captions': caps.tolist(), 'masks': masks, 'decoded': dec, 'decoded_plain': ce['decode_captions'](caps, i2w),
                   'decoded_1d': ce['decode_captions'](caps[3], i2w), 'masks_1d': ce['decode_captions_masks'](caps[5], i2w)[0]}
    g['get_captions'] = {v: rl['get_captions'](sents, v) for v in ('vid1', 'vid7', 'vid1200')}
    with open(os.path.join(REF, 'train_most_freq_vocab_400_truncated.txt')) as f:
        attr = [l.strip() for l in f]
    by = {}
    for vid, s in sents:
        by.setdefault(vid, []).append(s)
    sub = {v: by[v] for v in ('vid1', 'vid2', 'vid3', 'vid600', 'vid1200')}
    lab = al['get_multilabel'](sub, attr)
    g['get_multilabel'] = {'n_attributes': len(attr), 'videos': list(sub), 'labels': {v: [int(x) for x in lab[v]] for v in sub}}

    with gzip.GzipFile(OUT, 'wb', mtime=0) as f:
        f.write(json.dumps(g, sort_keys=True).encode())
    print('wrote', OUT, os.path.getsize(OUT))


if __name__ == '__main__':
    main()
