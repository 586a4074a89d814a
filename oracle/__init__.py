"""CPU oracle for the S2VT caption hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

This package is a NumPy restatement of the algorithms of
adwardlee/multitask-end-to-end-video-captioning (TF-1.1 / Python-2 scripts that
cannot run in this image: no python2, no tensorflow wheel, no pyciderevalcap).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product package
(``multitask-end-to-end-video-captioning_b200``) never imports it and has no CPU
fallback.

PARITY STATUS
* TF-1.1 numerics (BasicLSTMCell, softmax CE with label smoothing, Adam, global
  norm clip, tf.multinomial): **parity unpinned** -- the reference ships no
  golden vectors or tests and TensorFlow 1.1 is not installable here.  The
  restatement follows the reference call sites (file:line cited per function)
  and the TF r1.1 library semantics listed in SURVEY.md Q1-Q8 / R1-R9.
  Hand-derived gradients are cross-checked against torch.autograd (CPU) in
  tests/test_oracle_model.py.
* CIDEr-D: the scorer is the un-vendored third-party ``pyciderevalcap``
  (vrama91/cider, no version pinned; imported at cider_evaluation.py:9,12) and
  its ``data/msvd.p`` document-frequency pickle is missing.  The algorithm is
  restated from the published source; document frequencies are rebuilt from
  msvd_sents_train_noval_lc_nopunc.txt.  **Soft pin**: the reference's own
  artefact ``msvd_best_captions`` (output of choose_best_cider.py:126-143) is
  reproduced for >= 1100 of 1200 videos (tests/test_oracle_ciderd.py).
* Pure-Python helpers (sentence_padding_toix, decode_captions[_masks],
  Caption/TopN heap) are restated statement by statement; those are exact.
"""
