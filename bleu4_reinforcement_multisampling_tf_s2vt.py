#!/usr/bin/env python
"""Stage 2 with the BLEU-4 reward instead of CIDEr-D (drop-in for the reference script of the same name: the same K-sample
REINFORCE loop with `from bleu_evaluation import *`).

    python bleu4_reinforcement_multisampling_tf_s2vt.py --task {train,evaluate,test} [--gpu N] [--<constant> value ...]
"""
import s2vt_b200  # noqa: F401  (alias of the package directory multitask-end-to-end-video-captioning_b200)
from s2vt_b200 import cli

if __name__ == '__main__':
    parser = cli.build_parser(__doc__, dict(model_name='bleu4_reinforce_multisample8_model', start_learning_rate=1e-6, decay_steps=1000, clip_norm=5.0, batch_size=256, n_samples=8,
                                            reward='bleu4'))
    cli.run_rl(parser.parse_args())
