#!/usr/bin/env python
"""Stage 2, single sample: REINFORCE with CIDEr-D reward and greedy baseline, one sampled caption per video (drop-in for the reference
script of the same name: its train() draws ONE multinomial caption per row, reinforcement_tf_s2vt.py:743-753, batch 16, clip 5).

    python reinforcement_tf_s2vt.py --task {train,evaluate,test} [--gpu N] [--<constant> value ...]
"""
import s2vt_b200  # noqa: F401  (alias of the package directory multitask-end-to-end-video-captioning_b200)
from s2vt_b200 import cli

if __name__ == '__main__':
    parser = cli.build_parser(__doc__, dict(model_name='reinforce_model', start_learning_rate=1e-6, decay_steps=1000, clip_norm=5.0, batch_size=16, n_samples=1,
                                            n_epochs=20))
    cli.run_rl(parser.parse_args())
