#!/usr/bin/env python
"""Stage 3 with the attribute head on precomputed features: -(1-alpha) REINFORCE + alpha * attribute sigmoid-CE (drop-in for the
LSTM / RL / attribute-head part of the reference script; the in-graph CNN is out of scope).

    python reinforce_multitask_e2e_attribute_loss.py --task {train,evaluate,test} [--gpu N] [--<constant> value ...]
"""
import s2vt_b200  # noqa: F401  (alias of the package directory multitask-end-to-end-video-captioning_b200)
from s2vt_b200 import cli

if __name__ == '__main__':
    parser = cli.build_parser(__doc__, dict(model_name='reinforce_multitask_model_alpha005', start_learning_rate=1e-6, decay_steps=15000, clip_norm=10.0,
                                            batch_size=16, n_video_lstm_step=5, alpha=0.05, n_epochs=20))
    cli.run_attribute_loss(parser.parse_args())
