#!/usr/bin/env python
"""Stage 3 on precomputed features: -(1-lambda) REINFORCE + lambda XE (drop-in for the LSTM/RL/XE part of the reference script; the in-graph CNN is out of scope).

    python reinforce_multitask_e2e_attribute_s2vt.py --task {train,evaluate,test} [--gpu N] [--<constant> value ...]
"""
import s2vt_b200  # noqa: F401  (alias of the package directory multitask-end-to-end-video-captioning_b200)
from s2vt_b200 import cli

if __name__ == '__main__':
    parser = cli.build_parser(__doc__, dict(model_name='reinforce_multitask_model', start_learning_rate=1e-6, decay_steps=300000, clip_norm=5.0, batch_size=2, n_video_lstm_step=10, alpha=0.5, n_epochs=40))
    cli.run_stage3(parser.parse_args())
