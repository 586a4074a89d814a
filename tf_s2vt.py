#!/usr/bin/env python
"""Stage 1: cross-entropy S2VT on precomputed frame features (drop-in for the reference's tf_s2vt.py).

    python tf_s2vt.py --task {train,evaluate,test} [--gpu N] [--<constant> value ...]
"""
import s2vt_b200  # noqa: F401  (alias of the package directory multitask-end-to-end-video-captioning_b200)
from s2vt_b200 import cli

if __name__ == '__main__':
    parser = cli.build_parser(__doc__, dict(model_name='s2vt_model', start_learning_rate=1e-3, decay_steps=5000, clip_norm=10.0, batch_size=64))
    cli.run_xe(parser.parse_args())
