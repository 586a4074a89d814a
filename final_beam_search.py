#!/usr/bin/env python
"""Beam-search captioning on precomputed features (drop-in for final_beam_search.py: beam 3, no length normalisation).

    python final_beam_search.py --task {train,evaluate,test} [--gpu N] [--<constant> value ...]
"""
import s2vt_b200  # noqa: F401  (alias of the package directory multitask-end-to-end-video-captioning_b200)
from s2vt_b200 import cli

if __name__ == '__main__':
    parser = cli.build_parser(__doc__, dict(task='test', beam_size=3, length_normalization_factor=0.0, batch_size=64, out_file='best5frame_beam3.txt'))
    cli.run_beam(parser.parse_args())
