#!/usr/bin/env python
"""Beam-search captioning (drop-in for the post-CNN part of e2e_beam_search.py: beam 4, length normalisation 1).

    python e2e_beam_search.py --task {train,evaluate,test} [--gpu N] [--<constant> value ...]
"""
import s2vt_b200  # noqa: F401  (alias of the package directory multitask-end-to-end-video-captioning_b200)
from s2vt_b200 import cli

if __name__ == '__main__':
    parser = cli.build_parser(__doc__, dict(task='test', beam_size=4, length_normalization_factor=1.0, batch_size=64, out_file='e2e_beam4.txt'))
    cli.run_beam(parser.parse_args())
