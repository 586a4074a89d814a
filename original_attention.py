#!/usr/bin/env python
"""Temporal-attention caption model (drop-in for the reference script of the same name): single LSTM3 decoder with additive
attention over the frame embeddings, tanh MLP head, hinge regulariser on the attention weights of the first 8 frames.

    python original_attention.py --task {train,test,evaluate} [--gpu N] [--<constant> value ...]
"""
import s2vt_b200  # noqa: F401  (alias of the package directory multitask-end-to-end-video-captioning_b200)
from s2vt_b200 import cli

if __name__ == '__main__':
    parser = cli.build_parser(__doc__, dict(model_path='./attention_models', model_name='_beta10_m05_32img_attention_model', start_learning_rate=1e-4,
                                            decay_steps=10000, clip_norm=10.0, batch_size=1, n_epochs=20, seed_num=16, n_video_lstm_step=5,
                                            out_file='beta10_m05_batch64_32img_attention_model_val'))
    cli.run_attention(parser.parse_args())
