"""B200-native S2VT caption hot path (drop-in for the reference's Video_Caption_Generator graphs).

Import with ``importlib.import_module('multitask-end-to-end-video-captioning_b200')`` or through the alias
module ``s2vt_b200`` at the repo root.  All compute runs in libs2vt_b200.so (hand-written sm_100a CUDA behind
the C ABI of include/s2vt.h); there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .model import Video_Caption_Generator, LSTM1_W, LSTM1_B, LSTM2_W, LSTM2_B  # noqa: F401
from . import text, cider, rewards, trainer, checkpoint, cli, ingest, attention  # noqa: F401,E402
