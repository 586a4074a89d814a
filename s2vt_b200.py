"""Importable alias of the package directory `multitask-end-to-end-video-captioning_b200` (hyphens are not valid
in an `import` statement)."""
import importlib
import sys

_pkg = importlib.import_module('multitask-end-to-end-video-captioning_b200')
sys.modules[__name__] = _pkg
