"""B200-native (sm_100a) implementation of the NaturalSpeech2 denoiser hot path.

Public names mirror naturalspeech2_pytorch/__init__.py:8-24 for the path this package accelerates.
"""
from . import _lib, ops  # noqa: F401

__version__ = "0.1.0"
