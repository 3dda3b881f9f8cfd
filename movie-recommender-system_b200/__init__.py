"""mrs_b200 -- B200-native rating-prediction engine (hot path of shared/predictions.scala).

Layout: ``csrc/`` holds the CUDA kernels and the C-ABI (``libmrs_b200.so``, declared in
``include/mrs_b200.h``); ``engine.py`` is the ctypes binding; ``predictions.py`` mirrors the
reference's ``shared.predictions`` function names on top of it; ``synth.py`` generates the
MovieLens-shaped synthetic inputs.  There is no CPU fallback: everything that computes
raises if the CUDA library is missing.
"""
__version__ = "0.1.0"
