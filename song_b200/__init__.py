"""song_b200 -- B200-native second-order line-of-sight transfer stage of SONG (transfer2 / bessel2).

Layout: ``csrc/`` CUDA kernels + C-ABI (``include/song_los.h``), ``host/`` the C drop-in module
with SONG's own API, and thin Python mirrors used by the tests and ``bench.py``.
"""
__version__ = "0.1.0"
