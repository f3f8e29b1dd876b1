"""arecsys_b200 — B200-native implementation of the A-RecSys training hot path.

The directory is named `a-recsys_b200/` (not importable as written); `arecsys_b200.py`
at the repository root loads it under the module name `arecsys_b200`.
Layout: csrc/ (sm_100a kernels + C ABI), _lib.py (ctypes binding), attributes/ hmf/ lstm/
word2vec/ utils/ (host-side mirrors of the reference's interfaces for the hot path).
"""
__version__ = '0.1.0'
