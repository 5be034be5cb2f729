"""mgld-vsr_b200 — B200-native (sm_100a) implementation of the MGLD-VSR hot path.

Host side: Python/PyTorch classes that keep the reference's call signatures.
Device side: hand-written CUDA kernels in ``csrc/`` behind the C ABI declared in ``include/mgld.h``.
"""
__version__ = "0.1.0"
