"""
bx_python_b200 -- B200-native (sm_100a) drop-in for bx-python's interval-intersection / binned-bitset hot path.

Host side: pure Python mirroring bx.intervals.intersection / bx.bitset; all arithmetic runs in hand-written CUDA
kernels inside libbxb200.so, reached through the ctypes C-ABI declared in include/bxb200.h.  There is no CPU
fallback: any compute call raises RuntimeError when the library or a CUDA device is missing.
"""
__version__ = "0.1.0"
