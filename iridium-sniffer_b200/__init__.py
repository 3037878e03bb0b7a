"""B200-native Iridium burst detect -> downmix -> DQPSK path (host-side Python mirror).

The product is the C-ABI CUDA library built from csrc/ (libiridium_b200.so);
this package only loads it through ctypes for the tests and bench.py.
"""
