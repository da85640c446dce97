"""regda_b200 -- B200-native (sm_100a) implementation of RegDA's self-training hot path.

Mirrors the reference's Python surface for that path (same module paths below `regda.`,
same class / function names and argument meaning) on top of a C-ABI CUDA library
(include/regda_b200.h).  There is no CPU fallback: every op raises if the library or a CUDA
device is missing.
"""
__version__ = "0.1.0"
