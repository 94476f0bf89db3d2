"""B200-native ORB descriptor-matching hot path (ORBmatcher family + DBoW2 transform).

Host-side mirror of the reference interface lives in `matcher` (loads the CUDA C-ABI library
liborbmatch_b200.so and fails loudly when it is missing -- there is no CPU fallback).
`synth` holds the seeded synthetic workloads, `_abi` the ctypes view of include/orbmatch_b200.h.
"""
__version__ = "0.1.0"
