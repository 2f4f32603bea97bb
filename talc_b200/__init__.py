"""talc_b200 -- B200-native implementation of TALC's long-read correction hot path.

The product is libtalc_b200.so (hand-written sm_100a kernels behind the C ABI of include/talc_b200.h)
and the drop-in `talc` command line.  This package holds the sources (csrc/), the build recipe, the
ctypes harness used by tests and bench.py, and the synthetic workload generator.
"""
