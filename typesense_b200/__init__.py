"""tsgpu — B200-native query hot path for Typesense.

The product is the CUDA library `libtsgpu.so` (typesense_b200/csrc, C-ABI in include/tsgpu.h). This Python package is
only the harness around it: ctypes bindings (capi), numpy batch builders (structs) and synthetic collections (synth).
"""
