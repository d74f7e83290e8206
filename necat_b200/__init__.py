"""necat_b200 -- B200-native implementation of NECAT's oc2pmov -> oc2cns hot path.

The product is libnecat_b200.so (CUDA, sm_100a) and the drop-in executables under
necat_b200/bin; this package is the Python-side mirror used by tests and bench.py:
`capi` (ctypes binding of include/necat_b200.h) and `formats` (PackedDB volumes,
candidate records, FASTA outputs).
"""
from . import formats  # noqa: F401

__all__ = ["formats", "capi"]
