"""swat_b200 -- B200-native retrieval hot path for SWAT (score -> per-class top-k -> T2I filter).

Host-side mirror of ``retrieval/sample_retrieval.py``'s hot-path interface over a C-ABI CUDA
library (``include/swat_b200.h``).  There is no CPU fallback: every compute entry point raises if
``libswat_b200.so`` is missing or no CUDA device is present.
"""
__version__ = "0.1.0"
