"""aletsch_b200: B200-native per-bundle read-evidence path of the Aletsch meta-assembler.

The product is the C-ABI library ``libaletsch_gpu.so`` (include/aletsch_gpu.h, hand-written
sm_100a CUDA in aletsch_b200/csrc).  This Python package only binds it with ctypes for the
tests and bench.py, and binds the host-side packer / synthetic generator
(``libaletsch_host.so``).  There is no CPU fallback: ``aletsch_b200.gpu`` raises if the CUDA
library is missing or no device is usable.
"""

__all__ = ["hostlib", "gpu"]
