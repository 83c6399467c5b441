"""swem_b200 -- B200-native sequential weighted EM memory (SWEM) behind the reference's API.

``SWEM`` / ``SWEMCore`` mirror ``methods/SWEM/swem.py`` / ``methods/SWEM/modules.py`` of lmm077/SWEM;
the EM update and the readout run in hand-written sm_100a CUDA kernels (``libswem_b200.so``,
C ABI in ``include/swem_b200.h``).  There is no CPU fallback.
"""
from .core import MemoryBank, SWEMCore
from .model import SWEM, make_config

__all__ = ['SWEM', 'SWEMCore', 'MemoryBank', 'make_config']
__version__ = '0.1.0'
