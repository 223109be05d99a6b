"""periodicity_b200 -- B200-native brute-force trial-frequency search.

Drop-in ``GLS`` (``periodicity_b200.spectral``), ``PDM`` and ``StringLength``
(``periodicity_b200.phase``) with the call signatures of dioph/periodicity,
dispatching through a C ABI (``include/periodicity_b200.h``) into hand-written
sm_100a CUDA kernels.  No CPU fallback.
"""
from .core import FSeries, TSeries  # noqa: F401
from .phase import AOV, CE, GL, PDM, ConditionalEntropy, GregoryLoredo, StringLength  # noqa: F401
from .spectral import GLS  # noqa: F401

__version__ = "0.2.0"
__all__ = ["GLS", "PDM", "StringLength", "AOV", "CE", "ConditionalEntropy", "GL", "GregoryLoredo", "TSeries", "FSeries"]
