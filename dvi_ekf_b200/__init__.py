"""dvi_ekf_b200 -- B200-native batched VI-ESKF engine (drop-in for the
propagate/update hot path of salehahr/dvi-ekf).  The CUDA library is loaded
lazily by ``dvi_ekf_b200._lib.load()``; there is no CPU fallback."""
from .config import Config  # noqa: F401
from .engine import BatchFilter  # noqa: F401
from .filter import Filter, Simulator, State  # noqa: F401

__all__ = ["BatchFilter", "Config", "Filter", "Simulator", "State"]
