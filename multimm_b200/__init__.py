"""multimm_b200 — B200-native energy-minimisation engine behind MultiMM's driver interface.

Importing the package does not load CUDA; ``Engine`` / ``MultiMM`` load
``libmultimm_b200.so`` (built in-tree by ``python -m multimm_b200.build``) and fail loudly if it
is missing.  There is no CPU fallback.
"""
__version__ = "0.1.0"

from ._lib import Error  # noqa: F401


def __getattr__(name):
    if name == "Engine":
        from .engine import Engine
        return Engine
    if name == "MultiMM":
        from .model import MultiMM
        return MultiMM
    if name == "SimulationConfig":
        from .config import SimulationConfig
        return SimulationConfig
    raise AttributeError(name)
