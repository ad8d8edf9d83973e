"""normalizingflow_b200 — B200 (sm_100a) implementation of the normalizing-flow transform hot
path of sherryli59/NormalizingFlow behind the reference's own class API.

    from normalizingflow_b200.flows import FCNN, RealNVP, NSF_CL, Planar, Radial
    from normalizingflow_b200.models import NormalizingFlowModel
    from normalizingflow_b200.hmc import HMC, FlowSimulation

The top-level ``nf`` package of this repository re-exports the same modules under the
reference's import names (``nf.flows``, ``nf.models``, ``nf.utils``, ``nf.hmc``).
All compute runs in libnfk.so (hand-written CUDA, include/nfk.h); there is no CPU path.
"""
from . import _lib  # noqa: F401  (fails loudly when libnfk.so is missing)
from . import flows, models, utils  # noqa: F401

__all__ = ["flows", "models", "utils", "hmc", "dist"]
