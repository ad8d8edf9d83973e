from normalizingflow_b200.flows import *  # noqa: F401,F403
from normalizingflow_b200.flows import FCNN, RealNVP, NSF_AR, NSF_CL, Planar, Radial  # noqa: F401
