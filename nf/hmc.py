from normalizingflow_b200.hmc import *  # noqa: F401,F403
from normalizingflow_b200.hmc import HMC, FlowSimulation  # noqa: F401
