from normalizingflow_b200.models import *  # noqa: F401,F403
from normalizingflow_b200.models import NormalizingFlowModel, GaussianPrior  # noqa: F401
