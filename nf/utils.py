from normalizingflow_b200.utils import *  # noqa: F401,F403
from normalizingflow_b200.utils import (unconstrained_RQS, DEFAULT_MIN_BIN_WIDTH, DEFAULT_MIN_BIN_HEIGHT,  # noqa: F401
                                        DEFAULT_MIN_DERIVATIVE)
