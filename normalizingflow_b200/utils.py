"""Drop-in for the spline functions of the reference's nf/utils.py (:13-152), on libnfk."""
from __future__ import annotations

from . import _ops

DEFAULT_MIN_BIN_WIDTH = 1e-3      # nf/utils.py:13
DEFAULT_MIN_BIN_HEIGHT = 1e-3     # nf/utils.py:14
DEFAULT_MIN_DERIVATIVE = 1e-3     # nf/utils.py:15


def _check_minima(K, min_bin_width, min_bin_height, min_derivative):
    for name, v in (("min_bin_width", min_bin_width), ("min_bin_height", min_bin_height),
                    ("min_derivative", min_derivative)):
        if abs(v - 1e-3) > 1e-12:
            raise NotImplementedError(f"{name} is fixed at the reference default 1e-3 in the CUDA kernels")
    # nf/utils.py:68-71
    if min_bin_width * K > 1.0:
        raise ValueError("Minimal bin width too large for the number of bins")
    if min_bin_height * K > 1.0:
        raise ValueError("Minimal bin height too large for the number of bins")


def unconstrained_RQS(inputs, unnormalized_widths, unnormalized_heights, unnormalized_derivatives,
                      inverse=False, tail_bound=1., min_bin_width=DEFAULT_MIN_BIN_WIDTH,
                      min_bin_height=DEFAULT_MIN_BIN_HEIGHT, min_derivative=DEFAULT_MIN_DERIVATIVE,
                      arith=_ops.DEFAULT_ARITH):
    """nf/utils.py:27-56: identity outside [-tail_bound, tail_bound], RQS inside.
    Returns (outputs, logabsdet) shaped like ``inputs``.  Unlike the reference this does not
    raise when no element is inside the interval."""
    K = unnormalized_widths.shape[-1]
    _check_minima(K, min_bin_width, min_bin_height, min_derivative)
    out, lad, _ = _ops.unconstrained_rqs(inputs, unnormalized_widths, unnormalized_heights,
                                         unnormalized_derivatives, inverse, float(tail_bound), arith)
    return out, lad
