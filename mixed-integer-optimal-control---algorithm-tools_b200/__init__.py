"""B200-native trust-region subproblem solver (bellman_TRM! / eval_u_TRM! of the reference).

The compute lives in libbellman_b200.so (hand-written sm_100a CUDA behind a C ABI, include/bellman_b200.h);
this package is the Python mirror of the reference's operator interface for that path.  Importing the
package does not need a GPU; calling into it does -- there is no CPU fallback.
"""
from . import _lib
from ._lib import BellmanB200Error, InexactError, StaleCellError
from .api import Comm, MultiPlan, TRMPlan, bellman_TRM, eval_u_TRM, fp64_peak, nccl_version
from .iterators import bounded_sum_iterator, flatten, jump_cost_table, product_iterator

__all__ = ["TRMPlan", "MultiPlan", "Comm", "nccl_version", "bellman_TRM", "eval_u_TRM", "product_iterator", "bounded_sum_iterator", "flatten",
           "jump_cost_table", "fp64_peak", "BellmanB200Error", "InexactError", "StaleCellError", "device_count"]


def device_count() -> int:
    return int(_lib.load().bb200_device_count())
