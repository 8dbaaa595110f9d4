"""osqp-solver_b200: B200-native batched OSQP-style ADMM for GOMP trajectory QPs.

The directory name carries a hyphen (it mirrors the reference's repo name), so it is loaded
through `load_package()` in the repo-root `__graft_entry__.py` / tests under the module name
`osqp_solver_b200`.  Sub-modules:
  problem  -- ctypes view of libgomp_host (synthetic GOMP QP batches via the C++ facade builder)
  solver   -- ctypes view of libgomp_cuda (the C ABI in include/gomp_cuda.h)
  sharding -- contiguous batch shards per rank + the final result gather (torch.distributed)
  sqp      -- batched GOMPSolver::run (one horizon): SQP rounds with the data plane on the GPU
"""
from . import problem  # noqa: F401
from . import solver  # noqa: F401
from . import sharding  # noqa: F401
from . import sqp  # noqa: F401

__all__ = ["problem", "solver", "sharding", "sqp"]
