"""vfvm_b200 -- B200-native Newton hot path behind the VoronoiFVM.jl API (host-side mirror in Python).

The compute path is libvfvmb200.so (csrc/, hand-written sm_100a CUDA behind the C ABI in include/vfvm_b200.h).
Nothing in this package evaluates physics, assembles or solves on the CPU; if the CUDA library or a GPU is
missing, creating a SystemState fails loudly.
"""
from .grid import Grid, simplexgrid, cartesian, circular_symmetric, spherical_symmetric, cellmask  # noqa: F401
from . import physics  # noqa: F401
from .physics import Physics, BCondition, UnregisteredPhysicsError  # noqa: F401
from .system import (System, enable_species, enable_boundary_species, boundary_dirichlet, boundary_neumann, boundary_robin, unknowns, num_dof,  # noqa: F401
                     DIRICHLET)
from .system import physics as set_physics  # noqa: F401
from .state import SystemState  # noqa: F401,E402
from .solver import (SolverControl, NewtonSolverHistory, TransientSolution, solve, solve_state, solve_step, solve_transient,  # noqa: F401,E402
                     evaluate_residual_and_jacobian, fixed_timesteps, ConvergenceError, AssemblyError, LinearSolverError, EmbeddingError,
                     KrylovJL_BICGSTAB, KrylovJL_CG, KrylovJL_GMRES, JacobiPreconBuilder, BlockPreconBuilder, ILUZeroPreconBuilder, AMGPreconBuilder, SmoothedAggregationPreconBuilder,
                     DeviceDirectLike)
from . import postprocess  # noqa: F401,E402
from .postprocess import (integrate, edgeintegrate, lpnorm, l2norm, w1pseminorm, h1seminorm, w1pnorm, h1norm, nodevolumes, eval_rhs, eval_jacobian, mass_matrix, nodeflux)  # noqa: F401,E402
