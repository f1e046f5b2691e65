"""nlsolver_b200 — B200 (sm_100a) engine for the DE / PSO population loop (and SANN chain batches) of JSzitas/nlsolver.

The product is libnls_b200.so (C ABI: include/nls_b200.h; kernels: nlsolver_b200/csrc/).  This package is the
Python host binding: `solvers` mirrors the reference's DE / PSO interface, `distributed` shards swarms and islands
across GPUs with torch.distributed."""
from ._lib import (BEALE, BOOTH, BUKIN_N6, GOLDSTEIN_PRICE, LEVI_N13, MATYAS, MCCORMICK, SCHAFFER_N2, SHEKEL,  # noqa: F401
                   STYBLINSKI_TANG, THREE_HUMP_CAMEL)
from ._lib import (ACKLEY, DE_BEST, DE_RANDOM, F32, F64, FLAG_RECORD_MASKS, FLAG_SOCIAL_INDEX_J, PSO_ACCELERATED,  # noqa: F401
                   PSO_VANILLA, RASTRIGIN, ROSENBROCK, ROSENBROCK_EX, SPHERE, NlsError, lib)
from .solvers import (Beale, Booth, BukinN6, Goldstein_Price, LeviN13, Matyas, McCormick, SchafferN2, Shekel,  # noqa: F401
                      StyblinskiTang, ThreeHumpCamel)
from .solvers import (DE, PSO, Ackley, Context, DEPopulation, DESolver, ExchangeWindow, PSOSolver, PSOSwarm, PSOType,  # noqa: F401
                      Rastrigin, RecombinationStrategy, Rosenbrock, RosenbrockExample, SolverStatus, Sphere, de_cfg,
                      default_context, pso_cfg)
from .solvers import SANN, SANNChains, sann_cfg  # noqa: F401
from .solvers import DEIslands, DeviceGroup, ShardedSwarm  # noqa: F401
from .solvers import NelderMeadPSO, nmpso_cfg, nmpso_solve  # noqa: F401
