"""meshgraphnets.jl_b200 - B200-native (sm_100a) Encode-Process-Decode hot path of MeshGraphNets.jl.

csrc/            CUDA kernels + the C ABI (libmgn_b200.so, include/mgn_b200.h)
core.py          host mirror of the GraphNetCore.jl names MeshGraphNets.jl calls
graph.py         mirror of src/graph.jl      (create_base_graph, build_graph)
solve.py         mirror of src/solve.jl      (ode_step, ode_func_eval, rollout)
strategies.py    mirror of src/strategies.jl (DerivativeTraining, SolverTraining, MultipleShooting steps)
shooting.py      lock-step shooting intervals: fixed-step Runge-Kutta solve + exact reverse sweep over the C ABI
partition.py     graph partitioning + halo exchange for meshes larger than one GPU
workloads.py     synthetic BASELINE workloads + the driver's mask helpers (src/MeshGraphNets.jl:352-358)
parallel.py      data-parallel plumbing (window sharding, gradient / normaliser all-reduce)
"""
from ._lib import COMPUTE_BF16, COMPUTE_FP32, LIB_PATH, MgnError, load  # noqa: F401
from .core import (Adam, FeatureGraph, step_dp_, GraphIndex, GraphNetwork, Model, NormaliserOfflineMeanStd,  # noqa: F401
                   NormaliserOfflineMinMax, NormaliserOnline, build_model, edge_features, init_params,
                   inverse_data, mse_reduce, one_hot, parse_edges, profile_begin, profile_end, profile_tag,
                   shift_one_based, step_,
                   triangles_to_edges)
from .graph import build_graph, create_base_graph, create_base_graph_device  # noqa: F401
from .parallel import Communicator, pack_normaliser_states, allreduce_mean_, allreduce_normaliser_, allreduce_sum_, shard_windows  # noqa: F401
from .partition import (AbiExchange, DistExchange, LocalExchange, LocalGraph, PartitionedModel, build_partition,  # noqa: F401
                        build_partition_rank,
                        masked_mse_partial, partition_bounds, run_partitioned_step)
from ._lib import NORM_FORWARD, NORM_FORWARD_VJP, NORM_INVERSE, NORM_INVERSE_VJP  # noqa: F401
from ._lib import (HALO_GRAD, HALO_LATENT, ROWS_ADD, ROWS_PACK, ROWS_PACK_ZERO, ROWS_UNPACK, STAGE_DECODE,  # noqa: F401
                   STAGE_ENCODE)
from .fused import FusedGraph, backward_fused, forward_fused, update_online  # noqa: F401
from .solve import CapturedRollout, ode_func_eval, ode_step, ode_step_unfused, rk_step, rollout  # noqa: F401
from .workloads import (chain_edges, cylinder_flow_mesh, node_mask, synthetic_velocity, tet_grid_edges,  # noqa: F401
                        val_mask)
from .strategies import (DerivativeTraining, MultipleShooting, SolverTraining, get_delta, init_train_step,  # noqa: F401
                         train_step)
from .shooting import (DeviceAlgebra, DeviceRhs, RK_TABLEAUS, ShootingEngine, multiple_shooting_step,  # noqa: F401
                       shard_intervals, shooting_ranges, solver_training_step, time_steps)
