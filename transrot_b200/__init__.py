"""transrot_b200 — B200-native backend for TransRot's rigid-body Metropolis annealing hot path.

The CUDA library (libtransrot_b200.so, C ABI in include/transrot_b200.h) does all the work; this
package is the host-side mirror of the reference's data model plus a ctypes binding.
"""
from .host import (  # noqa: F401
    AtomDef, Config, ParticleDef, System, SystemBuilder, TransRotError, center_of_mass, read_dbase, read_input,
    read_params, setup_pair_vals, system_from_config,
)
from .engine import (  # noqa: F401
    Engine, EXPORTED_SYMBOLS, LIB_PATH, MINIMUM_DTYPE, MultiEngine, POINT_DTYPE, SCHEDULE_DTYPE, SUMMARY_DTYPE, TRACE_DTYPE,
    TR_KERNEL_AUTO, TR_KERNEL_CREW, TR_KERNEL_CREW2, TR_KERNEL_TEAM, load_library, schedule_from_config,
)
from .workloads import WORKLOADS, sweep_schedules, workload_config  # noqa: F401
from .sharding import chain_range, gather_minima, global_minimum, owner_of  # noqa: F401
from .replay import replay_outputs  # noqa: F401
from .writers import (  # noqa: F401
    format_acceptance_line, format_energies, format_heat_capacity_line, format_movie_frame, format_output_xyz,
    java_double_to_string, precision_round,
)
