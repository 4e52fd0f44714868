"""B200-native Monte Carlo pressure projection (walk-on-stars) behind the `zombie_bindings` API.

Layout:
  csrc/            CUDA kernels (sm_100a), host scene builder, C ABI (include/nmcfs.h), pybind11 module
  libnmcfs.so      built in-tree by `make -C <this dir>` (see __graft_entry__.build)
  capi.py          ctypes view of the C ABI (what tests and bench.py call)
  zombie.py        Python mirror of the reference's module surface: Scene(config, sourceValue), wost(...)
  zombie2d/, zombie3d/   the compiled drop-in modules named `zombie_bindings` (one per dimension)
  sharding.py      multi-GPU point sharding (one process per GPU, gather of the estimates)
  workloads.py     synthetic workloads of each example configuration's shape (bench.py, bench-size parity tests)
  siren.py         fused SIREN velocity network (drop-in for the reference's MLP) + fused Adam (csrc/siren*.cu)
  stepper.py       device-resident operator-split time step (advect fit, divergence grid, wost, projection fit)
  fields.py        density advection + Taylor-Green error on the device (csrc/fields.cu, include/nmcfs_fields.h)

The directory name contains '-', so import it with
    importlib.import_module("neural-monte-carlo-fluid-simulation_b200")
or through __graft_entry__.load_package().
"""
from . import capi, zombie, sharding, workloads  # noqa: F401


def load_fields():
    """Lazy import of the grid post-processing ops (needs torch)."""
    from . import fields as _f
    return _f


def load_siren():
    """Lazy import of the fused SIREN ops (needs torch)."""
    from . import siren as _s
    return _s
from .zombie import Scene, wost  # noqa: F401
