"""B200-native hot path of the Landau-Poisson solver (collision operator + DG advection).

The directory name carries hyphens (it mirrors the reference's repository name), so import it
through ``__graft_entry__.load_package()`` / ``tests/conftest.py`` which register it as the
module ``lpsolver_b200``.  Public surface:

* ``lpgpu``   -- ctypes binding of the C ABI in include/lpgpu.h (liblpgpu.so, CUDA kernels)
* ``solver``  -- host-side mirror of the reference driver for this path (time loop, sharding)
"""
from . import lpgpu  # noqa: F401
from .lpgpu import LPGpu, LPGpuError, Params, library_path  # noqa: F401
