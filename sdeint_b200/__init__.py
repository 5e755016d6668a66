"""sdeint_b200 -- B200-native ensemble SDE integrator with the API of mattja/sdeint.

Exports the 13 public names of sdeint/__init__.py:3-5 (same positional order and keyword
names) plus the operator set (`ops`), the tracer that turns plain Python callables into such operators
(`trace.trace_drift`, `trace.trace_diffusion`: the integrators call it themselves), ensemble helpers and
`launch_count()`.  All numerical
work happens in hand-written sm_100a CUDA kernels reached through the C ABI of
include/sdb.h (libsdb.so, bound with ctypes in `_lib.py`); importing the package does not
need a GPU, computing does, and there is no CPU fallback."""
from . import ops, trace  # noqa: F401  (trace: plain Python f(y, t), G(y, t) -> device operators)
from .wiener import deltaW, Ikpw, Jkpw, Iwik, Jwik, Icomm, Jcomm  # noqa: F401
from .integrate import (Error, SDEValueError, itoint, stratint, itoEuler, stratHeun,  # noqa: F401
                        itoSRI2, stratSRS2, stratKP2iS, itoKP2i, stratR2, last_status)
from ._lib import jit_cache_stats, launch_count, last_kernel  # noqa: F401
from . import ensemble  # noqa: F401

__version__ = '0.1.0'
