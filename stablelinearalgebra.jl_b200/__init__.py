"""sla-b200: B200-native (sm_100a) stabilised matrix products for determinant QMC -- the hot path of
StableLinearAlgebra.jl behind a C ABI (include/sla_b200.h), with this thin host-side mirror of the
reference's exported API.  No CPU fallback exists: everything numerical runs in libsla_b200.so."""
import importlib

from . import synthetic  # noqa: F401  (numpy only)
from ._lib import LIB_PATH, SIGNATURES, load  # noqa: F401


def __getattr__(name):
    # `api` needs torch; import lazily so that `synthetic` and the symbol checks work without it
    if name.startswith("__"):
        raise AttributeError(name)
    api = importlib.import_module(__name__ + ".api")
    if name == "api":
        return api
    if hasattr(api, name):
        return getattr(api, name)
    raise AttributeError(name)
