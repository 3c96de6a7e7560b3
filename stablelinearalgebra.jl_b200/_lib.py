"""ctypes binding of libsla_b200.so -- exactly the symbols include/sla_b200.h declares.

There is NO fallback: if the CUDA library is missing or a symbol is absent, importing/using the
package raises.  (The CPU oracle under oracle/ is test infrastructure and is never reached from here.)
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SLA_B200_LIB: another build of the same library (kernel A/B experiments under profiles/); never a fallback
LIB_PATH = os.environ.get("SLA_B200_LIB") or os.path.join(HERE, "csrc", "libsla_b200.so")

SLA_F64, SLA_C64, SLA_F32, SLA_C32 = 0, 1, 2, 3
SLA_OP_N, SLA_OP_C = 0, 1
SLA_INV_IPA, SLA_INV_IPUV = 0, 1
SLA_ENOTIMPL, SLA_EDEVICE = -100, -101
SLA_INFO_RANGE = -1
LDR_ALGOS = {0: "none", 1: "reg", 2: "cluster", 3: "streaming", 4: "hybrid", 5: "blocked", 6: "multicta", 7: "panel", 8: "small"}

_vp, _i, _ll, _sz, _dbl = C.c_void_p, C.c_int, C.c_longlong, C.c_size_t, C.c_double
_common = [_vp, _i, _i, _i]  # handle, dtype, n, batch

SIGNATURES = {
    "sla_version": (_i, []),
    "sla_launch_count": (C.c_ulonglong, []),
    "sla_debug_set_ldr_prof": (None, [_vp]),
    "sla_create": (_i, [C.POINTER(_vp), _vp]),
    "sla_destroy": (_i, [_vp]),
    "sla_set_stream": (_i, [_vp, _vp]),
    "sla_last_ldr_algo": (_i, [_vp]),
    "sla_ldr_algo_launches": (C.c_ulonglong, [_vp, _i]),
    "sla_greens_chain_workspace_info": (_vp, [_vp, _i, _i, _i, _i, _i]),
    "sla_max_n": (_i, [_i]),
    "sla_workspace_bytes": (_sz, [_i, _i, _i]),
    "sla_workspace_info": (_vp, [_vp, _i, _i, _i]),
    "sla_ldr": (_i, _common + [_vp, _vp, _vp, _vp]),
    "sla_ldr_from": (_i, _common + [_vp, _ll, _vp, _vp, _vp, _vp]),
    "sla_set_identity": (_i, _common + [_vp, _vp, _vp]),
    "sla_copy": (_i, _common + [_vp] * 6),
    "sla_ldr_to_mat": (_i, _common + [_vp] * 5),
    "sla_adjoint": (_i, _common + [_vp] * 5),
    "sla_lmul_mat_ldr": (_i, _common + [_vp, _ll, _vp, _vp, _vp, _vp]),
    "sla_rmul_ldr_mat": (_i, _common + [_vp, _vp, _vp, _vp, _ll, _vp]),
    "sla_lmul_ldr_ldr": (_i, _common + [_vp] * 7),
    "sla_rmul_ldr_ldr": (_i, _common + [_vp] * 7),
    "sla_lmul_ldr_mat": (_i, _common + [_vp] * 5),
    "sla_rmul_mat_ldr": (_i, _common + [_vp] * 5),
    "sla_det_lu": (_i, _common + [_vp] * 4),
    "sla_inv_lu": (_i, _common + [_vp] * 4),
    "sla_logabsdet": (_i, _common + [_vp] * 6),
    "sla_ldiv_lu": (_i, _common + [_vp] * 5),
    "sla_ldiv_ldr_mat": (_i, _common + [_vp] * 5),
    "sla_ldiv_ldr_ldr": (_i, _common + [_vp] * 7),
    "sla_ldiv_mat_ldr": (_i, _common + [_vp, _ll, _vp, _vp, _vp, _vp]),
    "sla_rdiv_mat_ldr": (_i, _common + [_vp] * 5),
    "sla_rdiv_ldr_ldr": (_i, _common + [_vp] * 7),
    "sla_rdiv_ldr_mat": (_i, _common + [_vp, _vp, _vp, _vp, _ll, _vp]),
    "sla_inv_IpA": (_i, _common + [_vp] * 7),
    "sla_inv_IpUV": (_i, _common + [_vp] * 10),
    "sla_inv_UpV": (_i, _common + [_vp] * 10),
    "sla_inv_invUpV": (_i, _common + [_vp] * 10),
    "sla_gemm": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _i, _ll, _vp, _i, _ll, _vp, _i, _ll, _i, _i]),
    "sla_walker_sums": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "sla_greens_chain_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "sla_greens_chain": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _dbl, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz]),
    "sla_sweep_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "sla_sweep_workspace_info": (_vp, [_vp, _i, _i, _i, _i, _i]),
    "sla_sweep": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _dbl, _vp, _vp, _vp, _vp, _vp, _vp, _sz]),
}

_lib = None


def load():
    """Load the shared library (once) and attach the prototypes.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
