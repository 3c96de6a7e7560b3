"""ctypes wrapper of oracle/_build/liboracle.so (the C restatement, sla_oracle_c.c).

TEST INFRASTRUCTURE ONLY: used by tests/ and by bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        _lib = C.CDLL(LIB)
        _lib.oracle_eval_batch.restype = C.c_int
        _lib.oracle_eval_batch.argtypes = [C.c_int, C.c_long, C.c_long, C.c_long, C.c_long, C.c_void_p, C.c_void_p,
                                           C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib.oracle_eval_batch_bt.restype = C.c_int
        _lib.oracle_eval_batch_bt.argtypes = _lib.oracle_eval_batch.argtypes + [C.c_int]
        _lib.oracle_blas_config.restype = C.c_char_p
    return _lib


def blas_config() -> str:
    return load().oracle_blas_config().decode()


def eval_batch(expK, fields, lam, n_stab, inverse="IpA", nthreads=1, blas_threads=1):
    """fields int8 (chains, L, n) -> (G (chains,n,n) ordinary indexing, logdet, sign); one chain per
    worker thread, `blas_threads` BLAS threads each (default 1)."""
    lib = load()
    chains, L, n = fields.shape
    cplx = np.iscomplexobj(expK)
    dt = np.complex128 if cplx else np.float64
    eK = np.asfortranarray(expK.astype(dt))
    f = np.ascontiguousarray(fields.astype(np.int8))
    G = np.zeros((chains, n, n), dtype=dt)          # each [c] is filled column-major -> transpose below
    logdet = np.zeros(chains)
    sgn = np.zeros(chains, dtype=dt)
    rc = lib.oracle_eval_batch_bt(int(cplx), n, L, n_stab, chains, eK.ctypes.data, f.ctypes.data, float(lam),
                                  0 if inverse == "IpA" else 1, G.ctypes.data, logdet.ctypes.data, sgn.ctypes.data,
                                  int(nthreads), int(blas_threads))
    assert rc == 0
    return np.swapaxes(G, 1, 2).copy(), logdet, sgn
