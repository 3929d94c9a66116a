"""oracle -- TEST INFRASTRUCTURE ONLY (parity checker, never the product path).

ctypes front ends for

* ``liboracle.so``  -- the CPU restatement of the reference algorithm (oracle/spgemm_oracle.cpp);
* ``_ref/libosp_ref_{a,b}.so`` -- the UNMODIFIED reference sources compiled by oracle/Makefile
  (present when the build container had /root/reference; the built files travel to the GPU box).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  Parity pin: the reference ships no golden vectors; the
restatement is pinned against the compiled reference itself (tests/test_oracle_vs_ref.py) and the
fixtures under tests/golden/ that were generated from it (tests/golden/make_golden.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ELEM = np.dtype([("idx", "<u4"), ("val", "<f4")])

_vp, _u64 = C.c_void_p, C.c_uint64


def build(quiet: bool = True) -> None:
    """Compiles the restatement and, when the reference sources are present, oracle/_ref."""
    subprocess.run(["make", "-C", _HERE] + (["-s"] if quiet else []), check=True)


def _load(name: str) -> Optional[C.CDLL]:
    path = os.path.join(_HERE, name)
    return C.CDLL(path) if os.path.exists(path) else None


_port = None
_ref_a = None
_ref_b = None


def port_lib() -> C.CDLL:
    global _port
    if _port is None:
        if not os.path.exists(os.path.join(_HERE, "liboracle.so")):
            build()
        _port = _load("liboracle.so")
        _port.orc_mtx_open.restype = _vp
        _port.orc_mtx_open.argtypes = [C.c_char_p, C.c_int]
        _port.orc_mtx_dims.argtypes = [_vp, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64)]
        _port.orc_mtx_copy.argtypes = [_vp, _vp, _vp, _vp]
        _port.orc_mtx_free.argtypes = [_vp]
        _port.orc_coo2csr.argtypes = [_u64, _vp, _vp, _vp, _u64, C.c_int, C.c_int, _vp, _vp]
        _port.orc_csr2csc.argtypes = [_u64, _u64, _vp, _vp, _vp, _vp]
        _port.orc_flops.restype = _u64
        _port.orc_flops.argtypes = [_u64, _vp, _vp]
        _port.orc_spgemm.restype = _vp
        _port.orc_spgemm.argtypes = [_u64, _vp, _vp, _vp, _vp, _u64]
        _port.orc_spgemm_rowblocks.restype = _vp
        _port.orc_spgemm_rowblocks.argtypes = [_u64, _vp, _vp, _vp, _vp, _u64]
        _port.orc_result_dims.argtypes = [_vp, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64)]
        _port.orc_result_copy.argtypes = [_vp, _vp, _vp]
        _port.orc_result_free.argtypes = [_vp]
    return _port


def ref_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libosp_ref_a.so")) and os.path.exists(
        os.path.join(_HERE, "_ref", "libosp_ref_b.so"))


def ref_libs() -> Tuple[C.CDLL, C.CDLL]:
    global _ref_a, _ref_b
    if _ref_a is None:
        if not ref_available():
            raise RuntimeError("oracle/_ref is not built (needs /root/reference at build time)")
        a = _load("_ref/libosp_ref_a.so")
        b = _load("_ref/libosp_ref_b.so")
        a.ref_mtx_open.restype = _vp
        a.ref_mtx_open.argtypes = [C.c_char_p, C.c_int]
        a.ref_mtx_dims.argtypes = [_vp, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64)]
        a.ref_mtx_copy.argtypes = [_vp, _vp, _vp, _vp]
        a.ref_mtx_free.argtypes = [_vp]
        a.ref_coo2csr.argtypes = [_u64, _vp, _vp, _vp, _u64, C.c_int, _vp, _vp]
        a.ref_spgemm.restype = _vp
        a.ref_spgemm.argtypes = [_u64, _vp, _vp, _vp, _vp]
        a.ref_compare.argtypes = [_u64, _vp, _vp, _u64, _vp, _vp]
        a.ref_result_dims.argtypes = [_vp, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64), C.POINTER(C.c_double)]
        a.ref_result_copy.argtypes = [_vp, _vp, _vp]
        a.ref_result_free.argtypes = [_vp]
        a.ref_compact.restype = _u64
        a.ref_compact.argtypes = [C.c_int, _u64, _vp, _vp, _vp, _vp, _vp, _vp]
        a.ref_compact_spgemm.restype = _vp
        a.ref_compact_spgemm.argtypes = [_u64, _vp, _vp, _vp, _vp, _u64, _vp, _vp, C.POINTER(C.c_int)]
        b.ref_taskprovider.restype = _vp
        b.ref_taskprovider.argtypes = [_u64, _vp, _vp, _vp, _vp]
        b.ref_tp_dims.argtypes = [_vp, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64),
                                  C.POINTER(C.c_double)]
        b.ref_tp_copy.argtypes = [_vp, _vp, _vp, _vp, _vp]
        b.ref_tp_free.argtypes = [_vp]
        _ref_a, _ref_b = a, b
    return _ref_a, _ref_b


def _p(a: np.ndarray):
    return a.ctypes.data if a.size else None


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


# ------------------------------------------------------------------------------ restatement ----
def readcoo(path: str, sym: bool = False, impl: str = "port"):
    """-> (rows u32, cols u32, vals f32, NRow, NCol)"""
    if impl == "port":
        lib, pre = port_lib(), "orc"
    else:
        lib, pre = ref_libs()[0], "ref"
    h = getattr(lib, pre + "_mtx_open")(os.fsencode(path), int(sym))
    if not h:
        raise FileNotFoundError(path)
    nr, nc, nz = _u64(), _u64(), _u64()
    getattr(lib, pre + "_mtx_dims")(h, C.byref(nr), C.byref(nc), C.byref(nz))
    rows, cols, vals = np.empty(nz.value, np.uint32), np.empty(nz.value, np.uint32), np.empty(nz.value, np.float32)
    getattr(lib, pre + "_mtx_copy")(h, _p(rows), _p(cols), _p(vals))
    getattr(lib, pre + "_mtx_free")(h)
    return rows, cols, vals, nr.value, nc.value


def coo2csr(rows, cols, vals, N: int, transpose: bool = False, impl: str = "port", single_row_quirk: bool = False):
    """-> (status, pos u64[N+1], data ELEM[nnz]); status 233 on duplicates."""
    rows, cols, vals = _c(rows, np.uint32), _c(cols, np.uint32), _c(vals, np.float32)
    n = len(rows)
    pos, data = np.zeros(N + 1, np.uint64), np.zeros(n, ELEM)
    if impl == "port":
        rc = port_lib().orc_coo2csr(n, _p(rows), _p(cols), _p(vals), N, int(transpose), int(single_row_quirk),
                                    _p(pos), _p(data))
    else:
        rc = ref_libs()[0].ref_coo2csr(n, _p(rows), _p(cols), _p(vals), N, int(transpose), _p(pos), _p(data))
    return rc, pos, data


def csr2csc(n_major: int, n_minor: int, pos, data):
    pos, data = _c(pos, np.uint64), _c(data, ELEM)
    pos_out, data_out = np.zeros(n_minor + 1, np.uint64), np.zeros(len(data), ELEM)
    rc = port_lib().orc_csr2csc(n_major, n_minor, _p(pos), _p(data), _p(pos_out), _p(data_out))
    if rc:
        raise ValueError("index out of range")
    return pos_out, data_out


def flops(a_pos, b_pos) -> int:
    a_pos, b_pos = _c(a_pos, np.uint64), _c(b_pos, np.uint64)
    return int(port_lib().orc_flops(len(a_pos) - 1, _p(a_pos), _p(b_pos)))


def spgemm(a_csc_pos, a_csc_data, b_pos, b_data, rows_override: int = 0, impl: str = "port"):
    """C = A*B, A as CSC, B as CSR -> (pos u64, data ELEM, products[, seconds for impl='ref'])."""
    a_pos, a_data = _c(a_csc_pos, np.uint64), _c(a_csc_data, ELEM)
    b_pos, b_data = _c(b_pos, np.uint64), _c(b_data, ELEM)
    n_k = len(a_pos) - 1
    assert len(b_pos) - 1 == n_k, "k dimensions differ"
    rows, nnz, prod = _u64(), _u64(), _u64()
    if impl == "port":
        lib = port_lib()
        h = lib.orc_spgemm(n_k, _p(a_pos), _p(a_data), _p(b_pos), _p(b_data), rows_override)
        lib.orc_result_dims(h, C.byref(rows), C.byref(nnz), C.byref(prod))
        pos, data = np.zeros(rows.value + 1, np.uint64), np.zeros(nnz.value, ELEM)
        lib.orc_result_copy(h, _p(pos), _p(data))
        lib.orc_result_free(h)
        return pos, data, prod.value
    lib = ref_libs()[0]
    sec = C.c_double()
    h = lib.ref_spgemm(n_k, _p(a_pos), _p(a_data), _p(b_pos), _p(b_data))
    lib.ref_result_dims(h, C.byref(rows), C.byref(nnz), C.byref(prod), C.byref(sec))
    pos, data = np.zeros(rows.value + 1, np.uint64), np.zeros(nnz.value, ELEM)
    lib.ref_result_copy(h, _p(pos), _p(data))
    lib.ref_result_free(h)
    return pos, data, prod.value, sec.value


def spgemm_rowblocks(a_csr_pos, a_csr_data, b_pos, b_data, rows_per_block: int = 4096):
    """Same result as spgemm() but from CSR(A), bounded memory (for the CPU baseline at size)."""
    a_pos, a_data = _c(a_csr_pos, np.uint64), _c(a_csr_data, ELEM)
    b_pos, b_data = _c(b_pos, np.uint64), _c(b_data, ELEM)
    lib = port_lib()
    h = lib.orc_spgemm_rowblocks(len(a_pos) - 1, _p(a_pos), _p(a_data), _p(b_pos), _p(b_data), rows_per_block)
    rows, nnz, prod = _u64(), _u64(), _u64()
    lib.orc_result_dims(h, C.byref(rows), C.byref(nnz), C.byref(prod))
    pos, data = np.zeros(rows.value + 1, np.uint64), np.zeros(nnz.value, ELEM)
    lib.orc_result_copy(h, _p(pos), _p(data))
    lib.orc_result_free(h)
    return pos, data, prod.value


def ref_compact(pos, data, raw: bool = False):
    """The reference's csr2compact (SimSpGEMM.cpp:154-219) or, with raw, csc2rawcompact (:221-243), unmodified
    -> (group_pos u64, rows u32, cols u32, vals f32)."""
    pos, data = _c(pos, np.uint64), _c(data, ELEM)
    lib = ref_libs()[0]
    n = len(pos) - 1
    npos = lib.ref_compact(int(raw), n, _p(pos), _p(data), None, None, None, None)
    gpos = np.zeros(npos, np.uint64)
    nnz = len(data)
    rows, cols, vals = np.zeros(nnz, np.uint32), np.zeros(nnz, np.uint32), np.zeros(nnz, np.float32)
    lib.ref_compact(int(raw), n, _p(pos), _p(data), _p(gpos), _p(rows), _p(cols), _p(vals))
    return gpos, rows, cols, vals


def ref_compact_spgemm(group_pos, rows, cols, vals, b_pos, b_data):
    """The reference's compactMulcsr (SimSpGEMM.cpp:247-263) on a compact operand, its partial products folded by the
    deduplicateCOO rule with a stable sort -> (pos u64, data ELEM, products).  Raises ValueError(233) when its
    dupcheck throws."""
    group_pos, b_pos, b_data = _c(group_pos, np.uint64), _c(b_pos, np.uint64), _c(b_data, ELEM)
    rows, cols, vals = _c(rows, np.uint32), _c(cols, np.uint32), _c(vals, np.float32)
    lib = ref_libs()[0]
    status = C.c_int()
    h = lib.ref_compact_spgemm(len(group_pos) - 1, _p(group_pos), _p(rows), _p(cols), _p(vals), len(b_pos) - 1, _p(b_pos), _p(b_data),
                               C.byref(status))
    if not h:
        raise ValueError(status.value)
    nrows, nnz, prod, sec = _u64(), _u64(), _u64(), C.c_double()
    lib.ref_result_dims(h, C.byref(nrows), C.byref(nnz), C.byref(prod), C.byref(sec))
    pos, data = np.zeros(nrows.value + 1, np.uint64), np.zeros(nnz.value, ELEM)
    lib.ref_result_copy(h, _p(pos), _p(data))
    lib.ref_result_free(h)
    return pos, data, prod.value


def ref_compare(pos_a, data_a, pos_b, data_b) -> bool:
    """The reference's own compareCOO (SimSpGEMM.cpp:283-297; NB its `abs` truncates to int)."""
    pos_a, data_a, pos_b, data_b = _c(pos_a, np.uint64), _c(data_a, ELEM), _c(pos_b, np.uint64), _c(data_b, ELEM)
    return bool(ref_libs()[0].ref_compare(len(pos_a) - 1, _p(pos_a), _p(data_a), len(pos_b) - 1, _p(pos_b), _p(data_b)))


def ref_taskprovider(a_csc_pos, a_csc_data, b_pos, b_data):
    """The reference's as-written TaskProvider -> dict(pos, data, mult_sizes[n,2], merge_ways[rows,2], seconds)."""
    a_pos, a_data = _c(a_csc_pos, np.uint64), _c(a_csc_data, ELEM)
    b_pos, b_data = _c(b_pos, np.uint64), _c(b_data, ELEM)
    lib = ref_libs()[1]
    h = lib.ref_taskprovider(len(a_pos) - 1, _p(a_pos), _p(a_data), _p(b_pos), _p(b_data))
    rows, nnz, nm, ng, sec = _u64(), _u64(), _u64(), _u64(), C.c_double()
    lib.ref_tp_dims(h, C.byref(rows), C.byref(nnz), C.byref(nm), C.byref(ng), C.byref(sec))
    pos, data = np.zeros(rows.value + 1, np.uint64), np.zeros(nnz.value, ELEM)
    mult, merge = np.zeros((nm.value, 2), np.uint32), np.zeros((ng.value, 2), np.uint32)
    lib.ref_tp_copy(h, _p(pos), _p(data), _p(mult), _p(merge))
    lib.ref_tp_free(h)
    return dict(pos=pos, data=data, mult_sizes=mult, merge_ways=merge, seconds=sec.value)
