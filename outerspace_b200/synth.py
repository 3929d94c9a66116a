"""Synthetic operands of the named BASELINE.json configs (SURVEY.md 8d "Synthetic inputs").

All generators are seeded, remove duplicate (row, col) draws, never emit an explicit zero and
return row-compressed operands sorted by (row, col) -- what ``coo2csr`` would build from the same
triplets (simulator/SimSpGEMM.cpp:102-152).  The generator is numpy's PCG64, not the mt19937_64
the survey's probe used, so counts differ from SURVEY.md Appendix A by sampling noise only.
"""
from __future__ import annotations

import numpy as np

from .formats import CSRMatrix


def _values(rng: np.random.Generator, n: int, nonneg: bool = False) -> np.ndarray:
    v = rng.random(n, dtype=np.float32)
    if nonneg:
        v = v + np.float32(2.0 ** -20)          # (0, 1]
    else:
        v = v * np.float32(2.0) - np.float32(1.0)
        v[v == 0] = np.float32(0.5)             # never an explicit zero
    return v.astype(np.float32)


def _csr_from_keys(keys: np.ndarray, n_rows: int, n_cols: int, rng, nonneg=False) -> CSRMatrix:
    keys = np.unique(keys)                       # sorted by (row, col), duplicates removed
    rows = (keys // np.uint64(n_cols)).astype(np.int64)
    cols = (keys % np.uint64(n_cols)).astype(np.uint32)
    pos = np.zeros(n_rows + 1, dtype=np.uint64)
    np.cumsum(np.bincount(rows, minlength=n_rows), out=pos[1:])
    return CSRMatrix.from_arrays(pos, cols, _values(rng, len(cols), nonneg))


def erdos_renyi(n_rows: int, n_cols: int, draws: int, seed: int, nonneg: bool = False) -> CSRMatrix:
    """`draws` uniform (i, j) samples with duplicates removed."""
    rng = np.random.default_rng(seed)
    keys = rng.integers(0, np.uint64(n_rows) * np.uint64(n_cols), size=draws, dtype=np.uint64)
    return _csr_from_keys(keys, n_rows, n_cols, rng, nonneg)


def rmat(scale: int, edge_factor: int, seed: int, a=0.57, b=0.19, c=0.19, d=0.05) -> CSRMatrix:
    """R-MAT (Graph500 parameters by default), 2^scale vertices, edge_factor * 2^scale draws,
    duplicates removed, not symmetrised."""
    rng = np.random.default_rng(seed)
    n = 1 << scale
    m = edge_factor * n
    rows = np.zeros(m, dtype=np.uint64)
    cols = np.zeros(m, dtype=np.uint64)
    for _ in range(scale):
        r = rng.random(m)
        row_bit = r >= (a + b)                              # quadrants c, d
        col_bit = ((r >= a) & (r < a + b)) | (r >= a + b + c)   # quadrants b, d
        rows = (rows << np.uint64(1)) | row_bit.astype(np.uint64)
        cols = (cols << np.uint64(1)) | col_bit.astype(np.uint64)
    keys = rows * np.uint64(n) + cols
    return _csr_from_keys(keys, n, n, rng)


def pruned_dense(n_rows: int, n_cols: int, density: float, seed: int, nonneg: bool = False) -> CSRMatrix:
    """Magnitude-pruned dense layer: N(0,1) weights, all but the largest `density` fraction by |w|
    zeroed (what NN_models/main.py:191-238 does with a quantile threshold); `nonneg` gives a
    post-ReLU-like activation matrix instead."""
    rng = np.random.default_rng(seed)
    w = rng.standard_normal((n_rows, n_cols), dtype=np.float32)
    if nonneg:
        w = np.abs(w)
    k = int(round(n_rows * n_cols * density))
    if k <= 0:
        thr = np.inf
    else:
        thr = np.partition(np.abs(w).ravel(), w.size - k)[w.size - k]
    mask = np.abs(w) >= thr
    rows, cols = np.nonzero(mask)
    pos = np.zeros(n_rows + 1, dtype=np.uint64)
    np.cumsum(np.bincount(rows, minlength=n_rows), out=pos[1:])
    return CSRMatrix.from_arrays(pos, cols.astype(np.uint32), w[mask].astype(np.float32))


def transpose_host(m: CSRMatrix, n_minor: int) -> CSRMatrix:
    """Host CSR<->CSC (stable), for building inputs only."""
    order = np.argsort(m.data["idx"], kind="stable")
    major = np.repeat(np.arange(m.NRow(), dtype=np.uint32), np.diff(m.pos.astype(np.int64)))
    pos = np.zeros(n_minor + 1, dtype=np.uint64)
    np.cumsum(np.bincount(m.data["idx"], minlength=n_minor), out=pos[1:])
    return CSRMatrix.from_arrays(pos, major[order], m.data["val"][order])


# ---- the named configs ------------------------------------------------------------------------
WORKLOADS = {
    # name: (description, builder) ; builder() -> dict(a_csr, b_csr, rows, n_k, cols)
    "mlp_fc2": "config 1: pruned MLP fc2 weight 1000x1000 at 1% density, C = A*A",
    "er16k": "config 2: Erdos-Renyi 16384x16384 density 1e-3, C = A*A",
    "rmat20": "config 3: R-MAT scale 20, edge factor 16, C = A*A",
    "er8m": "config 4: Erdos-Renyi 2^23 x 2^23, 8 nnz/row, C = A*A",
    "mlp_batch": "config 5: activation 65536x4096 (10%) times weight^T 4096x4096 (10%)",
}


def build_workload(name: str, scale_down: int = 1):
    """Returns (A_csr, B_csr, dims) for a named workload; `scale_down` > 1 shrinks the big ones
    by that factor (rows / scale) for tests and the bounded CPU baseline."""
    if name == "mlp_fc2":
        a = pruned_dense(1000, 1000, 0.01, seed=43)
        return a, a, dict(rows=1000, n_k=1000, cols=1000)
    if name == "er16k":
        n = 16384 // scale_down
        a = erdos_renyi(n, n, int(round(n * n * 1e-3)), seed=44)
        return a, a, dict(rows=n, n_k=n, cols=n)
    if name == "rmat20":
        s = 20 - int(np.log2(scale_down))
        a = rmat(s, 16, seed=45)
        return a, a, dict(rows=1 << s, n_k=1 << s, cols=1 << s)
    if name == "er8m":
        n = (1 << 23) // scale_down
        a = erdos_renyi(n, n, 8 * n, seed=46)
        return a, a, dict(rows=n, n_k=n, cols=n)
    if name == "mlp_batch":
        batch = 65536 // scale_down
        x = pruned_dense(batch, 4096, 0.10, seed=47, nonneg=True)       # activations [batch x in]
        w = pruned_dense(4096, 4096, 0.10, seed=48)                      # weight [out x in]
        wt = transpose_host(w, 4096)                                     # B = W^T as CSR [in x out]
        return x, wt, dict(rows=batch, n_k=4096, cols=4096)
    raise KeyError(name)
