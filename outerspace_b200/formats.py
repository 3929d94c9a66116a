"""Operand containers that mirror the reference's structs (simulator/common.h).

* ``ELEM`` is the packed 8-byte ``CSRElement {index_t idx; value_t val}`` (common.h:10-16).
* :class:`CSRMatrix` is ``struct CSRMatrix {vector<size_t> pos; vector<CSRElement> data;}``
  (common.h:39-47) and, as in the reference, holds either CSR (``idx`` = column ids) or CSC
  (``idx`` = row ids).
* ``COO`` is ``COOElement {row, col, val}`` (common.h:18-33) as three parallel arrays.

The arrays are exactly the byte layouts the C ABI (include/osp_b200.h) takes, so no conversion
happens between Python and the library.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

ELEM = np.dtype([("idx", "<u4"), ("val", "<f4")])
assert ELEM.itemsize == 8


@dataclass
class COO:
    rows: np.ndarray  # uint32
    cols: np.ndarray  # uint32
    vals: np.ndarray  # float32

    def __len__(self) -> int:
        return int(self.rows.shape[0])


@dataclass
class CSRMatrix:
    pos: np.ndarray   # uint64[n_slices + 1]
    data: np.ndarray  # ELEM[nnz]

    def NRow(self) -> int:  # same name as the reference (common.h:46)
        return int(self.pos.shape[0]) - 1

    @property
    def nnz(self) -> int:
        return int(self.data.shape[0])

    @staticmethod
    def from_arrays(pos, idx, val) -> "CSRMatrix":
        data = np.empty(len(idx), dtype=ELEM)
        data["idx"] = idx
        data["val"] = val
        return CSRMatrix(np.ascontiguousarray(pos, dtype=np.uint64), data)

    @staticmethod
    def from_scipy(m) -> "CSRMatrix":
        """From a scipy.sparse csr_matrix/csc_matrix with sorted indices (compressed axis = slices)."""
        m = m.copy()
        m.sort_indices()
        return CSRMatrix.from_arrays(m.indptr.astype(np.uint64), m.indices.astype(np.uint32), m.data.astype(np.float32))

    def to_scipy_csr(self, n_minor: int):
        import scipy.sparse as sp

        return sp.csr_matrix(
            (self.data["val"].astype(np.float32), self.data["idx"].astype(np.int64), self.pos.astype(np.int64)),
            shape=(self.NRow(), n_minor),
        )

    def equal_structure(self, other: "CSRMatrix") -> bool:
        return (
            self.pos.shape == other.pos.shape
            and np.array_equal(self.pos, other.pos)
            and np.array_equal(self.data["idx"], other.data["idx"])
        )

    def equal_bits(self, other: "CSRMatrix") -> bool:
        """Bit-exact equality of structure and values (values compared as raw 32-bit patterns)."""
        return self.equal_structure(other) and np.array_equal(
            self.data["val"].view(np.uint32), other.data["val"].view(np.uint32)
        )
