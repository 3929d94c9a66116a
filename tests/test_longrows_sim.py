"""k_long_count / k_long_fill (outerspace_b200/csrc/osp_longrows.cuh) on the CPU emulation of the CUDA
execution model (tests/cusim), bit for bit against the oracle.

The very source nvcc compiles for sm_100a is compiled here by g++ with OSP_CUSIM defined; the threads of a
block run as fibers that switch at the synchronising built-ins, which is one legal CUDA schedule.  This
checks the kernel's LOGIC (band index, run groups, band hand-over, arbitration order = the reference's left
fold in ascending k, SimSpGEMM.cpp:265-281 + :519-535) where no GPU exists; it cannot see data races, so
the GPU parity tests (tests/test_gpu_parity.py) stay the gate for the product.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from helpers import assert_bit_exact, oracle_spgemm, operands, pack, rand_sparse
from outerspace_b200.formats import ELEM

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "outerspace_b200", "csrc")

CONFIGS = {0: (64, 64, 8), 1: (128, 256, 32), 2: (32, 32, 64), 3: (96, 1024, 5)}   # threads, band, runs per group


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cusim") / "liblongrows_sim.so")
    src = os.path.join(HERE, "cusim", "longrows_sim.cpp")
    # CUSIM_ASAN=1 (with LD_PRELOAD=$(g++ -print-file-name=libasan.so) for the interpreter): AddressSanitizer build
    asan = ["-g", "-fsanitize=address", "-fno-omit-frame-pointer"] if os.environ.get("CUSIM_ASAN") == "1" else []
    cmd = ["g++", "-O1"] + asan + ["-std=c++17", "-x", "c++", "-fPIC", "-shared", "-ffp-contract=off", "-Wall", "-Wno-unknown-pragmas",
           "-I", os.path.join(HERE, "cusim"), "-I", CSRC, "-o", out, src]
    subprocess.run(cmd, check=True)
    lib = C.CDLL(out)
    vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
    lib.lr_sim.restype = C.c_int
    lib.lr_sim.argtypes = [C.c_int, vp, vp, vp, vp, u64, u64, vp, u32, C.c_uint, vp, vp, vp, u64]
    lib.lr_sim_bins.restype = C.c_int
    lib.lr_sim_bins.argtypes = [C.c_int, vp, vp, vp, vp, u64, u64, vp, u32, vp, u64, vp, vp, u64, u64, u64, vp, C.c_uint]
    return lib


def run_sim(lib, config, a_csr, b_csr, cols, rows, grid=2):
    rows = np.ascontiguousarray(rows, np.uint32)
    n = len(rows)
    count = np.full(n, 0xFFFFFFFF, np.uint32)
    off = np.zeros(n + 1, np.uint64)
    b_len = np.diff(b_csr.pos.astype(np.int64))
    bound = int(sum(int(b_len[a_csr.data["idx"][int(a_csr.pos[r]):int(a_csr.pos[r + 1])]].sum()) for r in rows)) + 1
    out = np.zeros(bound, ELEM)
    out["idx"] = 0xABABABAB
    ptr = lambda x: x.ctypes.data_as(C.c_void_p)
    rc = lib.lr_sim(config, ptr(a_csr.pos), ptr(a_csr.data), ptr(b_csr.pos), ptr(b_csr.data), b_csr.NRow(), cols, ptr(rows), n, grid,
                    ptr(count), ptr(off), ptr(out), bound)
    assert rc == 0, rc
    return count, off, out


def check(lib, config, A, B, rows=None, grid=2):
    a_csc, a_csr, b_csr = operands(A, B)
    want, _ = oracle_spgemm(a_csc, b_csr)
    if rows is None:
        rows = np.arange(A.shape[0], dtype=np.uint32)
    count, off, out = run_sim(lib, config, a_csr, b_csr, B.shape[1], rows, grid)
    wpos = want.pos.astype(np.int64)
    want_len = (wpos[1:] - wpos[:-1])[rows]
    assert np.array_equal(count.astype(np.int64), want_len), "k_long_count differs from the oracle's row lengths"
    sel = np.concatenate([np.arange(wpos[r], wpos[r + 1]) for r in rows]) if len(rows) else np.zeros(0, np.int64)
    got = pack(off, out[: int(off[-1])])
    exp = pack(np.concatenate([[0], np.cumsum(want_len)]).astype(np.uint64), want.data[sel])
    assert_bit_exact(got, exp, f"config {config}")


@pytest.mark.parametrize("config", sorted(CONFIGS))
def test_random_operands(sim, config):
    rng = np.random.default_rng(100 + config)
    A = rand_sparse(rng, 12, 40, 0.5)
    B = rand_sparse(rng, 40, 300, 0.25)
    check(sim, config, A, B)


@pytest.mark.parametrize("config", sorted(CONFIGS))
def test_every_column_hit_by_every_run(sim, config):
    """Dense operands: every chunk is full of same-column products, the arbitration decides the order."""
    rng = np.random.default_rng(7)
    A = sp.csr_matrix((rng.standard_normal((3, 70)) * 1e3).astype(np.float32))
    B = sp.csr_matrix((rng.standard_normal((70, 150)) * 1e-3).astype(np.float32))
    check(sim, config, A, B)


def test_cancellation_keeps_explicit_zero(sim):
    """x + (-x) leaves an explicit 0.0 in the row, like the reference's dedup (SimSpGEMM.cpp:519-535)."""
    A = sp.csr_matrix(np.array([[1.0, -1.0, 0.0], [2.0, 0.0, 3.0]], np.float32))
    B = sp.csr_matrix(np.array([[5.0, 0.0, 7.0, 0.0], [5.0, 0.0, 0.0, 1.0], [0.0, 0.0, 0.0, 0.0]], np.float32))
    for config in CONFIGS:
        check(sim, config, A, B)


def test_negative_zero_product(sim):
    """acc starts at -0.0: a single product of -0.0 stays -0.0, (-0.0) + (+0.0) = +0.0, as the left fold gives."""
    A = sp.csr_matrix(np.array([[-1.0, 1.0]], np.float32))
    # B holds explicit zeros at column 0 of both rows: products -0.0 (k=0) then +0.0 (k=1)
    Bz = sp.csr_matrix((np.array([0.0, 1.0, 0.0, 2.0], np.float32), np.array([0, 1, 0, 2]), np.array([0, 2, 4])), shape=(2, 3))
    for config in CONFIGS:
        check(sim, config, A, Bz)


@pytest.mark.parametrize("config", [0, 2])
def test_subset_of_rows_in_any_order_and_empty_rows(sim, config):
    rng = np.random.default_rng(5)
    A = rand_sparse(rng, 20, 30, 0.3).tolil()
    A[4, :] = 0                                   # an empty row of A
    A = sp.csr_matrix(A)
    A.eliminate_zeros()
    B = rand_sparse(rng, 30, 500, 0.1).tolil()
    B[3, :] = 0                                   # an empty row of B (a run of length 0)
    B = sp.csr_matrix(B)
    B.eliminate_zeros()
    check(sim, config, A, B, rows=np.array([17, 4, 0, 9, 19], np.uint32), grid=3)
    check(sim, config, A, B, rows=np.zeros(0, np.uint32), grid=1)


def test_columns_not_a_multiple_of_the_band_and_last_band_single_column(sim):
    rng = np.random.default_rng(11)
    for cols in (1, 31, 33, 64, 65, 257):
        A = rand_sparse(rng, 5, 9, 0.7)
        B = rand_sparse(rng, 9, cols, 0.6)
        for config in (0, 2):
            check(sim, config, A, B)


def test_long_runs(sim):
    """Runs much longer than a band: hundreds of elements per (run, band) segment, several chunks per group."""
    rng = np.random.default_rng(13)
    A = rand_sparse(rng, 2, 6, 0.9)
    B = rand_sparse(rng, 6, 5000, 0.6)
    check(sim, 3, A, B)
    check(sim, 1, A, B)


@pytest.mark.parametrize("config", [0, 1])
def test_engine_hand_over_into_the_bins(sim, config):
    """LongRowsInBins + k_mark_swept: listed rows of a row block with at least min_len partial products are merged
    into the start of their own bins (uniq[row] = nnz) and their tasks are flagged for the multiply; every other
    bin, uniq entry and task bit stays untouched."""
    rng = np.random.default_rng(21)
    A = rand_sparse(rng, 16, 24, 0.4)
    B = rand_sparse(rng, 24, 400, 0.3)
    a_csc, a_csr, b_csr = operands(A, B)
    want, _ = oracle_spgemm(a_csc, b_csr)
    b_len = np.diff(b_csr.pos.astype(np.int64))
    a_pos = a_csr.pos.astype(np.int64)
    row_len = np.array([int(b_len[a_csr.data["idx"][a_pos[r]:a_pos[r + 1]]].sum()) for r in range(16)], np.int64)
    row_bin = np.concatenate([[0], np.cumsum(row_len)]).astype(np.uint64)
    row_lo, row_hi, min_len = 3, 13, int(np.median(row_len))
    bin_base = int(row_bin[row_lo])
    xl_list = np.array([12, 1, 5, 9, 3, 14, 7], np.uint32)           # unordered, some outside the block
    bins = np.zeros(int(row_bin[row_hi]) - bin_base, ELEM)
    bins["idx"] = 0xABABABAB
    uniq = np.full(16, 0xEEEEEEEE, np.uint32)
    swept = np.zeros((a_csr.nnz + 31) // 32 + 1, np.uint32)
    ptr = lambda x: x.ctypes.data_as(C.c_void_p)
    rc = sim.lr_sim_bins(config, ptr(a_csr.pos), ptr(a_csr.data), ptr(b_csr.pos), ptr(b_csr.data), 24, 400, ptr(xl_list), len(xl_list),
                         ptr(row_bin), bin_base, ptr(bins), ptr(uniq), row_lo, row_hi, min_len, ptr(swept), 2)
    assert rc == 0
    wpos = want.pos.astype(np.int64)
    marked = np.unpackbits(swept.view(np.uint8), bitorder="little")[: a_csr.nnz].astype(bool)
    for r in range(16):
        listed_long = r in xl_list and row_len[r] >= min_len
        assert marked[a_pos[r]:a_pos[r + 1]].all() == listed_long or a_pos[r] == a_pos[r + 1]
        assert marked[a_pos[r]:a_pos[r + 1]].any() == (listed_long and a_pos[r] < a_pos[r + 1])
        taken = listed_long and row_lo <= r < row_hi
        if row_lo <= r < row_hi:
            b = bins[int(row_bin[r]) - bin_base: int(row_bin[r + 1]) - bin_base]
            if taken:
                n = wpos[r + 1] - wpos[r]
                assert uniq[r] == n
                assert np.array_equal(b[:n]["idx"], want.data["idx"][wpos[r]:wpos[r + 1]])
                assert np.array_equal(b[:n]["val"].view(np.uint32), want.data["val"][wpos[r]:wpos[r + 1]].view(np.uint32))
                assert np.all(b[n:]["idx"] == 0xABABABAB)
            else:
                assert np.all(b["idx"] == 0xABABABAB)
        if not taken:
            assert uniq[r] == 0xEEEEEEEE
