"""GPU parity of the OPT-IN fused band sweep of the long rows (OSP_LONGROW_SWEEP, outerspace_b200/csrc/osp_longrows.cuh).

Round 2: bit-exact on a B200 (profiles/r02_optin_paths.md) but not faster than k_multiply + k_merge_xl on config 3
(982 ms either way), so it stays opt-in; these tests run with the rest of the GPU suite so that every kernel in the
shipped library has a green hardware test.

    OSP_LONGROW_SWEEP=1 python -m pytest tests -m gpu -x -q        # the whole suite through the sweep
"""
import os

import numpy as np
import pytest

import oracle
import outerspace_b200 as osp
from outerspace_b200 import api, synth
from helpers import assert_bit_exact, check_csr_invariants, operands, oracle_spgemm, pack
from test_gpu_parity import _row_lengths_case

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]

SWEEP = api.OSP_LONGROW_SWEEP


def _kernels(res):
    return {name for name, _ in res.kernel_times()}


def test_rmat_small_through_the_sweep(engine):
    """config 3 shape at scale 12 (4096 columns <= 16384: the sweep must stay out) and at scale 15 (32768 columns: in)."""
    for sd, expect in ((256, False), (32, True)):
        a, b, dims = synth.build_workload("rmat20", scale_down=sd)
        a_csc = synth.transpose_host(a, dims["n_k"])
        want, prod = oracle_spgemm(a_csc, b)
        for is_csr, op in ((True, a), (False, a_csc)):
            res = engine.spgemm(op, b, a_is_csr=is_csr, cols_b=dims["cols"], flags=SWEEP | api.OSP_PROFILE_KERNELS)
            got = res.to_host(); st = res.stats(); names = _kernels(res); res.free()
            assert st["products"] == prod
            assert any("k_long_fill" in n for n in names) == expect, names
            assert_bit_exact(got, want, f"rmat20/{sd} sweep a_is_csr={is_csr}")


@pytest.mark.parametrize("cols,dup_rate", [(1 << 15, 0.0), (1 << 17, 0.3), (1 << 20, 0.0), (1 << 20, 0.3), ((1 << 24) + 5, 0.2)])
def test_every_row_length_class_through_the_sweep(engine, cols, dup_rate):
    """Short, medium and swept rows side by side in the same call; column ranges on both sides of the limit up to
    which k_merge_xl takes the medium rows (131072), bands that end inside the last word, 64-bit chain keys."""
    rng = np.random.default_rng(cols % 1000 + int(dup_rate * 10))
    big = 30000 if cols < (1 << 17) else 70000          # 2^24 columns = 1025 bands: only rows from 65 600 partial products are swept
    lens = [0, 1, 8, 33, 129, 512, 513, 700, 1500, 4096, 4097, 5000, 9000, big] * 3
    rng.shuffle(lens)
    A, B = _row_lengths_case(rng, lens, cols, dup_rate)
    a_csc, a_csr, b_csr = operands(A, B)
    want, prod = oracle_spgemm(a_csc, b_csr)
    res = engine.spgemm(a_csr, b_csr, a_is_csr=True, cols_b=cols, flags=SWEEP | api.OSP_PROFILE_KERNELS)
    got = res.to_host(); names = _kernels(res); res.free()
    assert any("k_long_fill" in n for n in names), names
    assert_bit_exact(got, want, f"sweep cols={cols} dup={dup_rate}")
    check_csr_invariants(got, cols)
    res = engine.spgemm(a_csr, b_csr, a_is_csr=True, cols_b=cols, flags=SWEEP | api.OSP_KSLICE_ORDER | api.OSP_PROFILE_KERNELS)
    got = res.to_host(); names = _kernels(res); res.free()
    assert not any("k_long_fill" in n for n in names), "the k-slice order keeps the bins for every row"
    assert_bit_exact(got, want, "k-slice order ignores the sweep flag")


def test_sweep_with_row_blocks_and_bounded_capacity(engine):
    """Row blocks (workspace limit) and C below the plan's bound: swept rows are filtered per block."""
    a, b, dims = synth.build_workload("rmat20", scale_down=32)
    want_pos, want_data, prod = oracle.spgemm_rowblocks(a.pos, a.data, b.pos, b.data, 4096)
    want = pack(want_pos, want_data)
    cs = np.concatenate([[0], np.cumsum(np.diff(b.pos.astype(np.int64))[a.data["idx"]])])
    plen = cs[a.pos[1:].astype(np.int64)] - cs[a.pos[:-1].astype(np.int64)]
    block = max(int(plen.max()), prod // 16)
    eng = osp.Engine(0)
    try:
        eng.set_workspace_limit(block * 8)
        eng.set_result_limit((len(want_data) + block + 64) * 8)
        res = eng.spgemm(a, b, a_is_csr=True, cols_b=dims["cols"], flags=SWEEP)
        got = res.to_host(); st = res.stats(); res.free()
        assert st["row_chunks"] > 4 and st["nnz_c"] == len(want_data)
        assert_bit_exact(got, want, "sweep in row blocks")
    finally:
        eng.close()


def test_sweep_threshold_from_the_environment():
    """OSP_LONGROW_SWEEP_MIN: rows below the threshold stay with k_multiply + k_merge_xl, rows above are swept."""
    rng = np.random.default_rng(3)
    lens = [100, 5000, 6000, 30000, 50000, 700] * 2
    A, B = _row_lengths_case(rng, lens, 1 << 18, 0.2)
    a_csc, a_csr, b_csr = operands(A, B)
    want, _ = oracle_spgemm(a_csc, b_csr)
    os.environ["OSP_LONGROW_SWEEP"] = "1"
    os.environ["OSP_LONGROW_SWEEP_MIN"] = "20000"
    try:
        eng = osp.Engine(0)
    finally:
        del os.environ["OSP_LONGROW_SWEEP"], os.environ["OSP_LONGROW_SWEEP_MIN"]
    try:
        res = eng.spgemm(a_csr, b_csr, a_is_csr=True, cols_b=1 << 18, flags=api.OSP_PROFILE_KERNELS)
        got = res.to_host(); names = _kernels(res); res.free()
        assert any("k_long_fill" in n for n in names) and any("k_merge_xl" in n for n in names), names
        assert_bit_exact(got, want, "sweep threshold 20000")
    finally:
        eng.close()
