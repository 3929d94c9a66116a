"""GPU parity of the OPT-IN fused short rows (OSP_FUSED_SHORT: k_merge_chain_fused computes the tiles of short rows
straight into its stage, no bins for them).  Like the long-row sweep it was written after the round's GPU budget was
spent: bit-exact on the CPU emulation (tests/test_engine_sim.py, tools/fuzz_engine_sim.py), NOT yet run on a B200, off
by default.  Runs only with OSP_TEST_FUSED_SHORT=1:

    OSP_TEST_FUSED_SHORT=1 python -m pytest tests/test_gpu_zzz_fused_short.py -m gpu -x -q
    OSP_FUSED_SHORT=1 python -m pytest tests -m gpu -x -q            # the whole suite through the fused chain
"""
import os

import pytest

import outerspace_b200 as osp
import test_gpu_parity as gp

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("OSP_TEST_FUSED_SHORT") != "1",
                                 reason="opt-in path not yet verified on a B200: run with OSP_TEST_FUSED_SHORT=1"),
              pytest.mark.timeout(300)]


@pytest.fixture(scope="module")
def engine():
    os.environ["OSP_FUSED_SHORT"] = "1"                   # read at osp_create
    try:
        eng = osp.Engine(0)
    finally:
        del os.environ["OSP_FUSED_SHORT"]
    yield eng
    eng.close()


test_golden = gp.test_golden
test_random_vs_oracle = gp.test_random_vs_oracle
test_er_config2_scaled = gp.test_er_config2_scaled
test_rmat_small = gp.test_rmat_small
test_edge_cases = gp.test_edge_cases
test_every_row_length_class = gp.test_every_row_length_class
test_many_tiles_chain_order = gp.test_many_tiles_chain_order
test_config2_full_size_bit_exact = gp.test_config2_full_size_bit_exact
test_config4_shape_properties = gp.test_config4_shape_properties
