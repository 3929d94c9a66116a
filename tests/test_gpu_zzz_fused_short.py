"""GPU parity of the OPT-IN fused short rows (OSP_FUSED_SHORT: k_merge_chain_fused computes the tiles of short rows
straight into its stage, no bins for them).  Round 2: bit-exact on a B200 (profiles/r02_optin_paths.md); runs with
the rest of the GPU suite.

    OSP_FUSED_SHORT=1 python -m pytest tests -m gpu -x -q            # the whole suite through the fused chain
"""
import os

import pytest

import outerspace_b200 as osp
import test_gpu_parity as gp

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


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
