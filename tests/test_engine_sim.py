"""The whole single-GPU engine on the CPU emulation of the CUDA execution model (tests/cusim), against the oracle.

tests/cusim/engine_sim.cpp compiles outerspace_b200/csrc/osp_engine.cu -- the C ABI, the host orchestration and every
kernel, the very sources nvcc builds for sm_100a -- with g++ against cusim.h (threads of a block = fibers, one legal
CUDA schedule) and cusim_runtime.h (device memory = host memory).  The parity tests of tests/test_gpu_parity.py that
are small enough run here UNCHANGED through the same Python binding, so host logic (planning, row blocks, capacity
admission, error paths) and kernel logic are checked bit for bit where no GPU exists; the opt-in long-row sweep
(OSP_LONGROW_SWEEP), which has not run on a B200 yet, is exercised end to end.

This is TEST INFRASTRUCTURE: the emulated library is built into a temporary directory and loaded by this module only.
The product (outerspace_b200/libosp_b200.so) has no CPU path, and an emulation cannot see data races or timing: the
`-m gpu` suite stays the gate.
"""
import os
import subprocess

import numpy as np
import pytest

import oracle
import outerspace_b200 as osp
from outerspace_b200 import api, synth
from helpers import assert_bit_exact, check_csr_invariants, operands, oracle_spgemm, pack, rand_sparse
import test_gpu_parity as gp

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "outerspace_b200", "csrc")


@pytest.fixture(scope="module", autouse=True)
def emulated_library(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cusim_engine") / "libosp_b200_cusim.so")
    cmd = ["g++", "-O1", "-std=c++17", "-x", "c++", "-fPIC", "-shared", "-ffp-contract=off", "-Wall", "-Wno-unknown-pragmas",
           "-Wno-unused-function", "-I", os.path.join(HERE, "cusim"), "-I", CSRC, "-o", out,
           os.path.join(HERE, "cusim", "engine_sim.cpp"), os.path.join(CSRC, "osp_host.cpp"), "-lpthread"]
    subprocess.run(cmd, check=True)
    saved = (api._LIB_PATH, api._lib)
    api._LIB_PATH, api._lib = out, None          # this module only; restored below
    try:
        lib = api.load_library()

        yield lib
    finally:
        api._LIB_PATH, api._lib = saved


@pytest.fixture(scope="module")
def engine(emulated_library):
    eng = osp.Engine(0)
    yield eng
    eng.close()


# ---- the GPU parity tests that are small enough, unchanged --------------------------------------------------------
test_golden = gp.test_golden
test_mtx_pipeline_like_reference_main = gp.test_mtx_pipeline_like_reference_main
test_random_vs_oracle = gp.test_random_vs_oracle
test_task_sizes_match_reference_structure = gp.test_task_sizes_match_reference_structure
test_row_chunking_gives_same_bits = gp.test_row_chunking_gives_same_bits
test_edge_cases = gp.test_edge_cases
test_error_codes = gp.test_error_codes
test_csr2csc_device_stable = gp.test_csr2csc_device_stable
test_csr2csc_random_and_duplicates = gp.test_csr2csc_random_and_duplicates
test_coo_ingest_on_device = gp.test_coo_ingest_on_device
