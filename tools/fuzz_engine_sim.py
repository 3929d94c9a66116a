#!/usr/bin/env python3
"""Fuzz campaign: the emulated engine (tests/cusim, the product's sources on the CPU emulation of the CUDA execution
model) against the oracle on random operands of every shape class -- a campaign the GPU budget never has room for.

    python tools/fuzz_engine_sim.py --cases 300 --seed 1 [--schedule random:7] [--resident 4] [--asan]

Every case draws: dimensions, densities, a column range (small: bitmap/dense kernels; medium; > 2^23: 64-bit chain
keys), a few rows made long on purpose (medium and xl rows), duplicates-free operands in the reference layout, CSR or
CSC hand-over of A, multiply order, fused-dense on/off, the opt-in long-row sweep and fused short rows, and sometimes a workspace / result
limit that forces row blocks.  The result must match the oracle bit for bit.  TEST INFRASTRUCTURE ONLY.
"""
import argparse
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import scipy.sparse as sp


def build(out, asan=False):
    csrc = os.path.join(ROOT, "outerspace_b200", "csrc")
    sim = os.path.join(ROOT, "tests", "cusim")
    subprocess.run(["g++", "-O1"] + (["-g", "-fsanitize=address", "-fno-omit-frame-pointer"] if asan else []) + ["-std=c++17", "-x", "c++", "-fPIC", "-shared", "-ffp-contract=off", "-w", "-I", sim, "-I", csrc, "-o", out,
                    os.path.join(sim, "engine_sim.cpp"), os.path.join(csrc, "osp_host.cpp"), "-lpthread"], check=True)


def draw_wide_case(rng):
    """A few rows of A with hundreds to > 1000 non-zeros (more runs than one sweep group of 512) over a wide column range."""
    k = int(rng.integers(600, 1500))
    cols = int(rng.choice([rng.integers(16385, 70000), rng.integers(70000, 300000)]))
    m = int(rng.integers(2, 8))
    A = sp.lil_matrix((m, k), dtype=np.float32)
    for r in range(m):
        n = int(rng.choice([0, rng.integers(1, 60), rng.integers(400, 520), rng.integers(513, k)]))
        sel = rng.choice(k, size=n, replace=False)
        A[r, sel] = (rng.standard_normal(n) + 2).astype(np.float32)
    B = sp.random(k, cols, density=float(rng.choice([4, 12, 30])) / cols, format="csr", random_state=int(rng.integers(1 << 31)), dtype=np.float32,
                  data_rvs=lambda n: (rng.standard_normal(n) * 2 + 0.1).astype(np.float32))
    A = sp.csr_matrix(A)
    A.eliminate_zeros()
    B.eliminate_zeros()
    return A, B, cols


DENSE_SHARE = 0.15


def draw_dense_case(rng):
    """Long rows over a small column range: the fused dense paths (bank-aligned k_fused_lanes, band kernel k_fused_dense).
    Rows of B pile up in few shared-memory banks now and then (runs of more than one piece), some are empty, some rows of A
    are empty; explicit +-0.0 values."""
    cols = int(rng.choice([rng.integers(1, 33), rng.integers(33, 1800), rng.integers(1800, 4097), rng.integers(4097, 8161), rng.integers(8161, 16385)]))
    k = int(rng.integers(8, 120))
    m = int(rng.integers(1, 40))
    b_rows, b_cols = [], []
    per_row = min(cols, int(rng.choice([40, 120, 400])))
    for r in range(k):
        kind = rng.random()
        if kind < 0.08:
            c = np.zeros(0, np.int64)
        elif kind < 0.2 and cols >= 64:
            bank = int(rng.integers(0, 32))                      # many columns of one bank: more groups than a piece holds
            c = np.unique(rng.integers(0, (cols - bank + 31) // 32, size=int(rng.integers(20, 70)))) * 32 + bank
            c = c[c < cols]
        else:
            c = np.nonzero(rng.random(cols) < per_row / cols)[0]
        b_rows += [r] * len(c)
        b_cols += c.tolist()
    vals = (rng.standard_normal(len(b_rows)) * 2).astype(np.float32)
    vals[::17] = 0.0
    vals[5::29] = -0.0
    B = sp.csr_matrix((vals, (b_rows, b_cols)), shape=(k, cols))
    dens = float(rng.choice([0.3, 0.7, 1.0]))
    mask = rng.random((m, k)) < dens
    mask[rng.random(m) < 0.1] = False                          # empty rows of A
    av = (rng.standard_normal((m, k)) * 3).astype(np.float32)
    av[av == 0] = 1
    A = sp.csr_matrix(np.where(mask, av, 0).astype(np.float32))
    return A, B, cols


def draw_case(rng):
    if rng.random() < DENSE_SHARE:
        return draw_dense_case(rng)
    if rng.random() < 0.1:
        return draw_wide_case(rng)
    cols = int(rng.choice([rng.integers(1, 64), rng.integers(64, 4096), rng.integers(4096, 16385), rng.integers(16385, 200000),
                           rng.integers(200000, 1 << 21), (1 << 23) + int(rng.integers(1, 1 << 22))]))
    m = int(rng.integers(1, 120))
    k = int(rng.integers(1, 160))
    da = float(rng.choice([0.02, 0.1, 0.4]))
    A = sp.random(m, k, density=da, format="lil", random_state=int(rng.integers(1 << 31)), dtype=np.float32,
                  data_rvs=lambda n: (rng.standard_normal(n) * 3 + 0.1).astype(np.float32))
    # rows of B: mostly short, a few long
    b_rows, b_cols = [], []
    for r in range(k):
        kind = rng.random()
        n = int(rng.integers(0, 6)) if kind < 0.6 else int(rng.integers(6, 200)) if kind < 0.92 else int(rng.integers(200, 3000))
        n = min(n, cols)
        c = np.unique(rng.integers(0, cols, size=n))
        b_rows += [r] * len(c)
        b_cols += c.tolist()
    vals = (rng.standard_normal(len(b_rows)) * 2).astype(np.float32)
    vals[vals == 0] = 1
    B = sp.csr_matrix((vals, (b_rows, b_cols)), shape=(k, cols))
    # a few output rows made long on purpose: dense rows of A
    for _ in range(int(rng.integers(0, 3))):
        r = int(rng.integers(0, m))
        sel = rng.random(k) < float(rng.choice([0.3, 0.9]))
        A[r, np.nonzero(sel)[0]] = (rng.standard_normal(int(sel.sum())) + 2).astype(np.float32)
    A = sp.csr_matrix(A)
    A.eliminate_zeros()
    return A, B, cols


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=100)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--dense-share", type=float, default=0.15, help="share of cases drawn for the fused dense paths (small column range, long rows)")
    ap.add_argument("--schedule", default="")
    ap.add_argument("--asan", action="store_true", help="AddressSanitizer build: every access of a kernel or of the host code outside an "
                    "allocation of the emulated device aborts the run (re-executes itself with libasan preloaded)")
    ap.add_argument("--resident", type=int, default=1, help="blocks resident at a time (CUSIM_RESIDENT = CUSIM_SMS): > 1 runs them on OS threads")
    args = ap.parse_args()
    global DENSE_SHARE
    DENSE_SHARE = args.dense_share
    if args.asan and "libasan" not in os.environ.get("LD_PRELOAD", ""):
        asan = subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True, check=True).stdout.strip()
        env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1")
        os.execve(sys.executable, [sys.executable] + sys.argv, env)
    if args.schedule:
        os.environ["CUSIM_SCHEDULE"] = args.schedule
    if args.resident > 1:
        os.environ["CUSIM_RESIDENT"] = os.environ["CUSIM_SMS"] = str(args.resident)
    tmp = tempfile.mkdtemp(prefix="osp_fuzz_")
    lib = os.path.join(tmp, "libosp_b200_cusim.so")
    build(lib, args.asan)
    import outerspace_b200 as osp
    from outerspace_b200 import api
    from helpers import assert_bit_exact, operands, oracle_spgemm
    api._LIB_PATH, api._lib = lib, None
    rng = np.random.default_rng(args.seed)
    t0 = time.time()
    seen = {"sweep": 0, "blocks": 0, "fused": 0, "xl": 0, "long": 0, "products": 0, "fused_short": 0, "lanes": 0, "compact": 0}
    for case in range(args.cases):
        A, B, cols = draw_case(rng)
        a_csc, a_csr, b_csr = operands(A, B)
        want, prod = oracle_spgemm(a_csc, b_csr, rows_override=A.shape[0])
        flags = int(rng.choice([0, api.OSP_KSLICE_ORDER, api.OSP_ROWWISE_ORDER, api.OSP_LONGROW_SWEEP, api.OSP_LONGROW_SWEEP, api.OSP_NO_FUSED_DENSE,
                                api.OSP_FUSED_SHORT, api.OSP_FUSED_SHORT | api.OSP_LONGROW_SWEEP, api.OSP_FUSED_SHORT | api.OSP_NO_FUSED_DENSE]))
        as_csr = bool(rng.integers(0, 2))
        os.environ["OSP_FUSED_LANES"] = str(rng.choice(["1", "2", "2"]))      # read at osp_create: automatic / bank-aligned whatever B's regrouped size
        os.environ["OSP_FL_DIRECT"] = str(rng.choice(["0", "1", "1"]))        # rows of C chained by the look-back / at the prefix of their bounds
        eng = osp.Engine(0)
        try:
            if rng.random() < 0.3 and prod > 64:
                eng.set_workspace_limit(max(int(prod // rng.integers(2, 9)), 512) * 8)
                if rng.random() < 0.5:
                    eng.set_result_limit((want.nnz + max(int(prod // 2), 64)) * 8)
            try:
                res = eng.spgemm(a_csr if as_csr else a_csc, b_csr, a_is_csr=as_csr, cols_b=cols if rng.random() < 0.8 else 0,
                                 rows_c=A.shape[0], flags=flags | api.OSP_PROFILE_KERNELS)
            except osp.OspError as e:
                if e.code == api.OSP_ERR_OOM and "does not fit" in str(e):      # a legal refusal of a too small result limit
                    continue
                raise
            got = res.to_host(); st = res.stats(); names = {n for n, _ in res.kernel_times()}; res.free()
            what = f"case {case} seed {args.seed}: A {A.shape} nnz {A.nnz}, B nnz {B.nnz}, cols {cols}, flags {flags}, csr {as_csr}"
            assert st["products"] == prod, what
            assert_bit_exact(got, want, what)
            seen["sweep"] += any("k_long_fill" in n for n in names)
            seen["fused"] += any("k_fused_dense" in n for n in names)
            seen["lanes"] += any("k_fused_lanes" in n for n in names)
            seen["compact"] += any("k_fl_compact" in n for n in names)
            seen["fused_short"] += any("k_merge_chain_fused" in n for n in names)
            seen["blocks"] += st["row_chunks"] > 1
            seen["xl"] += st["rows_long"] > 0
            seen["long"] += st["rows_medium"] > 0
            seen["products"] += prod
        finally:
            eng.close()
    print(f"fuzz ok: {args.cases} cases, seed {args.seed}, schedule '{args.schedule or 'forward'}', {args.resident} resident block(s){', AddressSanitizer' if args.asan else ''}, {time.time() - t0:.0f} s; "
          f"calls with sweep {seen['sweep']}, fused short rows {seen['fused_short']}, fused dense (band kernel) {seen['fused']}, bank-aligned {seen['lanes']} (of which with compaction {seen['compact']}), row blocks {seen['blocks']}, xl rows {seen['xl']}, "
          f"medium rows {seen['long']}; {seen['products']} partial products in total")


if __name__ == "__main__":
    main()
