"""A named BASELINE config at its FULL size on one GPU: timing + size-independent parity.

The element-wise oracle cannot finish configs 3 and 5 at full size (P = 2.1e10 / 1.1e10 partial products), so the
check is per row -- rows of C are independent of each other: a seeded sample of rows plus the heaviest rows are
recomputed by the oracle from A[rows, :] and B and compared bit for bit (row_ptr differences, col_idx, values), next
to the global invariants (P equal to the host's count from the two pos arrays, row_ptr monotone and ending at nnz(C),
row lengths within min(partials of the row, cols), a row is empty exactly when it has no partial product).

`run_check` is what tests/test_gpu_zz_fullsize.py calls; `main` prints the same as a log (profiles/r01_v8_full_*.log).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch  # noqa: E402

import oracle  # noqa: E402  (checker only)
import outerspace_b200 as osp  # noqa: E402
from outerspace_b200 import api, synth  # noqa: E402


def host_row_partials(a, b):
    """Partial products per output row, from the two pos arrays and A's indices alone."""
    a_pos, b_len = a.pos.astype(np.int64), np.diff(b.pos.astype(np.int64))
    cs = np.concatenate([[0], np.cumsum(b_len[a.data["idx"]])])
    return cs[a_pos[1:]] - cs[a_pos[:-1]]


def check_result(res, a, b, dims, plen, sample_rows, heavy_rows, log=print):
    """(global invariants hold, rows that differ from the oracle, rows compared, partial products compared)."""
    a_pos = a.pos.astype(np.int64)
    st = res.stats()
    P = int(plen.sum())
    ok = st["products"] == P
    pos = res.pos_to_host().astype(np.int64)
    rl = np.diff(pos)
    n = len(rl)
    ok &= bool(pos[0] == 0 and pos[-1] == res.nnz and np.all(rl >= 0))
    ok &= bool(np.all(rl <= np.minimum(plen[:n], dims["cols"]))) and bool(np.all((rl > 0) == (plen[:n] > 0)))
    ok &= bool(np.all(plen[n:] == 0))                                   # rows beyond max row id + 1 hold nothing
    log("global invariants:", "OK" if ok else "VIOLATED", f"(rows {n}, nnz(C) {res.nnz}, P {st['products']} vs host {P})")
    rng = np.random.default_rng(7)
    pick = np.unique(np.concatenate([rng.integers(0, n, size=sample_rows), np.argsort(plen[:n])[-heavy_rows:]])).astype(np.int64)
    sub_pos = np.zeros(len(pick) + 1, np.uint64)
    np.cumsum(a_pos[pick + 1] - a_pos[pick], out=sub_pos[1:])
    sub_data = np.concatenate([a.data[a_pos[r]:a_pos[r + 1]] for r in pick])
    t1 = time.time()
    w_pos, w_data, w_prod = oracle.spgemm_rowblocks(sub_pos, sub_data, b.pos, b.data, 64)
    w_pos = np.concatenate([w_pos, np.full(len(pick) + 1 - len(w_pos), w_pos[-1], np.uint64)]).astype(np.int64)   # trailing empty rows
    bad = 0
    for j, r in enumerate(pick):
        got = res.rows_to_host(int(r), int(r) + 1)
        want = w_data[w_pos[j]:w_pos[j + 1]]
        if len(got.data) != len(want) or not np.array_equal(got.data.view(np.uint64), want.view(np.uint64)):
            bad += 1
            if bad <= 3:
                log(f"  row {r}: got {len(got.data)} entries, oracle {len(want)}")
    log(f"sampled rows vs oracle: {'BIT-EXACT' if bad == 0 else f'MISMATCH in {bad} rows'} "
        f"({len(pick)} rows, {int(w_prod)} partial products, heaviest {int(plen[pick].max())}; oracle {time.time() - t1:.1f} s)")
    return ok, bad, len(pick), int(w_prod)


def run_check(workload, scale_down=1, iters=2, sample_rows=256, heavy_rows=4, kernels=False, no_check=False, log=print):
    """Runs `iters` products of the workload with HBM-resident operands; returns dict(stats of the last iteration,
    invariants_ok, bad_rows, rows_checked, products_checked).  Raises osp.OspError if the engine refuses."""
    t0 = time.time()
    a, b, dims = synth.build_workload(workload, scale_down)
    plen = host_row_partials(a, b)
    P = int(plen.sum())
    bound = int(np.minimum(plen, dims["cols"]).sum())
    log(json.dumps(dict(workload=workload, scale_down=scale_down, nnz_a=a.nnz, nnz_b=b.nnz, dims=dims, products=P,
                        bins_gb=round(P * 8 / 1e9, 2), c_bound_gb=round(bound * 8 / 1e9, 2), max_row_partials=int(plen.max()),
                        build_s=round(time.time() - t0, 1))))
    dev = torch.device("cuda:0")

    def up(x):
        return torch.from_numpy(x.view(np.uint8).reshape(-1)).to(dev)
    t = [up(a.pos), up(a.data)]
    t += t if b is a else [up(b.pos), up(b.data)]          # C = A*A: both operands are the same arrays in HBM
    torch.cuda.synchronize()
    free_b, total_b = torch.cuda.mem_get_info(0)
    log(json.dumps(dict(hbm_total_gb=round(total_b / 1e9, 1), hbm_free_gb=round(free_b / 1e9, 1))))
    eng = osp.Engine(0)
    res = None
    out = {}
    try:
        for it in range(iters):
            if res is not None:
                res.free()
                res = None
            flags = api.OSP_PROFILE_KERNELS if (kernels and it == iters - 1) else 0
            w0 = time.perf_counter()
            res = eng.spgemm_device(a.NRow(), t[0].data_ptr(), t[1].data_ptr(), b.NRow(), t[2].data_ptr(), t[3].data_ptr(),
                                    a_is_csr=True, cols_b=dims["cols"], flags=flags, a_nnz=a.nnz, b_nnz=b.nnz)
            wall = (time.perf_counter() - w0) * 1e3
            st = res.stats()
            log(json.dumps(dict(it=it, wall_ms=round(wall, 2), gflops=round(2 * st["products"] / (st["ms_total"] * 1e6), 2),
                                alg_gbs=round(st["algorithmic_bytes"] / (st["ms_total"] * 1e-3) / 1e9, 1),
                                **{k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()})))
            if flags:
                agg = {}
                for name, ms in res.kernel_times():
                    agg.setdefault(name, [0, 0.0]); agg[name][0] += 1; agg[name][1] += ms
                for name, (n, ms) in agg.items():
                    log(f"   {ms:10.3f} ms  x{n:<4d} {name}")
                per = [round(ms, 1) for name, ms in res.kernel_times() if "k_merge_xl" in name]
                if len(per) > 1:
                    log("   k_merge_xl per row block (ms):", per)
            out["stats"] = st
        if not no_check:
            ok, bad, n_rows, n_prod = check_result(res, a, b, dims, plen, sample_rows, heavy_rows, log)
            out.update(invariants_ok=ok, bad_rows=bad, rows_checked=n_rows, products_checked=n_prod)
    finally:
        if res is not None:
            res.free()
        eng.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="rmat20")
    ap.add_argument("--scale-down", type=int, default=1)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--sample-rows", type=int, default=256)
    ap.add_argument("--heavy-rows", type=int, default=4)
    ap.add_argument("--kernels", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="timing only (runs under ncu)")
    args = ap.parse_args()

    def log(*xs):
        print(*xs, flush=True)
    try:
        run_check(args.workload, args.scale_down, args.iters, args.sample_rows, args.heavy_rows, args.kernels, args.no_check, log)
    except osp.OspError as e:
        log("ENGINE ERROR:", e)
        free_b, _ = torch.cuda.mem_get_info(0)
        log(json.dumps(dict(hbm_free_gb_after_error=round(free_b / 1e9, 1))))


if __name__ == "__main__":
    main()
