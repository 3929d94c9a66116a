"""A named BASELINE config at its FULL size on one GPU: timing + size-independent parity.

The element-wise oracle cannot finish configs 3 and 5 at full size (P = 2.1e10 / 1.1e10 partial products), so the
check is per row -- rows of C are independent of each other: a seeded sample of rows plus the heaviest rows are
recomputed by the oracle from A[rows, :] and B and compared bit for bit (row_ptr differences, col_idx, values), next
to the global invariants (P equal to the host's count from the two pos arrays, row_ptr monotone and ending at nnz(C),
row lengths within min(partials of the row, cols)).
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
from outerspace_b200.formats import CSRMatrix  # noqa: E402


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
    t0 = time.time()
    a, b, dims = synth.build_workload(args.workload, args.scale_down)
    a_pos, b_len = a.pos.astype(np.int64), np.diff(b.pos.astype(np.int64))
    cs = np.concatenate([[0], np.cumsum(b_len[a.data["idx"]])])
    plen = cs[a_pos[1:]] - cs[a_pos[:-1]]                      # partial products per output row
    P = int(plen.sum())
    bound = int(np.minimum(plen, dims["cols"]).sum())
    print(json.dumps(dict(workload=args.workload, scale_down=args.scale_down, nnz_a=a.nnz, nnz_b=b.nnz, dims=dims, products=P,
                          bins_gb=round(P * 8 / 1e9, 2), c_bound_gb=round(bound * 8 / 1e9, 2), max_row_partials=int(plen.max()),
                          build_s=round(time.time() - t0, 1))), flush=True)
    dev = torch.device("cuda:0")

    def up(x):
        return torch.from_numpy(x.view(np.uint8).reshape(-1)).to(dev)
    t = [up(a.pos), up(a.data), up(b.pos), up(b.data)]
    torch.cuda.synchronize()
    free_b, total_b = torch.cuda.mem_get_info(0)
    print(json.dumps(dict(hbm_total_gb=round(total_b / 1e9, 1), hbm_free_gb=round(free_b / 1e9, 1))), flush=True)
    eng = osp.Engine(0)
    res = None
    try:
        for it in range(args.iters):
            if res is not None:
                res.free()
                res = None
            flags = api.OSP_PROFILE_KERNELS if (args.kernels and it == args.iters - 1) else 0
            w0 = time.perf_counter()
            res = eng.spgemm_device(a.NRow(), t[0].data_ptr(), t[1].data_ptr(), b.NRow(), t[2].data_ptr(), t[3].data_ptr(),
                                    a_is_csr=True, cols_b=dims["cols"], flags=flags, a_nnz=a.nnz, b_nnz=b.nnz)
            wall = (time.perf_counter() - w0) * 1e3
            st = res.stats()
            print(json.dumps(dict(it=it, wall_ms=round(wall, 2), gflops=round(2 * st["products"] / (st["ms_total"] * 1e6), 2),
                                  alg_gbs=round(st["algorithmic_bytes"] / (st["ms_total"] * 1e-3) / 1e9, 1),
                                  **{k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()})), flush=True)
            if flags:
                agg = {}
                for name, ms in res.kernel_times():
                    agg.setdefault(name, [0, 0.0]); agg[name][0] += 1; agg[name][1] += ms
                for name, (n, ms) in agg.items():
                    print(f"   {ms:10.3f} ms  x{n:<4d} {name}", flush=True)
                per = [round(ms, 1) for name, ms in res.kernel_times() if "k_merge_xl" in name]
                if len(per) > 1:
                    print("   k_merge_xl per row block (ms):", per, flush=True)
        if args.no_check:
            return
        # ---- global invariants ----
        ok = st["products"] == P
        pos = res.pos_to_host().astype(np.int64)
        rl = np.diff(pos)
        n = len(rl)
        ok &= pos[0] == 0 and pos[-1] == res.nnz and bool(np.all(rl >= 0))
        ok &= bool(np.all(rl <= np.minimum(plen[:n], dims["cols"]))) and bool(np.all((rl > 0) == (plen[:n] > 0)))
        print("global invariants:", "OK" if ok else "VIOLATED", f"(rows {n}, nnz(C) {res.nnz}, P {st['products']} vs host {P})", flush=True)
        # ---- sampled rows against the oracle ----
        rng = np.random.default_rng(7)
        pick = np.unique(np.concatenate([rng.integers(0, n, size=args.sample_rows), np.argsort(plen[:n])[-args.heavy_rows:]]))
        sub_len = (a_pos[pick + 1] - a_pos[pick])
        sub_pos = np.zeros(len(pick) + 1, np.uint64)
        np.cumsum(sub_len, out=sub_pos[1:])
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
                    print(f"  row {r}: got {len(got.data)} entries, oracle {len(want)}", flush=True)
        print(f"sampled rows vs oracle: {'BIT-EXACT' if bad == 0 else f'MISMATCH in {bad} rows'} "
              f"({len(pick)} rows, {int(w_prod)} partial products, heaviest {int(plen[pick].max())}; oracle {time.time() - t1:.1f} s)", flush=True)
    except osp.OspError as e:
        print("ENGINE ERROR:", e, flush=True)
        free_b, total_b = torch.cuda.mem_get_info(0)
        print(json.dumps(dict(hbm_free_gb_after_error=round(free_b / 1e9, 1))), flush=True)
    finally:
        if res is not None:
            res.free()
        eng.close()


if __name__ == "__main__":
    main()
