"""Ad-hoc phase timing of named workloads with HBM-resident operands (development aid, not bench.py)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch  # noqa: E402

import outerspace_b200 as osp  # noqa: E402
from outerspace_b200 import api, synth  # noqa: E402


def cached_workload(name, scale_down):
    """Operands of a named workload through /dev/shm: several library variants are timed on one box without regenerating them."""
    from outerspace_b200.formats import CSRMatrix
    path = f"/dev/shm/osp_wl_{name}_{scale_down}.npz"
    if os.path.exists(path):
        z = np.load(path)
        a = CSRMatrix(z["a_pos"], z["a_data"])
        b = a if int(z["same"]) else CSRMatrix(z["b_pos"], z["b_data"])
        return a, b, {k: int(v) for k, v in zip(("rows", "n_k", "cols"), z["dims"])}
    a, b, dims = synth.build_workload(name, scale_down)
    same = b is a
    np.savez(path, a_pos=a.pos, a_data=a.data, b_pos=(a.pos[:1] if same else b.pos), b_data=(a.data[:1] if same else b.data),
             same=np.int64(same), dims=np.array([dims["rows"], dims["n_k"], dims["cols"]], np.int64))
    return a, b, dims


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="er16k")
    ap.add_argument("--scale-down", type=int, default=1)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--kernels", action="store_true", help="print per-kernel event times of the last iteration")
    ap.add_argument("--flush", action="store_true", help="flush L2 (256 MiB write) before every iteration")
    ap.add_argument("--cache", action="store_true", help="keep the generated operands in /dev/shm for the next run of the same workload")
    args = ap.parse_args()
    t0 = time.time()
    a, b, dims = cached_workload(args.workload, args.scale_down) if args.cache else synth.build_workload(args.workload, args.scale_down)
    print(f"built {args.workload}/{args.scale_down}: nnzA={a.nnz} nnzB={b.nnz} dims={dims} in {time.time()-t0:.1f}s", flush=True)
    dev = torch.device("cuda:0")

    def up(x):
        return torch.from_numpy(x.view(np.uint8).reshape(-1)).to(dev)
    t = [up(a.pos), up(a.data)]
    t += t if b is a else [up(b.pos), up(b.data)]          # C = A*A: both operands are the same arrays in HBM
    torch.cuda.synchronize()
    eng = osp.Engine(0)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if args.flush else None
    for it in range(args.iters):
        if args.flush:
            flush_buf.fill_(it & 1)
            torch.cuda.synchronize()
        if args.kernels and it == args.iters - 1:
            args.flags |= api.OSP_PROFILE_KERNELS
        w0 = time.perf_counter()
        res = eng.spgemm_device(a.NRow(), t[0].data_ptr(), t[1].data_ptr(), b.NRow(), t[2].data_ptr(), t[3].data_ptr(),
                                a_is_csr=True, cols_b=dims["cols"], flags=args.flags, a_nnz=a.nnz, b_nnz=b.nnz)
        wall = (time.perf_counter() - w0) * 1e3
        st = res.stats()
        gbs = st["algorithmic_bytes"] / (st["ms_total"] * 1e-3) / 1e9
        print(json.dumps(dict(it=it, wall_ms=round(wall, 3), gflops=round(2 * st["products"] / (st["ms_total"] * 1e6), 2),
                              alg_gbs=round(gbs, 1), **{k: (round(v, 4) if isinstance(v, float) else v) for k, v in st.items()})), flush=True)
        if args.check and it == 0:
            import oracle
            got = res.to_host()
            pos, data, prod = oracle.spgemm_rowblocks(a.pos, a.data, b.pos, b.data, 4096)
            ok = np.array_equal(got.pos, pos) and np.array_equal(got.data["idx"], data["idx"]) and \
                np.array_equal(got.data["val"].view(np.uint32), data["val"].view(np.uint32))
            print("oracle check:", "BIT-EXACT" if ok else "MISMATCH", flush=True)
        if args.kernels and it == args.iters - 1:
            for name, ms in res.kernel_times():
                print(f"   {ms*1e3:9.1f} us  {name}", flush=True)
        res.free()
    eng.close()


if __name__ == "__main__":
    main()
