"""Sparse MLP inference chained on the device (config 5's use case at 1/8 batch): act_0 (8192 x 4096, 10 %) through
three pruned 4096 x 4096 layers (10 %), relu(x W^T + b) kept sparse between them.  Development aid, not bench.py."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch  # noqa: E402

import outerspace_b200 as osp  # noqa: E402
from outerspace_b200 import synth  # noqa: E402


def main():
    batch, width, layers = 8192, 4096, 3
    dev = torch.device("cuda:0")
    x = synth.pruned_dense(batch, width, 0.10, seed=1, nonneg=True)
    ws = [synth.transpose_host(synth.pruned_dense(width, width, 0.10, seed=10 + i), width) for i in range(layers)]
    rng = np.random.default_rng(0)
    bs = [(rng.standard_normal(width) * 0.5 - 2.0).astype(np.float32) for _ in range(layers)]   # negative bias: sparse activations

    def up(a):
        return torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).to(dev)
    tx = [up(x.pos), up(x.data)]
    tw = [(up(w.pos), up(w.data), w.nnz) for w in ws]
    eng = osp.Engine(0)
    stream = torch.cuda.ExternalStream(eng.stream, device=dev)
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        a_pos, a_data, a_nnz = tx[0].data_ptr(), tx[1].data_ptr(), x.nnz
        keep, report = [], []
        for li in range(layers):
            r = eng.spgemm_device(batch, a_pos, a_data, width, tw[li][0].data_ptr(), tw[li][1].data_ptr(), a_is_csr=True,
                                  rows_c=batch, cols_b=width, a_nnz=a_nnz, b_nnz=tw[li][2])
            g = eng.bias_relu(r, width, bs[li])
            report.append((r.stats()["products"], r.nnz, g.nnz))
            keep += [r, g]
            a_pos, a_data = g.device_pointers()
            a_nnz = g.nnz
        e1.record(stream)
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        prods = sum(p for p, _, _ in report)
        print(f"it {it}: {ms:.3f} ms for {layers} layers, P = {prods:.3e}, {2 * prods / ms / 1e6:.1f} GFLOP/s; "
              f"per layer (P, nnz(C), nnz(act)) = {report}", flush=True)
        for k in keep:
            k.free()
    eng.close()


if __name__ == "__main__":
    main()
