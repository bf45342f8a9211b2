#!/usr/bin/env python
"""The projection shapes of one config-3 step in isolation (CUDA events, L2 flushed between launches, median): time,
algorithmic GB/s, and -- for the NT / NN shapes with K <= 128 -- the same launch with the resident weight panel switched
off (``tagan_gemm_set_tuning``).  Each result is also checked against a float64 matmul on a row sample."""
import argparse
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_600_000)
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--reps", type=int, default=9)
    ap.add_argument("--out", default="gpurun_out/gemm_shapes.jsonl")
    args = ap.parse_args()
    from tagan_b200 import _lib, ops
    lib = _lib.load()
    dev = torch.device("cuda:0")
    r, h = args.rows, args.hidden
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    torch.manual_seed(0)

    def timeit(fn):
        ts = []
        for i in range(args.reps + 1):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            if i:
                ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    cases = [("NT", 0, r, 3 * h, h), ("NT", 0, r, h, h), ("NN", 1, r, h, h), ("NN", 1, r, h, 3 * h), ("TN", 2, 3 * h, h, r),
             ("TN", 2, h, h, r), ("NT", 0, r // 16, 2 * h, 2 * h), ("NT", 0, r // 16, h, h), ("NT", 0, r // 16, 2 * h, h)]
    out = open(args.out, "w")
    for name, op, m, n, k in cases:
        if op == 0:
            a, b = torch.randn(m, k, device=dev), torch.randn(n, k, device=dev) / k ** 0.5
        elif op == 1:
            a, b = torch.randn(m, k, device=dev), torch.randn(k, n, device=dev) / k ** 0.5
        else:
            a, b = torch.randn(k, m, device=dev), torch.randn(k, n, device=dev) / k ** 0.5
        bias = torch.randn(n, device=dev) if op != 2 else None
        c = torch.empty(m, n, device=dev)
        lda, ldb = a.shape[1], b.shape[1]
        rec = {"case": "%s_%dx%dx%d" % (name, m, n, k)}
        nbytes = 4 * (a.numel() + c.numel() + b.numel())
        # (256-bit epilogue stores, lean instantiations: bit 0 pre-split K-major weights, bit 1 dW products)
        modes = [(1, 3), (1, 0), (1, 3), (1, 0)]
        for rep_i, mode in enumerate(modes):
            for key, val in zip((5, 6), mode):
                lib.tagan_gemm_set_tuning(key, val)
            c.fill_(float("nan"))
            ms = timeit(lambda: ops.gemm(op, m, n, k, a, lda, b, ldb, bias, c, n))
            key = "st256_%d_lean%d_run%d" % (mode + (rep_i // 2,)) 
            rec[key + "_ms"] = round(ms, 4)
            # float64 check on a sample of rows
            idx = torch.randint(0, m, (256,), device=dev)
            if op == 0:
                ref = a[idx].double() @ b.double().t()
            elif op == 1:
                ref = a[idx].double() @ b.double()
            else:
                ref = a[:, idx].double().t() @ b.double()
            if bias is not None:
                ref = ref + bias.double()
            err = float((c[idx].double() - ref).abs().max() / ref.abs().max())
            rec[key + "_relerr"] = err
            assert err < 2e-5, (rec, err)
        for key, val in ((0, 1), (2, 0), (3, 0), (4, 0x989680), (5, 1), (6, 3)):
            lib.tagan_gemm_set_tuning(key, val)
        print(json.dumps(rec), flush=True)
        out.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
