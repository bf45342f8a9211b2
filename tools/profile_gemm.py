#!/usr/bin/env python
"""GEMM microbenchmark: the projection shapes of the TAGAN layer, FFMA (precision 0) vs tcgen05 3xTF32
(precision 1), CUDA-event timed.  Reports ms, algorithmic GB/s (operands + result once) and TFLOP/s."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tagan_b200 import _lib, ops  # noqa: E402


def run(op, m, n, k, precision, iters=10):
    dev = torch.device("cuda:0")
    lib = _lib.load()
    if op == 0:
        a, b = torch.randn(m, k, device=dev), torch.randn(n, k, device=dev)
    elif op == 1:
        a, b = torch.randn(m, k, device=dev), torch.randn(k, n, device=dev)
    else:
        a, b = torch.randn(k, m, device=dev), torch.randn(k, n, device=dev)
    c = torch.empty(m, n, device=dev)
    wsb = lib.tagan_gemm_workspace_bytes(op, m, n, k)
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for it in range(iters + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.tagan_gemm(op, m, n, k, ops._ptr(a), a.stride(0), ops._ptr(b), b.stride(0), None, ops._ptr(c),
                            c.stride(0), 0, precision, ops._ptr(ws), ws.numel(), ops._stream())
        e1.record()
        torch.cuda.synchronize()
        assert rc == 0, rc
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    byts = 4 * (a.numel() + b.numel() + c.numel())
    return {"op": ["NT", "NN", "TN"][op], "m": m, "n": n, "k": k, "precision": precision, "ms": round(ms, 4),
            "GBs": round(byts / ms / 1e6, 1), "TFLOPs": round(2 * m * n * k / ms / 1e9, 2)}


def main():
    rows = 1_600_000        # T*N of config 3
    shapes = [(0, rows, 384, 128), (0, rows, 128, 128), (0, rows, 256, 256), (1, rows, 128, 384), (1, rows, 256, 128),
              (2, 384, 128, rows), (2, 128, 128, rows), (2, 256, 256, rows), (0, 100_000, 384, 128), (0, 1_000_000, 768, 256)]
    for op, m, n, k in shapes:
        for prec in (1, 3, 2):
            print(json.dumps(run(op, m, n, k, prec)), flush=True)


if __name__ == "__main__":
    main()
