#!/usr/bin/env python
"""Fused-epilogue GEMMs in isolation (for ncu / CUDA-event timing): RES_LN at the stacked size T*N, the two GRU-step
GEMMs (GATES, BLEND; two-source A) and the backward GATES_BWD at the per-step size N, next to the plain GEMM of the
same shape.  One JSON line per case."""
import argparse
import ctypes as C
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000)
    ap.add_argument("--stack", type=int, default=16)
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--fast", action="store_true")
    ap.add_argument("--out", default="gpurun_out/fused_gemm.jsonl")
    args = ap.parse_args()
    from tagan_b200 import _lib, fused, ops
    dev = torch.device("cuda:0")
    fused.EPI_FAST_MATH = args.fast
    n, hd = args.rows, args.hidden
    big = n * args.stack
    torch.manual_seed(0)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)

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

    recs = []
    x = torch.randn(big, hd, device=dev)
    w = torch.randn(hd, hd, device=dev) / hd ** 0.5
    b = torch.randn(hd, device=dev)
    res = torch.randn(big, hd, device=dev)
    g, be = torch.ones(hd, device=dev), torch.zeros(hd, device=dev)
    o = torch.empty(big, hd, device=dev)
    recs.append(("plain_NT_%dx%dx%d" % (big, hd, hd), timeit(lambda: ops.gemm(0, big, hd, hd, x, hd, w, hd, b, o, hd)), 4 * big * hd * 2))
    fused.FUSED_RES_LN = True
    recs.append(("res_ln_%dx%dx%d" % (big, hd, hd), timeit(lambda: fused.linear_res_ln(x, w, b, res, g, be, True)), 4 * big * hd * 4))
    fused.FUSED_RES_LN = False
    recs.append(("gemm_then_ln_%dx%dx%d" % (big, hd, hd), timeit(lambda: fused.linear_res_ln(x, w, b, res, g, be, True)), 4 * big * hd * 6))
    del x, res, o
    kk = 2 * hd
    xh, hh = torch.randn(n, hd, device=dev), torch.randn(n, hd, device=dev)
    w_rz = torch.randn(2 * hd, kk, device=dev) / kk ** 0.5
    b_rz = torch.randn(2 * hd, device=dev)
    w_c = torch.randn(hd, kk, device=dev) / kk ** 0.5
    r, z, rs, cand, hn, dhh = (torch.empty(n, hd, device=dev) for _ in range(6))
    g2 = torch.empty(n, 2 * hd, device=dev)
    dg = torch.randn(n, 3 * hd, device=dev)
    cat = torch.cat([xh, hh], 1)
    recs.append(("plain_NT_%dx%dx%d" % (n, 2 * hd, kk), timeit(lambda: ops.gemm(0, n, 2 * hd, kk, cat, kk, w_rz, kk, b_rz, g2, 2 * hd)), 4 * n * hd * 4))
    e1 = fused._epi(_lib.EPI_GATES, split=hd, in0=hh, ld_in0=hd, out0=r, ld_out0=hd, out1=rs, ld_out1=hd, out2=z, ld_out2=hd)
    recs.append(("gates_2src_%dx%dx%d" % (n, 2 * hd, kk), timeit(lambda: fused.gemm_fused(0, n, 2 * hd, kk, xh, hd, hh, hd, hd, w_rz, kk, b_rz, e1, dev)), 4 * n * hd * 6))
    e2 = fused._epi(_lib.EPI_BLEND, in0=z, ld_in0=hd, in1=hh, ld_in1=hd, out0=cand, ld_out0=hd, out1=hn, ld_out1=hd)
    recs.append(("blend_2src_%dx%dx%d" % (n, hd, kk), timeit(lambda: fused.gemm_fused(0, n, hd, kk, xh, hd, rs, hd, hd, w_c, kk, b, e2, dev)), 4 * n * hd * 6))
    e3 = fused._epi(_lib.EPI_GATES_BWD, in0=r, ld_in0=hd, in1=hh, ld_in1=hd, out0=dg, ld_out0=3 * hd, out1=dhh, ld_out1=hd)
    w_c_h = C.c_void_p(w_c.data_ptr() + hd * 4)
    recs.append(("gates_bwd_%dx%dx%d" % (n, hd, hd), timeit(lambda: fused.gemm_fused(1, n, hd, hd, C.c_void_p(dg.data_ptr() + 2 * hd * 4), 3 * hd, None, 0, 0, w_c_h, kk, None, e3, dev)), 4 * n * hd * 6))
    with open(args.out, "w") as f:
        for name, ms, nbytes in recs:
            rec = {"case": name, "ms": ms, "algorithmic_gbs": nbytes / ms / 1e6, "fast_math": args.fast}
            print(json.dumps(rec), flush=True)
            f.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
