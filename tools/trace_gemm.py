#!/usr/bin/env python
"""Pipeline timeline of the tcgen05 GEMM's CTA 0 (clock64 stamps written by the kernel through tagan_gemm_set_trace):
per k-block the gaps TMA issue -> bytes landed -> split done -> MMA start -> MMA issued, per tile accumulator free ->
full -> epilogue done.  Prints medians in cycles for each c3 shape."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    from tagan_b200 import _lib, ops
    lib = _lib.load()
    dev = torch.device("cuda:0")
    r, h = 1_600_000, 128
    tr = torch.zeros(16, 512, dtype=torch.int64, device=dev)
    out = open("gpurun_out/gemm_trace.jsonl", "w")
    for name, op, m, n, k in (("NT", 0, r, h, h), ("NT", 0, r, 3 * h, h), ("NN", 1, r, h, 3 * h), ("TN", 2, 3 * h, h, r), ("TN", 2, h, h, r)):
        if op == 0:
            a, b = torch.randn(m, k, device=dev), torch.randn(n, k, device=dev)
        elif op == 1:
            a, b = torch.randn(m, k, device=dev), torch.randn(k, n, device=dev)
        else:
            a, b = torch.randn(k, m, device=dev), torch.randn(k, n, device=dev)
        c = torch.empty(m, n, device=dev)
        for _ in range(2):
            ops.gemm(op, m, n, k, a, a.shape[1], b, b.shape[1], None, c, n)
        tr.zero_()
        lib.tagan_gemm_set_trace(tr.data_ptr())
        ops.gemm(op, m, n, k, a, a.shape[1], b, b.shape[1], None, c, n)
        torch.cuda.synchronize()
        lib.tagan_gemm_set_trace(None)
        t = tr.cpu()
        torch.save(t, "gpurun_out/gemm_trace_%s_%dx%dx%d.pt" % (name, m, n, k))
        nkb = int((t[0] > 0).sum())
        ntile = int((t[6] > 0).sum())
        kb = slice(64, min(nkb, 448))
        med = lambda x: float(x.float().median())
        issue, landed, split, mstart, missued = (t[i, kb] for i in range(5))
        rec = {"case": "%s_%dx%dx%d" % (name, m, n, k), "kblocks_traced": nkb, "tiles_traced": ntile,
               "cyc_per_kblock": med(issue[1:] - issue[:-1]),
               "tma_issue_to_landed": med(landed - issue), "landed_to_split_done": med(split - landed),
               "split_done_to_mma_start": med(mstart - split), "mma_start_to_issued": med(missued - mstart),
               "mma_issued_to_next_tma_issue_same_stage": med(issue[4:] - missued[:-4])}
        if ntile > 8:
            tl = slice(4, min(ntile, 100))
            free, full, done = t[5, tl], t[6, tl], t[7, tl]
            rec.update({"cyc_per_tile": med(full[1:] - full[:-1]), "acc_free_to_full": med(full - free), "full_to_epilogue_done": med(done - full)})
        print(json.dumps(rec), flush=True)
        out.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
