#!/usr/bin/env python
"""One GEMM shape, a few launches -- the command profiled with `ncu --set full -k regex:gemm_tc`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tagan_b200 import _lib, ops
op, m, n, k = (int(v) for v in sys.argv[1:5])
dev = torch.device("cuda:0")
lib = _lib.load()
if op == 0: a, b = torch.randn(m, k, device=dev), torch.randn(n, k, device=dev)
elif op == 1: a, b = torch.randn(m, k, device=dev), torch.randn(k, n, device=dev)
else: a, b = torch.randn(k, m, device=dev), torch.randn(k, n, device=dev)
c = torch.empty(m, n, device=dev)
ws = torch.empty(max(lib.tagan_gemm_workspace_bytes(op, m, n, k), 16), dtype=torch.uint8, device=dev)
for _ in range(4):
    rc = lib.tagan_gemm(op, m, n, k, ops._ptr(a), a.stride(0), ops._ptr(b), b.stride(0), None, ops._ptr(c), c.stride(0), 0, 1,
                        ops._ptr(ws), ws.numel(), ops._stream())
    assert rc == 0
torch.cuda.synchronize()
print("ok")
