"""Shared tolerance helper of the GPU parity tests.

``close(..., scaled=False)`` asserts the LITERAL north-star tolerance (rtol 1e-4 / atol 1e-5) -- used for outputs and
attention weights.  ``scaled=True`` takes the atol relative to the reference's max magnitude when that exceeds 1 -- used only
for gradients, which are sums of many O(1..10) terms (a sum of magnitude 20 cannot be resolved to 1e-5 absolute in fp32).
Every comparison is recorded (max abs error, max error relative to the tensor scale, tolerance used); conftest.py writes
the records to ``gpurun_out/parity_r02.json`` at session end, so the parity claim is a set of measured numbers."""
import inspect

import torch

RECORDS = []


def _caller():
    for fr in inspect.stack()[2:8]:
        if fr.function.startswith("test_"):
            return f"{fr.filename.rsplit('/', 1)[-1]}::{fr.function}:{fr.lineno}"
    fr = inspect.stack()[2]
    return f"{fr.filename.rsplit('/', 1)[-1]}::{fr.function}:{fr.lineno}"


def close(got, ref, rtol=1e-4, atol=1e-5, scaled=False, msg=None, kind=None):
    got_c, ref_c = got.detach().float().cpu(), ref.detach().float().cpu()
    scale = max(1.0, float(ref_c.abs().max())) if (scaled and ref_c.numel()) else 1.0
    if ref_c.numel():
        diff = (got_c - ref_c).abs()
        RECORDS.append({"where": _caller(), "kind": kind or ("gradient" if scaled else "output"),
                        "max_abs_err": float(diff.max()), "ref_max_abs": float(ref_c.abs().max()),
                        "max_err_over_scale": float(diff.max()) / max(float(ref_c.abs().max()), 1e-30),
                        "rtol": rtol, "atol": atol * scale, "numel": ref_c.numel()})
    torch.testing.assert_close(got_c, ref_c, rtol=rtol, atol=atol * scale, msg=msg)
