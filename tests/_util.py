"""Shared helpers for the parity tests (tolerances are the ones BASELINE.json:north_star states)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star: "within 1e-5 relative (fp32) or 2e-2 relative (bf16 GEMM variant)".
# "relative" is taken against the tensor's scale (max |reference|): element-wise relative error is
# meaningless for gradient entries that are themselves rounding noise around zero.
RTOL_FP32 = 1e-5
RTOL_BF16 = 2e-2


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
    return {k: z[k] for k in z.files}


def rel_err(got, ref):
    got = torch.as_tensor(got).detach().double().cpu().reshape(-1)
    ref = torch.as_tensor(ref).detach().double().cpu().reshape(-1)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if ref.numel() == 0:
        return 0.0
    scale = ref.abs().max().item()
    if scale == 0.0:
        scale = 1.0
    return (got - ref).abs().max().item() / scale


def assert_close(got, ref, rtol=RTOL_FP32, what=""):
    e = rel_err(got, ref)
    assert e <= rtol, f"{what}: scaled max error {e:.3e} > {rtol:.1e}"


def sub(d, prefix):
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


def assert_params_close(got, ref32, ref64, what=""):
    """Post-Adam parameters.  Adam divides by sqrt(v): on entries whose gradient is itself rounding noise the
    step direction of ANY two fp32 implementations differs, so the yardstick is the fp32 reference's own
    distance to the float64 run of the same algorithm: |cuda - ref32| <= max(1e-5, 2 * |ref32 - ref64|)
    (all scaled by max|ref|)."""
    e = rel_err(got, ref32)
    anchor = rel_err(ref32, ref64)
    assert e <= max(RTOL_FP32, 2.0 * anchor), f"{what}: scaled max error {e:.3e} (fp32-vs-fp64 anchor {anchor:.3e})"


def rel_l2(got, ref):
    got = torch.as_tensor(got).detach().double().cpu().reshape(-1)
    ref = torch.as_tensor(ref).detach().double().cpu().reshape(-1)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    den = ref.norm().item()
    return (got - ref).norm().item() / (den if den > 0 else 1.0)


def assert_close_l2(got, ref, rtol, what=""):
    """bf16 variant: relative error in the L2 norm of the tensor (north_star: "2e-2 relative (bf16 GEMM variant)").
    A max-norm test is meaningless there: one ReLU gate or one near-zero entry flipped by bf16 rounding is an
    O(1/B) outlier although the tensor as a whole agrees to bf16 precision."""
    e = rel_l2(got, ref)
    assert e <= rtol, f"{what}: relative L2 error {e:.3e} > {rtol:.1e}"
