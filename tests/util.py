"""Comparators shared by the parity tests.

Tolerances (BASELINE.json north_star): integer outputs (match indices, labels, classes, masks, NMS keep
indices) bit-exact; fp32 losses, deltas and gradients within 1e-5 relative."""
import torch

RTOL = 1e-5


def assert_equal_int(a, b, what):
    a, b = a.cpu(), b.cpu()
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, tuple(a.shape), tuple(b.shape))
    assert a.dtype == b.dtype, "%s: dtype %s vs %s" % (what, a.dtype, b.dtype)
    bad = (a != b).nonzero()
    assert bad.numel() == 0, "%s: %d mismatches, first at %s" % (what, bad.shape[0], bad[0].tolist())


def assert_close_scalar(a, b, what, rtol=RTOL):
    a, b = float(a), float(b)
    assert abs(a - b) <= rtol * max(abs(b), 1e-30), "%s: %.9g vs %.9g (rel %.3g)" % (
        what, a, b, abs(a - b) / max(abs(b), 1e-30))


def assert_close_tensor(a, b, what, rtol=RTOL, atol_scale=1e-7):
    """|a-b| <= rtol*|b| + atol_scale*max|b| elementwise: 1e-5 relative with an absolute floor of 1e-7 of
    the tensor's largest magnitude (fp32 rounding of the reference itself)."""
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, tuple(a.shape), tuple(b.shape))
    scale = float(b.abs().max()) if b.numel() else 0.0
    err = (a - b).abs()
    tol = rtol * b.abs() + atol_scale * scale
    bad = err > tol
    if bad.any():
        i = int((err - tol).argmax())
        raise AssertionError("%s: %d/%d elements out of tolerance; worst got %.9g want %.9g (scale %.3g)" % (
            what, int(bad.sum()), a.numel(), a.flatten()[i], b.flatten()[i], scale))
