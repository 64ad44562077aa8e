"""The backprojection kernel divides by the pixel size as  q0 = num * r;  q = fma(fma(-q0, px, num), r, q0)  with
r = RN(1 / px)  (csrc/backproject_tma.cu: consume).  This checks, in float32 arithmetic emulated with numpy, that the
result is bit-identical to the IEEE division num / px the reference performs (src/openmp/backprojection.cpp:49) on
2e7 random numerators per pixel size -- the pixel sizes of the BASELINE configurations and a few odd ones."""
import numpy as np

rng = np.random.default_rng(1)


def fma(a, b, c):
    # a*b is exact in float64 for float32 inputs; the sum fits float64 for operands of similar magnitude
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


worst = 0
for px in (0.1, 0.2, 0.4, 0.8, 0.05, 0.0993, 0.127, 0.3, 0.074):
    s = np.float32(px)
    r = np.float32(1.0) / s
    num = rng.uniform(0, 2 * 2048 * px * 1.2, 20_000_000).astype(np.float32)
    q0 = num * r
    q = fma(fma(-q0, np.full_like(num, s), num), np.full_like(num, r), q0)
    bad = int((q != num / s).sum())
    worst = max(worst, bad)
    print(f"px {px}: {bad} of {num.size} quotients differ from IEEE division (plain reciprocal multiply: {int((q0 != num / s).sum())})")
assert worst == 0
