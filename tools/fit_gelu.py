"""Fit of the GEGLU epilogue's erf-GELU form (difashion_b200/csrc/dfb_common.cuh: gelu_sigmoid_f):
gelu(x) = x * Phi(x) ~= x * sigmoid(2 x q(x^2)), q = a0 + a1 x^2 + a2 x^4, fitted against the exact erf form,
then checked in float32 arithmetic with the kernel's clamp (x^2 <= 64).  Prints the coefficients scaled by
-2*log2(e) as they appear in the kernel and the max abs error."""
import numpy as np
from scipy.optimize import least_squares
from scipy.special import erf

x = np.linspace(-9, 9, 200001)
g = x * 0.5 * (1 + erf(x / np.sqrt(2)))


def model(c, x):
    x2 = x * x
    p = np.zeros_like(x)
    for a in c[::-1]:
        p = p * x2 + a
    return x / (1 + np.exp(np.clip(-2 * x * p, -80, 80)))


c = np.array([0.7978845608, 0.7978845608 * 0.044715, 0.0])
c = least_squares(lambda c: model(c, x) - g, c, xtol=1e-15, ftol=1e-15).x
for pw in (4, 8, 16):       # push the least-squares fit towards minimax
    c = least_squares(lambda c: np.sign(model(c, x) - g) * (np.abs(model(c, x) - g) * 1e3) ** (pw / 2), c, xtol=1e-15, ftol=1e-15).x
cp = (c * (-2 * np.log2(np.e))).astype(np.float32)
print("q coefficients:", c, "\nkernel constants:", [repr(float(v)) for v in cp])
xf = np.linspace(-12, 12, 2000001).astype(np.float32)
x2 = np.minimum(xf * xf, np.float32(64))
q = (cp[2] * x2 + cp[1]).astype(np.float32)
q = (q * x2 + cp[0]).astype(np.float32)
with np.errstate(over="ignore"):
    e = np.exp2((xf * q).astype(np.float32)).astype(np.float32)
got = (xf / (np.float32(1) + e)).astype(np.float32)
ref = xf.astype(np.float64) * 0.5 * (1 + erf(xf.astype(np.float64) / np.sqrt(2)))
print("max abs error (float32 evaluation):", np.abs(got - ref).max())
