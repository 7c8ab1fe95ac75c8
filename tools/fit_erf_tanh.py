"""Coefficients of gelu_fast (csrc/gemm_tc.cuh): erf(z) ~ tanh(z (a + b z^2 + c z^4)), minimax over z in [0, 5] (Nelder-Mead on the
max error), rewritten in x = sqrt(2) z, and the resulting GELU error against the exact erf form (models/vilbert_dialog.py:121)."""
import numpy as np
from scipy.optimize import minimize
from scipy.special import erf

z = np.linspace(0, 5, 20001)


def err(c):
    return np.max(np.abs(np.tanh(z * (c[0] + c[1] * z ** 2 + c[2] * z ** 4)) - erf(z)))


c = np.array([2 / np.sqrt(np.pi), 0.1, 0.0])
for _ in range(6):
    c = minimize(err, c, method="Nelder-Mead", options=dict(xatol=1e-10, fatol=1e-12, maxiter=20000)).x
r2 = np.sqrt(2)
cx = (c[0] / r2, c[1] / r2 ** 3, c[2] / r2 ** 5)
print("erf fit (a, b, c):", c.tolist(), "max |error|", err(c))
print("in x:", cx)
x = np.linspace(-12, 12, 600001).astype(np.float32)
x2 = np.minimum(x * x, np.float32(50))
q = x2 * (x2 * np.float32(cx[2]) + np.float32(cx[1])) + np.float32(cx[0])
ref = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / r2))
for rel in (0.0, 2.0 ** -11):              # exact tanh, and the worst case of tanh.approx.f32 (relative error 2^-11)
    t = np.tanh((x * q).astype(np.float64)) * (1 + rel)
    g = 0.5 * x * (1 + t)
    print(f"tanh relative error {rel:.1e}: max |GELU error| {np.max(np.abs(g - ref)):.2e}, max |error| / |x| {np.max(np.abs(g - ref)[np.abs(x) > 1e-3] / np.abs(x)[np.abs(x) > 1e-3]):.2e}")
