"""Fit of the polynomial used by the GELU epilogue (keep_b200/csrc/gemm_tcgen05.cu: gelu_erf).

gelu(x) = relu(x) - |x| * q(|x|),  q(a) = 0.5 * erfc(a / sqrt(2)) = exp2(P(a)) on [0, 6], P of degree 4 (default) or 6.
Weighted (error measured on gelu, not on log2 q) iteratively re-weighted least squares -> near-minimax.
"""
import numpy as np
from scipy.special import erf, erfc

import sys

A, DEG = 6.0, int(sys.argv[1]) if len(sys.argv) > 1 else 4  # 4 = shipped default, 6 = KB_GELU_DEG 6
xs = (np.cos(np.linspace(0, np.pi, 40001)) + 1) / 2 * A
q = 0.5 * erfc(xs / np.sqrt(2))
L = np.log2(q)
w = np.maximum(q * np.maximum(xs, 0.05 if DEG < 6 else 0.3), 1e-12 if DEG < 6 else 1e-9)
V = np.vander(xs / A, DEG + 1, increasing=True)
ww = w.copy()
for _ in range(200):
    coef, *_ = np.linalg.lstsq(V * ww[:, None], L * ww, rcond=None)
    err = (V @ coef - L) * w
    ww = ww * (1 + 4 * np.abs(err) / np.abs(err).max())
    ww /= ww.max()
c = coef / (A ** np.arange(DEG + 1))
print(f"coefficients a^0..a^{DEG}:", ", ".join(f"{v:.9e}f" for v in c))
x = np.linspace(-10, 10, 4000001).astype(np.float32)
m = np.minimum(np.abs(x), np.float32(A))
p = np.full_like(m, np.float32(c[DEG]))
for k in range(DEG - 1, -1, -1):
    p = (p * m + np.float32(c[k])).astype(np.float32)
g = (np.maximum(x, 0) - np.abs(x) * np.exp2(p).astype(np.float32)).astype(np.float32)
ref = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
print("fp32 evaluation: max |gelu error| =", np.abs(g - ref).max())
