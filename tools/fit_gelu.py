"""Fit of the erfc form the encoder-stem epilogue uses (asr-ttl-mtl_b200/csrc/stem_conv.cu, gelu_from_half):

    erfc(u) ~= 2 ** (-u q(u)),  0 <= u <= 4.3,  q a polynomial

q is fitted to -log2(erfc(u)) / u by iteratively re-weighted least squares, the weight being the sensitivity of
GELU(v) = h + |h| (1 - erfc(|h| sqrt 2)), h = v / 2, to an error in q.  Prints, per degree, the float32 coefficients
(constant term first) and the largest |erfc| and |GELU| error when q is evaluated in float32.  CPU only (numpy, scipy).
"""
import numpy as np
from scipy.special import erf, erfc

U = 4.3


def fit(degree):
    u = np.linspace(1e-6, U, 200001)
    g = -np.log2(erfc(u)) / u
    w = erfc(u) * u * np.log(2) * (u / np.sqrt(2) + 0.05)
    wt = w.copy()
    a = np.vander(u, degree + 1, increasing=True)
    for _ in range(60):
        c = np.linalg.lstsq(a * wt[:, None], g * wt, rcond=None)[0]
        err = (a @ c - g) * w
        wt = wt * (1 + 4 * np.abs(err) / np.abs(err).max())
        wt /= wt.max()
    return c.astype(np.float32)


def check(c):
    v = np.linspace(-8.5, 8.5, 1700001)
    h = (0.5 * v).astype(np.float32)
    a = np.abs(h)
    u = np.minimum(a * np.float32(np.sqrt(2)), np.float32(U)).astype(np.float32)
    q = np.full_like(u, c[-1])
    for k in range(len(c) - 2, -1, -1):
        q = (q * u + c[k]).astype(np.float32)
    e = np.exp2(-(u * q).astype(np.float32).astype(np.float64))
    gelu = h.astype(np.float64) + a - a * e
    v = 2.0 * h.astype(np.float64)      # the pre-activation this float32 `h` stands for
    true = 0.5 * v * (1 + erf(v / np.sqrt(2)))
    return np.abs(e - erfc(np.minimum(np.abs(v) / np.sqrt(2), 30.0))).max(), np.abs(gelu - true).max()


if __name__ == "__main__":
    for degree in (5, 6, 7):
        c = fit(degree)
        e_err, g_err = check(c)
        print(f"degree {degree}: |erfc error| {e_err:.2e}  |GELU error| {g_err:.2e}")
        print("   ", ", ".join(repr(float(x)) for x in c))
