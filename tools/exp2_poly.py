"""Design aid for the tower-attention softmax: coefficients and accuracy of a polynomial 2^f on
[-0.5, 0.5] evaluated on the FMA pipe (x = n + f with round-to-nearest n; result = p(f) with n added to the exponent field), as an
alternative to MUFU.EX2 for a share of the scores.  Prints, per degree, the minimax-style relative error and the fp32 coefficients."""
import numpy as np


def fit(deg, lo=-0.5, hi=0.5, iters=40):
    x = np.cos(np.pi * (np.arange(4000) + 0.5) / 4000) * (hi - lo) / 2 + (hi + lo) / 2
    y = np.exp2(x)
    w = np.ones_like(x)
    for _ in range(iters):                       # iteratively re-weighted least squares on the RELATIVE error -> near-minimax
        A = np.vander(x, deg + 1, increasing=True) / y[:, None]
        c, *_ = np.linalg.lstsq(A * w[:, None], w, rcond=None)
        err = A @ c - 1
        w = w * (1 + 8 * np.abs(err) / np.abs(err).max())
    return c, float(np.abs(err).max())


if __name__ == "__main__":
    for deg in (2, 3, 4):
        c, e = fit(deg)
        print(f"degree {deg}: max rel err {e:.3e}  (bf16 half-ulp 2^-9 = {2 ** -9:.3e})  coeffs c0..c{deg} = "
              + ", ".join(f"{np.float32(v):.9g}f" for v in c))
    c = fit(3)[0].astype(np.float32)
    x = np.random.default_rng(0).uniform(-40, 8, 2_000_000).astype(np.float32)
    n = np.rint(x).astype(np.float32)
    f = (x - n).astype(np.float32)
    p = ((c[3] * f + c[2]) * f + c[1]) * f + c[0]
    r = (p.view(np.int32) + (n.astype(np.int32) << 23)).view(np.float32)
    print("degree 3 in fp32 with exponent insertion, x in [-40, 8]: max rel err", float(np.abs(r / np.exp2(x.astype(np.float64)) - 1).max()))
