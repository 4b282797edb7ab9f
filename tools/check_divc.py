"""Exhaustive check behind csrc/affine_sample.cu::base_coord: for every base-grid value m = lin(i, n) * (n - 1) of every
output size n <= 4096, the corrected-reciprocal quotient  q = RN(m r), r = RN(1/n);  RN(q + fma(-q, n, m) r)  equals
the IEEE division m / n bit for bit (fma emulated in float64: products of float32 are exact there)."""
import numpy as np

f32, f64 = np.float32, np.float64


def fma(a, b, c):
    return (f64(a) * f64(b) + f64(c)).astype(f32)


def main(max_n=4096):
    bad = tot = 0
    for n in range(2, max_n + 1):
        i = np.arange(n)
        step = f32(f32(2) / f32(n - 1))
        lin = np.where(i < n // 2, fma(step, i.astype(f32), f32(-1)), fma(-step, (n - 1 - i).astype(f32), f32(1))).astype(f32)
        m = (lin * f32(n - 1)).astype(f32)
        r = f32(f32(1) / f32(n))
        q = (m * r).astype(f32)
        bad += int((fma(fma(-q, f32(n), m), r, q) != (m / f32(n)).astype(f32)).sum())
        tot += n
    print(f"{bad} mismatches in {tot} quotients (n = 2..{max_n})")
    return bad


if __name__ == "__main__":
    raise SystemExit(1 if main() else 0)
