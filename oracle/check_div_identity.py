"""Test infrastructure (not shipped): host check of the identity the gather-mean kernel uses to
divide a row's fp32 sums by its token count without an IEEE division (csrc/rowops.cu):
    r = RN(1/n);  q = RN(x r);  q' = RN(q + RN(x - n q) r)  ==  RN(x / n)
sampled over n = 1 .. 1000 and 40 decades of |x| (longdouble stands in for the exact FMA)."""
import numpy as np


def mismatches(samples_per_n: int = 100_000, seed: int = 0) -> int:
    rng = np.random.default_rng(seed)
    bad = 0
    for n in list(range(1, 129)) + [255, 256, 257, 1000]:
        fn = np.float32(n)
        r = (np.float32(1.0) / fn).astype(np.float32)
        x = (rng.standard_normal(samples_per_n) * np.exp(rng.uniform(-40, 40, samples_per_n))).astype(np.float32)
        x = x[(np.abs(x) > 1e-30) & (np.abs(x) < 1e30)]
        q0 = (x * r).astype(np.float32)
        e = (np.longdouble(x) - np.longdouble(fn) * np.longdouble(q0)).astype(np.float32)
        q1 = (np.longdouble(q0) + np.longdouble(e) * np.longdouble(r)).astype(np.float32)
        bad += int((q1 != (x / fn).astype(np.float32)).sum())
    return bad


if __name__ == "__main__":
    print("mismatches:", mismatches())
