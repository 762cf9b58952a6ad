"""The exact arg-max of the fused kernels takes ONE square root: with r = sqrt_rn(pmax) and r- its predecessor, a candidate
p <= pmax rounds to the same root exactly when p >= roundup(((r- + r) / 2)^2), the square taken exactly in double
(usc_warpfft.cuh `same_root_threshold`).  This file restates the threshold in numpy and checks the claim
against the definition — sqrt of every candidate — around the boundary, where an off-by-one-ulp error would show."""
import numpy as np


def same_root_threshold(pmax):
    """numpy twin of same_root_threshold(): (threshold, root) as float32."""
    pmax = np.float32(pmax)
    r = np.sqrt(pmax, dtype=np.float32)                       # IEEE, correctly rounded
    if r == 0:
        return np.float32(0), r
    rm = np.nextafter(r, np.float32(0), dtype=np.float32)
    s = np.float64(r) + np.float64(rm)                        # 25 significant bits: exact
    q = 0.25 * (s * s)                                        # 50 significant bits: exact in double
    t = np.float32(q)                                         # round to nearest, then up if it fell below q
    if np.float64(t) < q:
        t = np.nextafter(t, np.float32(np.inf), dtype=np.float32)
    return t, r


def _neighbours(x, n):
    out = [np.float32(x)]
    lo = hi = np.float32(x)
    for _ in range(n):
        lo = np.nextafter(lo, np.float32(0), dtype=np.float32)
        hi = np.nextafter(hi, np.float32(np.inf), dtype=np.float32)
        out += [lo, hi]
    return out


def test_threshold_separates_exactly_the_candidates_with_the_same_root():
    rng = np.random.default_rng(7)
    pmaxes = np.concatenate([
        np.exp(rng.uniform(np.log(1e-30), np.log(1e30), 3000)).astype(np.float32),
        np.float32([1, 2, 4, 0.25, 3, 2 ** 23, 2 ** 24, 2 ** 25, 1.0000001, 1.9999999, 3.9999998, 4.0000005]),
        np.ldexp(np.float32(1), np.arange(-100, 100, 7)).astype(np.float32),          # powers of two: the root changes binade
        (np.float32(1) + np.ldexp(np.float32(1), -np.arange(1, 24))).astype(np.float32),
        np.float32([1e-38, 2e-38, 1.1754944e-38, 5e-39, 1e-40, 1.4e-45]),             # around and below the smallest normal
    ])
    checked = 0
    for pmax in pmaxes:
        thr, r = same_root_threshold(pmax)
        assert thr <= pmax                                    # the maximum itself always qualifies
        for p in _neighbours(thr, 3) + _neighbours(pmax, 2):
            if p > pmax or p < 0:
                continue
            same = np.sqrt(p, dtype=np.float32) == r
            assert same == (p >= thr), (pmax, p, thr, r)
            checked += 1
    assert checked > 20000


def test_first_index_rule_equals_arm_max_over_the_roots():
    """arg-max over sqrt(p) with first-occurrence ties (arm_max_f32) == first k with p_k >= threshold(max p)."""
    rng = np.random.default_rng(11)
    for trial in range(400):
        n = int(rng.integers(5, 160))
        base = np.float32(np.exp(rng.uniform(-20, 40)))
        p = (base * (1 + rng.uniform(-1e-3, 0, n))).astype(np.float32)
        # plant near-ties: copies of the maximum moved by a few ulps
        top = p.max()
        for _ in range(int(rng.integers(1, 6))):
            v = top
            for _ in range(int(rng.integers(0, 4))):
                v = np.nextafter(v, np.float32(0), dtype=np.float32)
            p[int(rng.integers(0, n))] = v
        roots = np.sqrt(p, dtype=np.float32)
        want = int(np.argmax(roots))                          # numpy: first occurrence of the maximum
        thr, r = same_root_threshold(p.max())
        got = int(np.nonzero(p >= thr)[0][0])
        assert got == want and r == roots[want]


def test_zero_maximum():
    thr, r = same_root_threshold(0.0)
    assert thr == 0 and r == 0
