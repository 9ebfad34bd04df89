"""CPU check of the error bound behind the wide assignment pass (csrc/kmeans_solver.cu, assign_pass<TS, true>).

For k > 128 the CUDA solver evaluates the reference's final distance (model/kmeans_with_time.py:86-101: per-row min-max
normalisation of the fp32 feature distance and of the time distance, then sqrt(nf^2 + alpha nt^2)) exactly only for the clusters
whose fp32 ESTIMATE is within 2 E of the smallest estimate, E = 5e-7 fmax / (fmax - fmin) + 1e-6 (2 + sqrt(alpha)).  That is
sound iff every estimate is within E of the exact value.  This test replays both pipelines in numpy with the kernel's operation
order (the approximate square root is modelled as the correctly rounded one perturbed by up to 2 ulp, twice the PTX bound of
sqrt.approx.f32) on rows built to stress the bound: tiny min-max spreads, distances around 1e-3 ... 1e3, duplicate clusters,
float32 and float64 time.  The GPU side of the same claim is tests/test_gpu_parity.py::test_wide_assignment_filter_equals_exact.
"""
import numpy as np
import pytest

f32, f64 = np.float32, np.float64


def exact_final(d2, dt, alpha, ts_f64=True):
    """The reference's arithmetic for one row (what `consider` evaluates): returns the fp32 final distances."""
    d = np.sqrt(np.maximum(d2, 0.0)).astype(f32)  # (float)sqrt(double)
    fmin, fmax = d.min(), d.max()
    fspread = f32(fmax - fmin)
    nf = np.zeros_like(d) if not fmax > fmin else ((d - fmin).astype(f32) / fspread).astype(f32)
    T_ = f64 if ts_f64 else f32
    dt = dt.astype(T_)
    tmin, tmax = dt.min(), dt.max()
    tspread = T_(tmax - tmin)
    nt = np.zeros_like(dt) if not tmax > tmin else ((dt - tmin).astype(T_) / tspread).astype(T_)
    if ts_f64:
        v = (nf * nf).astype(f32).astype(f64) + f64(alpha) * (nt * nt)
        return np.sqrt(v).astype(f32), (fmin, fmax, fspread, tmin, tmax, tspread)
    ntf = nt.astype(f32)
    v = ((nf * nf).astype(f32) + (f32(alpha) * (ntf * ntf).astype(f32)).astype(f32)).astype(f32)
    return np.sqrt(v).astype(f32), (fmin, fmax, fspread, tmin, tmax, tspread)


def estimate_final(d2, dt, alpha, ext, rng):
    """The kernel's fp32 estimate (pass 2) and its bound E."""
    fmin, fmax, fspread, tmin, tmax, tspread = ext
    x = np.maximum(d2, 0.0).astype(f32)
    da = np.sqrt(x).astype(f32)
    da = (da * (f32(1.0) + rng.uniform(-1.0, 1.0, da.shape).astype(f32) * f32(2.0 ** -22))).astype(f32)  # sqrt.approx: <= 2 ulp off
    inv_f = f32(1.0) / fspread if fmax > fmin else f32(0.0)
    inv_t = f32(1.0) / f32(tspread) if tmax > tmin else f32(0.0)
    nf = ((da - fmin).astype(f32) * inv_f).astype(f32)
    nt = ((dt.astype(type(tmin)) - tmin).astype(f32) * inv_t).astype(f32)
    alf = f32(alpha)
    v = ((alf * nt).astype(f32).astype(f64) * nt.astype(f64) + (nf * nf).astype(f32).astype(f64)).astype(f32)  # fmaf(alf * nt, nt, nf * nf)
    E = f32(5e-7) * (fmax * inv_f if fmax > fmin else f32(0.0)) + f32(1e-6) * (f32(2.0) + np.sqrt(alf))
    return np.sqrt(v.astype(f64)), float(E)


@pytest.mark.parametrize("ts_f64", [True, False])
def test_estimate_is_within_the_bound(ts_f64):
    rng = np.random.default_rng(7)
    worst = 0.0
    for trial in range(400):
        k = int(rng.choice([130, 200, 960]))
        scale = 10.0 ** rng.uniform(-3, 3)                       # distances from 1e-3 to 1e3
        rel_spread = 10.0 ** rng.uniform(-4, 0)                  # (fmax - fmin) / fmax down to 1e-4: R up to 1e4
        d = scale * (1.0 - rel_spread * rng.random(k))
        d2 = (d * d).astype(f64)
        if trial % 5 == 0:
            d2[rng.integers(0, k, k // 4)] = d2[0]               # duplicated clusters: exact ties
        if trial % 7 == 0:
            d2[rng.integers(0, k)] = -1e-9                       # cancellation: slightly negative squared distance
        horizon = 10.0 ** rng.uniform(0, 4)
        dt = np.abs(rng.uniform(0, horizon) - np.sort(rng.uniform(0, horizon, k)))
        alpha = float(rng.choice([2.0, 0.5, 8.0]))
        F, ext = exact_final(d2, dt, alpha, ts_f64)
        Fa, E = estimate_final(d2, dt, alpha, ext, rng)
        err = float(np.abs(Fa - F.astype(f64)).max())
        worst = max(worst, err / E)
        assert err <= E, f"trial {trial}: |estimate - exact| = {err:.3e} > E = {E:.3e} (k={k}, scale={scale:.2e}, spread={rel_spread:.1e})"
        # the selection rule: the exact argmin (lowest index on ties) is among the candidates
        best = int(np.argmin(F))
        assert Fa[best] <= Fa.min() + 2 * E
    assert worst < 1.0
    print(f"worst |estimate - exact| / E over 400 rows: {worst:.3f}")


def test_degenerate_rows():
    """All clusters equidistant (spread 0: nf = 0 by the reference's rule) and all centre times equal (nt = 0)."""
    rng = np.random.default_rng(1)
    k = 150
    d2 = np.full(k, 17.25)
    dt = np.full(k, 3.0)
    F, ext = exact_final(d2, dt, 2.0)
    Fa, E = estimate_final(d2, dt, 2.0, ext, rng)
    assert np.all(F == 0) and np.all(Fa == 0) and E > 0
    dt = np.abs(10.0 - np.arange(k, dtype=f64))
    F, ext = exact_final(d2, dt, 2.0)
    Fa, E = estimate_final(d2, dt, 2.0, ext, rng)
    assert float(np.abs(Fa - F).max()) <= E and int(np.argmin(F)) == 10
