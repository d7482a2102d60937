"""CPU: the arithmetic of the one-pass variance op (mxb::OpVar in matx_b200/csrc/mxb_device.cuh) restated in numpy fp32,
operation for operation — shifted sums about a pivot, the compact in-loop re-centring after 8 elements and then every
64, pivot shifts with exact algebra in the merges, M2 = s2 - |s1|^2 / n at the end — and run on the data shapes that
break naive one-pass formulas: |mean| >> stddev, an outlier as the very first element (the initial pivot), an outlier
in the middle, sorted data, a ramp.  This is a model of the device code for its numerical claims,
not a checker of device results: the GPU tests compare the kernels with the oracle and with fp64 truth."""
import numpy as np
import pytest

f = np.float32


def shift(a, m):
    k, s1, s2, c = a
    n = f(c)
    return (f(k + m), f(s1 - f(m * n)), f(s2 + f(f(n * f(m * m)) - f(f(2) * f(m * s1)))), c)


def rcp_approx(c):
    return f(f(1) / f(c)) * f(1 + 2.0 ** -22)      # MUFU.RCP: about one ulp off


def recentre(a):
    k, s1, s2, c = a
    return a if c == 0 else shift(a, f(s1 * rcp_approx(c)))


def lane_state(v):
    k, s1, s2, c = f(0), f(0), f(0), 0
    for x in v:
        if c == 0:
            k = x
        d = f(x - k)
        s1 = f(s1 + d)
        s2 = f(s2 + f(d * d))
        c += 1
        if (c & 63) == 8:                           # the compact in-loop form drops the O(2^-23) residual of s1
            m = f(s1 * rcp_approx(c))
            k, s2, s1 = f(k + m), f(s2 - f(s1 * m)), f(0)
    return (k, s1, s2, c)


def merge(a, b):
    if b[3] == 0:
        return a
    if a[3] == 0:
        return b
    a = recentre(a)
    b = shift(b, f(a[0] - b[0]))
    return (a[0], f(a[1] + b[1]), f(a[2] + b[2]), a[3] + b[3])


def finish(a):
    a = recentre(a)
    return f(a[2] - f(f(a[1] * a[1]) / f(a[3])))


def var_onepass(x, lanes):
    st = [lane_state(x[l::lanes]) for l in range(lanes)]
    while len(st) > 1:                              # warp / CTA / grid stages: a fixed tree
        st = [merge(st[i], st[i + 1]) if i + 1 < len(st) else st[i] for i in range(0, len(st), 2)]
    return finish(st[0]) / f(len(x) - 1)


def truth(x):
    x = x.astype(np.float64)
    return ((x - x.mean()) ** 2).sum() / (len(x) - 1)


def cases():
    rng = np.random.default_rng(1)
    n = 20000
    yield "normal", rng.standard_normal(n).astype(f), 1e-6
    yield "uniform", (rng.random(n) + 0.5).astype(f), 1e-6
    yield "large_mean", (rng.random(n) + 1e4).astype(f), 1e-4       # fp32 resolves these values to 1e-3: data-limited
    y = rng.standard_normal(n).astype(f)
    y[0] = 1e6
    yield "outlier_is_the_first_pivot", y, 5e-6
    y = rng.standard_normal(n).astype(f)
    y[7777] = -1e6
    yield "outlier_in_the_middle", y, 5e-6
    yield "sorted", np.sort(rng.standard_normal(n)).astype(f), 1e-6
    yield "ramp", (np.arange(n) * 0.37 + 5).astype(f), 1e-6


@pytest.mark.parametrize("lanes", [64, 1024])
def test_onepass_variance_model_is_stable(lanes):
    for name, x, bar in cases():
        got, want = float(var_onepass(x, lanes)), truth(x)
        assert abs(got - want) <= bar * want, (name, lanes, got, want, abs(got - want) / want)
