"""Property tests (hypothesis) of the aggregation primitive as restated in the oracle
(RPBCAC_agent._resilient_aggregation, agents/resilient_CAC_agents.py:42-58) -- including the single-pass identity the
CUDA kernel relies on: sum clip(v) = sum v - sum_{small<lo}(small-lo) - sum_{large>hi}(large-hi)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import rpbcac_oracle as O


@st.composite
def cases(draw):
    n = draw(st.integers(1, 16))
    H = draw(st.integers(0, n - 1))
    P = draw(st.integers(1, 6))
    vals = draw(st.lists(st.floats(-50, 50, allow_nan=False, width=32), min_size=n * P, max_size=n * P))
    v = np.array(vals, np.float64).reshape(n, P)
    if draw(st.booleans()):
        v = np.round(v)                                   # many ties
    return v, H


@given(cases())
@settings(max_examples=300, deadline=None)
def test_single_pass_identity_and_bounds(c):
    v, H = c
    n = v.shape[0]
    want = O.resilient_aggregation(v, H)
    s = np.sort(v, axis=0)
    own = v[0]
    lo, hi = np.minimum(s[H], own), np.maximum(s[n - H - 1], own)
    small, large = s[:H + 1], s[::-1][:H + 1]
    total = v.sum(0) - np.minimum(small - lo, 0).sum(0) - np.maximum(large - hi, 0).sum(0)
    np.testing.assert_allclose(total / n, want, rtol=1e-12, atol=1e-12)
    assert np.all(want >= v.min(0) - 1e-12) and np.all(want <= v.max(0) + 1e-12)
    assert np.all(lo <= own) and np.all(own <= hi)        # clip never sees lo > hi
    if H == 0:
        np.testing.assert_allclose(want, v.mean(0), rtol=1e-12, atol=1e-12)


@given(cases(), st.randoms(use_true_random=False))
@settings(max_examples=150, deadline=None)
def test_neighbour_order_does_not_matter(c, rnd):
    v, H = c
    idx = list(range(1, v.shape[0]))
    rnd.shuffle(idx)
    np.testing.assert_allclose(O.resilient_aggregation(v[[0] + idx], H), O.resilient_aggregation(v, H), rtol=1e-12, atol=1e-12)


@given(cases())
@settings(max_examples=100, deadline=None)
def test_bounded_influence_of_H_outliers(c):
    """Replacing up to H neighbours (not the own row) by arbitrarily large values moves the aggregate by at most the
    spread of the remaining values -- the resilience property the algorithm is built on."""
    v, H = c
    n = v.shape[0]
    if H == 0 or n < 2 * H + 2:
        return
    clean = O.resilient_aggregation(v, H)
    bad = v.copy()
    bad[1:1 + H] = 1e9
    out = O.resilient_aggregation(bad, H)
    honest = np.delete(v, np.s_[1:1 + H], axis=0)
    assert np.all(out <= honest.max(0) + 1e-9) and np.all(out >= honest.min(0) - 1e-9)
    assert np.all(np.isfinite(clean))
