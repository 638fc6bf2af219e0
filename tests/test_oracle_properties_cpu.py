"""Property tests of the CPU oracle (oracle/restate.py) — size-independent identities of the domain that the GPU parity
tests also rely on at full size: additivity of confusion matrices, partition of unity of the bilinear weights, bounds and
extremes of the normalised entropy, zero-sum softmax gradients, permutation invariance and perfect-prediction zero of
the Lovasz loss, the exit rule. hypothesis draws shapes / seeds; a few dozen small cases per property."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import restate as R

FAST = settings(max_examples=25, deadline=None)


@FAST
@given(st.integers(1, 4), st.integers(2, 9), st.integers(1, 60), st.integers(0, 2 ** 31 - 1))
def test_confusion_matrix_additivity_and_marginals(N, C, P, seed):
    r = np.random.default_rng(seed)
    pred = r.integers(0, C, (N, P))
    tgt = r.integers(-1, C + 2, (N, P))                       # includes void / out-of-range labels
    cm = R.confusion_matrix(pred, tgt, C)                     # [N, C+1, C]
    assert cm.shape == (N, C + 1, C) and cm.sum() == N * P
    whole = R.confusion_matrix(pred.reshape(1, -1), tgt.reshape(1, -1), C)[0]
    np.testing.assert_array_equal(cm.sum(0), whole)           # images add up: what the cross-rank all-reduce relies on
    for c in range(C):
        assert whole[:, c].sum() == (pred == c).sum()         # column marginal = predictions of class c
        assert whole[c].sum() == (tgt == c).sum()             # row marginal = pixels labelled c
    assert whole[C].sum() == ((tgt < 0) | (tgt >= C)).sum()   # void row
    tp, fp, fn = R.basics_from_cm(whole)
    np.testing.assert_array_equal(tp + fp, whole.sum(0))
    np.testing.assert_array_equal(tp + fn, whole[:C].sum(1))


@FAST
@given(st.integers(1, 9), st.integers(1, 9), st.integers(1, 40), st.integers(1, 40), st.floats(-5, 5))
def test_bilinear_partition_of_unity_and_range(h, w, H, W, v):
    x = np.full((1, 2, h, w), v, np.float32)
    up = R.bilinear_upsample(x, (H, W))
    np.testing.assert_allclose(up, v, rtol=1e-6, atol=1e-6)   # constants are preserved: weights sum to one
    r = np.random.default_rng(h * 131 + w * 17 + H * 7 + W)
    y = r.standard_normal((1, 1, h, w)).astype(np.float32)
    u = R.bilinear_upsample(y, (H, W))
    assert u.min() >= y.min() - 1e-5 and u.max() <= y.max() + 1e-5      # convex combination
    if (H, W) == (h, w):
        np.testing.assert_allclose(u, y, atol=1e-6)           # identity at equal size


@FAST
@given(st.integers(2, 30), st.integers(1, 12), st.integers(1, 12), st.integers(0, 2 ** 31 - 1))
def test_normalised_entropy_bounds(C, H, W, seed):
    r = np.random.default_rng(seed)
    p = R.softmax_c(r.standard_normal((C, H, W)).astype(np.float32) * 3, 0)
    e = R.pixel_norm_entropy(p, C)
    assert e.shape == (H, W) and e.min() >= -1e-6 and e.max() <= 1 + 1e-5
    np.testing.assert_allclose(R.pixel_norm_entropy(np.full((C, H, W), 1.0 / C, np.float32), C), 1.0, atol=1e-5)
    onehot = np.zeros((C, H, W), np.float32)
    onehot[r.integers(0, C)] = 1.0
    np.testing.assert_allclose(R.pixel_norm_entropy(onehot, C), 0.0, atol=1e-7)      # entr(0) = 0
    s = int(r.integers(1, 5))
    a, b = R.img_norm_entropy(p, C, s=s), R.img_norm_entropy(p, C, s=s, pool_min=True)
    assert b <= a + 1e-6                                       # min pooling never exceeds max pooling
    assert abs(R.img_norm_entropy(p, C) - e.mean(dtype=np.float32)) < 1e-6


@FAST
@given(st.integers(1, 3), st.integers(2, 8), st.integers(1, 30), st.integers(0, 2 ** 31 - 1))
def test_cross_entropy_gradient_identities(N, C, P, seed):
    r = np.random.default_rng(seed)
    x = (r.standard_normal((N, C, P)) * 2).astype(np.float32)
    t = r.integers(0, C + 1, (N, P))                           # C = ignore_index
    loss, d = R.pixel_ce(x, t, C)
    valid = t != C
    if valid.sum() == 0:
        assert np.isnan(loss)
        return
    assert loss >= 0
    np.testing.assert_allclose(d.sum(axis=1), 0.0, atol=1e-6)  # softmax - onehot sums to zero over the classes
    assert np.all(d[np.broadcast_to(~valid[:, None, :], d.shape)] == 0)          # void pixels get no gradient
    l2, _ = R.pixel_ce(x + 7.5, t, C)
    assert abs(l2 - loss) < 1e-4 * max(1.0, abs(loss))         # shift invariance of log-softmax
    # E exits, b_reduction='sum' with weights = the weighted sum of the single-exit losses
    E = 3
    y = (r.standard_normal((E, N, C, P, 1)) * 2).astype(np.float32)
    w = [0.5, 1.0, 2.0]
    tot, _, per = R.br_xentropy(y, t.reshape(N, P, 1), ignore_index=C, b_reduction="sum", n_exits=E, weights=w)
    singles = [R.pixel_ce(y[e], t.reshape(N, P, 1), C)[0] for e in range(E)]
    np.testing.assert_allclose(per, singles, rtol=1e-6)
    np.testing.assert_allclose(tot, np.dot(w, singles), rtol=1e-5)


@FAST
@given(st.integers(2, 6), st.integers(2, 60), st.integers(0, 2 ** 31 - 1))
def test_lovasz_permutation_invariance_and_extremes(C, P, seed):
    r = np.random.default_rng(seed)
    probas = R.softmax_c(r.standard_normal((P, C)).astype(np.float32) * 2, 1)
    labels = r.integers(0, C, P)
    loss, g = R.lovasz_softmax_flat(probas, labels)
    assert 0.0 <= loss <= 1.0 + 1e-6
    perm = r.permutation(P)
    loss_p, g_p = R.lovasz_softmax_flat(probas[perm], labels[perm])
    assert abs(loss_p - loss) < 1e-6                           # a set function of the pixels
    perfect = np.eye(C, dtype=np.float32)[labels]
    assert R.lovasz_softmax_flat(perfect, labels)[0] == 0.0    # perfect prediction
    worst = np.eye(C, dtype=np.float32)[(labels + 1) % C]
    assert R.lovasz_softmax_flat(worst, labels, classes="present")[0] > 0.99    # every present class fully wrong
    jac = R.lovasz_grad(np.sort(r.integers(0, 2, P).astype(np.float32))[::-1].copy())
    assert abs(jac.sum() - 1.0) < 1e-5 or jac.sum() == 0.0     # the Jaccard gradient telescopes to J_last = 1


@FAST
@given(st.lists(st.floats(0, 1), min_size=0, max_size=5), st.floats(0, 1), st.integers(0, 5))
def test_exit_rule(entropies, tau, skip):
    k = R.first_confident_exit(entropies, tau, skip)
    n = len(entropies)
    assert 0 <= k <= n
    if k < n:
        assert k >= skip and entropies[k] < tau and all(e >= tau for e in entropies[skip:k])
    else:
        assert all(e >= tau for e in entropies[skip:])
