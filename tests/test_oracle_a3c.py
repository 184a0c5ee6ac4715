"""CPU: internal consistency of the REF-A3C oracle (parity unpinned by the reference itself:
no TF here and no reference tests -- so the oracle is cross-checked against independent
formulations: autograd vs closed form, loop vs recurrence, Random123 known answers)."""
import numpy as np
import torch

from oracle import a3c, philox


def test_param_count_matches_survey():
    assert a3c.param_count(6) == 677943 and a3c.param_count(18) == 681027
    p = a3c.init_params(6, seed=1)
    assert np.abs(p["l1_w"]).max() <= 0.04 and not p["l1_b"].any()
    flat = a3c.flatten_params(p)
    q = a3c.unflatten_params(flat, 6)
    assert all(np.array_equal(p[k], q[k]) for k in p)


def test_philox_known_answers():
    f = lambda *a: [int(x[0]) for x in philox.philox4x32_10(*[np.array([v], np.uint32) for v in a])]
    assert f(0, 0, 0, 0, 0, 0) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert f(*([0xffffffff] * 6)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert f(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_sampler_distribution_and_edges():
    p = np.tile(np.array([[0.1, 0.2, 0.3, 0.4]], np.float32), (200000, 1))
    a = philox.sample_actions(p, np.arange(200000), step=7, seed=123)
    freq = np.bincount(a, minlength=4) / len(a)
    assert np.abs(freq - p[0]).max() < 5e-3
    onehot = np.zeros((5, 4), np.float32); onehot[:, 2] = 1
    assert (philox.sample_actions(onehot, np.arange(5), 0, 1) == 2).all()
    zeros = np.zeros((5, 4), np.float32)                    # degenerate row -> last action
    assert (philox.sample_actions(zeros, np.arange(5), 0, 1) == 3).all()
    # sharding independence: env ids, not positions, key the stream
    a_all = philox.sample_actions(p[:64], np.arange(64), 3, 9)
    a_hi = philox.sample_actions(p[:32], np.arange(32, 64), 3, 9)
    assert np.array_equal(a_all[32:], a_hi)


def test_returns_recurrence_vs_explicit_sum():
    rng = np.random.default_rng(0)
    T, B = 5, 7
    r = rng.normal(size=(T, B)); term = rng.random((T, B)) < 0.3; vb = rng.normal(size=B)
    R = a3c.nstep_returns(r, term, vb, 0.99)
    for b in range(B):
        for t in range(T):
            acc, disc, alive = 0.0, 1.0, True
            for k in range(t, T):
                acc += disc * r[k, b]
                if term[k, b]:
                    alive = False
                    break
                disc *= 0.99
            if alive:
                acc += disc * vb[b]
            assert abs(acc - R[t, b]) < 1e-12


def test_closed_form_head_grads_equal_autograd():
    torch.manual_seed(0)
    N, A = 11, 6
    logits = torch.randn(N, A, dtype=torch.float64, requires_grad=True)
    value = torch.randn(N, dtype=torch.float64, requires_grad=True)
    acts = torch.randint(0, A, (N,)); R = torch.randn(N, dtype=torch.float64)
    total, _, _ = a3c.loss_per_sample(logits, value, acts, R, beta=0.01)
    (total.sum() / 4).backward()
    dl, dv = a3c.analytic_head_grads(logits.detach(), value.detach(), acts, R, 0.01, 0.25)
    assert (logits.grad - dl).abs().max() < 1e-14 and (value.grad - dv).abs().max() < 1e-14


def test_forward_layout_nhwc_flatten_and_crosscorrelation():
    """One-hot weights pin the layouts: [kh,kw,cin,cout] cross-correlation, NHWC flatten."""
    A = 3
    p = {k: np.zeros(s) for k, s in a3c.param_shapes(A).items()}
    p["l1_w"][2, 5, 1, 7] = 255.0                          # picks x[4oy+2, 4ox+5, c=1] into ch 7
    x = np.random.default_rng(0).integers(0, 256, (1, 84, 84, 4)).astype(np.float64)
    _, _, keep = a3c.forward(a3c.to_torch(p), x, keep=True)
    a1 = keep["a1"].numpy()
    assert np.allclose(a1[0, 3, 4, 7], x[0, 14, 21, 1]) and a1[0, :, :, :7].max() == 0
    p["l2_w"][1, 3, 7, 9] = 1.0                            # a2[oy,ox,9] = a1[2oy+1, 2ox+3, 7]
    p["l4_w"][(2 * 9 + 5) * 32 + 9, 100] = 1.0             # flatten index (h*9+w)*32+c
    _, _, keep = a3c.forward(a3c.to_torch(p), x, keep=True)
    assert np.allclose(keep["h"].numpy()[0, 100], x[0, 4 * (2 * 2 + 1) + 2, 4 * (2 * 5 + 3) + 5, 1])


def test_clip_and_rmsprop_semantics():
    g = np.full(100, 10.0)                                  # norm 100 -> scaled to norm 40
    assert np.isclose(np.linalg.norm(a3c.clip_by_norm(g, 40.0)), 40.0)
    g2 = np.full(4, 1.0)
    assert np.array_equal(a3c.clip_by_norm(g2, 40.0), g2)
    w, ms = a3c.rmsprop_apply(np.array([1.0]), np.array([1.0]), np.array([2.0]), 0.1)
    assert np.isclose(ms[0], 1.0 + (4.0 - 1.0) * 0.01)      # rms slot starts at 1 (TF)
    assert np.isclose(w[0], 1.0 - 0.1 * 2.0 / np.sqrt(1.03 + 0.1))   # epsilon inside the sqrt
    assert np.isclose(a3c.learning_rate(0), 0.0007 * (80000000 + 1) / 80000000)


def test_cycle_shapes_and_terminal_mask():
    rng = np.random.default_rng(3)
    A, T, B = 4, 2, 3
    p = a3c.init_params(A, 5)
    rms = {k: np.ones_like(v) for k, v in p.items()}
    screens = rng.integers(0, 256, (T + 4, B, 84, 84), dtype=np.uint8)
    acts = rng.integers(0, A, (T, B)); rew = np.array([[3.0, -2.0, 0.5]] * T)
    term = np.zeros((T, B), bool); term[T - 1, 0] = True
    newp, newr, aux = a3c.a3c_cycle(p, rms, screens, acts, rew, term, step=0)
    assert aux["R"].shape == (T, B)
    assert np.isclose(aux["R"][T - 1, 0], 1.0)              # clipped reward, bootstrap masked
    assert np.isclose(aux["R"][T - 1, 1], -1.0 + 0.99 * aux["v_boot"][1])
    assert all(newp[k].shape == p[k].shape for k in p)
    st = a3c.stacks_from_screens(screens, T)
    assert st.shape == (T + 1, B, 84, 84, 4) and np.array_equal(st[1, :, :, :, 0], screens[1])


def test_async_q_oracle_matches_closed_form():
    """agent.py:186-190, 310-314: autograd of mean(delta^2) == the closed form the CUDA kernel
    emits at the head (dq[a] = -2*delta/N), and the target uses the TARGET parameters."""
    A, N = 5, 6
    rng = np.random.default_rng(0)
    params, tparams = a3c.init_params(A, 1), a3c.init_params(A, 2)
    s_t = rng.integers(0, 256, (N, 84, 84, 4)).astype(np.uint8)
    s_tp1 = rng.integers(0, 256, (N, 84, 84, 4)).astype(np.uint8)
    acts = rng.integers(0, A, N)
    rew = rng.choice([-2.0, 0.0, 1.0], N)
    term = rng.random(N) < 0.5
    grads, aux = a3c.async_q_gradients(params, tparams, s_t, s_tp1, acts, rew, term, 0.99)
    with torch.no_grad():
        qn, _ = a3c.forward(a3c.to_torch(tparams), s_tp1)
    tgt = (1.0 - term) * 0.99 * qn.numpy().max(1) + np.clip(rew, -1, 1)
    assert np.allclose(aux["target"], tgt, rtol=0, atol=1e-12)
    # head gradient in closed form: d loss / d p_b = sum_n dq[n]
    delta = tgt - aux["q"][np.arange(N), acts]
    dq = np.zeros((N, A))
    dq[np.arange(N), acts] = -2.0 * delta / N
    assert np.allclose(grads["p_b"], dq.sum(0), rtol=1e-10, atol=1e-14)
    assert np.abs(grads["q_w"]).max() == 0.0                       # value head unused
    assert abs(a3c.epsilon(0) - 1.0) < 1e-12 and abs(a3c.epsilon(10 ** 9) - 0.1) < 1e-12


def test_egreedy_oracle_properties():
    from oracle import philox
    q = np.random.default_rng(1).normal(0, 1, (4096, 6)).astype(np.float32)
    ids = np.arange(4096)
    assert np.array_equal(philox.egreedy_actions(q, ids, 5, 123, 0.0), q.argmax(1))
    a = philox.egreedy_actions(q, ids, 5, 123, 1.0)
    counts = np.bincount(a, minlength=6)
    assert counts.min() > 4096 / 6 * 0.8 and counts.max() < 4096 / 6 * 1.2
    frac = (philox.egreedy_actions(q, ids, 5, 123, 0.25) != q.argmax(1)).mean()
    assert 0.15 < frac < 0.27                                      # 0.25 * 5/6 = 0.208
