"""GPU: loss_mode 'async_q' -- the learner the reference actually runs (agent.py:141-207,
298-314): epsilon-greedy selection, 1-step targets from a target network, MSE
gradient, through the C-ABI against the float64 oracle."""
import numpy as np
import pytest
import torch

from oracle import a3c, philox
from test_gpu_network import gpu_masks, make_params
from util import REL_TOL, norm_err, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("A,B,ep", [(6, 257, 0.3), (18, 64, 0.0), (4, 100, 1.0), (6, 33, 0.05)])
def test_egreedy_bit_exact(pkg, cuda, A, B, ep):
    rng = np.random.default_rng(A * B)
    q = rng.normal(0, 1, (B, A)).astype(np.float32)
    q[::7, 1] = q[::7, 0] = 3.0                                   # ties -> lowest index
    qd = torch.as_tensor(q, device=cuda)
    act = torch.empty(B, dtype=torch.int32, device=cuda)
    pkg._cabi.call("arl_egreedy_actions", pkg._cabi.ptr(qd), pkg._cabi.ptr(act), B, A, ep, 1000,
                   (1 << 33) + 17, 123, pkg._cabi.stream_ptr())
    torch.cuda.synchronize()
    ref = philox.egreedy_actions(q, np.arange(1000, 1000 + B), (1 << 33) + 17, 123, ep)
    assert np.array_equal(act.cpu().numpy(), ref)
    if ep == 0.0:
        assert np.array_equal(ref, q.argmax(1))
    if ep == 1.0:
        assert len(np.unique(ref)) == A and ref.min() >= 0 and ref.max() < A


@pytest.mark.parametrize("A,N", [(6, 1000), (18, 37)])
def test_q_lossgrad_vs_oracle(pkg, cuda, A, N):
    rng = np.random.default_rng(N)
    q = rng.normal(0, 1, (N, A)).astype(np.float32)
    qn = rng.normal(0, 1, (N, A)).astype(np.float32)
    rew = rng.choice([-3.0, 0.0, 0.5, 2.0], N).astype(np.float32)
    term = (rng.random(N) < 0.3)
    act = rng.integers(0, A, N).astype(np.int32)
    d = lambda x: torch.as_tensor(x, device=cuda)
    tgt, dq, sums = torch.empty(N, device=cuda), torch.empty(N, A, device=cuda), torch.zeros(2, device=cuda)
    scale = 1.0 / N
    ins = [d(rew), d(term.astype(np.uint8)), d(act), d(q), d(qn)]    # keep the device copies alive
    pkg._cabi.call("arl_q_lossgrad", *[pkg._cabi.ptr(x) for x in ins],
                   pkg._cabi.ptr(tgt), pkg._cabi.ptr(dq), pkg._cabi.ptr(sums), N, A, 0.99, -1.0, 1.0,
                   scale, pkg._cabi.stream_ptr())
    torch.cuda.synchronize()
    ref_t = a3c.q_targets(qn, rew, term, 0.99)
    assert rel_err(tgt.cpu(), ref_t) <= 1e-6
    delta = ref_t - q[np.arange(N), act].astype(np.float64)
    ref_dq = np.zeros((N, A))
    ref_dq[np.arange(N), act] = -2.0 * delta * scale
    assert rel_err(dq.cpu(), ref_dq) <= 1e-6
    assert abs(float(sums[0]) - float((delta ** 2).sum())) <= 1e-4 * float((delta ** 2).sum())


@pytest.mark.parametrize("A,B,T", [(6, 9, 4), (18, 5, 2)])
def test_async_q_cycle_vs_oracle(pkg, cuda, A, B, T):
    """One whole async-Q cycle through Agent: rollout with epsilon-greedy, target forward,
    targets, gradient of mean(delta^2), per-tensor clip + RMSProp."""
    cfg = pkg.config.get_config({"model": "m1", "num_envs": B, "t_max": T, "loss_mode": "async_q"})
    env = pkg.GymEnvironment(cfg, env=pkg.SyntheticAtari(B, A, seed=3, pool=T + 3, device=cuda,
                                                         p_terminal=0.2), device=cuda)
    agent = pkg.Agent(cfg, env, device=cuda)
    net = agent.network
    params, tparams = make_params(A, seed=21, scale=2.0), make_params(A, seed=22, scale=2.0)
    net.set_weights(params)
    agent.before_train()                                        # syncs the target (main.py:92) ...
    net.target_params.copy_(torch.as_tensor(a3c.flatten_params(tparams), device=cuda))  # ... then differ
    agent.step = 100                                            # somewhere inside the anneal
    hist = agent.history
    first = hist.first_slot(0)
    snap = hist.planes().cpu().numpy()
    screens = [snap[:, (first + k) % hist.ring_slots] for k in range(4)]
    rews, terms, eps = [], [], []
    for t in range(T):
        eps.append(agent.ep)
        action = agent.predict()
        ref_a = philox.egreedy_actions(net.q[t * B:(t + 1) * B].cpu().numpy(), np.arange(B),
                                       agent.step, agent.seed, agent.ep)
        assert np.array_equal(action.cpu().numpy(), ref_a)
        scr, rew, term = env.act(action, is_training=True, fused=True)
        hist.add(scr)
        agent.batch_reward[t].copy_(rew)
        agent.batch_terminal[t].copy_(term)
        rews.append(rew.cpu().numpy()); terms.append(term.cpu().numpy())
        screens.append(hist.planes(hist.head).cpu().numpy())
        agent.t += 1
        agent.step += 1
    assert abs(eps[0] - a3c.epsilon(100)) < 1e-12
    agent.step -= 1
    acts = agent.batch_action.cpu().numpy().copy()
    masks = gpu_masks(net)
    agent.batch_update()
    torch.cuda.synchronize()
    stacks = a3c.stacks_from_screens(np.stack(screens), T)
    s_t = stacks[:T].reshape(T * B, 84, 84, 4)
    s_tp1 = stacks[1:T + 1].reshape(T * B, 84, 84, 4)
    grads, aux = a3c.async_q_gradients(params, tparams, s_t, s_tp1, acts, np.stack(rews),
                                       np.stack(terms), 0.99, scale=1.0 / (T * B), masks=masks)
    assert rel_err(net.target_q.cpu(), aux["q_next"]) <= REL_TOL
    assert rel_err(net.target_q_t.cpu(), aux["target"]) <= REL_TOL
    errs = {k: rel_err(net.g[k].cpu(), grads[k]) for k in a3c.PARAM_NAMES if k not in ("q_w", "q_b")}
    print("async-Q grad rel-err", errs)
    assert max(errs.values()) <= REL_TOL, errs
    assert float(net.g["q_w"].abs().max()) == 0.0 and float(net.g["q_b"].abs().max()) == 0.0
    rms = {k: np.ones_like(v) for k, v in params.items()}
    new_p, _ = a3c.update(params, rms, grads, a3c.learning_rate(agent.step - (T - 1)))
    perr = {k: rel_err(net.w[k].cpu(), new_p[k]) for k in a3c.PARAM_NAMES}
    assert max(perr.values()) <= REL_TOL, perr
    # the target network moves only when asked to (agent.py:342-344)
    assert rel_err(net.target_params.cpu(), a3c.flatten_params(tparams)) == 0.0
    agent.update_target_q_network()
    assert bool((net.target_params == net.params).all())
