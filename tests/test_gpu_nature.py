"""GPU: the 'nature' trunk (network.py:30-42) through the C-ABI -- forward against
the golden vectors made by executing the reference's own lines (oracle/make_golden_nature.py) and
against the float64 oracle, all 12 gradients + the layer-to-layer gradients against float64
autograd, and an Agent trajectory with DQN_type='nature'.  Tolerance: north_star rel-err <= 1e-3."""
import os

import numpy as np
import pytest
import torch

from oracle import a3c
from oracle.make_golden_nature import golden_weights
from oracle.make_golden_network import golden_stacks
from util import REL_TOL, block, norm_err, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _net(pkg, A, B, T, params):
    net = pkg.make_network(DQN_type="nature", action_size=A, num_envs=B, t_max=T, device="cuda:0")
    net.set_weights(params)
    return net


def _history(pkg, cuda, B, T, screens):
    cfg = pkg.config.get_config({"model": "m1", "num_envs": B, "t_max": T})
    hist = pkg.History(cfg, num_envs=B, device=cuda)
    for k in range(4):
        hist.add(torch.as_tensor(screens[k], device=cuda))
    return hist


def test_nature_forward_vs_executed_reference_lines(pkg, cuda):
    g = np.load(os.path.join(ROOT, "tests", "golden", "nature_golden.npz"))
    A = int(g["action_size"])
    stacks = golden_stacks()                                       # [2,84,84,4]
    net = _net(pkg, A, 2, 1, golden_weights(A))
    screens = np.ascontiguousarray(stacks.transpose(3, 0, 1, 2))  # [4, B, 84, 84], oldest first
    hist = _history(pkg, cuda, 2, 1, screens)
    logits, probs, value = net.forward(hist, 0)
    torch.cuda.synchronize()
    assert np.array_equal(net.x.cpu().numpy(), stacks.astype(np.float32))      # History.get channel order
    errs = dict(a1=rel_err(net.l1.cpu(), g["a1"]), a2=rel_err(net.l2.cpu(), g["a2"]),
                a3=rel_err(net.l3.cpu(), g["a3"]), h=rel_err(net.l4.cpu(), g["h"]),
                logits=rel_err(logits.cpu(), g["logits"]), value=rel_err(value.cpu(), g["value"].reshape(-1)),
                probs=rel_err(probs.cpu(), g["policy"]))
    print("nature forward vs executed-reference golden", errs)
    assert max(errs.values()) <= REL_TOL, errs


@pytest.mark.parametrize("A,B,T", [(6, 5, 3), (18, 2, 1), (4, 37, 2)])
def test_nature_backward_vs_oracle(pkg, cuda, A, B, T):
    rng = np.random.default_rng(A * 100 + B)
    params = a3c.init_params(A, seed=9, trunk="nature")
    for k in params:                                               # non-zero biases, livelier weights
        params[k] = (params[k] * 2.0 if k.endswith("_w") else rng.normal(0, 0.01, params[k].shape)).astype(np.float32)
    net = _net(pkg, A, B, T, params)
    screens = rng.integers(0, 256, (T + 4, B, 84, 84), dtype=np.uint8)
    hist = _history(pkg, cuda, B, T, screens)
    for t in range(T):
        net.forward(hist, t)
        hist.add(torch.as_tensor(screens[4 + t], device=cuda))
    acts = rng.integers(0, A, (T, B)); rew = rng.choice([-2.0, 0.0, 1.0], (T, B)); term = rng.random((T, B)) < 0.2
    d = lambda x: torch.as_tensor(x, device=cuda)
    v_boot = net.bootstrap_value(hist)
    net.compute_gradients(hist, d(rew.astype(np.float32)), d(term.astype(np.uint8)), v_boot,
                          actions=d(acts.astype(np.int32).reshape(-1)), grad_scale=1.0 / B)
    torch.cuda.synchronize()
    stacks = a3c.stacks_from_screens(screens, T)
    with torch.no_grad():
        _, vb = a3c.forward(a3c.to_torch(params), stacks[T])
    assert rel_err(v_boot.cpu(), vb) <= REL_TOL
    R = a3c.nstep_returns(a3c.clip_rewards(rew), term, vb.numpy(), 0.99)
    masks = dict(a1=(net.l1 > 0).cpu().numpy(), a2=(net.l2 > 0).cpu().numpy(),
                 a3=(net.l3 > 0).cpu().numpy(), h=(net.l4 > 0).cpu().numpy())
    N = T * B
    grads, aux = a3c.gradients(params, stacks[:T].reshape(N, 84, 84, 4), acts.reshape(-1), R.reshape(-1),
                               0.01, B, masks=masks)
    ferr = dict(logits=rel_err(net.policy_logits.cpu(), aux["logits"]), value=rel_err(net.value.cpu(), aux["value"]))
    gerr = {k: rel_err(net.g[k].cpu(), grads[k]) for k in a3c.NATURE_PARAM_NAMES}
    ierr = dict(d_h=rel_err(net.d_l4.cpu(), aux["d_h"]), d_a3=rel_err(net.d_l3.cpu(), aux["d_a3"]),
                d_a2=rel_err(net.d_l2.cpu(), aux["d_a2"]), d_a1=rel_err(net.d_l1.cpu(), aux["d_a1"]))
    print("nature forward", ferr, "\ngrad", gerr, "\nlayer gradients", ierr)
    assert max(ferr.values()) <= REL_TOL and max(gerr.values()) <= REL_TOL and max(ierr.values()) <= REL_TOL


def test_nature_clip_rmsprop_layout(pkg, cuda):
    """arl_clip_rmsprop_layout over the 12-tensor layout (agent.py:316-321; main.py:63-65), with
    gradients large enough that the float32 update step is resolved (as in the nips test)."""
    A = 6
    rng = np.random.default_rng(A)
    params = a3c.init_params(A, seed=4, trunk="nature")
    net = _net(pkg, A, 2, 1, params)
    shapes = a3c.param_shapes(A, "nature")
    grads = {k: rng.normal(0, 1.0, s).astype(np.float32) for k, s in shapes.items()}
    grads["l4_w"] *= 0.02                     # norm ~25 -> untouched; l3_w norm ~192 -> clipped
    grads["l1_b"] *= 100.0                    # norm ~560 -> clipped
    rms0 = {k: rng.uniform(0.5, 2.0, s).astype(np.float32) for k, s in shapes.items()}
    net.grads.copy_(torch.as_tensor(a3c.flatten_params(grads)))
    net.rms.copy_(torch.as_tensor(a3c.flatten_params(rms0)))
    net.apply_gradients(0.0007)
    torch.cuda.synchronize()
    newp, newr = a3c.update(params, rms0, grads, 0.0007)
    for i, k in enumerate(a3c.NATURE_PARAM_NAMES):
        nrm = float(np.sqrt((grads[k].astype(np.float64) ** 2).sum()))
        assert abs(float(net.grad_norms[i]) - nrm) <= 1e-4 * nrm
        lo, hi = net.offsets[i], net.offsets[i + 1]
        assert rel_err(net.rms[lo:hi].cpu().reshape(shapes[k]), newr[k]) <= 1e-5
        step_ref = newp[k] - params[k].astype(np.float64)
        step_gpu = net.w[k].cpu().numpy().astype(np.float64) - params[k].astype(np.float64)
        assert rel_err(step_gpu, step_ref) <= REL_TOL, k


def test_nature_agent_trajectory(pkg, cuda):
    """Agent with DQN_type='nature': 4 updates (eager, then CUDA-graph replays) against the oracle
    cycle on the agent's own rollouts."""
    A, B, T = 6, 8, 5
    cfg = pkg.config.get_config({"model": "m1", "num_envs": B, "t_max": T, "DQN_type": "nature"})
    env = pkg.GymEnvironment(cfg, env=pkg.SyntheticAtari(B, A, seed=4, pool=T + 3, device=cuda,
                                                         p_terminal=0.05, reward_scale=2.0), device=cuda)
    agent = pkg.Agent(cfg, env, device=cuda)
    assert type(agent.network).__name__ == "NatureNetwork"
    params = a3c.init_params(A, seed=3, trunk="nature")
    agent.network.set_weights(params)
    p_ref = {k: v.astype(np.float64) for k, v in params.items()}
    r_ref = {k: np.ones_like(v, np.float64) for k, v in params.items()}
    agent.before_train()
    net = agent.network
    for u in range(4):
        ring = agent.history
        first = ring.first_slot(0)
        snap = ring.planes().cpu().numpy()
        screens = [snap[:, (first + k) % ring.ring_slots] for k in range(4)]
        step0 = agent.step
        for t in range(T):
            a = agent.predict()
            scr, rew, term = env.act(a, is_training=True, fused=True)
            if t == T - 1:                                         # (the update runs inside observe)
                acts = agent.batch_action.cpu().numpy().copy()
            agent.observe(scr, rew, a, term)
            screens.append(agent.history.planes(agent.history.head).cpu().numpy())
            agent.step += 1
        torch.cuda.synchronize()
        masks = dict(a1=(net.l1 > 0).cpu().numpy(), a2=(net.l2 > 0).cpu().numpy(),
                     a3=(net.l3 > 0).cpu().numpy(), h=(net.l4 > 0).cpu().numpy())
        p_ref, r_ref, aux = a3c.a3c_cycle(p_ref, r_ref, np.stack(screens), acts,
                                          agent.batch_reward.cpu().numpy(),
                                          agent.batch_terminal.cpu().numpy().astype(bool), step0,
                                          num_envs=B, masks=masks)
        flat_gpu = net.params.cpu().numpy().astype(np.float64)
        assert rel_err(flat_gpu, a3c.flatten_params(p_ref)) <= REL_TOL, u
        gerr = max(rel_err(net.g[k].cpu(), aux["grads"][k]) for k in a3c.NATURE_PARAM_NAMES if k.endswith("_w"))
        assert gerr <= REL_TOL, (u, gerr)
    assert agent.update_count == 4
