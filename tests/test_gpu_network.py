"""GPU: network forward/backward, sampling, returns/loss, clip+RMSProp and the full A3C cycle
through the C-ABI, against the float64 oracle on identical synthetic inputs and weights.
Tolerance: BASELINE.json north_star -- rel-err <= 1e-3 for logits, values, returns, gradients;
bit-exact for action sampling on identical probabilities; parameter trajectory within 1e-3
over 100 updates."""
import numpy as np
import pytest
import torch

from oracle import a3c, philox
from util import REL_TOL, norm_err, rel_err, unblock

pytestmark = pytest.mark.gpu


def make_params(A, seed=0, bias_std=0.01, scale=1.0):
    p = a3c.init_params(A, seed)
    rng = np.random.default_rng(seed + 1)
    for k in p:
        if k.endswith("_b"):
            p[k] = rng.normal(0, bias_std, p[k].shape).astype(np.float32)
        else:
            p[k] = (p[k] * scale).astype(np.float32)
    return p


def net_for(pkg, A, B, T, params):
    net = pkg.Network(action_size=A, num_envs=B, t_max=T, device="cuda:0")
    net.set_weights(params)
    return net


def gpu_masks(net):
    """relu activation pattern of the device forward (see oracle.a3c._relu)."""
    return dict(a1=(net.a1() > 0).cpu().numpy(), a2=(net.a2() > 0).cpu().numpy(),
                h=(net.l4 > 0).cpu().numpy())


def ring_stacks(ring_np, first_slot, steps):
    """oracle-side view of the ring: [steps*B, 84, 84, 4], t-major."""
    ring_np = unblock(ring_np)
    B, R = ring_np.shape[:2]
    out = []
    for t in range(steps):
        out.append(np.stack([ring_np[:, (first_slot + t + k) % R] for k in range(4)], axis=-1))
    return np.concatenate(out, axis=0)


@pytest.mark.parametrize("A,B,T,R,first", [(6, 5, 3, 8, 6), (18, 4, 1, 4, 2), (4, 7, 2, 9, 0)])
def test_forward_layers_vs_oracle(pkg, cuda, A, B, T, R, first):
    rng = np.random.default_rng(A * 100 + B)
    params = make_params(A, seed=A, scale=2.0)
    ring_np = rng.integers(0, 256, (B, R, 84, 84), dtype=np.uint8)
    ring = torch.as_tensor(ring_np, device=cuda)
    flat = torch.as_tensor(a3c.flatten_params(params), device=cuda)
    N = B * T
    f32 = dict(device=cuda, dtype=torch.float32)
    a1 = torch.empty(N, 20, 20, 16, **f32); a2 = torch.empty(N, 2592, **f32)
    h = torch.empty(N, 256, **f32); lg = torch.empty(N, A, **f32)
    pr = torch.empty(N, A, **f32); v = torch.empty(N, **f32)
    fc_w = torch.empty(pkg._cabi.prepared_floats(), **f32)
    pkg._cabi.call("arl_forward", flat.data_ptr(), fc_w.data_ptr(), 1, A, ring.data_ptr(), B, R, first,
                   T, a1.data_ptr(), a2.data_ptr(), h.data_ptr(), lg.data_ptr(), pr.data_ptr(),
                   v.data_ptr(), pkg._cabi.stream_ptr())
    torch.cuda.synchronize()
    a1 = pkg.network.decode_a1(a1)                                # device layout: fp16, blocked
    a2 = pkg.network.decode_split(a2, N, 2592)                    # one split block of the call's N rows
    assert rel_err(pkg.network.decode_split(fc_w[:2592 * 256], 2592, 256).cpu(), params["l4_w"]) <= 1e-5
    logits, value, keep = a3c.forward(a3c.to_torch(params), ring_stacks(ring_np, first, T), keep=True)
    pi, _, _ = a3c.policy_terms(logits)
    errs = dict(a1=rel_err(a1.cpu(), keep["a1"]), a2=rel_err(a2.cpu(), keep["a2"]),
                h=rel_err(h.cpu(), keep["h"]), logits=rel_err(lg.cpu(), logits),
                value=rel_err(v.cpu(), value), probs=rel_err(pr.cpu(), pi))
    print("forward rel-err", errs)
    assert max(errs.values()) <= REL_TOL, errs
    assert float(a1.min()) >= 0 and float(h.min()) >= 0 and float((a2 == 0).float().mean()) > 0.05


def test_ops_wrappers_layer_by_layer(pkg, cuda):
    """conv2d / linear / heads / batch_sample of src/ops.py against the oracle."""
    A, B = 6, 6
    params = make_params(A, seed=3, scale=2.0)
    cfg = pkg.config.get_config({"model": "m1", "num_envs": B, "t_max": 5})
    hist = pkg.History(cfg, num_envs=B, device=cuda)
    rng = np.random.default_rng(9)
    hist.ring.copy_(torch.as_tensor(rng.integers(0, 256, tuple(hist.ring.shape), dtype=np.uint8)))
    flat = torch.as_tensor(a3c.flatten_params(params), device=cuda)
    ops = pkg.ops
    l1 = ops.conv2d(hist, flat, 16, [8, 8], [4, 4], name='l1')
    l2 = ops.conv2d(l1, flat, 32, [4, 4], [2, 2], name='l2')
    l4 = ops.linear(l2, flat, 256, name='l4')
    # a plain float32 activation matrix is accepted too (encoded into a split block on the host)
    l4_plain = ops.linear(pkg.network.decode_split(l2, B, 2592), flat, 256, name='l4', input_is_split=False)
    assert rel_err(l4_plain.cpu(), l4.cpu().numpy()) <= 1e-6
    lg, pr, v = ops.heads(l4, flat, A)
    act = ops.batch_sample(pr, step=5, seed=123, env_id_base=10)
    torch.cuda.synchronize()
    stacks = hist.get().cpu().numpy()
    logits, value, keep = a3c.forward(a3c.to_torch(params), stacks, keep=True)
    assert rel_err(pkg.network.decode_a1(l1).cpu(), keep["a1"]) <= REL_TOL and rel_err(l4.cpu(), keep["h"]) <= REL_TOL
    assert rel_err(lg.cpu(), logits) <= REL_TOL and rel_err(v.cpu(), value) <= REL_TOL
    ref_act = philox.sample_actions(pr.cpu().numpy(), np.arange(10, 10 + B), 5, 123)
    assert np.array_equal(act.cpu().numpy(), ref_act)
    with pytest.raises(NotImplementedError):
        ops.conv2d(l1, flat, 64, [3, 3], [1, 1], name='l3')


@pytest.mark.parametrize("A,B", [(6, 4096), (18, 1000), (4, 33)])
def test_sampler_bit_exact_on_identical_probs(pkg, cuda, A, B):
    g = torch.Generator(device=cuda).manual_seed(A)
    probs = torch.softmax(torch.randn(B, A, device=cuda, generator=g) * 2, dim=1).contiguous()
    for step, seed, base in [(0, 123, 0), (77, 123, 4096), (2 ** 33 + 5, 2 ** 40 + 9, 12345)]:
        act = pkg.ops.batch_sample(probs, step=step, seed=seed, env_id_base=base)
        ref = philox.sample_actions(probs.cpu().numpy(), np.arange(base, base + B), step, seed)
        assert np.array_equal(act.cpu().numpy(), ref)
    # degenerate rows
    z = torch.zeros(5, A, device=cuda)
    assert (pkg.ops.batch_sample(z).cpu().numpy() == A - 1).all()
    # argmax: ties -> lowest index
    s = torch.tensor([[1., 3., 3., 0.], [2., 2., 2., 2.]], device=cuda)
    assert pkg.ops.argmax(s).cpu().tolist() == [1, 0]


@pytest.mark.parametrize("A,T,B", [(6, 5, 37), (18, 20, 8), (4, 1, 3)])
def test_returns_and_loss_grads_vs_oracle(pkg, cuda, A, T, B):
    rng = np.random.default_rng(T * 10 + B)
    logits = rng.normal(0, 1.5, (T, B, A)).astype(np.float32)
    value = rng.normal(0, 1, (T, B)).astype(np.float32)
    vboot = rng.normal(0, 1, B).astype(np.float32)
    rew = rng.choice([-3.0, -1.0, 0.0, 0.5, 1.0, 7.0], (T, B)).astype(np.float32)
    term = (rng.random((T, B)) < 0.25)
    acts = rng.integers(0, A, (T, B)).astype(np.int32)
    d = lambda x: torch.as_tensor(x, device=cuda)
    R = torch.empty(T, B, device=cuda); dl = torch.empty(T, B, A, device=cuda)
    dv = torch.empty(T, B, device=cuda); sums = torch.zeros(3, device=cuda)
    scale = 1.0 / B
    g_rew, g_term, g_acts = d(rew), d(term.astype(np.uint8)), d(acts)      # keep alive
    g_logits, g_value, g_vboot = d(logits), d(value), d(vboot)
    pkg._cabi.call("arl_returns_lossgrad", g_rew.data_ptr(), g_term.data_ptr(),
                   g_acts.data_ptr(), g_logits.data_ptr(), g_value.data_ptr(),
                   g_vboot.data_ptr(), R.data_ptr(), dl.data_ptr(), dv.data_ptr(),
                   sums.data_ptr(), T, B, A, 0.99, 0.01, -1.0, 1.0, scale, pkg._cabi.stream_ptr())
    torch.cuda.synchronize()
    Rref = a3c.nstep_returns(a3c.clip_rewards(rew), term, vboot.astype(np.float64), 0.99)
    lt = torch.tensor(logits.reshape(-1, A), dtype=torch.float64)
    vt = torch.tensor(value.reshape(-1), dtype=torch.float64)
    at = torch.tensor(acts.reshape(-1)); Rt = torch.tensor(Rref.reshape(-1))
    dlr, dvr = a3c.analytic_head_grads(lt, vt, at, Rt, 0.01, scale)
    tot, pl, vl = a3c.loss_per_sample(lt, vt, at, Rt, 0.01)
    _, _, ent = a3c.policy_terms(lt)
    errs = dict(R=rel_err(R.cpu(), Rref), dlogits=rel_err(dl.cpu().reshape(-1, A), dlr),
                dvalue=rel_err(dv.cpu().reshape(-1), dvr),
                sums=rel_err(sums.cpu(), [float(pl.sum()), float(vl.sum()), float(ent.sum())]))
    print("returns/loss rel-err", errs)
    assert max(errs.values()) <= REL_TOL, errs


def _gpu_cycle_grads(pkg, cuda, net, hist, rew, term, acts, scale):
    T, B = net.t_max, net.num_envs
    d = lambda x: torch.as_tensor(x, device=cuda)
    v_boot = net.bootstrap_value(hist)
    net.compute_gradients(hist, d(rew.astype(np.float32)), d(term.astype(np.uint8)), v_boot,
                          actions=d(acts.astype(np.int32).reshape(-1)), grad_scale=scale)
    torch.cuda.synchronize()
    return v_boot


@pytest.mark.parametrize("A,B,T", [(6, 7, 5), (18, 3, 2), (6, 100, 5)])
def test_backward_gradients_vs_oracle(pkg, cuda, A, B, T):
    """Whole backward (heads -> fc -> conv2 -> conv1) on a rollout, all 10 tensors, plus the
    intermediate input-gradients checked layer by layer with float64 autograd."""
    rng = np.random.default_rng(A + B)
    params = make_params(A, seed=11, scale=2.0)
    net = net_for(pkg, A, B, T, params)
    cfg = pkg.config.get_config({"model": "m1", "num_envs": B, "t_max": T})
    hist = pkg.History(cfg, num_envs=B, device=cuda)
    screens = rng.integers(0, 256, (T + 4, B, 84, 84), dtype=np.uint8)
    for k in range(4):
        hist.add(torch.as_tensor(screens[k], device=cuda))
    for t in range(T):
        net.forward(hist, t)
        hist.add(torch.as_tensor(screens[4 + t], device=cuda))
    acts = rng.integers(0, A, (T, B)); rew = rng.choice([-2.0, 0.0, 1.0], (T, B))
    term = rng.random((T, B)) < 0.2
    v_boot = _gpu_cycle_grads(pkg, cuda, net, hist, rew, term, acts, 1.0 / B)

    stacks = a3c.stacks_from_screens(screens, T)
    p64 = a3c.to_torch(params)
    with torch.no_grad():
        _, vb = a3c.forward(p64, stacks[T])
    assert rel_err(v_boot.cpu(), vb) <= REL_TOL
    R = a3c.nstep_returns(a3c.clip_rewards(rew), term, vb.numpy(), 0.99)
    assert rel_err(net.R.cpu().reshape(T, B), R) <= REL_TOL
    # (1) with the device's relu pattern forced in the oracle: every arithmetic path, max-norm
    grads, aux = a3c.gradients(params, stacks[:T].reshape(T * B, 84, 84, 4), acts.reshape(-1),
                               R.reshape(-1), 0.01, B, masks=gpu_masks(net))
    errs = {k: rel_err(net.g[k].cpu(), grads[k]) for k in a3c.PARAM_NAMES}
    print("grad rel-err", errs)
    assert max(errs.values()) <= REL_TOL, errs
    # the gradients handed from layer to layer, decoded from their device layouts
    N = T * B
    assert bool((net.d_h() == net.d_h_transposed()).all())     # the two device copies of d_h agree
    ierrs = dict(d_h=rel_err(net.d_h().cpu(), aux["d_h"]), d_a2=rel_err(net.d_a2().cpu(), aux["d_a2"]),
                 d_a1=rel_err(net.d_a1().cpu(), aux["d_a1"]))
    print("layer-gradient rel-err", ierrs)
    assert max(ierrs.values()) <= REL_TOL, ierrs
    raw = net.d_l1.reshape(-1)[:N * 3528].view(torch.float16).reshape(2, N, 21, 21, 8)
    assert float(raw[:, :, 20].abs().max()) == 0.0 and float(raw[:, :, :, 20].abs().max()) == 0.0
    # (2) free oracle (REPORTED; loosely gated): a pre-activation within rounding distance of 0 --
    # ~2e-4 relative now that a1 is stored as fp16 -- flips one relu and with it one gradient
    # column; with only T*B samples in the sums a handful of flips is visible in the 2-norm
    # (tests/precision_study.py: 250 flips in 11.8 M relus, 1.3e-2 on l2_w, against 4.4e-4 with the
    # pattern forced).  The gated comparison is (1); the 100-update free trajectory is its own test.
    grads_free, _ = a3c.gradients(params, stacks[:T].reshape(T * B, 84, 84, 4), acts.reshape(-1),
                                  R.reshape(-1), 0.01, B)
    nerrs = {k: norm_err(net.g[k].cpu(), grads_free[k]) for k in a3c.PARAM_NAMES}
    print("grad 2-norm rel-err (free relu)", nerrs)
    assert max(nerrs.values()) <= 5e-2, nerrs
    assert rel_err(net.policy_logits.cpu(), aux["logits"]) <= REL_TOL
    assert rel_err(net.value.cpu(), aux["value"]) <= REL_TOL


def test_backward_single_sample_and_zero_samples(pkg, cuda):
    A, B, T = 6, 1, 1
    params = make_params(A, seed=2, scale=2.0)
    net = net_for(pkg, A, B, T, params)
    cfg = pkg.config.get_config({"model": "m1", "num_envs": B, "t_max": T})
    hist = pkg.History(cfg, num_envs=B, device=cuda)
    rng = np.random.default_rng(1)
    screens = rng.integers(0, 256, (T + 4, B, 84, 84), dtype=np.uint8)
    for k in range(4):
        hist.add(torch.as_tensor(screens[k], device=cuda))
    net.forward(hist, 0)
    hist.add(torch.as_tensor(screens[4], device=cuda))
    acts = np.array([[2]]); rew = np.array([[1.0]]); term = np.array([[False]])
    _gpu_cycle_grads(pkg, cuda, net, hist, rew, term, acts, 1.0)
    stacks = a3c.stacks_from_screens(screens, T)
    with torch.no_grad():
        _, vb = a3c.forward(a3c.to_torch(params), stacks[T])
    R = a3c.nstep_returns(a3c.clip_rewards(rew), term, vb.numpy(), 0.99)
    grads, _ = a3c.gradients(params, stacks[0], acts.reshape(-1), R.reshape(-1), 0.01, 1)
    errs = {k: rel_err(net.g[k].cpu(), grads[k]) for k in a3c.PARAM_NAMES}
    assert max(errs.values()) <= REL_TOL, errs
    # zero samples: every backward entry point must zero its gradient slice and return OK
    lib = pkg._cabi.load()
    net.grads.fill_(5.0)
    st = pkg._cabi.stream_ptr()
    P = pkg._cabi.ptr
    assert lib.arl_backward(P(net.params), P(net.fc_w), A, P(hist.ring), 0, hist.ring_slots, 0, 1, P(net.l1),
                            P(net.l2), P(net.l4), P(net.d_logits), P(net.d_value), P(net.d_l4),
                            P(net.d_l2), P(net.d_l1), P(net.grads), P(net.workspace), 1.0, 0, st) == 0
    torch.cuda.synchronize()
    assert float(net.grads.abs().max()) == 0.0


@pytest.mark.parametrize("A", [6, 18])
def test_clip_rmsprop_vs_oracle(pkg, cuda, A):
    rng = np.random.default_rng(A)
    params = make_params(A, seed=4)
    net = net_for(pkg, A, 2, 1, params)
    shapes = a3c.param_shapes(A)
    grads = {k: rng.normal(0, 1.0, s).astype(np.float32) for k, s in shapes.items()}
    grads["l4_w"] *= 0.2                      # norm ~163 -> clipped; q_b norm ~1 -> untouched
    grads["l1_b"] *= 100.0                    # norm ~400 -> clipped
    rms0 = {k: rng.uniform(0.5, 2.0, s).astype(np.float32) for k, s in shapes.items()}
    net.grads.copy_(torch.as_tensor(a3c.flatten_params(grads)))
    net.rms.copy_(torch.as_tensor(a3c.flatten_params(rms0)))
    lr = 0.0007
    net.apply_gradients(lr)
    torch.cuda.synchronize()
    newp, newr = a3c.update(params, rms0, grads, lr)
    for i, k in enumerate(a3c.PARAM_NAMES):
        nrm = float(np.sqrt((grads[k].astype(np.float64) ** 2).sum()))
        assert abs(float(net.grad_norms[i]) - nrm) <= 1e-4 * nrm
        lo, hi = net.offsets[i], net.offsets[i + 1]
        assert rel_err(net.rms[lo:hi].cpu().reshape(shapes[k]), newr[k]) <= 1e-5
        step_ref = newp[k] - params[k].astype(np.float64)
        step_gpu = net.w[k].cpu().numpy().astype(np.float64) - params[k].astype(np.float64)
        assert rel_err(step_gpu, step_ref) <= REL_TOL, k       # the update itself, not just w


def _run_trajectory(pkg, cuda, A, B, T, updates, seed=123, free=False):
    """``free=False``: the oracle is given the device's relu pattern (every arithmetic path, max
    norm).  ``free=True``: the oracle runs on its own (no pattern forced); the comparison is in the
    2-norm and the relus whose state differs between device and oracle are counted."""
    params = make_params(A, seed=21, bias_std=0.0)
    cfg = pkg.config.get_config({"model": "m1", "num_envs": B, "t_max": T, "seed": seed})
    env = pkg.GymEnvironment(cfg, env=pkg.SyntheticAtari(B, action_size=A, seed=seed, pool=T + 3,
                                                         device=cuda, p_terminal=0.05,
                                                         reward_scale=2.0), device=cuda)
    agent = pkg.Agent(cfg, env, device=cuda)
    agent.network.set_weights(params)
    p_ref = {k: v.astype(np.float64) for k, v in params.items()}
    r_ref = {k: np.ones_like(v, np.float64) for k, v in params.items()}
    p0 = {k: v.copy() for k, v in p_ref.items()}
    agent.before_train()
    worst = dict(param=0.0, disp=0.0, grad=0.0, per_tensor={})
    for u in range(updates):
        ring = agent.history
        screens = []
        first = ring.first_slot(0)
        snap = ring.planes().cpu().numpy()
        screens = [snap[:, (first + k) % ring.ring_slots] for k in range(4)]
        step0 = agent.step
        for t in range(T):
            action = agent.predict()
            scr, rew, term = env.act(action, is_training=True, fused=True)
            agent.history.add(scr)
            agent.batch_reward[t].copy_(rew)
            agent.batch_terminal[t].copy_(term)
            screens.append(agent.history.planes(agent.history.head).cpu().numpy())
            agent.t += 1
            agent.step += 1
        agent.step -= 1                                          # lr uses the step of the last frame
        agent.batch_update()
        agent.step += 1
        torch.cuda.synchronize()
        acts = agent.batch_action.cpu().numpy()
        dev_masks = gpu_masks(agent.network)
        if free:
            stacks = a3c.stacks_from_screens(np.stack(screens), T)
            with torch.no_grad():
                _, _, keep = a3c.forward(a3c.to_torch(p_ref), stacks[:T].reshape(T * B, 84, 84, 4), keep=True)
            flips = sum(int(((keep[k].numpy() > 0) != dev_masks[k]).sum()) for k in ("a1", "a2", "h"))
            worst["flips"] = worst.get("flips", 0) + flips
            worst["relus"] = worst.get("relus", 0) + sum(int(m.size) for m in dev_masks.values())
        p_ref, r_ref, aux = a3c.a3c_cycle(p_ref, r_ref, np.stack(screens), acts,
                                          agent.batch_reward.cpu().numpy(),
                                          agent.batch_terminal.cpu().numpy().astype(bool), step0,
                                          num_envs=B, masks=None if free else dev_masks)
        flat_gpu = agent.network.params.cpu().numpy().astype(np.float64)
        flat_ref = a3c.flatten_params(p_ref)
        worst["flat"] = max(worst.get("flat", 0.0), rel_err(flat_gpu, flat_ref))
        worst["flat_2norm"] = max(worst.get("flat_2norm", 0.0), norm_err(flat_gpu, flat_ref))
        worst["grad_2norm"] = max(worst.get("grad_2norm", 0.0), max(
            norm_err(agent.network.g[k].cpu(), aux["grads"][k]) for k in a3c.PARAM_NAMES if k.endswith("_w")))
        flat0 = a3c.flatten_params(p0)
        worst["disp_2norm"] = norm_err(flat_gpu - flat0, flat_ref - flat0)
        for k in a3c.PARAM_NAMES:
            w = agent.network.w[k].cpu().numpy().astype(np.float64)
            abs_err = float(np.abs(w - p_ref[k]).max())
            if k.endswith("_w"):
                worst["param"] = max(worst["param"], rel_err(w, p_ref[k]))
                ge = rel_err(agent.network.g[k].cpu(), aux["grads"][k])
                if ge > worst["grad"]:
                    worst["grad"], worst["grad_at"] = ge, (u, k, float(np.abs(aux["grads"][k]).max()))
            else:
                # zero-initialised biases: scale = larger of the tensor and one lr-sized step
                scale = max(float(np.abs(p_ref[k]).max()), 0.0007)
                worst["bias"] = max(worst.get("bias", 0.0), abs_err / scale)
            if u == updates - 1:
                worst["disp"] = max(worst["disp"], rel_err(w - p0[k], p_ref[k] - p0[k]))
                worst["per_tensor"][k] = (rel_err(w, p_ref[k]), abs_err)
    return worst


def test_agent_cycle_short_trajectory(pkg, cuda):
    worst = _run_trajectory(pkg, cuda, A=6, B=16, T=5, updates=5)
    print("5-update trajectory", worst)
    assert max(worst["param"], worst["flat"], worst["bias"], worst["grad"]) <= REL_TOL, worst


def test_rmsprop_trajectory_100_updates_config2(pkg, cuda):
    """BASELINE config 2: 256 envs, history 4, t_max 5; parameters within 1e-3 of the float64
    oracle trajectory over 100 updates (teacher-forced actions, independent parameter copies)."""
    worst = _run_trajectory(pkg, cuda, A=6, B=256, T=5, updates=100)
    print("100-update trajectory", worst)
    # whole parameter vector and every weight tensor, at every one of the 100 updates
    assert max(worst["flat"], worst["param"], worst["bias"]) <= REL_TOL, worst
    # every tensor (biases included) at the end of the run
    assert max(v[0] for v in worst["per_tensor"].values()) <= REL_TOL, worst


def test_rmsprop_trajectory_100_updates_free_oracle(pkg, cuda):
    """VERDICT r1 #8a: the same 100 updates against the FREE float64 oracle (no relu pattern is
    forced; each side keeps its own parameters).  Gate: parameters within 1e-3 in the 2-norm at
    every update.  Reported: the number of relus whose state differs (a pre-activation within
    rounding distance of 0), the worst gradient 2-norm error (one flipped relu moves a whole
    gradient column, so this is NOT gated at 1e-3: tests/precision_study.py shows 1.3e-3 for the
    float32-grade storage against float64 from 6 flips in 11.8 M) and the error of the
    displacement from the start."""
    worst = _run_trajectory(pkg, cuda, A=6, B=256, T=5, updates=100, free=True)
    print("100-update FREE trajectory", {k: v for k, v in worst.items() if k != "per_tensor"})
    assert worst["flat_2norm"] <= REL_TOL, worst
    assert max(v[0] for k, v in worst["per_tensor"].items() if k.endswith("_w")) <= REL_TOL, worst
    # zero-initialised biases are a few lr-sized steps large: against the FREE oracle their error is
    # the flipped relus' (measured 1.0e-3 .. 1.3e-3 of the larger of the tensor and one lr step)
    assert worst["bias"] <= 5e-3, worst
    # a1 is stored as fp16 (2^-12 per element): a pre-activation within ~2e-4 of 0 can land on the
    # other side -- measured 3.2e-5 of the 1.2e9 relus (with the bf16 hi+lo storage of round 1: 5e-7)
    assert worst["flips"] <= 1e-4 * worst["relus"], worst
    assert worst["disp_2norm"] <= 2e-2, worst


@pytest.mark.parametrize("A,B,T", [(6, 4096, 5), (18, 4096, 20)])
def test_full_size_properties(pkg, cuda, A, B, T):
    """BASELINE config 3 size (4096 envs, t_max 5, A=6) and config 4's largest head / rollout
    (18 actions, t_max 20), size-independent properties:
    (1) forward of a random subset of the 20 480 samples equals the oracle,
    (2) the gradient is additive over samples: full batch == first half + second half of the
        envs (different tiling, split-K ranges and partial counts on the device),
    (3) two runs are bit-identical (no float atomics on the gradient path)."""
    params = make_params(A, seed=5, scale=2.0)
    g = torch.Generator(device=cuda).manual_seed(3)

    def run(envs, ring_full=None, lo=0):
        cfg = pkg.config.get_config({"model": "m1", "num_envs": envs, "t_max": T})
        net = net_for(pkg, A, envs, T, params)
        hist = pkg.History(cfg, num_envs=envs, device=cuda)
        if ring_full is None:
            hist.ring.copy_(torch.randint(0, 256, tuple(hist.ring.shape), dtype=torch.uint8,
                                          device=cuda, generator=g))
        else:
            hist.ring.copy_(ring_full[lo:lo + envs])
        hist.head = T + 3                                         # slots 0..T+3 = f_-3 .. f_T
        for t in range(T):
            pkg._cabi.call("arl_forward", pkg._cabi.ptr(net.params), pkg._cabi.ptr(net.fc_w), 1, A,
                           pkg._cabi.ptr(hist.ring), envs, hist.ring_slots, t, 1, *[pkg._cabi.ptr(x[t * envs:(t + 1) * envs])
                                                          for x in (net.l1, net.l2, net.l4,
                                                                    net.policy_logits, net.policy,
                                                                    net.value)],
                           pkg._cabi.stream_ptr())
        return net, hist

    net, hist = run(B)
    rng = np.random.default_rng(0)
    acts = torch.as_tensor(rng.integers(0, A, T * B).astype(np.int32), device=cuda)
    rew = torch.as_tensor(rng.choice([-1.0, 0.0, 1.0], (T, B)).astype(np.float32), device=cuda)
    term = torch.as_tensor((rng.random((T, B)) < 0.05).astype(np.uint8), device=cuda)
    v_boot = torch.zeros(B, device=cuda)

    def grads_of(net, hist, a, r, tm, vb):
        net.compute_gradients(hist, r.contiguous(), tm.contiguous(), vb, actions=a.contiguous(),
                              grad_scale=1.0 / B)
        torch.cuda.synchronize()
        return net.grads.clone()

    g_full = grads_of(net, hist, acts, rew, term, v_boot)
    g_again = grads_of(net, hist, acts, rew, term, v_boot)
    assert bool((g_full == g_again).all())                        # (3)

    # (1) forward of 48 random samples vs the oracle
    ring = hist.planes().cpu().numpy()
    pick = rng.choice(T * B, 48, replace=False)
    stacks = np.stack([np.stack([ring[n % B, (n // B) + k] for k in range(4)], axis=-1) for n in pick])
    logits, value = a3c.forward(a3c.to_torch(params), stacks)
    assert rel_err(net.policy_logits[torch.as_tensor(pick, device=cuda)].cpu(), logits) <= REL_TOL
    assert rel_err(net.value[torch.as_tensor(pick, device=cuda)].cpu(), value) <= REL_TOL

    # (2) additivity over env halves
    a2d, half = acts.view(T, B), B // 2
    total = torch.zeros_like(g_full)
    for lo in (0, half):
        n2, h2 = run(half, hist.ring, lo)
        total += grads_of(n2, h2, a2d[:, lo:lo + half].reshape(-1), rew[:, lo:lo + half],
                          term[:, lo:lo + half], v_boot[lo:lo + half])
    err = float((total - g_full).abs().max()) / float(g_full.abs().max())
    print("additivity rel-err", err)
    # fp32 partial sums in a different order: 1e-5 at 20 480 samples (measured 3e-6), 1.3e-5 at 81 920
    assert err <= (1e-5 if T * B <= 20480 else 5e-5)


def test_train_with_summary_and_checkpoint_resume(pkg, cuda, tmp_path):
    """agent.py:69-139 statistics (reference tag names, JSON lines) and agent.py:29/34 +
    main.py:74-80 checkpoint/resume: a restored agent continues bit-identically."""
    import json
    A, B, T = 6, 12, 5
    cfg = pkg.config.get_config({"model": "m1", "num_envs": B, "t_max": T, "_test_step": 10})

    def make():
        import random
        random.seed(123)
        env = pkg.GymEnvironment(cfg, env=pkg.SyntheticAtari(B, A, seed=9, pool=64, device=cuda,
                                                             p_terminal=0.1), device=cuda)
        return pkg.Agent(cfg, env, device=cuda), env

    agent, env = make()
    log = tmp_path / "stats.jsonl"
    recs = agent.train_with_summary(num_steps=20, log_path=str(log))
    assert len(recs) == 2 and agent.step_op == 20 and agent.update_count == 4
    lines = [json.loads(l) for l in open(log)]
    assert lines == recs
    for key in ("average.reward", "average.loss", "average.q", "episode.max reward",
                "episode.min reward", "episode.avg reward", "episode.num of game",
                "training.learning_rate"):
        assert key in recs[0]
    # the reward statistic equals a host recomputation from the synthetic env's reward table
    r = env.env._rewards.clamp(-1, 1)
    first = env.env._i - 20                                        # pool index of the first timed step
    ref = float(sum(r[(first + i) % env.env.pool].sum() for i in range(10))) / (10 * B)
    assert abs(recs[0]["average.reward"] - ref) < 1e-6
    assert sum(recs[0]["episode.actions"]) == 10 * B and recs[0]["episode.num of game"] >= 0

    # checkpoint -> a fresh agent -> identical continuation
    ck = tmp_path / "ck"
    agent.save_checkpoint(str(ck))
    agent2, env2 = make()
    assert agent2.load_checkpoint(str(ck)) and agent2.step_op == 20
    assert bool((agent2.network.params == agent.network.params).all())
    assert bool((agent2.network.rms == agent.network.rms).all())


def test_chunked_forward_for_very_large_batches(pkg, cuda):
    """More than 16 384 envs per step run as several forward launches over env ranges, each with its
    own a2 block (arl_a2_block_rows): the outputs are those of the two halves run on their own (bit
    for bit: a sample's forward does not depend on its batch) and the gradient is their sum."""
    A, B, T = 6, 18432, 1
    assert pkg._cabi.a2_block_rows(B) == 9216 and pkg._cabi.a2_block_rows(4096) == 4096
    assert pkg._cabi.a2_block_rows(65536) == 16384 and pkg._cabi.a2_block_rows(16385) == 16385
    params = make_params(A, seed=8, scale=2.0)
    g = torch.Generator(device=cuda).manual_seed(5)
    cfg = pkg.config.get_config({"model": "m1", "num_envs": B, "t_max": T})
    hist = pkg.History(cfg, num_envs=B, device=cuda)
    hist.ring.copy_(torch.randint(0, 256, tuple(hist.ring.shape), dtype=torch.uint8, device=cuda, generator=g))
    rng = np.random.default_rng(1)
    acts = torch.as_tensor(rng.integers(0, A, B).astype(np.int32), device=cuda)
    rew = torch.as_tensor(rng.choice([-1.0, 0.0, 1.0], (T, B)).astype(np.float32), device=cuda)
    term = torch.as_tensor((rng.random((T, B)) < 0.05).astype(np.uint8), device=cuda)

    def run(envs, lo):
        c = pkg.config.get_config({"model": "m1", "num_envs": envs, "t_max": T})
        net = net_for(pkg, A, envs, T, params)
        h = pkg.History(c, num_envs=envs, device=cuda)
        h.ring.copy_(hist.ring[lo:lo + envs])
        h.head = hist.head
        sampled = net.forward_sample(h, 0, step=7, seed=123, env_id_base=lo).clone()
        h.add(torch.zeros(envs, 84, 84, dtype=torch.uint8, device=cuda))        # one push: s_0 is 1 back
        net.compute_gradients(h, rew[:, lo:lo + envs].contiguous(), term[:, lo:lo + envs].contiguous(),
                              torch.zeros(envs, device=cuda), actions=acts[lo:lo + envs].contiguous(),
                              grad_scale=1.0 / B)
        torch.cuda.synchronize()
        return net, sampled

    full, s_full = run(B, 0)
    total = torch.zeros_like(full.grads)
    for lo in (0, B // 2):
        half, s_half = run(B // 2, lo)
        r = slice(lo, lo + B // 2)
        assert torch.equal(half.policy_logits, full.policy_logits[r]) and torch.equal(half.value, full.value[r])
        assert torch.equal(s_half, s_full[r])
        total += half.grads
    err = float((total - full.grads).abs().max()) / float(full.grads.abs().max())
    print("chunked forward: gradient additivity", err)
    assert err <= 1e-5
