"""CPU: the network oracle (oracle/a3c.py::forward) against the golden vectors made by EXECUTING
the reference's own layer functions (src/ops.py conv2d / linear, over oracle/tf_stub.py) with the
literal arguments of its call sites -- oracle/make_golden_network.py.  This pins the wiring the
reference contributes (kernel shapes, stride lists, NHWC flatten order, [in,out] matrices, bias
and activation order); the arithmetic inside tf.nn.conv2d / tf.matmul stays a restatement (the
stub's header says so), which is why DESIGN.md keeps "parity unpinned" for TF's numerics."""
import os

import numpy as np
import pytest
import torch

from oracle import a3c
from oracle.make_golden_network import golden_stacks, golden_weights
from util import rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def net_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "network_golden.npz"))


def test_inputs_are_reproducible(net_golden):
    A = int(net_golden["action_size"])
    p, s = golden_weights(A), golden_stacks()
    assert int(s.astype(np.int64).sum()) == int(net_golden["stacks_sum"])
    sums = np.array([float(np.abs(p[k].astype(np.float64)).sum()) for k in sorted(p)])
    assert np.allclose(sums, net_golden["weights_sum"], rtol=1e-12, atol=0)
    assert {k: v.shape for k, v in p.items()} == dict(a3c.param_shapes(A))


def test_reference_requested_the_flat_buffer_layouts(net_golden):
    # the shapes ops.py asked tf.get_variable for == the C-ABI's parameter layout (TF layouts)
    A = int(net_golden["action_size"])
    req = dict(r.rsplit(" (", 1) for r in net_golden["requested"].tolist())
    want = {"l1/w": a3c.param_shapes(A)["l1_w"], "l2/w": a3c.param_shapes(A)["l2_w"],
            "l3/Matrix": a3c.param_shapes(A)["l4_w"], "policy/linear/Matrix": a3c.param_shapes(A)["p_w"],
            "value/linear/Matrix": a3c.param_shapes(A)["q_w"]}
    for path, shape in want.items():
        got = tuple(int(v) for v in req[path].rstrip(")").split(",") if v.strip())
        assert got == tuple(shape), (path, got, shape)


def test_forward_matches_executed_reference_layers(net_golden):
    A = int(net_golden["action_size"])
    p = a3c.to_torch(golden_weights(A), dtype=torch.float64)
    logits, value, keep = a3c.forward(p, golden_stacks(), keep=True)
    assert rel_err(keep["a1"].numpy(), net_golden["a1"]) < 1e-12
    assert rel_err(keep["a2"].numpy(), net_golden["a2"]) < 1e-12
    assert rel_err(keep["h"].numpy(), net_golden["h"]) < 1e-12
    assert rel_err(logits.numpy(), net_golden["logits"]) < 1e-12
    assert rel_err(value.numpy(), net_golden["value"].reshape(-1)) < 1e-12
    pi, logpi, ent = a3c.policy_terms(logits)
    assert rel_err(pi.numpy(), net_golden["policy"]) < 1e-12
    # non-degenerate fixture: about half of every relu layer is active
    for k in ("a1", "a2", "h"):
        assert 0.3 < float((net_golden[k] > 0).mean()) < 0.7


def test_loss_matches_executed_reference_lines(net_golden):
    # network.py:60-94 executed verbatim, one sample at a time (see make_golden_network.py)
    logits = torch.as_tensor(net_golden["logits"])
    value = torch.as_tensor(net_golden["value"]).reshape(-1)
    actions = torch.as_tensor(net_golden["actions"])
    R = torch.as_tensor(net_golden["returns"])
    total, pol, val = a3c.loss_per_sample(logits, value, actions, R, beta=float(net_golden["beta"]))
    pi, logpi, ent = a3c.policy_terms(logits)
    rows = net_golden["loss_rows"]
    assert rel_err(pol.numpy(), rows[:, 0]) < 1e-12
    assert rel_err(val.numpy(), rows[:, 1]) < 1e-12
    assert rel_err(total.numpy(), rows[:, 2]) < 1e-12
    assert rel_err(ent.numpy(), rows[:, 3]) < 1e-12
    logp_a = logpi.gather(1, actions.reshape(-1, 1).long()).reshape(-1)
    assert rel_err(logp_a.numpy(), rows[:, 4]) < 1e-12
    # the closed-form gradients the CUDA loss kernel emits are the autograd gradients of that loss
    lg = logits.clone().requires_grad_(True)
    vv = value.clone().requires_grad_(True)
    a3c.loss_per_sample(lg, vv, actions, R, beta=float(net_golden["beta"]))[0].sum().backward()
    dl, dv = a3c.analytic_head_grads(logits, value, actions, R, beta=float(net_golden["beta"]))
    assert rel_err(dl.numpy(), lg.grad.numpy()) < 1e-12 and rel_err(dv.numpy(), vv.grad.numpy()) < 1e-12


def test_agent_scalar_formulas_match_executed_reference_lines(net_golden):
    # agent.py:142-144, 154, 188-190, 395 and 310-314 executed verbatim with the reference's config.M1
    g = net_golden
    assert np.allclose([a3c.epsilon(int(s)) for s in g["steps"]], g["eps"], rtol=1e-15, atol=0)
    assert np.allclose([a3c.learning_rate(int(s)) for s in g["steps"]], g["lrs"], rtol=1e-15, atol=0)
    assert np.array_equal(a3c.clip_rewards(g["raw_rewards"]), g["clipped"])
    tq = a3c.q_targets(g["q_next"], g["tq_rewards"], g["tq_terminals"])
    assert np.allclose(tq, g["target_q"], rtol=1e-15, atol=0)
    loss, delta = a3c.async_q_loss(torch.as_tensor(g["q_values"]), torch.as_tensor(g["q_actions"]),
                                   torch.as_tensor(g["target_q"]))
    assert abs(float(loss) - float(g["q_loss"])) <= 1e-15 * abs(float(g["q_loss"])) + 1e-18
    assert np.allclose(delta.numpy(), g["q_delta"], rtol=1e-15, atol=0)


def test_hyperparameters_are_the_ones_the_reference_passes(net_golden, pkg):
    # main.py:64-65 (RMSPropOptimizer arguments), agent.py:319 (clip_by_norm), config.py (M1):
    # the oracle's defaults and the product's config mirror carry the same numbers
    import inspect
    g = net_golden
    d = {k: v.default for k, v in inspect.signature(a3c.rmsprop_apply).parameters.items()
         if v.default is not inspect.Parameter.empty}
    assert d["decay"] == float(g["rms_decay"]) and d["eps"] == float(g["rms_epsilon"])
    assert float(g["rms_momentum"]) == 0.0                      # the momentum slot stays zero: not modelled
    assert inspect.signature(a3c.clip_by_norm).parameters["clip"].default == float(g["clip_norm"])
    lr = inspect.signature(a3c.learning_rate).parameters
    assert lr["max_step"].default == int(g["max_step"]) and lr["base"].default == float(g["base_lr"])
    assert inspect.signature(a3c.nstep_returns).parameters["gamma"].default == float(g["discount"])
    assert inspect.signature(a3c.loss_per_sample).parameters["beta"].default == float(g["cfg_beta"])
    cfg = pkg.config.M1
    assert (cfg.decay, cfg.epsilon, cfg.momentum, cfg.clip_norm) == (
        float(g["rms_decay"]), float(g["rms_epsilon"]), float(g["rms_momentum"]), float(g["clip_norm"]))
    assert (cfg.max_step, cfg.learning_rate, cfg.discount, cfg.beta) == (
        int(g["max_step"]), float(g["base_lr"]), float(g["discount"]), float(g["cfg_beta"]))


# ---- the 'nature' trunk (network.py:30-42), fixture made by oracle/make_golden_nature.py ---------
def test_nature_forward_matches_executed_reference_lines():
    """The reference's own statements network.py:31-40 (three conv2d calls under 'Nature_DQN') +
    linear(512) on the NHWC flatten + the heads, executed over the TF stub, against the oracle's
    nature branch."""
    from oracle.make_golden_nature import golden_weights as nature_weights
    g = np.load(os.path.join(ROOT, "tests", "golden", "nature_golden.npz"))
    A = int(g["action_size"])
    w = nature_weights(A)
    assert {k: v.shape for k, v in w.items()} == dict(a3c.param_shapes(A, "nature"))
    assert list(w) == list(a3c.NATURE_PARAM_NAMES) == list(a3c.names_of(w))
    sums = np.array([float(np.abs(w[k].astype(np.float64)).sum()) for k in sorted(w)])
    assert np.allclose(sums, g["weights_sum"], rtol=1e-12, atol=0)
    logits, value, keep = a3c.forward(a3c.to_torch(w, dtype=torch.float64), golden_stacks(), keep=True)
    assert rel_err(keep["a1"].numpy(), g["a1"]) < 1e-6 and rel_err(keep["a2"].numpy(), g["a2"]) < 1e-6   # stored as float32
    assert rel_err(keep["a3"].numpy(), g["a3"]) < 1e-12 and rel_err(keep["h"].numpy(), g["h"]) < 1e-12
    assert rel_err(logits.numpy(), g["logits"]) < 1e-12 and rel_err(value.numpy(), g["value"].reshape(-1)) < 1e-12
    req = dict(r.rsplit(" (", 1) for r in g["requested"].tolist())
    assert req["Nature_DQN/l3_conv/w"].startswith("3, 3, 64, 64") and req["Nature_DQN/l4_linear/Matrix"].startswith("3136, 512")


def test_nature_oracle_gradients_are_consistent():
    """Closed-form head gradients == autograd for the nature branch too, and every tensor gets one."""
    rng = np.random.default_rng(3)
    A, N = 5, 3
    p = a3c.init_params(A, seed=2, trunk="nature")
    stacks = rng.integers(0, 256, (N, 84, 84, 4), dtype=np.uint8)
    acts, R = rng.integers(0, A, N), rng.normal(0, 1, N)
    grads, aux = a3c.gradients(p, stacks, acts, R, 0.01, N)
    assert set(grads) == set(a3c.NATURE_PARAM_NAMES)
    assert aux["d_a1"].shape == (N, 20, 20, 32) and aux["d_a2"].shape == (N, 9, 9, 64)
    assert aux["d_a3"].shape == (N, 3136) and aux["d_h"].shape == (N, 512)
    dl, dv = a3c.analytic_head_grads(torch.as_tensor(aux["logits"]), torch.as_tensor(aux["value"]),
                                     torch.as_tensor(acts), torch.as_tensor(R), 0.01, 1.0 / N)
    h = a3c.forward(a3c.to_torch(p), stacks, keep=True)[2]["h"].numpy()
    assert rel_err(h.T @ dl.numpy(), grads["p_w"]) < 1e-10 and rel_err(dv.numpy().sum().reshape(1), grads["q_b"]) < 1e-10
    new_p, new_r = a3c.update(p, {k: np.ones_like(v) for k, v in p.items()}, grads, 0.0007)
    assert set(new_p) == set(a3c.NATURE_PARAM_NAMES)
