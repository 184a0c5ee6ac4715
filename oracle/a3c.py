"""REF-A3C step, CPU restatement in torch (float64 by default).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  "parity unpinned": the
arithmetic below lives in TensorFlow 0.x, which is neither vendored under
/root/reference nor installable here, and the reference ships no tests or golden
vectors for it.  Each function follows the cited reference call site; where the
reference's A3C ``Network`` is broken/dead the repair is the
paper's Algorithm 3 (assets/a3c.png) and is stated inline.

Layouts are the reference's own, so TF-style weights load unpermuted:
  l1_w [8,8,4,16]  l1_b [16]     conv 8x8 s4 VALID, NHWC, cross-correlation  (ops.py:19-25)
  l2_w [4,4,16,32] l2_b [32]     conv 4x4 s2 VALID                           (agent.py:228-229)
  l4_w [2592,256]  l4_b [256]    fc after NHWC flatten (h*9+w)*32+c          (agent.py:231-232,251)
  p_w  [256,A]     p_b  [A]      policy logits                               (network.py:62)
  q_w  [256,1]     q_b  [1]      value (named q_* in Network, network.py:79)
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

PARAM_NAMES = ("l1_w", "l1_b", "l2_w", "l2_b", "l4_w", "l4_b", "p_w", "p_b", "q_w", "q_b")
# the 'nature' trunk of network.py:30-42: conv32 8x8 s4, conv64 4x4 s2, conv64 3x3 s1, fc512
NATURE_PARAM_NAMES = ("l1_w", "l1_b", "l2_w", "l2_b", "l3_w", "l3_b", "l4_w", "l4_b", "p_w", "p_b",
                      "q_w", "q_b")


def param_shapes(action_size, trunk="nips"):
    if trunk == "nature":
        return OrderedDict([
            ("l1_w", (8, 8, 4, 32)), ("l1_b", (32,)),               # network.py:34-35
            ("l2_w", (4, 4, 32, 64)), ("l2_b", (64,)),              # network.py:36-37
            ("l3_w", (3, 3, 64, 64)), ("l3_b", (64,)),              # network.py:38-39
            ("l4_w", (3136, 512)), ("l4_b", (512,)),                # network.py:40-42 (flatten repaired)
            ("p_w", (512, action_size)), ("p_b", (action_size,)),
            ("q_w", (512, 1)), ("q_b", (1,)),
        ])
    return OrderedDict([
        ("l1_w", (8, 8, 4, 16)), ("l1_b", (16,)),
        ("l2_w", (4, 4, 16, 32)), ("l2_b", (32,)),
        ("l4_w", (2592, 256)), ("l4_b", (256,)),
        ("p_w", (256, action_size)), ("p_b", (action_size,)),
        ("q_w", (256, 1)), ("q_b", (1,)),
    ])


def param_count(action_size, trunk="nips"):
    return sum(int(np.prod(s)) for s in param_shapes(action_size, trunk).values())


def init_params(action_size, seed=123, dtype=np.float32, trunk="nips"):
    """agent.py:214 / network.py:10 truncated_normal(0, .02) for the convs (values beyond
    2 sigma are re-drawn), ops.py:36-39 random_normal(stddev=.02) for ``linear`` matrices,
    biases 0 (ops.py:24, 38-39).  RNG is numpy's, not TF's: only the distribution matches."""
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for name, shape in param_shapes(action_size, trunk).items():
        if name.endswith("_b"):
            out[name] = np.zeros(shape, dtype)
        elif len(shape) == 4:                                     # conv weights: truncated normal
            w = rng.normal(0.0, 0.02, shape)
            bad = np.abs(w) > 0.04
            while bad.any():
                w[bad] = rng.normal(0.0, 0.02, int(bad.sum()))
                bad = np.abs(w) > 0.04
            out[name] = w.astype(dtype)
        else:
            out[name] = rng.normal(0.0, 0.02, shape).astype(dtype)
    return out


def names_of(params):
    """The flat-buffer tensor order of a parameter dict (nips or nature trunk)."""
    return NATURE_PARAM_NAMES if "l3_w" in params else PARAM_NAMES


def flatten_params(params):
    return np.concatenate([np.asarray(params[n]).reshape(-1) for n in names_of(params)])


def unflatten_params(flat, action_size, trunk="nips"):
    out, o = OrderedDict(), 0
    for name, shape in param_shapes(action_size, trunk).items():
        n = int(np.prod(shape))
        out[name] = np.asarray(flat[o:o + n]).reshape(shape)
        o += n
    assert o == len(flat)
    return out


def to_torch(params, dtype=torch.float64, requires_grad=False):
    return OrderedDict((k, torch.tensor(np.asarray(v), dtype=dtype, requires_grad=requires_grad))
                       for k, v in params.items())


def _relu(x, mask):
    """relu, or -- when ``mask`` is given -- x * mask.  Forcing the activation pattern observed
    on the device removes the one discontinuity of the net from a parity comparison: a
    pre-activation within rounding distance of 0 would otherwise flip a whole gradient column."""
    if mask is None:
        return F.relu(x)
    return x * torch.as_tensor(mask).to(x.dtype)


def forward(p, s_nhwc, keep=False, masks=None):
    """Trunk + heads.  ``s_nhwc``: [N,84,84,4] holding u8 values (any real dtype),
    channel k=0 oldest ... 3 newest (history.py:20-24).

    agent.py:226  x = s_t / 255.
    agent.py:226-229 conv2d(16,[8,8],[4,4]) relu, conv2d(32,[4,4],[2,2]) relu   (ops.py:21-28)
    agent.py:231-232 flatten in NHWC order
    agent.py:251  linear(256, relu)                                             (ops.py:36-44)
    network.py:62 policy_logits = linear(l4, A);  network.py:79 value = linear(l4, 1)
    (value used as [N], repairing the [B]-[B,1] broadcast of network.py:88,91 -- SURVEY D3)
    """
    dt = p["l1_w"].dtype
    x = torch.as_tensor(s_nhwc).to(dt).permute(0, 3, 1, 2) / 255.0
    m = masks or {}
    m1 = None if "a1" not in m else torch.as_tensor(m["a1"]).permute(0, 3, 1, 2)
    a1 = _relu(F.conv2d(x, p["l1_w"].permute(3, 2, 0, 1), p["l1_b"], stride=4), m1)
    if "l3_w" in p:
        # network.py:30-42 'nature': conv2d(64,[4,4],[2,2]) relu, conv2d(64,[3,3],[1,1]) relu,
        # linear(512, relu) -- on the NHWC flatten (the reference hands linear the 4-D tensor,
        # which cannot run: same repair as agent.py:231-232)
        m2 = None if "a2" not in m else torch.as_tensor(m["a2"]).permute(0, 3, 1, 2)
        a2 = _relu(F.conv2d(a1, p["l2_w"].permute(3, 2, 0, 1), p["l2_b"], stride=2), m2)
        a3 = F.conv2d(a2, p["l3_w"].permute(3, 2, 0, 1), p["l3_b"], stride=1)
        flat = _relu(a3.permute(0, 2, 3, 1).reshape(a3.shape[0], -1), m.get("a3"))
        h = _relu(flat @ p["l4_w"] + p["l4_b"], m.get("h"))
        logits = h @ p["p_w"] + p["p_b"]
        value = (h @ p["q_w"] + p["q_b"]).reshape(-1)
        if keep:
            return logits, value, dict(a1=a1.permute(0, 2, 3, 1), a2=a2.permute(0, 2, 3, 1), a3=flat, h=h,
                                       _a1=a1, _a2=a2)
        return logits, value
    a2 = F.conv2d(a1, p["l2_w"].permute(3, 2, 0, 1), p["l2_b"], stride=2)
    flat = _relu(a2.permute(0, 2, 3, 1).reshape(a2.shape[0], -1), m.get("a2"))
    h = _relu(flat @ p["l4_w"] + p["l4_b"], m.get("h"))
    logits = h @ p["p_w"] + p["p_b"]
    value = (h @ p["q_w"] + p["q_b"]).reshape(-1)
    if keep:
        # _a1 is the tensor the graph continues from (NCHW): the one whose .grad is d loss / d a1
        return logits, value, dict(a1=a1.permute(0, 2, 3, 1), a2=flat, h=h, _a1=a1)
    return logits, value


def policy_terms(logits):
    """network.py:65-69: softmax, log OF the softmax, entropy = -sum p log p."""
    pi = torch.softmax(logits, dim=1)
    logpi = torch.log(pi)
    ent = -(pi * logpi).sum(dim=1)
    return pi, logpi, ent


def clip_rewards(r, lo=-1.0, hi=1.0):
    """agent.py:154 with config.py:40-41."""
    return np.clip(np.asarray(r, np.float64), lo, hi)


def nstep_returns(rewards, terminals, v_boot, gamma=0.99):
    """Algorithm 3 (assets/a3c.png) with the reference's terminal mask
    (agent.py:188-190): R_T = V(s_T); R_t = r_t + gamma*(1-term_t)*R_{t+1}.
    rewards/terminals [T,B], v_boot [B] -> R [T,B].  Same dtype as ``rewards``."""
    r = np.asarray(rewards)
    term = np.asarray(terminals).astype(r.dtype)
    T = r.shape[0]
    R = np.zeros_like(r)
    nxt = np.asarray(v_boot).astype(r.dtype)
    g = r.dtype.type(gamma)
    one = r.dtype.type(1)
    for t in range(T - 1, -1, -1):
        nxt = r[t] + g * (one - term[t]) * nxt
        R[t] = nxt
    return R


def loss_per_sample(logits, value, actions, R, beta=0.01):
    """network.py:81-94 repaired per SURVEY D3:
       policy_loss = -log pi(a) * (R - V) - beta * H     (advantage is a constant;
                     log pi(a) is the graph's own log-policy, not a placeholder)
       value_loss  = (R - V)^2 / 2
       total       = policy_loss + value_loss"""
    pi, logpi, ent = policy_terms(logits)
    adv = (R - value).detach()
    logp_a = logpi.gather(1, actions.reshape(-1, 1).long()).reshape(-1)
    policy_loss = -(logp_a * adv) - beta * ent
    value_loss = (R - value) ** 2 / 2
    return policy_loss + value_loss, policy_loss, value_loss


def analytic_head_grads(logits, value, actions, R, beta=0.01, scale=1.0):
    """d total / d logits and d total / d value in closed form (what the CUDA
    loss kernel emits).  dlogits_j = -adv*(1[a=j]-p_j) + beta*p_j*(logp_j + H);
    dV = -(R - V).  Multiplied by ``scale`` (1/num_envs for the mean-over-envs
    reduction)."""
    pi, logpi, ent = policy_terms(logits)
    adv = (R - value)
    onehot = F.one_hot(actions.long(), logits.shape[1]).to(logits.dtype)
    dlogits = -adv[:, None] * (onehot - pi) + beta * pi * (logpi + ent[:, None])
    dv = -(R - value)
    return dlogits * scale, dv * scale


def gradients(params_np, stacks, actions, R, beta=0.01, num_envs=None, dtype=torch.float64,
              masks=None):
    """Gradient of  sum_t mean_env total_loss   w.r.t. every
    parameter, by autograd of the loss expression (agent.py:317 compute_gradients).
    stacks [N,84,84,4] (N = T*B, t-major), actions [N], R [N].
    ``num_envs`` = B of the mean (global env count); None -> pure sum."""
    p = to_torch(params_np, dtype, requires_grad=True)
    logits, value, keep = forward(p, stacks, masks=masks, keep=True)
    nature = "l3_w" in p
    for k in (("_a1", "_a2", "a3", "h") if nature else ("_a1", "a2", "h")):
        keep[k].retain_grad()
    Rt = torch.as_tensor(np.asarray(R), dtype=dtype)
    at = torch.as_tensor(np.asarray(actions))
    total, pl, vl = loss_per_sample(logits, value, at, Rt, beta)
    denom = 1.0 if num_envs is None else float(num_envs)
    loss = total.sum() / denom
    loss.backward()
    grads = OrderedDict((k, v.grad.detach().numpy().copy()) for k, v in p.items())
    aux = dict(logits=logits.detach().numpy(), value=value.detach().numpy(),
               loss=float(loss.detach()), policy_loss=pl.detach().numpy(),
               value_loss=vl.detach().numpy())
    # gradients w.r.t. the PRE-relu layer outputs (what the backward kernels hand from layer to
    # layer): d loss / d post-relu activation, times the activation pattern
    pairs = ((("d_a1", "_a1"), ("d_a2", "_a2"), ("d_a3", "a3"), ("d_h", "h")) if nature else
             (("d_a1", "_a1"), ("d_a2", "a2"), ("d_h", "h")))
    for name, k in pairs:
        act, g = keep[k].detach(), keep[k].grad
        on = (act > 0)
        if masks is not None and k.lstrip("_") in masks:
            # a forced activation pattern IS the relu derivative (act = pre * mask can be a tiny
            # negative number where the device saw a tiny positive one)
            fm = torch.as_tensor(masks[k.lstrip("_")])
            on = (fm.permute(0, 3, 1, 2) if k.startswith("_") else fm) > 0
        d = (g * on.to(g.dtype)).numpy()
        aux[name] = d.transpose(0, 2, 3, 1) if k.startswith("_") else d
    return grads, aux


def clip_by_norm(g, clip=40.0):
    """agent.py:318-319 tf.clip_by_norm(grad, 40): g * clip / max(||g||_2, clip)."""
    n = float(np.sqrt((np.asarray(g, np.float64) ** 2).sum()))
    return g * (clip / max(n, clip))


def rmsprop_apply(w, ms, g, lr, decay=0.99, eps=0.1):
    """TF ApplyRMSProp as configured at main.py:63-65 (decay .99, momentum 0, epsilon .1),
    slot ``rms`` initialised to 1.0:
        ms  += (g*g - ms) * (1 - decay)
        mom  = 0*mom + lr * g / sqrt(ms + eps)        (epsilon inside the sqrt)
        var -= mom"""
    ms = ms + (g * g - ms) * (1.0 - decay)
    w = w - lr * g / np.sqrt(ms + eps)
    return w, ms


def learning_rate(step, max_step=80000000, base=0.0007):
    """agent.py:393-395 with config.py:5,11."""
    return (max_step - step + 1.0) / max_step * base


def update(params, rms, grads, lr, clip=40.0, decay=0.99, eps=0.1):
    """Per-tensor clip (agent.py:316-319) then RMSProp (agent.py:321)."""
    new_p, new_r = OrderedDict(), OrderedDict()
    for k in names_of(params):
        g = clip_by_norm(np.asarray(grads[k], np.float64), clip)
        w, m = rmsprop_apply(np.asarray(params[k], np.float64), np.asarray(rms[k], np.float64),
                             g, lr, decay, eps)
        new_p[k], new_r[k] = w, m
    return new_p, new_r


def stacks_from_screens(screens, t_max):
    """screens u8 [T+4, B, 84, 84] (f_{-3}..f_T) -> stacks [T+1, B, 84, 84, 4] (s_0..s_T),
    channel k = frame t-3+k, oldest first (history.py:13-24)."""
    Tn = t_max + 1
    return np.stack([np.stack([screens[t + k] for k in range(4)], axis=-1) for t in range(Tn)])


def a3c_cycle(params, rms, screens, actions, rewards, terminals, step, *,
              beta=0.01, gamma=0.99, num_envs=None, max_step=80000000, base_lr=0.0007,
              reduce_mean=True, masks=None):
    """One REF-A3C cycle (SURVEY §8 normative order, steps 2-10) with the actions
    given (teacher forcing: sampling parity is tested separately on identical probs).

    screens [T+4,B,84,84] u8, actions/rewards/terminals [T,B], ``step`` = env steps
    taken per env before this cycle (agent.py:55,395).  Returns new params/rms + aux."""
    T, B = np.asarray(actions).shape
    stacks = stacks_from_screens(np.asarray(screens), T)
    p64 = to_torch(params, torch.float64)
    with torch.no_grad():
        _, v_boot = forward(p64, stacks[T])
    r = clip_rewards(rewards)
    # bootstrap: R_T = V(s_T); the (1 - term) mask of the last step zeroes it when terminal
    R = nstep_returns(r, np.asarray(terminals), v_boot.numpy(), gamma)
    n_env = (num_envs if num_envs is not None else B) if reduce_mean else None
    grads, aux = gradients(params, stacks[:T].reshape((T * B,) + stacks.shape[2:]),
                           np.asarray(actions).reshape(-1), R.reshape(-1), beta, n_env,
                           masks=masks)
    lr = learning_rate(step, max_step, base_lr)
    new_p, new_r = update(params, rms, grads, lr)
    aux.update(R=R, v_boot=v_boot.numpy(), grads=grads, lr=lr)
    return new_p, new_r, aux


# ---- the reference's AS-RUNNING learner: asynchronous 1-step Q-learning ------------
def epsilon(step, ep_start=1.0, ep_end=0.1, ep_end_t=4000000, learn_start=32):
    """agent.py:142-144."""
    return ep_end + max(0.0, (ep_start - ep_end) * (ep_end_t - max(0.0, step - learn_start)) / ep_end_t)


def q_targets(q_next, rewards, terminals, discount=0.99):
    """agent.py:186-190: target_q_t = (1 - terminal) * discount * max_a Q_target(s_{t+1}) + reward
    (the reward is already clipped by observe, agent.py:154)."""
    q_next = np.asarray(q_next, np.float64)
    term = np.asarray(terminals).astype(np.float64)
    return (1.0 - term) * discount * q_next.max(axis=1) + clip_rewards(rewards)


def async_q_loss(q, actions, targets, scale=None):
    """agent.py:310-314: q_acted = sum(q * onehot(action)), delta = target - q_acted,
    loss = mean(delta^2)  (``scale`` replaces the 1/N of the mean).  Returns (loss, delta)."""
    q_acted = q.gather(1, actions.reshape(-1, 1).long()).reshape(-1)
    delta = targets - q_acted
    n = delta.shape[0]
    return (delta ** 2).sum() * (1.0 / n if scale is None else scale), delta


def async_q_gradients(params_np, target_params_np, stacks_t, stacks_tp1, actions, rewards,
                      terminals, discount=0.99, scale=None, dtype=torch.float64, masks=None):
    """agent.py:169-207 + 306-317 with the Q head in the p_w/p_b slot (q = the policy-logit
    head; the value head is unused and gets zero gradient):
        q_t_plus_1 = target_q(s_{t+1})            agent.py:186   (no gradient)
        target     = q_targets(...)               agent.py:188-190
        q_acted    = sum(q * onehot(action))      agent.py:310-311
        loss       = mean((target - q_acted)^2)   agent.py:312-314
    ``scale`` replaces the 1/N of the mean (1/(T * global envs) under env-sharded sync DP)."""
    p = to_torch(params_np, dtype, requires_grad=True)
    with torch.no_grad():
        q_next, _ = forward(to_torch(target_params_np, dtype), stacks_tp1)
    tgt = torch.as_tensor(q_targets(q_next.numpy(), np.asarray(rewards).reshape(-1),
                                    np.asarray(terminals).reshape(-1), discount), dtype=dtype)
    q, _ = forward(p, stacks_t, masks=masks)
    at = torch.as_tensor(np.asarray(actions).reshape(-1)).long()
    loss, delta = async_q_loss(q, at, tgt, scale)
    loss.backward()
    grads = OrderedDict((k, (v.grad if v.grad is not None else torch.zeros_like(v)).detach().numpy().copy())
                        for k, v in p.items())
    aux = dict(q=q.detach().numpy(), q_next=q_next.numpy(), target=tgt.numpy(),
               loss=float(loss.detach()), delta=delta.detach().numpy())
    return grads, aux
