"""Storage-precision study: what does it cost at the OUTPUT of the path
(logits, values, all 10 gradients, the 100-update RMSProp trajectory) if a tensor that lives in
HBM between two kernels is kept in a narrower format?

TEST INFRASTRUCTURE: runs the float64 oracle (oracle/a3c.py) with the stored tensors rounded the
way a device kernel would round them when it writes them (round-to-nearest-even of every element,
forward tensors on the way forward, layer-to-layer gradients on the way back), and compares with
the SAME oracle unrounded ("free" oracle: no relu pattern is forced, every number is a 2-norm
relative error; flipped relus are counted).  Nothing here touches the product path.

    python tests/precision_study.py [--envs 64] [--updates 100]      -> a markdown table

Formats: f32 (reference), split = bf16 hi + bf16 lo (what round 1 stored), bf16, fp16 (gradients
scaled by the power of two ``gscale`` before rounding, exactly undone afterwards).
"""
import argparse
import os
import sys
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import a3c  # noqa: E402


def rnd(x, fmt, scale=1.0):
    """Round every element of ``x`` (float64 tensor) to the storage format."""
    if fmt == "f64":
        return x
    if fmt == "f32":
        return x.to(torch.float32).to(x.dtype)
    if fmt == "bf16":
        return x.to(torch.float32).to(torch.bfloat16).to(x.dtype)
    if fmt == "fp16":
        return (x * scale).to(torch.float32).to(torch.float16).to(x.dtype) / scale
    if fmt == "split":
        x32 = x.to(torch.float32)
        hi = x32.to(torch.bfloat16).to(torch.float32)
        lo = (x32 - hi).to(torch.bfloat16).to(torch.float32)
        return (hi + lo).to(x.dtype)
    raise ValueError(fmt)


class Stored(torch.autograd.Function):
    """y = round_fwd(x) on the way forward, dx = round_bwd(dy) on the way back."""

    @staticmethod
    def forward(ctx, x, ffmt, bfmt, gscale):
        ctx.bfmt, ctx.gscale = bfmt, gscale
        return rnd(x, ffmt)

    @staticmethod
    def backward(ctx, g):
        return rnd(g, ctx.bfmt, ctx.gscale), None, None, None


def forward_q(p, s_nhwc, v, wfmt="f64"):
    """oracle.a3c.forward with the tensors a variant ``v`` narrows rounded where a kernel would
    store them.  v = dict(a1=(fwd, bwd), a2=(fwd, bwd), h=(fwd, bwd), gscale=...): the forward
    format of the activation and the format of the gradient w.r.t. it."""
    gs = v.get("gscale", 1.0)
    x = torch.as_tensor(s_nhwc).to(torch.float64).permute(0, 3, 1, 2) / 255.0
    w = {k: rnd(t, wfmt) for k, t in p.items()} if wfmt != "f64" else p
    a1 = F.relu(F.conv2d(x, w["l1_w"].permute(3, 2, 0, 1), w["l1_b"], stride=4))
    a1 = Stored.apply(a1, v["a1"][0], v["a1"][1], gs)
    a2 = F.conv2d(a1, w["l2_w"].permute(3, 2, 0, 1), w["l2_b"], stride=2)
    flat = F.relu(a2.permute(0, 2, 3, 1).reshape(a2.shape[0], -1))
    flat = Stored.apply(flat, v["a2"][0], v["a2"][1], gs)
    h = F.relu(flat @ w["l4_w"] + w["l4_b"])
    h = Stored.apply(h, v["h"][0], v["h"][1], gs)
    logits = h @ w["p_w"] + w["p_b"]
    value = (h @ w["q_w"] + w["q_b"]).reshape(-1)
    return logits, value, dict(a1=a1, a2=flat, h=h)


FREE = dict(a1=("f64", "f64"), a2=("f64", "f64"), h=("f64", "f64"))
VARIANTS = OrderedDict([
    # name: stored formats (forward, gradient) per tensor
    ("r1: split a1/a2/d_a1/d_h, f32 h/d_a2", dict(a1=("split", "split"), a2=("split", "f32"), h=("f32", "split"))),
    ("bf16 a1 + d_a1 only", dict(a1=("bf16", "bf16"), a2=("split", "f32"), h=("f32", "split"))),
    ("bf16 d_a1 only", dict(a1=("split", "bf16"), a2=("split", "f32"), h=("f32", "split"))),
    ("bf16 a1 only", dict(a1=("bf16", "split"), a2=("split", "f32"), h=("f32", "split"))),
    ("bf16 everything stored", dict(a1=("bf16", "bf16"), a2=("bf16", "bf16"), h=("f32", "bf16"))),
    ("fp16 a1 + d_a1 only", dict(a1=("fp16", "fp16"), a2=("split", "f32"), h=("f32", "split"))),
    ("fp16 a1/a2, fp16 d_a1/d_a2/d_h", dict(a1=("fp16", "fp16"), a2=("fp16", "fp16"), h=("f32", "fp16"))),
])


def nerr(x, ref):
    x, ref = np.asarray(x, np.float64).ravel(), np.asarray(ref, np.float64).ravel()
    n = np.linalg.norm(ref)
    return float(np.linalg.norm(x - ref) / n) if n > 0 else float(np.linalg.norm(x - ref))


def cycle(params, rms, stacks, acts, rew, term, step, v, B, gscale):
    """One REF-A3C cycle (oracle.a3c.a3c_cycle) through forward_q."""
    T = acts.shape[0]
    v = dict(v, gscale=gscale)
    p = a3c.to_torch(params, torch.float64, requires_grad=True)
    with torch.no_grad():
        _, vb, _ = forward_q(p, stacks[T], v)
    R = a3c.nstep_returns(a3c.clip_rewards(rew), term, vb.numpy(), 0.99)
    logits, value, keep = forward_q(p, stacks[:T].reshape((T * B,) + stacks.shape[2:]), v)
    total, _, _ = a3c.loss_per_sample(logits, value, torch.as_tensor(acts.reshape(-1)),
                                      torch.as_tensor(R.reshape(-1)), 0.01)
    (total.sum() / B).backward()
    grads = OrderedDict((k, t.grad.numpy().copy()) for k, t in p.items())
    new_p, new_r = a3c.update(params, rms, grads, a3c.learning_rate(step))
    masks = {k: (t.detach() > 0).numpy() for k, t in keep.items()}
    return new_p, new_r, dict(logits=logits.detach().numpy(), value=value.detach().numpy(),
                              grads=grads, masks=masks, R=R)


def study(B=64, T=5, A=6, updates=100, seed=5, gscale=None, variants=None, log=None):
    """Every variant runs its OWN trajectory from the same start on the same rollout data (the
    actions are given: sampling parity is a separate, bit-exact test); errors are against the free
    float64 trajectory at every update."""
    rng = np.random.default_rng(seed)
    gscale = float(gscale or 2.0 ** int(np.ceil(np.log2(B)) + 8))
    variants = variants or VARIANTS
    params0 = a3c.init_params(A, seed=21)
    names = ["free"] + list(variants)
    P = {n: {k: w.astype(np.float64) for k, w in params0.items()} for n in names}
    Rm = {n: {k: np.ones_like(w, np.float64) for k, w in params0.items()} for n in names}
    worst = {n: dict(logits=0.0, value=0.0, grad=0.0, traj=0.0, flips=0, grad_by={}) for n in variants}
    screens = rng.integers(0, 256, (T + 4, B, 84, 84), dtype=np.uint8)
    for u in range(updates):
        new = rng.integers(0, 256, (T, B, 84, 84), dtype=np.uint8)
        screens = np.concatenate([screens[-4:], new]) if u else screens
        stacks = a3c.stacks_from_screens(screens, T)
        acts = rng.integers(0, A, (T, B))
        rew = rng.choice([-2.0, 0.0, 1.0], (T, B), p=[0.05, 0.9, 0.05])
        term = rng.random((T, B)) < 0.05
        out = {}
        for n in names:
            v = FREE if n == "free" else variants[n]
            P[n], Rm[n], out[n] = cycle(P[n], Rm[n], stacks, acts, rew, term, u * T, v, B, gscale)
        ref = out["free"]
        for n in variants:
            w, o = worst[n], out[n]
            if u == 0:                                   # same parameters: errors of ONE step
                w["logits"], w["value"] = nerr(o["logits"], ref["logits"]), nerr(o["value"], ref["value"])
                w["grad_by"] = {k: nerr(o["grads"][k], ref["grads"][k]) for k in a3c.PARAM_NAMES}
                w["grad"] = max(w["grad_by"].values())
                w["flips"] = int(sum((o["masks"][k] != ref["masks"][k]).sum() for k in ref["masks"]))
                w["relus"] = int(sum(ref["masks"][k].size for k in ref["masks"]))
                # the same gradients against the float64 oracle with THIS variant's relu pattern
                # forced (oracle.a3c._relu): the arithmetic error without the flip discontinuity
                m = dict(o["masks"], a1=o["masks"]["a1"].transpose(0, 2, 3, 1))
                gf, _ = a3c.gradients(params0, stacks[:T].reshape((T * B,) + stacks.shape[2:]),
                                      acts.reshape(-1), o["R"].reshape(-1), 0.01, B, masks=m)
                w["grad_forced"] = max(nerr(o["grads"][k], gf[k]) for k in a3c.PARAM_NAMES)
            w["traj"] = max(w["traj"], nerr(a3c.flatten_params(P[n]), a3c.flatten_params(P["free"])))
            d_free = a3c.flatten_params(P["free"]) - a3c.flatten_params(params0)
            d_var = a3c.flatten_params(P[n]) - a3c.flatten_params(params0)
            w["disp"] = nerr(d_var, d_free)              # error of the DISPLACEMENT from the start
        if log and (u + 1) % 10 == 0:
            log("update %d: " % (u + 1) + ", ".join("%s traj %.1e" % (n[:12], worst[n]["traj"]) for n in variants))
    return worst, gscale


def table(worst, B, T, updates, gscale):
    rows = ["| stored formats | logits | value | worst gradient, free (tensor) | worst gradient, relu pattern forced | relus flipped (step 1) | "
            "params after %d updates | displacement after %d updates |" % (updates, updates),
            "|---|---|---|---|---|---|---|---|"]
    for n, w in worst.items():
        k = max(w["grad_by"], key=w["grad_by"].get)
        rows.append("| %s | %.1e | %.1e | %.1e (%s) | %.1e | %d / %d | %.1e | %.1e |" % (
            n, w["logits"], w["value"], w["grad"], k, w["grad_forced"], w["flips"], w["relus"],
            w["traj"], w["disp"]))
    head = ("2-norm relative errors against the free float64 oracle, %d envs x t_max %d, fp16 "
            "gradient scale 2^%d:\n\n" % (B, T, int(np.log2(gscale))))
    return head + "\n".join(rows)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=64)
    ap.add_argument("--t-max", type=int, default=5)
    ap.add_argument("--updates", type=int, default=100)
    args = ap.parse_args()
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    worst, gs = study(args.envs, args.t_max, 6, args.updates, log=lambda s: print(s, file=sys.stderr, flush=True))
    print(table(worst, args.envs, args.t_max, args.updates, gs))
