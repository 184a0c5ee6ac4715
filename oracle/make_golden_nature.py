"""Generate tests/golden/nature_golden.npz by EXECUTING the reference's 'nature' trunk lines.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_golden_nature.py

The statements network.py:31-40 (``with tf.variable_scope('Nature_DQN') ...`` down to the third
``conv2d`` call) are cut out of /root/reference/src/network.py by their text and exec'd, with
``conv2d`` / ``linear`` = the reference's own src/ops.py functions imported unmodified over
oracle/tf_stub.py (+ ``tf.div`` and a no-op ``tf.device`` for these lines).  network.py:41-42
hands the 4-D ``self.l3`` to ``linear``, whose ``shape[1]`` is then 7 and whose matmul cannot run
(the same defect as SURVEY D3 for the nips branch); the repair is the NHWC flatten of
agent.py:231-232, after which ``linear`` is called with the literal arguments of network.py:41-42
(512, activation_fn, name='l4_linear').  Heads: network.py:62 / :79 as in make_golden_network.py.

Inputs: the closed-form stacks of make_golden_network.golden_stacks and closed-form weights.
"""
import contextlib
import os
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden", "nature_golden.npz")


def golden_weights(action_size):
    """Deterministic, RNG-free 'nature' parameters (|w| <= 0.04 convs, fc scaled to keep the
    activations O(1); biases non-zero so that bias_add is exercised)."""
    shapes = [("l1_w", (8, 8, 4, 32)), ("l1_b", (32,)), ("l2_w", (4, 4, 32, 64)), ("l2_b", (64,)),
              ("l3_w", (3, 3, 64, 64)), ("l3_b", (64,)), ("l4_w", (3136, 512)), ("l4_b", (512,)),
              ("p_w", (512, action_size)), ("p_b", (action_size,)), ("q_w", (512, 1)), ("q_b", (1,))]
    out = {}
    for k, (name, shape) in enumerate(shapes):
        n = int(np.prod(shape))
        i = np.arange(n, dtype=np.float64)
        amp = 0.04 if name.endswith("_w") else 0.01
        v = amp * np.sin(0.5171 * (k + 1) + 1.6180339887 * i) * np.cos(0.0113 * i + k)
        out[name] = v.reshape(shape).astype(np.float32)
    return out


def main(action_size=6):
    sys.dont_write_bytecode = True
    sys.path.insert(0, HERE)
    import tf_stub
    from make_golden_network import golden_stacks
    tf = tf_stub.install()
    tf.div = lambda a, b, **k: a / b
    tf.device = lambda name: contextlib.nullcontext()
    sys.path.insert(0, os.path.join(REF, "src"))
    import ops                                     # /root/reference/src/ops.py, unmodified
    assert os.path.realpath(ops.__file__).startswith(REF), ops.__file__

    p = golden_weights(action_size)
    S = "Nature_DQN/"
    tf_stub.VARIABLES.update({
        S + "l1_conv/w": p["l1_w"], S + "l1_conv/biases": p["l1_b"],
        S + "l2_conv/w": p["l2_w"], S + "l2_conv/biases": p["l2_b"],
        S + "l3_conv/w": p["l3_w"], S + "l3_conv/biases": p["l3_b"],
        S + "l4_linear/Matrix": p["l4_w"], S + "l4_linear/bias": p["l4_b"],
        "policy/linear/Matrix": p["p_w"], "policy/linear/bias": p["p_b"],
        "value/linear/Matrix": p["q_w"], "value/linear/bias": p["q_b"]})
    stacks = golden_stacks()

    src = open(os.path.join(REF, "src", "network.py")).read().split("\n")
    first = next(i for i, l in enumerate(src) if l.strip() == "with tf.variable_scope('Nature_DQN'), tf.device(device):")
    last = next(i for i, l in enumerate(src) if i > first and "name='l3_conv')" in l)
    assert (first + 1, last + 1) == (31, 40), (first + 1, last + 1)
    block = src[first:last + 1]
    indent = len(block[0]) - len(block[0].lstrip())
    code = compile("\n".join(l[indent:] for l in block), os.path.join(REF, "src", "network.py") + ":31-40", "exec")

    class _Self(object):
        pass
    me = _Self()
    me.s_t = tf_stub.Tensor(stacks)
    ns = dict(tf=tf, conv2d=ops.conv2d, linear=ops.linear, self=me, device="/cpu:0",
              initializer=tf.truncated_normal_initializer(0, 0.02), activation_fn=tf.nn.relu,
              data_format="NHWC")
    exec(code, ns)
    # network.py:41-42 with the flatten repair (agent.py:231-232)
    shape = me.l3.get_shape().as_list()
    l3_flat = tf.reshape(me.l3, [-1, int(np.prod(shape[1:]))])
    with tf.variable_scope("Nature_DQN"):
        l4, w4, b4 = ops.linear(l3_flat, 512, activation_fn=tf.nn.relu, name="l4_linear")
    with tf.variable_scope("policy"):
        logits, pw, pb = ops.linear(l4, action_size, name="linear")          # network.py:62
        policy = tf.nn.softmax(logits)
    with tf.variable_scope("value"):
        value, qw, qb = ops.linear(l4, 1, name="linear")                      # network.py:79

    req = dict(tf_stub.REQUESTED)
    assert req[S + "l1_conv/w"] == (8, 8, 4, 32) and req[S + "l2_conv/w"] == (4, 4, 32, 64)
    assert req[S + "l3_conv/w"] == (3, 3, 64, 64) and req[S + "l4_linear/Matrix"] == (3136, 512)
    assert me.l1.value.shape == (2, 20, 20, 32) and me.l2.value.shape == (2, 9, 9, 64)
    assert me.l3.value.shape == (2, 7, 7, 64)
    np.savez_compressed(
        OUT, action_size=action_size, stacks_sum=int(stacks.astype(np.int64).sum()),
        weights_sum=np.array([float(np.abs(p[k].astype(np.float64)).sum()) for k in sorted(p)]),
        a1=me.l1.value.astype(np.float32), a2=me.l2.value.astype(np.float32), a3=l3_flat.value,
        h=l4.value, logits=logits.value, policy=policy.value, value=value.value,
        requested=np.array(sorted("%s %s" % kv for kv in req.items())))
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes;",
          "relu-active fractions a1 %.2f a2 %.2f a3 %.2f h %.2f" %
          ((me.l1.value > 0).mean(), (me.l2.value > 0).mean(), (l3_flat.value > 0).mean(), (l4.value > 0).mean()))


if __name__ == "__main__":
    main()
