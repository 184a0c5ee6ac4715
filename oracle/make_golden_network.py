"""Generate tests/golden/network_golden.npz by EXECUTING the reference's layer functions.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_golden_network.py

What it executes, unmodified, from /root/reference: ``src/ops.py`` ``conv2d`` (:4-30) and
``linear`` (:32-46), imported over oracle/tf_stub.py (TensorFlow itself is absent; the stub's
header says exactly which arithmetic is restated and which code is the reference's own).  They
are called with the literal arguments of the reference's call sites:

    agent.py:226-229   conv2d(s_t/255., 16, [8,8], [4,4], init, relu, 'NHWC', name='l1')
                       conv2d(l1,       32, [4,4], [2,2], init, relu, 'NHWC', name='l2')
    agent.py:231-232   l2_flat = reshape(l2, [-1, prod(shape[1:])])
    agent.py:251       linear(l2_flat, 256, activation_fn=relu, name='l3')     (network.py:51: 'l4_linear')
    network.py:62      linear(l4, action_size, name='linear')   under scope 'policy'
    network.py:79      linear(l4, 1, name='linear')             under scope 'value'

``src/agent.py`` / ``src/network.py`` themselves cannot be imported (py2 syntax, an undefined
``batch_sample``, a 4-D input to ``linear`` -- SURVEY.md D2-D6), so the call SEQUENCE above is
restated here; the callee bodies are the reference's.

Inputs: two u8 stacks from a closed-form pattern and weights from ``golden_weights`` (closed
form, no RNG), both reproducible anywhere; the fixture stores the inputs' checksums and the
outputs (a1, a2 flat, h, logits, value; float64).
"""
import os
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden", "network_golden.npz")


def golden_weights(action_size):
    """Deterministic, RNG-free parameters of realistic scale (|w| <= 0.04, like N(0, .02)
    truncated at 2 sigma; biases non-zero so that bias_add is exercised)."""
    shapes = [("l1_w", (8, 8, 4, 16)), ("l1_b", (16,)), ("l2_w", (4, 4, 16, 32)), ("l2_b", (32,)),
              ("l4_w", (2592, 256)), ("l4_b", (256,)), ("p_w", (256, action_size)),
              ("p_b", (action_size,)), ("q_w", (256, 1)), ("q_b", (1,))]
    out = {}
    for k, (name, shape) in enumerate(shapes):
        n = int(np.prod(shape))
        i = np.arange(n, dtype=np.float64)
        amp = 0.04 if name.endswith("_w") else 0.01
        v = amp * np.sin(0.7310 * (k + 1) + 1.6180339887 * i) * np.cos(0.0137 * i + k)
        out[name] = v.reshape(shape).astype(np.float32)           # float32 values, as the product holds them
    return out


def golden_stacks(n=2):
    """u8 [n, 84, 84, 4] NHWC stacks (channel 0 oldest), closed form."""
    b, y, x, c = np.meshgrid(np.arange(n), np.arange(84), np.arange(84), np.arange(4), indexing="ij")
    return ((b * 97 + y * 31 + x * 17 + c * 53 + (y * x) % 29 + (x * x + 3 * y) % 251) % 256).astype(np.uint8)


def main(action_size=6):
    sys.dont_write_bytecode = True
    sys.path.insert(0, HERE)
    import tf_stub
    tf = tf_stub.install()
    sys.path.insert(0, os.path.join(REF, "src"))
    import ops                                     # /root/reference/src/ops.py, unmodified
    assert os.path.realpath(ops.__file__).startswith(REF), ops.__file__

    p = golden_weights(action_size)
    tf_stub.VARIABLES.update({
        "l1/w": p["l1_w"], "l1/biases": p["l1_b"], "l2/w": p["l2_w"], "l2/biases": p["l2_b"],
        "l3/Matrix": p["l4_w"], "l3/bias": p["l4_b"],
        "policy/linear/Matrix": p["p_w"], "policy/linear/bias": p["p_b"],
        "value/linear/Matrix": p["q_w"], "value/linear/bias": p["q_b"]})
    stacks = golden_stacks()
    init, relu = tf.truncated_normal_initializer(0, 0.02), tf.nn.relu

    s_t = tf_stub.Tensor(stacks)
    l1, w1, b1 = ops.conv2d(s_t / 255., 16, [8, 8], [4, 4], init, relu, "NHWC", name="l1")
    l2, w2, b2 = ops.conv2d(l1, 32, [4, 4], [2, 2], init, relu, "NHWC", name="l2")
    shape = l2.get_shape().as_list()
    l2_flat = tf.reshape(l2, [-1, int(np.prod(shape[1:]))])
    l3, w3, b3 = ops.linear(l2_flat, 256, activation_fn=relu, name="l3")
    with tf.variable_scope("policy"):
        logits, pw, pb = ops.linear(l3, action_size, name="linear")
        policy = tf.nn.softmax(logits)
    with tf.variable_scope("value"):
        value, qw, qb = ops.linear(l3, 1, name="linear")

    req = dict(tf_stub.REQUESTED)
    assert req == {"l1/w": (8, 8, 4, 16), "l1/biases": (16,), "l2/w": (4, 4, 16, 32), "l2/biases": (32,),
                   "l3/Matrix": (2592, 256), "l3/bias": (256,),
                   "policy/linear/Matrix": (256, action_size), "policy/linear/bias": (action_size,),
                   "value/linear/Matrix": (256, 1), "value/linear/bias": (1,)}, req
    assert l1.value.shape == (2, 20, 20, 16) and l2.value.shape == (2, 9, 9, 32)
    np.savez_compressed(
        OUT, action_size=action_size, stacks_sum=int(stacks.astype(np.int64).sum()),
        weights_sum=np.array([float(np.abs(p[k].astype(np.float64)).sum()) for k in sorted(p)]),
        a1=l1.value, a2=l2_flat.value, h=l3.value, logits=logits.value, policy=policy.value,
        value=value.value, requested=np.array(sorted("%s %s" % kv for kv in req.items())))
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes;",
          "relu-active fractions a1 %.2f a2 %.2f h %.2f" %
          ((l1.value > 0).mean(), (l2_flat.value > 0).mean(), (l3.value > 0).mean()))


if __name__ == "__main__":
    main()
