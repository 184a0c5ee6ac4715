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

It also executes the SOURCE LINES network.py:60-94 (policy head, softmax, log, entropy, log pi of
the sampled action, value head, policy / value / total loss) verbatim: the block is cut out of
the file by its first and last statement, dedented and exec'd with ``self`` = a plain namespace
whose ``l4`` is the hidden layer computed above, ``linear`` = the reference's, ``batch_sample`` =
a function returning the actions of the fixture (it is undefined upstream), the
placeholders ``target_reward`` / ``true_action`` fed with R and with the block's own
``log_policy_of_sampled_action`` (the repair of SURVEY D3: the reference feeds it from outside).
The block is run ONE SAMPLE AT A TIME: with more, ``self.R - self.value`` broadcasts [N] - [N,1]
to [N,N], which the oracle deliberately does not reproduce.

And the scalar formulas of the as-running learner, again as the reference's own source lines
(``src/agent.py`` is py2 and cannot be imported; the cited statements are cut out by their text
and exec'd with ``self`` = the reference's ``config.M1``, imported unmodified):
    agent.py:142-144   epsilon schedule            agent.py:154       reward clip
    agent.py:188-190   1-step Q targets (numpy)    agent.py:395       learning-rate anneal
    agent.py:310-314   async-Q loss (one_hot, q_acted, delta, mean of squares; over the stub)
    main.py:64-65, agent.py:319   the arguments handed to RMSPropOptimizer / clip_by_norm

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

    # ---- network.py:60-94, the reference's own source lines, one sample at a time ----------
    src = open(os.path.join(REF, "src", "network.py")).read().split("\n")
    first = next(i for i, l in enumerate(src) if l.strip() == "with tf.variable_scope('policy'):")
    last = next(i for i, l in enumerate(src) if l.strip() == "self.total_loss = self.policy_loss + self.value_loss")
    block = src[first:last + 1]
    assert (first + 1, last + 1) == (60, 94), (first + 1, last + 1)
    indent = len(block[0]) - len(block[0].lstrip())
    code = compile("\n".join(l[indent:] for l in block), os.path.join(REF, "src", "network.py") + ":60-94", "exec")
    actions = np.array([1, 4])
    returns = np.array([0.75, -0.4])
    beta = 0.01                                                     # config.py:16
    loss_rows = []
    for n in range(2):
        class _Self(object):
            pass
        me = _Self()
        me.w = {}
        me.l4 = tf_stub.Tensor(l3.value[n:n + 1])
        ns = dict(tf=tf, linear=ops.linear, self=me, action_size=action_size, beta=beta,
                  batch_sample=lambda policy, n=n: tf_stub.Tensor(actions[n:n + 1]))
        tf_stub.FEEDS["target_reward"] = returns[n:n + 1]
        # first pass up to the placeholder needs true_action: feed the block's own log pi(a)
        tf_stub.FEEDS["true_action"] = np.zeros(1)
        exec(code, ns)
        tf_stub.FEEDS["true_action"] = me.log_policy_of_sampled_action.value
        exec(code, ns)
        assert np.allclose(me.policy_logits.value, logits.value[n:n + 1], rtol=0, atol=1e-15)
        loss_rows.append([float(me.policy_loss.value.reshape(-1)[0]), float(me.value_loss.value.reshape(-1)[0]),
                          float(me.total_loss.value.reshape(-1)[0]), float(me.policy_entropy.value[0]),
                          float(me.log_policy_of_sampled_action.value[0])])

    # ---- agent.py scalar formulas, the reference's own statements --------------------------
    sys.path.insert(0, REF)
    import config as ref_config                     # /root/reference/config.py, unmodified
    assert os.path.realpath(ref_config.__file__).startswith(REF)
    asrc = open(os.path.join(REF, "src", "agent.py")).read().split("\n")

    def cut(first_stmt, last_stmt, want_lines):
        a = next(i for i, l in enumerate(asrc) if l.strip() == first_stmt)
        b = next(i for i, l in enumerate(asrc) if i >= a and l.strip() == last_stmt)
        assert (a + 1, b + 1) == want_lines, (first_stmt, a + 1, b + 1)
        ind = len(asrc[a]) - len(asrc[a].lstrip())
        return "\n".join(l[ind:] for l in asrc[a:b + 1])

    class _Agent(ref_config.M1):                    # self.<flag> resolves to the reference's config
        pass
    me = _Agent()
    eps_code = cut("ep = test_ep or (self.ep_end +",
                   "* (self.ep_end_t - max(0., self.step - self.learn_start)) / self.ep_end_t))", (142, 144))
    clip_code = cut("reward = max(self.min_reward, min(self.max_reward, reward))",
                    "reward = max(self.min_reward, min(self.max_reward, reward))", (154, 154))
    tgt_code = cut("terminal = np.array(terminal) + 0.",
                   "target_q_t = (1. - terminal) * self.discount * max_q_t_plus_1 + reward", (188, 190))
    lr_code = cut("return (self.max_step - self.step + 1.) / self.max_step * self.learning_rate",
                  "return (self.max_step - self.step + 1.) / self.max_step * self.learning_rate", (395, 395))
    loss_code = cut("action_one_hot = tf.one_hot(self.action, self.env.action_size, 1.0, 0.0, name='action_one_hot')",
                    "self.loss = tf.reduce_mean(tf.square(self.delta), name='loss')", (310, 314))
    steps = np.array([0, 31, 32, 33, 1000, 123456, 3999999, 4000032, 4000033, 79999999], dtype=np.int64)
    eps, lrs = [], []
    for st in steps:
        me.step = int(st)
        ns = dict(self=me, test_ep=None)
        exec(eps_code, ns)
        eps.append(ns["ep"])
        lrs.append(eval(lr_code[len("return "):], dict(self=me)))
    raw_rewards = np.array([-7.5, -1.0, -0.25, 0.0, 0.5, 1.0, 3.0])
    clipped = []
    for r in raw_rewards:
        ns = dict(self=me, reward=float(r))
        exec(clip_code, ns)
        clipped.append(ns["reward"])
    q_next = np.cos(np.arange(4 * action_size, dtype=np.float64).reshape(4, action_size) * 0.37)
    tq_rewards, tq_terminals = [1.0, 0.0, -1.0, 1.0], [False, True, False, True]
    ns = dict(self=me, np=np, terminal=list(tq_terminals), q_t_plus_1=q_next, reward=list(tq_rewards))
    exec(tgt_code, ns)
    target_q = np.asarray(ns["target_q_t"], np.float64)

    class _Env(object):
        pass
    me.env = _Env()
    me.env.action_size = action_size
    q_actions = np.array([2, 0, 5, 3])
    me.q = tf_stub.Tensor(np.sin(np.arange(4 * action_size, dtype=np.float64).reshape(4, action_size) * 0.21))
    me.action = tf_stub.Tensor(q_actions)
    me.target_q_t = tf_stub.Tensor(target_q)
    exec(loss_code, dict(self=me, tf=tf))
    q_loss, q_delta = float(me.loss.value), me.delta.value

    # ---- optimizer hyper-parameters as the reference passes them (TF ops themselves: unpinned) --
    msrc = open(os.path.join(REF, "main.py")).read().split("\n")
    a = next(i for i, l in enumerate(msrc) if l.strip() == "optimizer = tf.train.RMSPropOptimizer(")
    assert a + 1 == 64 and msrc[a + 1].strip() == "lr_op, decay=0.99, momentum=0, epsilon=0.1)"
    rec = {}

    class _Train(object):
        @staticmethod
        def RMSPropOptimizer(lr, **kw):
            rec.update(kw)
            return "optimizer"
    tf.train = _Train
    exec("\n".join(l.strip() for l in msrc[a:a + 2]).replace("(\n", "("), dict(tf=tf, lr_op="lr_op"))
    clip_line = next(l.strip() for l in asrc if "tf.clip_by_norm(" in l)
    assert clip_line == "new_grads_and_vars.append((tf.clip_by_norm(grad, 40), var))"      # agent.py:319
    tf.clip_by_norm = lambda g, c: rec.setdefault("clip_norm", c)
    exec(clip_line, dict(tf=tf, new_grads_and_vars=[], grad="grad", var="var"))

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
        value=value.value, actions=actions, returns=returns, beta=beta,
        steps=steps, eps=np.array(eps), lrs=np.array(lrs), raw_rewards=raw_rewards,
        clipped=np.array(clipped), q_next=q_next, tq_rewards=np.array(tq_rewards),
        tq_terminals=np.array(tq_terminals), target_q=target_q, q_values=me.q.value,
        q_actions=q_actions, q_loss=q_loss, q_delta=q_delta,
        rms_decay=float(rec["decay"]), rms_momentum=float(rec["momentum"]), rms_epsilon=float(rec["epsilon"]),
        clip_norm=float(rec["clip_norm"]), base_lr=float(ref_config.M1.learning_rate),
        max_step=int(ref_config.M1.max_step), discount=float(ref_config.M1.discount),
        cfg_beta=float(ref_config.M1.beta),
        loss_rows=np.array(loss_rows),   # per sample: policy_loss, value_loss, total_loss, entropy, log pi(a)
        requested=np.array(sorted("%s %s" % kv for kv in req.items())))
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes;",
          "relu-active fractions a1 %.2f a2 %.2f h %.2f" %
          ((l1.value > 0).mean(), (l2_flat.value > 0).mean(), (l3.value > 0).mean()))


if __name__ == "__main__":
    main()
