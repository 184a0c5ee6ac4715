"""Network: the A3C policy/value net (reference src/network.py:6-127, 'nips' trunk as wired in
agent.py:226-252), with the reference's attribute names: ``w`` {l1_w..q_b}, ``policy_logits``,
``policy``, ``log_policy``, ``policy_entropy``, ``sampled_action``, ``value``, ``R``,
``total_loss``, ``copy_from_global``, ``save_model``, ``load_model``.

The reference builds a TF graph; this class owns flat f32 device buffers (parameters,
gradients, RMSProp slot) plus the rollout activations, and launches the sm_100a kernels
through the C-ABI.  Repairs of the reference's dead/broken graph follow
the paper's Algorithm 3 and are listed in DESIGN.md.
"""
import os
from collections import OrderedDict

import numpy as np
import torch

from .. import _cabi

PARAM_NAMES = ("l1_w", "l1_b", "l2_w", "l2_b", "l4_w", "l4_b", "p_w", "p_b", "q_w", "q_b")
A1_ELEMS, A2_ELEMS, FC, DA1_ELEMS = 6400, 2592, 256, 7056
A1_STORE = 3200                                               # float32 words of storage per sample of a1 (fp16)
DA1_STORE = 3528                                              # ... of d_a1 (fp16 on the 21x21 grid)


def decode_a1(raw, n=None):
    """conv1's output lives in HBM as fp16 in space-to-depth blocks (csrc/convs_tc.cu "a1s":
    [n][kc = (i*2+j)*2 + chalf][q = yp*10 + xp][8 ch], pixel (y,x) = (2yp+i, 2xp+j), 12 800 bytes
    per sample) -- the layout the conv2 tensor-core kernels consume without conversion.  ``raw`` is
    the storage the C-ABI filled (any shape whose first dimension is the sample count, or pass
    ``n``); returns the float32 [N,20,20,16] activations (network.py:47-48 l1)."""
    n = raw.shape[0] if n is None else n
    b = raw.reshape(-1)[:n * A1_STORE].view(torch.float16).reshape(n, 2, 2, 2, 10, 10, 8)
    return b.float().permute(0, 4, 1, 5, 2, 3, 6).reshape(n, 20, 20, 16)   # [n, yp, i, xp, j, chalf, 8]


def decode_split(raw, rows, cols):
    """A "split block" (include/asyncrl_b200.h: [hi|lo][cols/8 chunks][rows][8] bf16, the operand
    layout of the fc256 tensor-core kernels) -> float32 [rows, cols].  ``raw`` is the float32
    storage of rows*cols elements the C-ABI filled."""
    b = raw.reshape(-1).view(torch.bfloat16).reshape(2, cols // 8, rows, 8)
    return (b[0].float() + b[1].float()).permute(1, 0, 2).reshape(rows, cols)


def encode_split(x):
    """float32 [rows, cols] -> the split block decode_split reads (same element count, float32
    storage): hi = bf16(x), lo = bf16(x - hi).  Host-side helper for feeding a plain activation
    matrix to a kernel that expects a block (ops.linear(..., input_is_split=False)) and for tests;
    the kernels write blocks themselves."""
    rows, cols = x.shape
    hi = x.float().to(torch.bfloat16)
    lo = (x.float() - hi.float()).to(torch.bfloat16)
    b = torch.stack([hi, lo]).reshape(2, rows, cols // 8, 8).permute(0, 2, 1, 3).contiguous()
    return b.view(torch.float32).reshape(rows, cols)


def decode_da1(raw, n):
    """The gradient w.r.t. conv1's output (include/asyncrl_b200.h: one fp16 per value on the 21x21
    grid, [2 channel groups][n*441 + y*21 + x][8], 14 112 bytes per sample, still multiplied by
    tensor_scale) -> float32 [n,20,20,16]."""
    b = raw.reshape(-1)[:n * DA1_STORE].view(torch.float16).reshape(2, n, 21, 21, 8)
    return b.float().permute(1, 2, 3, 0, 4).reshape(n, 21, 21, 16)[:, :20, :20]


def param_shapes(action_size):
    return OrderedDict([
        ("l1_w", (8, 8, 4, 16)), ("l1_b", (16,)), ("l2_w", (4, 4, 16, 32)), ("l2_b", (32,)),
        ("l4_w", (A2_ELEMS, FC)), ("l4_b", (FC,)), ("p_w", (FC, action_size)),
        ("p_b", (action_size,)), ("q_w", (FC, 1)), ("q_b", (1,))])


def initial_weights(action_size, seed=123, stddev=0.02):
    """network.py:10 truncated_normal(0,.02) for convs, ops.py:36-39 normal(.02) for ``linear``
    matrices, zero biases (ops.py:24,38-39).  Returns an OrderedDict of CPU float32 tensors."""
    g = torch.Generator().manual_seed(int(seed))
    out = OrderedDict()
    for name, shape in param_shapes(action_size).items():
        if name.endswith("_b"):
            out[name] = torch.zeros(shape)
        elif name in ("l1_w", "l2_w"):
            w = torch.empty(shape)
            torch.nn.init.trunc_normal_(w, 0.0, stddev, -2 * stddev, 2 * stddev, generator=g)
            out[name] = w
        else:
            out[name] = torch.randn(shape, generator=g) * stddev
    return out


class Network(object):
    def __init__(self, sess=None, data_format='NHWC', history_length=4, screen_height=84,
                 screen_width=84, action_size=6, activation_fn='relu', initializer=None,
                 gamma=0.99, beta=0.01, global_network=None, global_optim=None, DQN_type='nips',
                 num_envs=256, t_max=5, device='cuda', seed=123, decay=0.99, epsilon=0.1,
                 clip_norm=40.0, min_reward=-1.0, max_reward=1.0):
        if DQN_type.lower() != 'nips':
            raise ValueError("Network is the 'nips' trunk (network.py:43-52); DQN_type='nature' "
                             "(network.py:30-42) is network_nature.NatureNetwork -- use "
                             "make_network(DQN_type=...); got %r" % (DQN_type,))
        if data_format != 'NHWC':
            raise ValueError("main.py:45 forces NHWC; NCHW is not built")
        if (history_length, screen_height, screen_width) != (4, 84, 84):
            raise ValueError("kernels are built for 84x84x4 stacks")
        if activation_fn not in ('relu', None) and getattr(activation_fn, '__name__', '') != 'relu':
            raise ValueError("only relu is built")
        self.sess = sess
        self.device = torch.device(device)
        _cabi.init(self.device)
        self.action_size, self.num_envs, self.t_max = int(action_size), int(num_envs), int(t_max)
        self.gamma, self.beta = float(gamma), float(beta)
        self.decay, self.epsilon, self.clip_norm = float(decay), float(epsilon), float(clip_norm)
        self.min_reward, self.max_reward = float(min_reward), float(max_reward)
        self.global_network = global_network

        A, B, T = self.action_size, self.num_envs, self.t_max
        self.offsets = _cabi.param_layout(A)
        n_params = self.offsets[-1]
        dev = self.device
        self.params = torch.zeros(n_params, device=dev)
        self.grads = torch.zeros(n_params, device=dev)
        self.rms = torch.ones(n_params, device=dev)           # TF RMSProp slot starts at 1.0
        self.w, self.g = OrderedDict(), OrderedDict()
        for i, (name, shape) in enumerate(param_shapes(A).items()):
            self.w[name] = self.params[self.offsets[i]:self.offsets[i + 1]].view(shape)
            self.g[name] = self.grads[self.offsets[i]:self.offsets[i + 1]].view(shape)
        # the weights as the tensor-core kernels keep them resident (arl_prepare_weights); refreshed
        # lazily: the key is (tensor version, number of C-ABI writes) of the parameters it was made from
        self.fc_w = torch.empty(_cabi.prepared_floats(), device=dev)
        self._fc_w_key = None
        self._param_writes = 0
        self.set_weights(initial_weights(A, seed))

        N = B * T
        f32 = dict(device=dev, dtype=torch.float32)
        # rollout activations, t-major: sample n = t*B + b
        self.l1 = torch.empty(N, A1_STORE, **f32)             # network.py:47-48, stored as blocked fp16: see a1()
        self.l2 = torch.empty(N, A2_ELEMS, **f32)             # network.py:49-50 (flattened NHWC), one split block per step: see a2()
        self.l4 = torch.empty(N, FC, **f32)                   # network.py:51-52
        self.policy_logits = torch.empty(N, A, **f32)         # network.py:62
        self.policy = torch.empty(N, A, **f32)                # network.py:65
        self.value = torch.empty(N, **f32)                    # network.py:79
        self.sampled_action = torch.zeros(N, dtype=torch.int32, device=dev)   # network.py:72
        self.R = torch.empty(N, **f32)                        # network.py:82
        # bootstrap-state scratch (not kept for backward)
        self._b = dict(l1=torch.empty(B, A1_STORE, **f32), l2=torch.empty(B, A2_ELEMS, **f32),
                       l4=torch.empty(B, FC, **f32), logits=torch.empty(B, A, **f32),
                       probs=torch.empty(B, A, **f32), value=torch.empty(B, **f32))
        # backward scratch
        self.d_logits = torch.empty(N, A, **f32)
        self.d_value = torch.empty(N, **f32)
        # d_h twice (include/asyncrl_b200.h): a split block of N rows, then its transposed copy
        # (groups of 8 samples x 256 columns) for the fc256 weight gradient: see d_h()
        self.d_l4 = torch.empty(N + (N + 7) // 8 * 8, FC, **f32)
        self.d_l2 = torch.empty(N, A2_ELEMS, **f32)
        self.d_l1 = torch.empty(N, DA1_STORE, **f32)          # fp16 on the 21x21 grid: see decode_da1
        self.workspace = torch.empty(_cabi.workspace_bytes(A), dtype=torch.uint8, device=dev)
        self.loss_sums = torch.zeros(3, **f32)                # sum policy / value loss, entropy
        self.grad_norms = torch.zeros(len(PARAM_NAMES), **f32)
        self.events = {}
        self.tensor_scale = 1.0                               # set by compute_gradients

    # -- parameters -----------------------------------------------------------------------
    def set_weights(self, weights):
        for name in PARAM_NAMES:
            self.w[name].copy_(torch.as_tensor(np.asarray(weights[name]), dtype=torch.float32))

    def get_weights(self):
        return OrderedDict((k, v.detach().cpu().numpy().copy()) for k, v in self.w.items())

    def copy_from_global(self):
        """network.py:96-107: theta' <- theta.  Replicas are updated synchronously with the same
        all-reduced gradient, so this is the identity unless a global_network was given."""
        if self.global_network is not None and self.global_network is not self:
            self.params.copy_(self.global_network.params)

    # -- forward ----------------------------------------------------------------------------
    def _rows(self, t):
        B = self.num_envs
        return slice(t * B, (t + 1) * B)

    # Optional per-entry timing (bench.py): ``timed`` is a set of C-ABI entry names (or {'*'});
    # the composed arl_forward/arl_backward are then issued as their per-layer entries -- the
    # same kernels in the same order -- with a CUDA-event pair around each selected entry.
    timed = None

    def _timed_call(self, name, *args):
        if self.timed is not None and (name in self.timed or '*' in self.timed):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _cabi.call(name, *args)
            e1.record()
            self.events.setdefault(name, []).append((e0, e1))
        else:
            _cabi.call(name, *args)

    def _fc_w_stale(self):
        """True (once) when l4_w has changed since fc_w was made: torch bumps ``_version`` on every
        in-place write to params or to a view of it, C-ABI writes are counted by hand."""
        key = (self.params._version, self._param_writes)
        stale = key != self._fc_w_key
        self._fc_w_key = key
        return stale

    def _forward_into(self, history, l1, l2, l4, logits, probs, value, refresh=None):
        """``refresh``: rebuild the prepared weight images first; None = decide here from the
        parameter version (host state).  A captured region passes the decision in."""
        P, st = _cabi.ptr, _cabi.stream_ptr()
        B, A = self.num_envs, self.action_size
        refresh = (1 if self._fc_w_stale() else 0) if refresh is None else int(bool(refresh))
        if self.timed is None:
            _cabi.call("arl_forward", P(self.params), P(self.fc_w), refresh, A, P(history.ring), B,
                       history.ring_slots, history.first_slot(0), 1, P(l1), P(l2), P(l4),
                       P(logits), P(probs), P(value), st)
            return
        if refresh:
            _cabi.call("arl_prepare_weights", P(self.params), P(self.fc_w), st)
        self._timed_layers(history, l1, l2, l4, logits, probs, value, None, 0, 0, None, 0)

    def _timed_layers(self, history, l1, l2, l4, logits, probs, value, actions, env_id_base, step, sd, seed):
        """The per-layer entries arl_forward / arl_forward_sample are made of, one event pair each
        (bench.py): the same launches in the same order, env range by env range when the batch is
        larger than one a2 block (arl_a2_block_rows)."""
        P, st = _cabi.ptr, _cabi.stream_ptr()
        B, A = self.num_envs, self.action_size
        c = _cabi.a2_block_rows(B)
        for o in range(0, B, c):
            r = slice(o, o + c)
            self._timed_call("arl_conv1_forward", P(self.fc_w), P(history.ring[r]), P(l1[r]), c,
                             history.ring_slots, history.first_slot(0), 1, st)
            self._timed_call("arl_conv2_forward", P(self.fc_w), P(l1[r]), P(l2[r]), c, st)
            # fc256, then heads + softmax (+ the action draw) as one launch
            self._timed_call("arl_fc_heads_forward", P(self.params), P(self.fc_w), A, P(l2[r]), P(l4[r]),
                             P(logits[r]), P(probs[r]), P(value[r]),
                             P(actions[r]) if actions is not None else None, int(env_id_base) + o,
                             int(step), sd, int(seed), c, st)

    def forward(self, history, t, refresh=None):
        """Forward of the current stack into rollout slot ``t``; returns (logits, policy, value)."""
        r = self._rows(t)
        self._forward_into(history, self.l1[r], self.l2[r], self.l4[r], self.policy_logits[r],
                           self.policy[r], self.value[r], refresh)
        return self.policy_logits[r], self.policy[r], self.value[r]

    def forward_sample(self, history, t, step, seed, env_id_base=0, step_dev=None, refresh=None):
        """``forward`` into rollout slot t + ``sample`` (network.py:72) as one composed call: the
        heads and the Philox draw run inside the fc256 kernel (arl_forward_sample).  The Philox step
        is ``step`` (+ the int64 device counter ``step_dev`` when given)."""
        r = self._rows(t)
        P, st = _cabi.ptr, _cabi.stream_ptr()
        B, A = self.num_envs, self.action_size
        refresh = (1 if self._fc_w_stale() else 0) if refresh is None else int(bool(refresh))
        sd = P(step_dev) if step_dev is not None else None
        if self.timed is None:
            _cabi.call("arl_forward_sample", P(self.params), P(self.fc_w), refresh, A, P(history.ring), B,
                       history.ring_slots, history.first_slot(0), P(self.l1[r]), P(self.l2[r]),
                       P(self.l4[r]), P(self.policy_logits[r]), P(self.policy[r]), P(self.value[r]),
                       P(self.sampled_action[r]), int(env_id_base), int(step), sd, int(seed), st)
            return self.sampled_action[r]
        if refresh:
            _cabi.call("arl_prepare_weights", P(self.params), P(self.fc_w), st)
        self._timed_layers(history, self.l1[r], self.l2[r], self.l4[r], self.policy_logits[r], self.policy[r],
                           self.value[r], self.sampled_action[r], env_id_base, step, sd, seed)
        return self.sampled_action[r]

    def a1(self):
        """conv1 activations of the rollout as float32 [N,20,20,16] (decoded from the device layout)."""
        return decode_a1(self.l1)

    def a2(self):
        """conv2 activations of the rollout as float32 [N,2592] (NHWC flatten, agent.py:231-232),
        decoded from the per-step split blocks."""
        B, c = self.num_envs, _cabi.a2_block_rows(self.num_envs)
        return torch.cat([decode_split(self.l2[t * B + o:t * B + o + c], c, A2_ELEMS)
                          for t in range(self.t_max) for o in range(0, B, c)])

    @staticmethod
    def pick_tensor_scale(grad_scale):
        """The power of two the layer-to-layer gradients are stored multiplied by
        (include/asyncrl_b200.h, arl_backward): 64 / grad_scale rounded to a power of two, i.e. the
        loss gradient as if it were multiplied by ~64 instead of divided by the env count."""
        import math
        e = int(round(math.log2(64.0 / max(float(grad_scale), 1e-30))))
        return float(2.0 ** max(-20, min(40, e)))

    def d_h(self):
        """Gradient w.r.t. the fc256 output of the last backward as float32 [N,256]."""
        N = self.num_envs * self.t_max
        return decode_split(self.d_l4[:N], N, FC) / self.tensor_scale

    def d_a2(self):
        """Gradient w.r.t. the conv2 output (pre-relu) of the last backward, float32 [N,2592]."""
        return self.d_l2 / self.tensor_scale

    def d_a1(self):
        """Gradient w.r.t. the conv1 output (pre-relu) of the last backward, float32 [N,20,20,16]."""
        return decode_da1(self.d_l1, self.num_envs * self.t_max) / self.tensor_scale

    def d_h_transposed(self):
        """The same gradient decoded from its second, transposed copy (what fc wgrad reads)."""
        N = self.num_envs * self.t_max
        n8 = (N + 7) // 8 * 8
        b = self.d_l4[N:].reshape(-1).view(torch.bfloat16).reshape(2, n8 // 8, FC, 8)
        return (b[0].float() + b[1].float()).permute(0, 2, 1).reshape(n8, FC)[:N] / self.tensor_scale

    def sample(self, t, step, seed, env_id_base=0):
        """network.py:72-73 sampled_action for rollout slot t."""
        r = self._rows(t)
        self._timed_call("arl_sample_actions", _cabi.ptr(self.policy[r]),
                         _cabi.ptr(self.sampled_action[r]), self.num_envs, self.action_size,
                         int(env_id_base), int(step), int(seed), _cabi.stream_ptr())
        return self.sampled_action[r]

    def bootstrap_value(self, history, refresh=None):
        """V(s_T) under the same theta (Algorithm 3: R = V(s_t, theta'_v))."""
        b = self._b
        self._forward_into(history, b['l1'], b['l2'], b['l4'], b['logits'], b['probs'], b['value'],
                           refresh)
        return b['value']

    def sample_dev(self, t, step_dev, seed, env_id_base=0):
        """``sample`` with the Philox step read from the int64 device counter ``step_dev`` (the
        form a captured CUDA graph can replay)."""
        r = self._rows(t)
        _cabi.call("arl_sample_actions_dev", _cabi.ptr(self.policy[r]),
                   _cabi.ptr(self.sampled_action[r]), self.num_envs, self.action_size,
                   int(env_id_base), _cabi.ptr(step_dev), int(seed), _cabi.stream_ptr())
        return self.sampled_action[r]

    def evaluate(self, history, step, seed, ep=None, env_id_base=0):
        """Forward of ``history``'s current stack into the scratch buffers (no rollout slot is
        touched) and one action per env: epsilon-greedy over the Q values when ``ep`` is given
        (agent.py:141-151), else a sample from the policy (network.py:72).  Used by Agent.play."""
        b = self._b
        self._forward_into(history, b['l1'], b['l2'], b['l4'], b['logits'], b['probs'], b['value'])
        if 'action' not in b:
            b['action'] = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
        if ep is not None:
            _cabi.call("arl_egreedy_actions", _cabi.ptr(b['logits']), _cabi.ptr(b['action']),
                       self.num_envs, self.action_size, float(ep), int(env_id_base), int(step),
                       int(seed), _cabi.stream_ptr())
        else:
            _cabi.call("arl_sample_actions", _cabi.ptr(b['probs']), _cabi.ptr(b['action']),
                       self.num_envs, self.action_size, int(env_id_base), int(step), int(seed),
                       _cabi.stream_ptr())
        return b['action']

    @property
    def log_policy(self):                                   # network.py:67 (log OF the softmax)
        return torch.log(self.policy)

    @property
    def policy_entropy(self):                               # network.py:69
        return -(self.policy * self.log_policy).sum(1)

    @property
    def total_loss(self):
        """network.py:94 summed over the last rollout (policy + value), as a host float."""
        s = self.loss_sums.tolist()
        return s[0] + s[1]

    # -- backward ---------------------------------------------------------------------------
    def compute_gradients(self, history, rewards, terminals, v_boot, actions=None,
                          grad_scale=1.0, allreduce=False, refresh=None):
        """Returns + loss grads (K4) then the full backward (agent.py:317) over the T*B samples
        of the rollout.  ``history`` must have had exactly t_max pushes since s_0."""
        T, B, A = self.t_max, self.num_envs, self.action_size
        acts = self.sampled_action if actions is None else actions
        self.loss_sums.zero_()
        st = _cabi.stream_ptr()
        self._timed_call("arl_returns_lossgrad", _cabi.ptr(rewards), _cabi.ptr(terminals),
                         _cabi.ptr(acts), _cabi.ptr(self.policy_logits), _cabi.ptr(self.value),
                         _cabi.ptr(v_boot), _cabi.ptr(self.R), _cabi.ptr(self.d_logits),
                         _cabi.ptr(self.d_value), _cabi.ptr(self.loss_sums), T, B, A, self.gamma,
                         self.beta, self.min_reward, self.max_reward, float(grad_scale), st)
        P = _cabi.ptr
        if self._fc_w_stale() if refresh is None else refresh:   # (a forward normally did this already)
            _cabi.call("arl_prepare_weights", P(self.params), P(self.fc_w), st)
        S = self.tensor_scale = self.pick_tensor_scale(grad_scale)
        if self.timed is None:
            _cabi.call("arl_backward", P(self.params), P(self.fc_w), A, P(history.ring), B,
                       history.ring_slots, history.first_slot(T), T, P(self.l1), P(self.l2),
                       P(self.l4), P(self.d_logits), P(self.d_value), P(self.d_l4), P(self.d_l2),
                       P(self.d_l1), P(self.grads), P(self.workspace), S, 1 if allreduce else 0, st)
            return self.grads
        N = T * B
        self._timed_call("arl_heads_backward", P(self.params), A, P(self.l4), P(self.d_logits),
                         P(self.d_value), P(self.d_l4), P(self.grads), P(self.workspace), N, S, st)
        self._timed_call("arl_fc_backward", P(self.fc_w), P(self.l2), _cabi.a2_block_rows(B), P(self.d_l4),
                         P(self.d_l2), P(self.grads), P(self.workspace), N, 1.0 / S, st)
        # the exchange as arl_backward does it: ONE all-reduce after the last backward kernel, or
        # (ARL_ALLREDUCE_OVERLAP=1) the l4_w..q_b bucket on the side stream from here on
        overlap = allreduce and os.environ.get("ARL_ALLREDUCE_OVERLAP", "")[:1] == "1"
        if overlap:
            lo = self.offsets[4]
            _cabi.call("arl_allreduce_begin", P(self.grads), lo, self.offsets[-1] - lo, st)
        self._timed_call("arl_conv2_backward", P(self.fc_w), P(self.l1), P(self.d_l2),
                         P(self.d_l1), P(self.grads), P(self.workspace), N, 1.0 / S, st)
        self._timed_call("arl_conv1_backward", P(history.ring), P(self.d_l1), P(self.grads),
                         P(self.workspace), B, history.ring_slots, history.first_slot(T), T, 1.0 / S, st)
        if overlap:
            _cabi.call("arl_allreduce_begin", P(self.grads), 0, self.offsets[4], st)
            _cabi.call("arl_allreduce_end", st)
        elif allreduce:
            self._timed_call("arl_allreduce_grads", P(self.grads), self.offsets[-1], st)
        return self.grads

    # -- async-Q mode (agent.py:169-207, 298-314): the Q head lives in the p_w/p_b slot ---------
    def make_target(self):
        """agent.py:257-296: a second parameter set, the target network."""
        self.target_params = self.params.clone()
        N, A = self.num_envs * self.t_max, self.action_size
        f32 = dict(device=self.device, dtype=torch.float32)
        self.target_q = torch.empty(N, A, **f32)              # agent.py:186 q_t_plus_1
        self.target_q_t = torch.empty(N, **f32)               # agent.py:190
        self._tq_scratch = (torch.empty(N, A, **f32), torch.empty(N, **f32))
        self.target_fc_w = torch.empty(_cabi.prepared_floats(), device=self.device)
        return self.target_params

    def update_target(self):
        """agent.py:298-303, 342-344: target <- prediction network."""
        self.target_params.copy_(self.params)

    def egreedy(self, t, step, seed, ep, env_id_base=0):
        """agent.py:141-151 on the Q values of rollout slot t."""
        r = self._rows(t)
        _cabi.call("arl_egreedy_actions", _cabi.ptr(self.policy_logits[r]),
                   _cabi.ptr(self.sampled_action[r]), self.num_envs, self.action_size, float(ep),
                   int(env_id_base), int(step), int(seed), _cabi.stream_ptr())
        return self.sampled_action[r]

    def compute_q_gradients(self, history, rewards, terminals, grad_scale, allreduce=False):
        """agent.py:186-197: target-network forward over s_1..s_T, 1-step targets, MSE gradient,
        backward.  The backward scratch buffers hold the target activations meanwhile."""
        T, B, A = self.t_max, self.num_envs, self.action_size
        P, st = _cabi.ptr, _cabi.stream_ptr()
        probs, value = self._tq_scratch
        _cabi.call("arl_forward", P(self.target_params), P(self.target_fc_w), 1, A, P(history.ring),
                   B, history.ring_slots, history.first_slot(T - 1), T, P(self.d_l1), P(self.d_l2),
                   P(self.d_l4), P(self.target_q), P(probs), P(value), st)
        self.loss_sums.zero_()
        _cabi.call("arl_q_lossgrad", P(rewards), P(terminals), P(self.sampled_action),
                   P(self.policy_logits), P(self.target_q), P(self.target_q_t), P(self.d_logits),
                   P(self.loss_sums), T * B, A, self.gamma, self.min_reward, self.max_reward,
                   float(grad_scale), st)
        self.d_value.zero_()                                   # the value head is unused
        if self._fc_w_stale():
            _cabi.call("arl_prepare_weights", P(self.params), P(self.fc_w), st)
        S = self.tensor_scale = self.pick_tensor_scale(grad_scale)
        _cabi.call("arl_backward", P(self.params), P(self.fc_w), A, P(history.ring), B,
                   history.ring_slots, history.first_slot(T), T, P(self.l1), P(self.l2), P(self.l4),
                   P(self.d_logits), P(self.d_value), P(self.d_l4), P(self.d_l2),
                   P(self.d_l1), P(self.grads), P(self.workspace), S, 1 if allreduce else 0, st)
        return self.grads

    @property
    def q(self):                                               # agent.py:252
        return self.policy_logits

    def _exchange_update(self, lr, step_dev, step_offset, base_lr, max_step):
        """K5 with the gradient exchange inside (arl_exchange_clip_rmsprop: every rank's gradient
        is read over NVLink peer memory in the norm pass; no separate all-reduce)."""
        import ctypes
        if not hasattr(self, '_offsets_c'):
            self._offsets_c = (ctypes.c_int64 * len(self.offsets))(*self.offsets)
        self._timed_call("arl_exchange_clip_rmsprop", _cabi.ptr(self.params), _cabi.ptr(self.rms),
                         _cabi.ptr(self.grads), self._offsets_c, len(self.offsets) - 1, float(lr),
                         _cabi.ptr(step_dev) if step_dev is not None else None, int(step_offset),
                         float(base_lr), int(max_step), self.decay, self.epsilon, self.clip_norm,
                         _cabi.ptr(self.grad_norms), _cabi.ptr(self.workspace), _cabi.stream_ptr())

    def apply_gradients(self, lr, exchange=False):
        """agent.py:316-321: per-tensor clip_by_norm(40) + shared RMSProp (K5).  ``exchange``: sum
        the gradient over the ranks first, inside the same kernels (NVLink peer memory)."""
        if exchange:
            self._exchange_update(lr, None, 0, 0.0, 1)
        else:
            self._timed_call("arl_clip_rmsprop", _cabi.ptr(self.params), _cabi.ptr(self.rms),
                             _cabi.ptr(self.grads), self.action_size, float(lr), self.decay, self.epsilon,
                             self.clip_norm, _cabi.ptr(self.grad_norms), _cabi.ptr(self.workspace),
                             _cabi.stream_ptr())
        self._param_writes += 1                                # the kernel wrote params: fc_w is stale

    def apply_gradients_sched(self, step_dev, step_offset, base_lr, max_step, count_write=True,
                              exchange=False):
        """``apply_gradients`` with the learning rate of agent.py:393-395 evaluated on the device
        from the int64 step counter ``step_dev`` (+ ``step_offset``): replayable by a CUDA graph."""
        if exchange:
            self._exchange_update(0.0, step_dev, step_offset, base_lr, max_step)
        else:
            self._timed_call("arl_clip_rmsprop_sched", _cabi.ptr(self.params), _cabi.ptr(self.rms),
                             _cabi.ptr(self.grads), self.action_size, _cabi.ptr(step_dev), int(step_offset),
                             float(base_lr), int(max_step), self.decay, self.epsilon, self.clip_norm,
                             _cabi.ptr(self.grad_norms), _cabi.ptr(self.workspace), _cabi.stream_ptr())
        if count_write:
            self._param_writes += 1

    # -- checkpoints (network.py:109-127), reference variable names + the rms slot ----------
    def save_model(self, saver=None, checkpoint_dir='checkpoints', step=None):
        os.makedirs(checkpoint_dir, exist_ok=True)
        path = os.path.join(checkpoint_dir, "Network-%s.npz" % (step if step is not None else 0))
        blobs = {k: v.detach().cpu().numpy() for k, v in self.w.items()}
        blobs["rms"] = self.rms.detach().cpu().numpy()
        blobs["step"] = np.int64(step if step is not None else 0)
        np.savez(path, **blobs)
        return path

    def load_model(self, saver=None, checkpoint_dir='checkpoints'):
        if not os.path.isdir(checkpoint_dir):
            return False
        files = sorted((f for f in os.listdir(checkpoint_dir) if f.startswith("Network-")),
                       key=lambda f: int(f[8:-4]))
        if not files:
            return False
        z = np.load(os.path.join(checkpoint_dir, files[-1]))
        self.set_weights({k: z[k] for k in PARAM_NAMES})
        self.rms.copy_(torch.as_tensor(z["rms"]))
        self.loaded_step = int(z["step"])
        return True


def make_network(DQN_type='nips', **kw):
    """network.py:30-55: the trunk is chosen by ``DQN_type`` ('nature' | 'nips'); anything else is
    the reference's ValueError('Wrong DQN type')."""
    kind = str(DQN_type).lower()
    if kind == 'nips':
        return Network(DQN_type='nips', **kw)
    if kind == 'nature':
        from .network_nature import NatureNetwork
        return NatureNetwork(DQN_type='nature', **kw)
    raise ValueError('Wrong DQN type: %s' % DQN_type)
