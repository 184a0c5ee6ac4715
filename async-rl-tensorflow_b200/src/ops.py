"""Layer ops (reference src/ops.py:4-46 ``conv2d`` / ``linear`` and the ``batch_sample``
that network.py:4 imports but ops.py never defines), as thin wrappers over the C-ABI.

The reference's functions build TF graph nodes and create variables; here the variables
live in one flat f32 buffer (``params``) in the reference's order and layouts, and each
call launches the sm_100a kernel for that layer of the conv16-conv32-fc256 ('nips') net.
"""
import torch

from .. import _cabi

A1_ELEMS, A2_ELEMS, FC = 6400, 2592, 256


def _prepared(params):
    """params -> the resident operand images of the kernels (arl_prepare_weights), made on every
    call: these stand-alone wrappers keep no state.  Network caches it per parameter version."""
    prepared = torch.empty(_cabi.prepared_floats(), device=params.device)
    _cabi.call("arl_prepare_weights", _cabi.ptr(params), _cabi.ptr(prepared), _cabi.stream_ptr())
    return prepared


def conv2d(x, params, output_dim, kernel_size, stride, out=None, name='l1', steps=1):
    """ops.py:4-30 (VALID, NHWC, bias, relu).

    name='l1': ``x`` is a History (u8 ring; /255 folded in, agent.py:226), 8x8 s4 -> [N,20,20,16]
    name='l2': ``x`` is a1 f32 [N,20,20,16], 4x4 s2 -> [N,9,9,32]"""
    ks, st = tuple(kernel_size), tuple(stride)
    if name == 'l1' and (output_dim, ks, st) == (16, (8, 8), (4, 4)):
        hist = x
        n = hist.num_envs * steps
        if out is None:
            out = torch.empty(n, 20, 20, 16, device=params.device)
        first = hist.first_slot(steps - 1)
        prep = _prepared(params)                                 # named: must outlive the launch
        _cabi.call("arl_conv1_forward", _cabi.ptr(prep), _cabi.ptr(hist.ring), _cabi.ptr(out),
                   hist.num_envs, hist.ring_slots, first, steps, _cabi.stream_ptr())
        return out
    if name == 'l2' and (output_dim, ks, st) == (32, (4, 4), (2, 2)):
        n = x.shape[0]
        if out is None:
            out = torch.empty(n, 9, 9, 32, device=params.device)
        prep = _prepared(params)
        _cabi.call("arl_conv2_forward", _cabi.ptr(prep), _cabi.ptr(x), _cabi.ptr(out), n,
                   _cabi.stream_ptr())
        return out
    raise NotImplementedError("conv2d %s: only the 'nips' trunk layers are built "
                              "(16,[8,8],[4,4]) and (32,[4,4],[2,2])" % name)


def linear(input_, params, output_size, out=None, name='l4', input_is_split=True):
    """ops.py:32-46.  name='l4': relu(x.W + b), [N,2592] -> [N,256] (agent.py:251).  ``input_`` is
    what conv2d(name='l2') returned (a split block, include/asyncrl_b200.h); pass
    ``input_is_split=False`` for a plain float32 [N,2592] matrix (it is encoded on the host)."""
    if name == 'l4' and output_size == FC:
        x = input_.reshape(input_.shape[0], -1)
        if not input_is_split:
            from .network import encode_split
            x = encode_split(x)
        n = x.shape[0]
        if out is None:
            out = torch.empty(n, FC, device=params.device)
        # x is the split block conv2d(name='l2') wrote
        prep = _prepared(params)
        _cabi.call("arl_fc_forward", _cabi.ptr(params), _cabi.ptr(prep), _cabi.ptr(x),
                   _cabi.ptr(out), n, _cabi.stream_ptr())
        return out
    raise NotImplementedError("linear %s: only the fc256 layer is exposed stand-alone; the "
                              "policy/value heads run fused in heads()" % name)


def heads(h, params, action_size, logits=None, probs=None, value=None):
    """network.py:62 (policy logits), :65 (softmax), :79 (value) in one kernel."""
    n = h.shape[0]
    dev = params.device
    logits = torch.empty(n, action_size, device=dev) if logits is None else logits
    probs = torch.empty(n, action_size, device=dev) if probs is None else probs
    value = torch.empty(n, device=dev) if value is None else value
    _cabi.call("arl_heads_forward", _cabi.ptr(params), action_size, _cabi.ptr(h),
               _cabi.ptr(logits), _cabi.ptr(probs), _cabi.ptr(value), n, _cabi.stream_ptr())
    return logits, probs, value


def batch_sample(policy, step=0, seed=123, env_id_base=0, out=None):
    """network.py:72 ``batch_sample(self.policy)``: one action per row, a ~ policy.
    Philox4x32-10 keyed (seed; env, step) + inverse CDF -- reproducible and independent of
    how the envs are sharded over GPUs."""
    n, a = policy.shape
    if out is None:
        out = torch.empty(n, dtype=torch.int32, device=policy.device)
    _cabi.call("arl_sample_actions", _cabi.ptr(policy), _cabi.ptr(out), n, a, int(env_id_base),
               int(step), int(seed), _cabi.stream_ptr())
    return out


def argmax(scores, out=None):
    """agent.py:254 tf.argmax(q, 1): ties -> lowest index."""
    n, a = scores.shape
    if out is None:
        out = torch.empty(n, dtype=torch.int32, device=scores.device)
    _cabi.call("arl_greedy_actions", _cabi.ptr(scores), _cabi.ptr(out), n, a, _cabi.stream_ptr())
    return out
