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
