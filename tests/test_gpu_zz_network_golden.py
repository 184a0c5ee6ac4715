"""GPU: the CUDA forward (arl_forward through the C-ABI) against tests/golden/network_golden.npz,
the outputs of the reference's own src/ops.py conv2d / linear executed over oracle/tf_stub.py
(oracle/make_golden_network.py).  Same call as test_gpu_network.py::test_forward_layers_vs_oracle;
tolerance: BASELINE.json north_star, rel-err <= 1e-3 (measured ~1e-5)."""
import os

import numpy as np
import pytest
import torch

from oracle import a3c
from oracle.make_golden_network import golden_stacks, golden_weights
from util import REL_TOL, block, rel_err, unblock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def golden_ring():
    """The two golden stacks as a device-layout ring: [B=2, R=4, 84*84] u8, plane k of env b =
    channel k of stack b (k = 0 oldest), stored in 4x4 blocks; window (first=0, t=0)."""
    stacks = golden_stacks()                                        # [2, 84, 84, 4]
    planes = np.ascontiguousarray(stacks.transpose(0, 3, 1, 2))     # [2, 4, 84, 84]
    return stacks, np.ascontiguousarray(block(planes))


def test_golden_ring_round_trip():
    stacks, ring = golden_ring()
    back = unblock(ring)                                            # [2, 4, 84, 84]
    assert np.array_equal(np.stack([back[:, k] for k in range(4)], axis=-1), stacks)


@pytest.mark.gpu
def test_cuda_forward_matches_executed_reference_layers(pkg, cuda):
    g = np.load(os.path.join(ROOT, "tests", "golden", "network_golden.npz"))
    A, B, T, R, first = int(g["action_size"]), 2, 1, 4, 0
    params = golden_weights(A)
    _, ring_np = golden_ring()
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
    a1 = pkg.network.decode_a1(a1)
    a2 = pkg.network.decode_split(a2, N, 2592)
    errs = dict(a1=rel_err(a1.cpu(), g["a1"]), a2=rel_err(a2.cpu(), g["a2"]),
                h=rel_err(h.cpu(), g["h"]), logits=rel_err(lg.cpu(), g["logits"]),
                value=rel_err(v.cpu(), g["value"].reshape(-1)), probs=rel_err(pr.cpu(), g["policy"]))
    print("forward vs executed-reference golden, rel-err", errs)
    assert max(errs.values()) <= REL_TOL, errs
