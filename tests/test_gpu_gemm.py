"""GPU: the tcgen05 GEMM behind the fc256 layer (arl_debug_gemm test hook) against float64
matmul on ragged shapes.  bf16x3 split => fp32-faithful: rel-err well below 1e-3."""
import numpy as np
import pytest
import torch

from util import rel_err

pytestmark = pytest.mark.gpu


def _run(pkg, variant, M, N, K, k_splits=1, seed=0):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(seed)
    rnd = lambda *s: torch.randn(*s, device=dev, generator=g)
    if variant in (0, 1):
        A, B, extra = rnd(M, K), rnd(K, N), rnd(N)
        ref = torch.relu(A.double() @ B.double() + extra.double())
        D = torch.full((M, N), 7.0, device=dev)
    elif variant == 2:
        A, B, extra = rnd(M, K), rnd(N, K), rnd(M, N)
        ref = (A.double() @ B.double().t()) * (extra > 0)
        D = torch.full((M, N), 7.0, device=dev)
    else:
        A, B, extra = rnd(K, M), rnd(K, N), None
        ref = A.double().t() @ B.double()
        D = torch.full((k_splits, M, N), 7.0, device=dev)
    pkg._cabi.call("arl_debug_gemm", variant, A.data_ptr(), B.data_ptr(), D.data_ptr(),
                   extra.data_ptr() if extra is not None else None, M, N, K, k_splits,
                   pkg._cabi.stream_ptr())
    torch.cuda.synchronize()
    if variant >= 3:
        kb = 32                                   # FcWgrad::KB (csrc/fc.cu)
        per = (K + k_splits - 1) // k_splits
        k_chunk = (per + kb - 1) // kb * kb
        used = (K + k_chunk - 1) // k_chunk
        D = D[:used].double().sum(0)
    return rel_err(D.cpu().numpy(), ref.cpu().numpy())


@pytest.mark.parametrize("variant,M,N,K", [
    (0, 128, 256, 32), (0, 128, 256, 2592), (0, 300, 256, 64), (0, 4096, 256, 2592),
    (1, 128, 64, 32), (1, 1000, 256, 2592), (1, 5, 256, 96),
    (2, 128, 256, 256), (2, 777, 2592, 256), (2, 64, 48, 40),
    # K == 256: the resident-weight dgrad (FcDgradRes); ragged N tile, row tiles that do not divide
    # evenly over the CTAs of an N tile (empty trailing items), one CTA per N tile
    (2, 3000, 2592, 256), (2, 1500, 208, 256), (2, 100, 2592, 256), (2, 20480, 2592, 256),
])
def test_gemm_variants_vs_float64(pkg, cuda, variant, M, N, K):
    e = _run(pkg, variant, M, N, K)
    print("gemm variant %d %dx%dx%d rel-err %.3e" % (variant, M, N, K, e))
    assert e <= 2e-5


@pytest.mark.parametrize("M,N,K,splits", [(128, 256, 32, 1), (2592, 256, 1280, 7),
                                          (2592, 256, 77, 7), (264, 256, 5000, 3)])
def test_gemm_wgrad_splitk_vs_float64(pkg, cuda, M, N, K, splits):
    # K (= samples) need not be a multiple of 8 for the sample-major (MN-major) operands; variants
    # 3 and 4 are the same instantiation
    for variant in (3, 4):
        e = _run(pkg, variant, M, N, K, k_splits=splits)
        print("gemm wgrad v%d %dx%dx%d/%d rel-err %.3e" % (variant, M, N, K, splits, e))
        assert e <= 2e-5
