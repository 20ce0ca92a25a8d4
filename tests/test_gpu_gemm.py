"""-m gpu: the dense contraction kernel (tcgen05, split-bf16 operand planes) against fp64."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    import gpu_harness
    return gpu_harness.PKG.ops


def _ref(A, B, bias, layout):
    A, B = A.double().cpu(), B.double().cpu()
    if layout == "nt":
        D = A @ B.T
    elif layout == "nn":
        D = A @ B
    else:
        D = A.T @ B
    if bias is not None:
        D = D + bias.double().cpu()
    return D


SHAPES = [(128, 128, 64), (256, 512, 512), (1000, 512, 512), (160, 1001, 1024), (77, 40, 200), (4160, 1536, 512), (130, 136, 1001)]
# expected normwise relative error per path: bf16x2 split ~2^-16, bf16x3 split fp32-grade
TOL = {1: 3e-5, 2: 8e-6}      # tensor-core fp32 accumulation truncates: ~K*2^-24 floor for bf16x3


@pytest.mark.parametrize("layout", ["nt", "nn", "tn"])
@pytest.mark.parametrize("path", [1, 2])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_layouts_and_paths(layout, path, M, N, K):
    ops = _ops()
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    if layout == "nt":
        A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    elif layout == "nn":
        A, B = torch.randn(M, K, generator=g), torch.randn(K, N, generator=g)
    else:
        A, B = torch.randn(K, M, generator=g), torch.randn(K, N, generator=g)
    bias = torch.randn(N, generator=g) if layout != "tn" else None
    D = ops.gemm(A.cuda(), B.cuda(), None if bias is None else bias.cuda(), layout=layout, path=path)
    torch.cuda.synchronize()
    ref = _ref(A, B, bias, layout)
    err = float((D.double().cpu() - ref).norm() / ref.norm())
    assert err < TOL[path], (layout, path, M, N, K, err)


def test_gemm_tc_large_k_splitk_wgrad_shape():
    """dW_v = dPV^T V at the headline size: M = N = 512, K = 160*196 (split-K with fp32 atomics)."""
    ops = _ops()
    g = torch.Generator().manual_seed(0)
    K, M, N = 160 * 196, 512, 512
    A, B = torch.randn(K, M, generator=g), torch.randn(K, N, generator=g)
    D = ops.gemm(A.cuda(), B.cuda(), None, layout="tn", path=1)
    ref = A.double().T @ B.double()
    assert float((D.double().cpu() - ref).norm() / ref.norm()) < 3e-5


@pytest.mark.parametrize("M,N,K", [(8192 + 100, 512, 512), (700, 256, 192), (31360, 512, 512)])
def test_gemm_cta_pair_mode_matches_single_cta(M, N, K):
    """CTA-pair (cta_group::2, 256x256 tiles) projection kernel: fp32 output through ops.gemm and bf16 hi/lo planes through
    ops.proj_planes, against fp64 and against the single-CTA kernel (HCA_TC_PAIR=0).  M tails that leave the second CTA of a
    pair partly / completely out of range are included."""
    import os
    ops = _ops()
    g = torch.Generator().manual_seed(M + N + K)
    A, W, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.05, torch.randn(N, generator=g)
    ref = A.double() @ W.double().T + b.double()
    prev = os.environ.get("HCA_TC_PAIR")
    try:
        res = {}
        for mode in ("1", "0"):
            os.environ["HCA_TC_PAIR"] = mode
            D = ops.gemm(A.cuda(), W.cuda(), b.cuda(), layout="nt", path=1)
            out = torch.full((2, M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
            ops.proj_planes(ops.split_planes(A.cuda()), ops.split_planes(W.cuda()), b.cuda(), out)
            torch.cuda.synchronize()
            res[mode] = (D.double().cpu(), (out[0].double() + out[1].double()).cpu())
            for r in res[mode]:
                assert float((r - ref).norm() / ref.norm()) < 3e-5, (mode, M, N, K)
        assert torch.equal(res["1"][0], res["0"][0]) and torch.equal(res["1"][1], res["0"][1])     # same k order, same accumulator
    finally:
        if prev is None:
            os.environ.pop("HCA_TC_PAIR", None)
        else:
            os.environ["HCA_TC_PAIR"] = prev


@pytest.mark.parametrize("M,N,K", [(512, 512, 4160), (2048, 512, 4160), (512, 1536, 4160)])
def test_gemm_cta_pair_weight_gradient_shapes(M, N, K):
    """Split-K weight-gradient products (both operands MN-major) through the CTA-pair kernel and through the single-CTA one."""
    import os
    ops = _ops()
    g = torch.Generator().manual_seed(M + N + K)
    A, B = torch.randn(K, M, generator=g), torch.randn(K, N, generator=g)
    ref = A.double().T @ B.double()
    prev = os.environ.get("HCA_TC_PAIR_WGRAD")
    try:
        for mode in ("1", "0"):
            os.environ["HCA_TC_PAIR_WGRAD"] = mode
            D = ops.gemm(A.cuda(), B.cuda(), None, layout="tn", path=1)
            torch.cuda.synchronize()
            assert float((D.double().cpu() - ref).norm() / ref.norm()) < 3e-5, (mode, M, N, K)
    finally:
        if prev is None:
            os.environ.pop("HCA_TC_PAIR_WGRAD", None)
        else:
            os.environ["HCA_TC_PAIR_WGRAD"] = prev
