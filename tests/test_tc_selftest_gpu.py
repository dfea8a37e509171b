"""GPU: pins the UMMA (tcgen05.mma) shared-memory descriptor conventions of csrc/tc_common.cuh on real
hardware with a one-CTA GEMM over fp16 tile-canonical blobs, in K-major and MN-major operand modes."""
import ctypes

import pytest
import torch

from lush_nerf_b200 import _lib

pytestmark = pytest.mark.gpu


def to_blob(mat: torch.Tensor) -> torch.Tensor:
    """[rows, cols] fp16 -> tile-canonical blob: offset(r,c) = (c/8)*(rows*16) + r*16 + (c%8)*2 bytes."""
    rows, cols = mat.shape
    return mat.reshape(rows, cols // 8, 8).permute(1, 0, 2).contiguous().reshape(-1)


def run(A, B, a_mode, b_mode, swap=0):
    """A [128,K], B [N,K] fp16 (logical).  MN-major views store the transposed matrix in the blob."""
    N, K = B.shape
    a_blob = to_blob(A if a_mode != 1 else A.t().contiguous()).cuda()
    b_blob = to_blob(B if b_mode == 0 else B.t().contiguous()).cuda()
    out = torch.full((128, N), float("nan"), device="cuda")
    rc = _lib.lib().lush_tc_selftest(ctypes.c_void_p(a_blob.data_ptr()), ctypes.c_void_p(b_blob.data_ptr()),
                                     ctypes.c_void_p(out.data_ptr()), N, K, a_mode, b_mode, swap, None)
    _lib.check(rc, "lush_tc_selftest")
    torch.cuda.synchronize()
    return out.cpu()


@pytest.mark.parametrize("a_mode,b_mode", [(0, 0), (1, 1), (0, 1), (1, 0)])
@pytest.mark.parametrize("N,K", [(256, 64), (128, 32), (64, 128), (256, 256), (16, 16)])
def test_umma_descriptor_conventions(a_mode, b_mode, N, K):
    g = torch.Generator().manual_seed(N * 7 + K)
    A = (torch.randn(128, K, generator=g) * 0.5).half()
    B = (torch.randn(N, K, generator=g) * 0.5).half()
    ref = A.float() @ B.float().t()
    out = run(A, B, a_mode, b_mode)
    err = (out - ref).abs().max().item()
    if not err < 2e-3:
        alt = run(A, B, a_mode, b_mode, swap=1)
        err_alt = (alt - ref).abs().max().item()
        pytest.fail(f"UMMA result wrong: max err {err:.3e} (with LBO/SBO swapped: {err_alt:.3e})")


@pytest.mark.parametrize("b_mode", [0, 1])
@pytest.mark.parametrize("N,K", [(256, 64), (128, 32), (128, 256), (64, 256), (16, 32)])
def test_umma_a_operand_from_tensor_memory(b_mode, N, K):
    """A written to TMEM with tcgen05.st (two fp16 per column, lane = row) and consumed by tcgen05.mma [d],[a],b."""
    g = torch.Generator().manual_seed(N * 11 + K)
    A = (torch.randn(128, K, generator=g) * 0.5).half()
    B = (torch.randn(N, K, generator=g) * 0.5).half()
    ref = A.float() @ B.float().t()
    out = run(A, B, 2, b_mode)
    err = (out - ref).abs().max().item()
    assert err < 2e-3, f"TS-mode UMMA wrong: max err {err:.3e}"


@pytest.mark.parametrize("N,K", [(256, 64), (128, 32), (256, 256), (64, 16)])
def test_umma_cta_pair(N, K):
    """cta_group::2: one M=256 MMA per K=16 on a cluster of two CTAs; CTA r stages rows [128r, 128r+128) of A and rows
    [N/2 r, N/2 r + N/2) of B in its own shared memory and finds its 128 rows of D in its own tensor memory."""
    g = torch.Generator().manual_seed(N * 13 + K)
    A = (torch.randn(256, K, generator=g) * 0.5).half()
    B = (torch.randn(N, K, generator=g) * 0.5).half()
    ref = A.float() @ B.float().t()
    a_blob = torch.cat([to_blob(A[:128].contiguous()), to_blob(A[128:].contiguous())]).cuda()
    b_blob = torch.cat([to_blob(B[:N // 2].contiguous()), to_blob(B[N // 2:].contiguous())]).cuda()
    out = torch.full((256, N), float("nan"), device="cuda")
    rc = _lib.lib().lush_tc_selftest_pair(ctypes.c_void_p(a_blob.data_ptr()), ctypes.c_void_p(b_blob.data_ptr()),
                                          ctypes.c_void_p(out.data_ptr()), N, K, None)
    _lib.check(rc, "lush_tc_selftest_pair")
    torch.cuda.synchronize()
    err = (out.cpu() - ref).abs().max().item()
    assert err < 2e-3, f"CTA-pair UMMA wrong: max err {err:.3e}"
