"""tcgen05 GEMM kernel (bf16 operands, fp32 TMEM accumulation) vs torch on the same bf16-rounded operands."""
import pytest
import torch

from mujoco_reinforcement_learning_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda"

CASES = [
    # M, N, K, a_mn, b_mn, bn, split
    (128, 128, 64, 0, 0, 128, 1),
    (128, 64, 16, 0, 0, 64, 1),
    (256, 256, 256, 0, 0, 256, 1),
    (4096, 256, 376, 0, 0, 128, 1),      # forward layer 1 (K not a multiple of 64: TMA zero fill)
    (500, 256, 256, 0, 0, 64, 1),        # ragged M
    (300, 200, 100, 0, 0, 128, 1),       # ragged everything
    (128, 128, 64, 1, 1, 128, 1),        # both operands MN-major (wgrad form)
    (128, 128, 128, 1, 0, 128, 1),
    (128, 128, 128, 0, 1, 64, 1),
    (256, 377, 4096, 1, 1, 128, 8),      # wgrad layer 1 with the ones-column, split-K
    (256, 257, 500, 1, 1, 64, 3),        # wgrad layer 2, ragged K
    (64, 27, 1000, 1, 1, 64, 2),
    (256, 377, 2048, 1, 1, 192, 4),      # 192-wide tiles: N = 377 in two tiles
    (256, 257, 1024, 1, 1, 192, 3),
    (17, 257, 4096, 1, 1, 192, 5),       # output-layer wgrad: M = act_dim
    (1000, 192, 200, 0, 0, 192, 1),
    (4096, 256, 17, 0, 0, 128, 1),       # output-layer dgrad: K = act_dim
    (4096, 17, 256, 0, 0, 64, 1),        # output-layer forward: N = act_dim
    (40000, 256, 376, 0, 0, -1, 1),      # weights-stationary persistent kernel: forward layer 1 (W = 192 KB in smem)
    (40000, 256, 256, 0, 0, -1, 1),      # forward / dgrad layer 2
    (33000, 17, 256, 0, 0, -1, 1),       # narrow output layer
    (20000, 100, 130, 0, 0, -1, 1),      # ragged N, K
    (300, 256, 256, 0, 0, -1, 1),        # fewer tiles than CTAs
]


@pytest.mark.parametrize("M,N,K,a_mn,b_mn,bn,split", CASES)
def test_tc_gemm_matches_torch(M, N, K, a_mn, b_mn, bn, split):
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(M * 7 + N * 3 + K)
    A = torch.randn((K, M) if a_mn else (M, K), device=DEV, generator=g)
    B = torch.randn((K, N) if b_mn else (N, K), device=DEV, generator=g)
    C = torch.full((M, N), float("nan"), device=DEV)
    _lib.check(lib.b200ppo_debug_tc_gemm(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), M, N, K, a_mn, b_mn, bn, split,
                                         _lib.stream_ptr()), "debug_tc_gemm")
    Ab = A.bfloat16().double()
    Bb = B.bfloat16().double()
    Am = Ab.t() if a_mn else Ab
    Bm = Bb.t() if b_mn else Bb
    ref = Am @ Bm.t()
    err = (C.double() - ref).abs().max().item() / ref.abs().max().item()
    assert torch.isfinite(C).all()
    assert err < 2e-5, f"scaled max error {err:.3e}"  # same bf16 operands, fp32 vs fp64 accumulation only


FWD_CASES = [
    # M, N, K — forward epilogue (tanh, bf16 bulk store); N > 128 takes the CTA-pair (cta_group::2) kernel
    (40000, 256, 376),
    (40000, 256, 256),
    (32768, 256, 64),
    (33000, 200, 130),    # ragged N and K, M not a multiple of 256
    (40000, 128, 376),    # N <= 128: single-CTA kernel, bulk store
    (513, 256, 256),      # fewer tiles than CTA pairs
]


@pytest.mark.parametrize("M,N,K", FWD_CASES)
def test_ws_forward_epilogue_matches_torch(M, N, K):
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(M + N * 5 + K)
    A = torch.randn((M, K), device=DEV, generator=g)
    B = torch.randn((N, K), device=DEV, generator=g) / K ** 0.5
    C = torch.full((M, N), float("nan"), device=DEV)
    _lib.check(lib.b200ppo_debug_tc_gemm(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), M, N, K, 0, 0, -3, 1,
                                         _lib.stream_ptr()), "debug_tc_gemm")
    ref = torch.tanh(A.bfloat16().double() @ B.bfloat16().double().t())
    assert torch.isfinite(C).all()
    err = (C.double() - ref).abs().max().item()
    assert err < 1e-2, f"max abs error {err:.3e}"  # bf16 output (2^-9 relative) + MUFU tanh (2^-11)


DGRAD_CASES = [
    # M, N, K, b_mn — dgrad epilogue dz = (A B^T) * (1 - h^2); N > 128 with 32-byte row pitch takes the CTA-pair kernel
    (40000, 256, 256, 1),   # the production layer-2 dgrad: W read MN-major ([out, in] as stored)
    (40000, 256, 17, 1),    # output-layer dgrad: K = act_dim
    (33000, 208, 96, 0),    # last 64-column chunk ragged (208 = 3*64 + 16): element-wise tail of the direct epilogue
    (33000, 200, 64, 0),    # row pitch not a multiple of 16: single-CTA kernel
    (700, 256, 256, 1),     # fewer row tiles than CTA pairs, ragged M
]


@pytest.mark.parametrize("M,N,K,b_mn", DGRAD_CASES)
def test_ws_dgrad_epilogue_matches_torch(M, N, K, b_mn):
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(M + N * 11 + K)
    A = torch.randn((M, K), device=DEV, generator=g)
    B = torch.randn((K, N) if b_mn else (N, K), device=DEV, generator=g) / K ** 0.5
    h = torch.tanh(torch.randn((M, N), device=DEV, generator=g))
    C = h.clone()
    _lib.check(lib.b200ppo_debug_tc_gemm(_lib.ptr(A), _lib.ptr(B), _lib.ptr(C), M, N, K, 0, b_mn, -4, 1,
                                         _lib.stream_ptr()), "debug_tc_gemm")
    Bm = B.bfloat16().double().t() if b_mn else B.bfloat16().double()
    hb = h.bfloat16().double()
    ref = (A.bfloat16().double() @ Bm.t()) * (1.0 - hb * hb)
    assert torch.isfinite(C).all()
    err = (C.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 6e-3, f"scaled max error {err:.3e}"  # bf16 output rounding (2^-9 relative)
