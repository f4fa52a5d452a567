"""tcgen05 GEMM parity (through the C ABI, spk_gemm) against an fp64 matmul of the same split operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [
    # planes, M, N, K, a_mn, b_mn, ksplit, block_n
    (1, 128, 64, 64, 0, 0, 1, 0),          # one tile, one k-block
    (1, 256, 256, 256, 0, 0, 1, 0),
    (1, 384, 128, 128, 0, 0, 1, 0),
    (1, 160, 192, 64, 0, 0, 1, 0),         # UMMA N = 192
    (1, 300, 184, 80, 0, 0, 1, 0),         # ragged M / N / K (TMA zero fill, column guards)
    (2, 256, 256, 256, 0, 0, 1, 0),        # split-bf16: Ah*Bh + Ah*Bl + Al*Bh
    (2, 4000, 768, 256, 0, 0, 1, 0),
    (1, 51200, 256, 256, 0, 0, 1, 0),      # 400 tiles on 148 persistent CTAs (TMEM double buffer)
    (1, 256, 256, 256, 0, 1, 1, 0),        # MN-major B (dgrad / P*V)
    (2, 300, 64, 160, 0, 1, 1, 0),
    (2, 38400, 1024, 256, 0, 1, 1, 0),
    (1, 256, 256, 256, 1, 1, 1, 0),        # MN-major A and B (wgrad)
    (2, 160, 64, 160, 1, 1, 1, 0),
    (2, 256, 1024, 4096, 1, 1, 8, 0),      # split-K with fp32 atomics
    (1, 256, 80, 1000, 1, 1, 3, 0),        # N = 80 (prenet wgrad)
    (1, 512, 256, 512, 0, 0, 1, 64),       # forced small tile
    (2, 512, 256, 512, 0, 0, 1, 128),
]


@pytest.mark.parametrize("planes,m,n,k,a_mn,b_mn,ksplit,block_n", CASES)
def test_gemm_matches_fp64(planes, m, n, k, a_mn, b_mn, ksplit, block_n):
    from speaker_embedding_torch_b200 import _native as N
    torch.manual_seed(m * 7 + n * 3 + k)
    dev = "cuda"
    A = torch.randn(m, k, device=dev)
    B = torch.randn(n, k, device=dev)
    a_s = N.split_pack(A.t().contiguous() if a_mn else A, planes)
    b_s = N.split_pack(B.t().contiguous() if b_mn else B, planes)
    a_eff, b_eff = N.split_unpack(a_s), N.split_unpack(b_s)
    a_eff = a_eff.t() if a_mn else a_eff
    b_eff = b_eff.t() if b_mn else b_eff
    ref = a_eff.double() @ b_eff.double().t()
    if ksplit > 1:
        out = torch.zeros(m, n, device=dev)
        N.gemm(a_s, b_s, planes, m, n, k, bool(a_mn), bool(b_mn), atomic_out=out, ksplit=ksplit, block_n=block_n)
    else:
        out = N.gemm(a_s, b_s, planes, m, n, k, bool(a_mn), bool(b_mn), out_f32=True, block_n=block_n)
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    err = (out.double() - ref).abs().max().item()
    # planes == 1: exact products of the bf16 operands, fp32 accumulation only
    # planes == 2: the dropped Al*Bl term is ~2^-18 of |A||B|
    tol = 2e-6 if planes == 1 else 2e-5
    assert err <= tol * scale * max(1.0, (k / 256) ** 0.5), (err, scale)
    if planes == 2:   # two planes reproduce the *fp32* operands to ~2^-16
        ref32 = A.double() @ B.double().t()
        assert (out.double() - ref32).abs().max().item() <= 5e-5 * ref32.abs().max().item() * max(1.0, (k / 256) ** 0.5)


def test_gemm_bias_relu_and_split_output():
    from speaker_embedding_torch_b200 import _native as N
    torch.manual_seed(5)
    m, n, k = 1000, 512, 256
    A, B, bias = torch.randn(m, k, device="cuda"), torch.randn(n, k, device="cuda"), torch.randn(n, device="cuda")
    for planes in (1, 2):
        a_s, b_s = N.split_pack(A, planes), N.split_pack(B, planes)
        ref = torch.relu(N.split_unpack(a_s).double() @ N.split_unpack(b_s).double().t() + bias.double())
        out = N.split_unpack(N.gemm(a_s, b_s, planes, m, n, k, bias=bias, relu=True))
        torch.cuda.synchronize()
        tol = 8e-3 if planes == 1 else 4e-5       # output rounding: bf16 vs hi+lo
        assert (out.double() - ref).abs().max().item() <= tol * ref.abs().max().item()


def test_gemm_rejects_bad_arguments():
    from speaker_embedding_torch_b200 import _native as N
    a = N.split_pack(torch.randn(64, 64, device="cuda"), 1)
    with pytest.raises(RuntimeError, match="multiple of 8"):
        N.gemm(a, a, 1, 64, 60, 64)
    with pytest.raises(RuntimeError, match="split-K"):
        N.gemm(a, a, 1, 64, 64, 64, ksplit=2)
    with pytest.raises(RuntimeError, match="CUDA"):
        N.gemm(a.cpu(), a, 1, 64, 64, 64)


@pytest.mark.parametrize("planes,m,n,k,b_mn,bias,relu,out_f32", [
    (3, 512, 256, 256, 0, 0, 0, 0),
    (3, 1000, 768, 256, 0, 1, 0, 0),           # ragged M: the peer CTA's rows run past M
    (3, 300, 384, 320, 0, 1, 1, 0),            # N = 1.5 pair tiles (the peer's half of the last tile is empty)
    (2, 777, 1024, 256, 1, 0, 0, 1),           # MN-major B (dgrad), fp32 out
    (2, 2048, 128, 192, 1, 0, 0, 0),           # N = 128: only the leader's half of B exists
    (3, 40000, 1024, 256, 0, 1, 1, 0),         # many tiles per pair (TMEM double buffer, ring wrap-around)
])
def test_cta_pair_kernel_is_bit_identical_to_single_cta(planes, m, n, k, b_mn, bias, relu, out_f32):
    """tcgen05.mma.cta_group::2 path (two CTAs per 256 x 256 tile, each stages half of B) against the single-CTA
    kernel: same MMAs in the same order per accumulator element, so the outputs must match bit for bit."""
    from speaker_embedding_torch_b200 import _native as N
    torch.manual_seed(m + n + k)
    a = N.split_pack(torch.randn(m, k, device="cuda"), planes)
    b = N.split_pack(torch.randn(k, n, device="cuda") if b_mn else torch.randn(n, k, device="cuda"), planes)
    bv = torch.randn(n, device="cuda") if bias else None
    outs = []
    try:
        for mode in (0, 1):
            N.set_option("gemm_cta_pairs", mode)
            outs.append(N.gemm(a, b, planes, m, n, k, b_mn=bool(b_mn), bias=bv, relu=bool(relu),
                               out_f32=bool(out_f32)).clone())
    finally:
        N.set_option("gemm_cta_pairs", 1)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
