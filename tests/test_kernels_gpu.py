"""Parity of every C-ABI kernel against the CPU oracle / golden vectors (run on the B200 box).

Integer / index work must be bit-exact; fp32 work is held to rel 1e-4 norm-wise
(allclose(rtol=1e-4, atol=1e-4*max|ref|)) and usually lands at ~1e-6; bf16-operand
GEMMs are compared with the same inputs rounded to bf16 and fp32 accumulation.
"""
import numpy as np
import pytest
import torch

from oracle import numpitron_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from numpitron_b200 import ops
    from numpitron_b200 import runtime as rt


def dev(a):
    return rt.from_numpy(np.ascontiguousarray(a))


def host(t):
    torch.cuda.synchronize()
    return t.float().cpu().numpy() if t.dtype == torch.bfloat16 else t.cpu().numpy()


def close(got, want, rtol=1e-4):
    want = np.asarray(want)
    np.testing.assert_allclose(got, want, rtol=rtol, atol=rtol * max(np.abs(want).max(), 1e-30))


def bf16_round(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).bfloat16().float().numpy()


# ---------------------------------------------------------------------------------------------
# GEMM
# ---------------------------------------------------------------------------------------------
def run_gemm(mode, M, N, K, ta, tb, seed=0, alpha=1.0, bias=False, relu=False, mask=False, bf16_out=False):
    rng = np.random.default_rng(seed)
    # leading dimensions TMA can address: multiples of 8 bf16 / 4 fp32 (mode 2 on odd pitches runs the FFMA kernel)
    pad = (lambda n: (n + 7) // 8 * 8) if mode == 1 else ((lambda n: (n + 3) // 4 * 4) if mode == 2 else (lambda n: n))
    a_shape = (K, M) if ta else (M, K)
    b_shape = (N, K) if tb else (K, N)
    lda, ldb = pad(a_shape[1]), pad(b_shape[1])
    A = np.zeros((a_shape[0], lda), np.float32)
    B = np.zeros((b_shape[0], ldb), np.float32)
    A[:, : a_shape[1]] = rng.standard_normal(a_shape)
    B[:, : b_shape[1]] = rng.standard_normal(b_shape)
    bias_v = rng.standard_normal(N).astype(np.float32) if bias else None
    mask_v = rng.standard_normal((M, N)).astype(np.float32) if mask else None
    if mode == 1:
        A, B = bf16_round(A), bf16_round(B)
        At, Bt = dev(A).bfloat16(), dev(B).bfloat16()
    else:
        At, Bt = dev(A), dev(B)
    C = rt.empty((M, N))
    Cb = rt.empty((M, N), torch.bfloat16) if bf16_out else None
    ops.gemm(At, Bt, M, N, K, lda=lda, ldb=ldb, trans_a=ta, trans_b=tb, C_out=C, ldc=N, Cb_out=Cb, ldcb=N,
             alpha=alpha, bias=dev(bias_v) if bias else None, relu=relu,
             mask=dev(mask_v) if mask else None, ldmask=N, mode=mode)
    opA = A[:, : a_shape[1]].astype(np.float64)
    opB = B[:, : b_shape[1]].astype(np.float64)
    ref = (opA.T if ta else opA) @ (opB.T if tb else opB) * alpha
    if bias:
        ref = ref + bias_v
    if relu:
        ref = np.maximum(ref, 0)
    if mask:
        ref = ref * (mask_v > 0)
    got = host(C)
    scale = np.abs(ref).max()
    err = np.abs(got - ref).max() / scale
    if bf16_out:
        errb = np.abs(host(Cb) - ref).max() / scale
        assert errb < 1e-2, f"bf16 output error {errb}"
    return err


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("shape", [(128, 128, 64), (128, 64, 128), (200, 72, 136), (1, 1, 1), (300, 257, 1000)])
def test_gemm_fp32_simt(shape, ta, tb):
    assert run_gemm(0, *shape, ta, tb) < 2e-6


def test_gemm_fp32_splitk_and_epilogue():
    assert run_gemm(0, 384, 1152, 8192, True, False) < 5e-6          # dW shape -> split-K path
    assert run_gemm(0, 256, 384, 96, False, False, alpha=0.5, bias=True, relu=True) < 2e-6
    assert run_gemm(0, 256, 384, 96, False, True, mask=True) < 2e-6
    assert run_gemm(0, 130, 70, 4096, True, False, bias=True) < 5e-6  # split-K + ragged + bias


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("shape", [(128, 128, 32), (128, 64, 128), (256, 384, 384), (200, 72, 136), (1, 1, 1), (300, 257, 1000),
                                   (4096, 1152, 384), (130, 40, 20)])
def test_gemm_tf32x3_tcgen05(shape, ta, tb):
    """fp32-accurate tensor-core mode: fp32 operands split hi/lo inside the kernel, three kind::tf32 MMAs per
    k-step, fp32 TMEM accumulator — fp32-level error (north star: per-layer rel 1e-4; FFMA lands at ~1e-6)."""
    assert run_gemm(2, *shape, ta, tb) < 5e-6


def test_gemm_tf32x3_splitk_and_epilogue():
    # long K: the tensor core adds into the TMEM accumulator with truncation, so the error grows with the chain
    # length; split-K caps a chain at 1024 k and the partials are added in fp32 round-to-nearest
    assert run_gemm(2, 384, 1152, 8192, True, False) < 1e-5          # dW shape -> split-K path
    assert run_gemm(2, 256, 384, 96, False, False, alpha=0.5, bias=True, relu=True) < 5e-6
    assert run_gemm(2, 256, 384, 96, False, True, mask=True) < 5e-6
    assert run_gemm(2, 130, 72, 4096, True, False, bias=True) < 5e-6  # split-K + ragged + bias
    # hard case for a split scheme: large common offset, tiny signal (the reference's activations: mean >> std)
    rng = np.random.default_rng(3)
    A = (1.0 + 1e-3 * rng.standard_normal((512, 384))).astype(np.float32)
    B = (0.003 + 1e-4 * rng.standard_normal((384, 256))).astype(np.float32)
    C = rt.empty((512, 256))
    ops.gemm(dev(A), dev(B), 512, 256, 384, lda=384, ldb=256, C_out=C, ldc=256, mode=2)
    ref = A.astype(np.float64) @ B.astype(np.float64)
    want32 = A @ B
    err, err32 = np.abs(host(C) - ref).max() / np.abs(ref).max(), np.abs(want32 - ref).max() / np.abs(ref).max()
    assert err < max(4 * err32, 1e-6), (err, err32)


@pytest.mark.parametrize("ta,tb", [(False, True), (False, False), (True, False), (True, True)])
@pytest.mark.parametrize("shape", [(128, 128, 64), (128, 64, 64), (128, 128, 256), (256, 384, 384), (200, 72, 136),
                                   (384, 1536, 192), (300, 520, 72), (128, 1152, 64)])
def test_gemm_bf16_tcgen05(shape, ta, tb):
    # inputs are exactly representable in bf16 -> only fp32 accumulation-order error remains
    assert run_gemm(1, *shape, ta, tb) < 1e-5


def test_gemm_bf16_splitk_and_epilogue():
    assert run_gemm(1, 384, 1152, 8192, True, False) < 2e-5
    assert run_gemm(1, 512, 384, 384, False, False, alpha=0.5, bias=True, relu=True, bf16_out=True) < 1e-5
    assert run_gemm(1, 512, 384, 1536, False, True, mask=True) < 1e-5
    assert run_gemm(1, 130, 72, 4096, True, False, bias=True) < 2e-5


def test_gemm_bf16_persistent_many_tiles_and_fused_epilogues():
    """Bench-size linear layers: > 148 tiles per launch (persistent loop, both TMEM accumulators),
    bf16-only outputs, bias+ReLU, bf16 ReLU-mask — checked against an fp64 product on the GPU."""
    rng = np.random.default_rng(9)
    M, K, N = 16384, 384, 1536
    x = bf16_round(rng.standard_normal((M, K)))
    w = bf16_round(rng.standard_normal((K, N)) * 0.05)
    bias = rng.standard_normal(N).astype(np.float32)
    xt, wt = dev(x).bfloat16(), dev(w).bfloat16()
    ref = torch.clamp(dev(x).double() @ dev(w).double() + dev(bias).double(), min=0)
    yb = rt.empty((M, N), torch.bfloat16)
    ops.gemm(xt, wt, M, N, K, lda=K, ldb=N, Cb_out=yb, ldcb=N, bias=dev(bias), relu=True, mode=1)   # bf16 only
    torch.cuda.synchronize()
    err = (yb.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 6e-3, err
    y = rt.empty((M, N))
    ops.gemm(xt, wt, M, N, K, lda=K, ldb=N, C_out=y, ldc=N, Cb_out=yb, ldcb=N, bias=dev(bias), relu=True, mode=1)
    torch.cuda.synchronize()
    assert (y.double() - ref).abs().max().item() / ref.abs().max().item() < 1e-5
    np.testing.assert_array_equal(host(yb), bf16_round(host(y)))
    # dX = (dY @ W2^T) * (hidden > 0) with the bf16 hidden tensor as mask, bf16 + fp32 outputs
    dy = bf16_round(rng.standard_normal((M, K)))
    w2 = bf16_round(rng.standard_normal((N, K)) * 0.05)            # stored (d_hidden, d_out) = (N, K)
    dyt, w2t = dev(dy).bfloat16(), dev(w2).bfloat16()
    dx, dxb = rt.empty((M, N)), rt.empty((M, N), torch.bfloat16)
    ops.gemm(dyt, w2t, M, N, K, lda=K, ldb=K, trans_b=True, C_out=dx, ldc=N, Cb_out=dxb, ldcb=N, mask=yb, ldmask=N, mode=1)
    torch.cuda.synchronize()
    ref2 = (dev(dy).double() @ dev(w2).double().T) * (yb.double() > 0)
    assert (dx.double() - ref2).abs().max().item() / ref2.abs().max().item() < 1e-5
    np.testing.assert_array_equal(host(dxb), bf16_round(host(dx)))
    dxb2 = rt.empty((M, N), torch.bfloat16)                          # masked, bf16-only output
    ops.gemm(dyt, w2t, M, N, K, lda=K, ldb=K, trans_b=True, Cb_out=dxb2, ldcb=N, mask=yb, ldmask=N, mode=1)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(host(dxb2), host(dxb))


@pytest.mark.parametrize("ta,tb", [(False, True), (False, False), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(16384, 384, 200), (9984, 1152, 136), (16384, 1536, 72), (20480, 128, 320)])
def test_gemm_bf16_cta_pairs(M, N, K, ta, tb, monkeypatch):
    """Shapes large enough for the CTA-pair kernels (tcgen05 cta_group::2: one 256 x BN MMA per pair, each
    CTA staging its own A rows and half of the B tile), all four operand layouts, BN = 128 / 192 / 256,
    ragged K tails, fp32 and bf16 outputs, bias.  The pair kernels must reproduce the single-CTA ones
    bit for bit (NUMPITRON_GEMM_PAIR=0, read at every launch)."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn((K, M) if ta else (M, K), device="cuda", generator=g).bfloat16()
    b = torch.randn((N, K) if tb else (K, N), device="cuda", generator=g).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    ref = (a.double().T if ta else a.double()) @ (b.double().T if tb else b.double())
    out = {}
    for pair in ("1", "0"):
        monkeypatch.setenv("NUMPITRON_GEMM_PAIR", pair)
        c = rt.empty((M, N))
        cb = rt.empty((M, N), torch.bfloat16)
        ops.gemm(a, b, M, N, K, lda=a.shape[1], ldb=b.shape[1], trans_a=ta, trans_b=tb, C_out=c, ldc=N, mode=1)
        ops.gemm(a, b, M, N, K, lda=a.shape[1], ldb=b.shape[1], trans_a=ta, trans_b=tb, Cb_out=cb, ldcb=N, bias=bias, mode=1)
        torch.cuda.synchronize()
        out[pair] = (host(c), host(cb))
    err = np.abs(out["1"][0] - host(ref)).max() / ref.abs().max().item()
    assert err < 1e-5, err
    np.testing.assert_array_equal(out["1"][0], out["0"][0])
    np.testing.assert_array_equal(out["1"][1], out["0"][1])
    np.testing.assert_array_equal(out["1"][1], bf16_round(out["1"][0] + host(bias)[None, :]))


@pytest.mark.parametrize("M,K,N", [(16384, 384, 1536), (300, 136, 200), (128, 64, 72)])
def test_gemm_packed_relu_bits(M, K, N):
    """The forward epilogue packs (relu output > 0) into 1 bit/element; the backward GEMM consuming those
    words must equal the one consuming the bf16 activation as mask (ReLU.backward, activation.py:42-43)."""
    rng = np.random.default_rng(M + N)
    x = dev(bf16_round(rng.standard_normal((M, K)))).bfloat16()
    w = dev(bf16_round(rng.standard_normal((K, N)) * 0.1)).bfloat16()
    bias = dev(rng.standard_normal(N).astype(np.float32))
    nw = (N + 31) // 32
    bits = torch.full((nw, M), -1, dtype=torch.int32, device="cuda")  # word-major: [n / 32][m]
    hb = rt.empty((M, N), torch.bfloat16)
    h = rt.empty((M, N))
    ops.gemm(x, w, M, N, K, lda=K, ldb=N, C_out=h, ldc=N, Cb_out=hb, ldcb=N, bias=bias, relu=True, relu_bits_out=bits, mode=1)
    got = np.ascontiguousarray(host(bits).view(np.uint32).T)
    want = np.zeros((M, nw), np.uint32)
    pos = host(h) > 0
    for j in range(N):
        want[:, j // 32] |= pos[:, j].astype(np.uint32) << np.uint32(j % 32)
    valid = np.full(nw, 0xFFFFFFFF, np.uint32)
    if N % 32:
        valid[-1] = (1 << (N % 32)) - 1
    np.testing.assert_array_equal(got & valid, want)
    dy = dev(bf16_round(rng.standard_normal((M, K)))).bfloat16()
    w2 = dev(bf16_round(rng.standard_normal((N, K)) * 0.1)).bfloat16()
    a, b = rt.empty((M, N)), rt.empty((M, N))
    ab, bb = rt.empty((M, N), torch.bfloat16), rt.empty((M, N), torch.bfloat16)
    ops.gemm(dy, w2, M, N, K, lda=K, ldb=K, trans_b=True, C_out=a, ldc=N, Cb_out=ab, ldcb=N, mask=hb, ldmask=N, mode=1)
    ops.gemm(dy, w2, M, N, K, lda=K, ldb=K, trans_b=True, C_out=b, ldc=N, Cb_out=bb, ldcb=N, mask_bits=bits, mode=1)
    np.testing.assert_array_equal(host(a), host(b))
    np.testing.assert_array_equal(host(ab), host(bb))


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_gemm_batched_head_layout(mode):
    """q@k^T and p@v addressed in place inside the (B,S,h,3,Dh) projection output (attention.py:43-52)."""
    rng = np.random.default_rng(5)
    B_, S, h, Dh = 2, 64, 3, 64
    qkv = rng.standard_normal((B_, S, h, 3, Dh)).astype(np.float32)
    if mode == 1:
        qkv = bf16_round(qkv)
    t = dev(qkv.reshape(B_, S, -1))
    op = t.bfloat16() if mode == 1 else t
    ld = h * 3 * Dh
    head, mat, merged = (S * ld, 3 * Dh), (h * S * S, S * S), (S * h * Dh, Dh)
    scores = rt.empty((B_, h, S, S))
    ops.gemm(op, op, S, S, Dh, lda=ld, ldb=ld, trans_b=True, b_off=Dh, a_strides=head, b_strides=head,
             C_out=scores, ldc=S, c_strides=mat, batch=B_ * h, batch_inner=h, mode=mode)
    q, k, v = (qkv[:, :, :, i, :].transpose(0, 2, 1, 3).astype(np.float64) for i in range(3))
    close(host(scores), q @ k.transpose(0, 1, 3, 2), 1e-5)
    p = rng.standard_normal((B_, h, S, S)).astype(np.float32)
    if mode == 1:
        p = bf16_round(p)
    pt = dev(p).bfloat16() if mode == 1 else dev(p)
    att = rt.empty((B_, S, h * Dh))
    ops.gemm(pt, op, S, Dh, S, lda=S, ldb=ld, b_off=2 * Dh, a_strides=mat, b_strides=head, C_out=att,
             ldc=h * Dh, c_strides=merged, batch=B_ * h, batch_inner=h, mode=mode)
    want = (p.astype(np.float64) @ v).transpose(0, 2, 1, 3).reshape(B_, S, h * Dh)
    close(host(att), want, 1e-5)
    # dK = dS^T @ q written into the k-slot of d_qkv
    dqkv = rt.zeros((B_, S, ld))
    ops.gemm(pt, op, S, Dh, S, lda=S, ldb=ld, trans_a=True, b_off=0, a_strides=mat, b_strides=head, C_out=dqkv,
             ldc=ld, c_off=Dh, c_strides=head, batch=B_ * h, batch_inner=h, mode=mode)
    want = (p.astype(np.float64).transpose(0, 1, 3, 2) @ q).transpose(0, 2, 1, 3)  # (B,S,h,Dh)
    got = host(dqkv).reshape(B_, S, h, 3, Dh)
    close(got[:, :, :, 1, :], want, 1e-5)
    assert not got[:, :, :, 0, :].any() and not got[:, :, :, 2, :].any()


@pytest.mark.parametrize("B,S,h", [(1, 128, 1), (2, 256, 3), (1, 512, 2), (3, 256, 6),
                                   (40, 256, 6),     # 480 work items: every persistent CTA walks several
                                   (50, 128, 4),     # 200 single-step items
                                   (13, 384, 5),     # 195 items of 1-3 steps
                                   (2, 1024, 1)])    # GPT-2 small under tp = 8: one local head, 8 key tiles
@pytest.mark.parametrize("Dh", [64, 96])
def test_fused_attention_fwd_bwd(B, S, h, Dh):
    """Flash-style tcgen05 attention against the reference's arithmetic (attention.py:43-52, 81-98)
    evaluated in fp64 on the same bf16-rounded inputs; scale = 1/sqrt(d_model) with d_model = 384.
    head_dim 96 = BASELINE configs[2] (12 heads of 64 re-cut as 1 local head of 96 per rank, attention.py:37-45)."""
    div = float(np.float32(384**0.5))
    rng = np.random.default_rng(S + h)
    qkv = bf16_round(rng.standard_normal((B, S, h, 3, Dh)) * 2.0)
    d_out = bf16_round(rng.standard_normal((B, S, h, Dh)))
    q, k, v = (qkv[:, :, :, i, :].transpose(0, 2, 1, 3).astype(np.float64) for i in range(3))
    mask = np.tri(S, S, dtype=bool)
    s = np.where(mask, q @ k.transpose(0, 1, 3, 2) / div, -np.inf)
    mx = s.max(-1, keepdims=True)
    e = np.exp(s - mx)
    p = e / e.sum(-1, keepdims=True)
    out_ref = (p @ v).transpose(0, 2, 1, 3)                          # (B,S,h,Dh)
    lse_ref = (mx + np.log(e.sum(-1, keepdims=True)))[..., 0]          # (B,h,S)
    qkv_t = dev(qkv.reshape(B, S, -1)).bfloat16()
    out, lse = ops.attention_fwd(qkv_t, B, S, h, Dh, div)
    close(host(out).reshape(B, S, h, Dh), out_ref, 1e-2)
    close(host(lse), lse_ref, 1e-4)
    # backward (uses the bf16 forward output for D = sum(dO*O), as the kernel does)
    do = d_out.transpose(0, 2, 1, 3).astype(np.float64)                # (B,h,S,Dh)
    dv = p.transpose(0, 1, 3, 2) @ do
    dp = do @ v.transpose(0, 1, 3, 2)
    ds = np.where(mask, p * (dp - (dp * p).sum(-1, keepdims=True)), 0) / div
    dq = ds @ k
    dk = ds.transpose(0, 1, 3, 2) @ q
    dqkv_ref = np.stack([dq, dk, dv], axis=3).transpose(0, 2, 1, 3, 4)  # (B,S,h,3,Dh)
    dqkv = ops.attention_bwd(qkv_t, out, dev(d_out.reshape(B, S, -1)).bfloat16(), lse, B, S, h, Dh, div)
    got = host(dqkv).reshape(B, S, h, 3, Dh)
    for i, name in enumerate(("dq", "dk", "dv")):
        ref = dqkv_ref[:, :, :, i]
        err = np.abs(got[:, :, :, i] - ref).max() / np.abs(ref).max()
        assert err < 2e-2, (name, err)


def test_cast_bf16_padding():
    x = np.random.default_rng(0).standard_normal((37, 65)).astype(np.float32)
    out = host(ops.cast_bf16(dev(x), pad_cols=True))
    assert out.shape == (37, 72)
    np.testing.assert_array_equal(out[:, :65], bf16_round(x))
    assert not out[:, 65:].any()
    y = np.random.default_rng(1).standard_normal((16, 384)).astype(np.float32)
    np.testing.assert_array_equal(host(ops.cast_bf16(dev(y))), bf16_round(y))


# ---------------------------------------------------------------------------------------------
# LayerNorm / ReLU / colsum / softmax against golden vectors of the reference
# ---------------------------------------------------------------------------------------------
def test_layernorm_golden(golden_layers):
    G = golden_layers
    y, mean, var = ops.layernorm_fwd(dev(G["ln/x"]), dev(G["ln/w"]), dev(G["ln/b"]), 1e-5)
    close(host(y), G["ln/y"], 1e-5)
    dx, dg, db = ops.layernorm_bwd(dev(G["ln/x"]), dev(G["ln/w"]), mean, var, dev(G["ln/dy"]), 1e-5)
    close(host(dx), G["ln/dx"], 2e-5)
    close(host(dg), G["ln/dw"], 1e-5)
    close(host(db), G["ln/db"], 1e-5)


@pytest.mark.parametrize("rows,d", [(1, 4), (77, 384), (1000, 768), (33, 1024), (19, 130), (5, 4096), (4099, 384),
                                    (700, 2048), (300, 4096), (50, 8192), (9, 1028), (3, 8196)])   # wide rows: block-per-row kernels
def test_layernorm_shapes_vs_oracle(rows, d):
    rng = np.random.default_rng(d + rows)
    x = (1.5 + 0.03 * rng.standard_normal((1, rows, d))).astype(np.float32)   # large mean, small std (Q12)
    g = rng.random(d).astype(np.float32) * 0.2
    b = (rng.standard_normal(d) * 0.1).astype(np.float32)
    res = rng.standard_normal((1, rows, d)).astype(np.float32)
    dy = rng.standard_normal((1, rows, d)).astype(np.float32)
    y_ref, ctx = O.layernorm_forward(x, g, b)
    dx_ref, dg_ref, db_ref = O.layernorm_backward(ctx, g, dy)
    xt = dev(x)
    y, mean, var = ops.layernorm_fwd(xt, dev(g), dev(b), 1e-5, residual=dev(res), want_bf16=True)
    close(host(y), y_ref + res, 2e-5)
    close(host(y._npt_bf16), y_ref + res, 1e-2)
    close(host(var), ctx["var"].reshape(-1), 1e-4)
    dx, dg, db = ops.layernorm_bwd(xt, dev(g), mean, var, dev(dy), 1e-5)
    close(host(dx), dx_ref, 1e-4)
    close(host(dg), dg_ref, 1e-4)
    close(host(db), db_ref, 1e-5)
    if d % 4 == 0 and d <= 1024:
        # fused bias gradient of the preceding Linear: column sums of dx from the same pass (fp32 parity
        # mode and the bf16 throughput mode, whose dx differs by the reciprocal-multiply rounding only)
        dx2, dg2, db2 = ops.layernorm_bwd(xt, dev(g), mean, var, dev(dy), 1e-5, want_colsum=True)
        np.testing.assert_array_equal(host(dx2), host(dx))
        np.testing.assert_array_equal(host(dg2), host(dg))
        close(host(dx2._npt_colsum), dx_ref.reshape(rows, d).astype(np.float64).sum(0), 2e-4)
        dx3, _, _ = ops.layernorm_bwd(xt, dev(g), mean, var, dev(dy), 1e-5, bf16_only=True, want_colsum=True)
        close(host(dx3), dx_ref, 1e-2)
        close(host(dx3._npt_colsum), dx_ref.reshape(rows, d).astype(np.float64).sum(0), 2e-4)
        # bf16 inputs (sub-layer output / dX that left their GEMM as bf16 alone): same math on the rounded values
        xb, dyb = bf16_round(x), bf16_round(dy)
        yb_ref, ctxb = O.layernorm_forward(xb, g, b)
        dxb_ref, dgb_ref, dbb_ref = O.layernorm_backward(ctxb, g, dyb)
        y4, mean4, var4 = ops.layernorm_fwd(dev(xb).bfloat16(), dev(g), dev(b), 1e-5, residual=dev(res), want_bf16=True)
        close(host(y4), yb_ref + res, 2e-5)
        dx4, dg4, db4 = ops.layernorm_bwd(dev(xb).bfloat16(), dev(g), mean4, var4, dev(dyb).bfloat16(), 1e-5, want_colsum=True)
        close(host(dx4), dxb_ref, 1e-4)
        close(host(dg4), dgb_ref, 1e-4)
        close(host(db4), dbb_ref, 1e-5)
        close(host(dx4._npt_colsum), dxb_ref.reshape(rows, d).astype(np.float64).sum(0), 2e-4)
        dx5, _, _ = ops.layernorm_bwd(dev(xb).bfloat16(), dev(g), mean4, var4, dev(dyb), 1e-5)   # mixed: dy is cast
        np.testing.assert_array_equal(host(dx5), host(dx4))


def test_relu_colsum_golden(golden_layers):
    G = golden_layers
    np.testing.assert_array_equal(host(ops.relu_fwd(dev(G["relu/x"]))), G["relu/y"])
    np.testing.assert_array_equal(host(ops.relu_bwd(dev(G["relu/x"]), dev(G["relu/dy"]))), G["relu/dx"])
    x = np.random.default_rng(3).standard_normal((5000, 1537)).astype(np.float32)
    close(host(ops.colsum(5000, 1537, dev(x))), x.astype(np.float64).sum(0), 1e-5)
    close(host(ops.colsum(16, 24, dev(G["linear/dy"].reshape(16, 24)))), G["linear/db"], 1e-6)
    close(host(ops.mean(dev(x[:, :100].copy()))), [x[:, :100].astype(np.float64).mean()], 1e-4)


def test_softmax_causal_golden(golden_layers):
    G = golden_layers
    x = np.where(np.isinf(G["softmax/x"]), 0.0, G["softmax/x"]).astype(np.float32)  # kernel applies the mask itself
    p = ops.softmax_causal_fwd(dev(x), 1.0)
    close(host(p), G["softmax/y"], 1e-6)
    # reference order: softmax.backward on the unmasked upstream gradient, then mask (attention.py:86-87)
    ds = ops.softmax_causal_bwd(p, dev(G["softmax/dy"]), 1.0)
    want = np.where(np.tri(12, 12, dtype=bool), G["softmax/dx"], 0)
    close(host(ds), want, 1e-5)


@pytest.mark.parametrize("S", [16, 64, 256, 1024, 1160])
def test_softmax_causal_vs_oracle(S):
    rng = np.random.default_rng(S)
    s = (rng.standard_normal((2, 2, S, S)) * 4).astype(np.float32)
    dp = rng.standard_normal((2, 2, S, S)).astype(np.float32)
    div = np.float32(384**0.5)
    mask = np.tri(S, S, dtype=bool)
    p_ref = O.softmax(np.where(mask, s / div, -np.inf).astype(np.float32))
    ds_ref = np.where(mask, O.softmax_backward_compact(p_ref, dp), 0) / div
    p = ops.softmax_causal_fwd(dev(s), float(div), want_bf16=True)
    close(host(p), p_ref, 1e-5)
    close(host(p._npt_bf16), p_ref, 1e-2)
    close(host(ops.softmax_causal_bwd(p, dev(dp), float(div))), ds_ref, 1e-5)


# ---------------------------------------------------------------------------------------------
# Embedding (bit-exact) / loss / Adam (bit-exact)
# ---------------------------------------------------------------------------------------------
def test_embedding_golden_bit_exact(golden_layers):
    G = golden_layers
    tok = dev(G["emb/tokens"])
    E = dev(G["emb/E"])
    y = ops.embedding_gather(tok, E, None, 16, 0, 16, True)
    np.testing.assert_array_equal(host(y), G["emb/y"])
    np.testing.assert_array_equal(host(ops.add_pos(y, dev(G["emb/pos"]), 16)), G["emb/y_pos"])
    y2 = ops.embedding_gather(tok, E, dev(G["emb/pos"]), 16, 0, 16, True)   # fused positions
    np.testing.assert_array_equal(host(y2), G["emb/y_pos"])
    grad = dev(G["emb/dE_gemm"].copy())
    ops.embedding_scatter_add(tok, dev(G["emb/dy"]), grad, 0, 16, True)
    np.testing.assert_array_equal(host(grad), G["emb/dE"])


@pytest.mark.parametrize("V,tp,T,d", [(65, 1, 4096, 384), (65, 2, 3000, 384), (50257, 8, 2048, 96), (17, 1, 1, 8), (300, 1, 5000, 33)])
def test_embedding_scatter_bit_exact_vs_numpy(V, tp, T, d):
    """np.add.at order (flattened (b,s), sequential) for every TP rank's chunk; heavy collisions at V=65."""
    rng = np.random.default_rng(V + T)
    tokens = rng.integers(0, V, (1, T)).astype(np.uint16)
    d_out = rng.standard_normal((1, T, d)).astype(np.float32)
    full = rng.standard_normal((d, V)).astype(np.float32)
    for r in range(tp):
        lo, hi = O.vocab_chunk_bounds(V, tp, r)
        E = O.shard(full, tp, r, 1)
        y_ref, ctx = O.embedding_forward(tokens.copy(), E, lo, hi, True)
        g_ref = O.embedding_backward(ctx, E.copy() * 0.5, d_out)
        tok = dev(tokens)
        np.testing.assert_array_equal(host(ops.embedding_gather(tok, dev(E), None, T, lo, hi, True)), y_ref)
        grad = dev(E * 0.5)
        ops.embedding_scatter_add(tok, dev(d_out), grad, lo, hi, True)
        np.testing.assert_array_equal(host(grad), g_ref)


def test_loss_golden(golden_layers):
    G = golden_layers
    loss, grad = ops.softmax_ce_fused(dev(G["loss/logits"]), dev(G["loss/labels"]), 0, 16)
    close(host(loss), G["loss/loss"], 1e-6)
    close(host(grad), G["loss/dlogits"], 1e-6)


@pytest.mark.parametrize("V,tp,T", [(65, 1, 1000), (65, 2, 777), (50257, 8, 64), (50257, 1, 16), (3000, 3, 50)])
def test_loss_vocab_parallel_vs_oracle(V, tp, T):
    rng = np.random.default_rng(V * tp)
    logits = (rng.standard_normal((1, T, V)) * 5 + 100).astype(np.float32)
    labels = rng.integers(0, V, (1, T)).astype(np.uint16)
    labels[0, 0] = V - 1
    if V % tp == 0:  # the reference's own bounds raise here (Q8); use array_split-compatible bounds
        sizes = [a.shape[0] for a in np.array_split(np.arange(V), tp)]
        bounds = [(sum(sizes[:r]), sum(sizes[: r + 1])) for r in range(tp)]
    else:
        bounds = [O.vocab_chunk_bounds(V, tp, r) for r in range(tp)]
    shards = [np.ascontiguousarray(s) for s in np.array_split(logits, tp, axis=-1)]
    loss_ref, grads_ref = O.softmax_cross_entropy_world(shards, labels.copy(), bounds)
    lab = dev(labels)
    if tp == 1:
        loss, grad = ops.softmax_ce_fused(dev(shards[0]), lab, *bounds[0], want_bf16=True)
        close(host(loss), loss_ref, 1e-5)
        close(host(grad), grads_ref[0], 1e-5)
        gb = host(grad._npt_bf16)
        close(gb[..., :V], grads_ref[0], 1e-2)
        assert not gb[..., V:].any()
        return
    dl = [dev(s) for s in shards]
    rowmax = [host(ops.softmax_ce_stage1(x)) for x in dl]
    gmax = dev(np.max(rowmax, axis=0))                       # all-reduce(max) done by hand
    st2 = [ops.softmax_ce_stage2(x, lab, gmax, *bounds[r]) for r, x in enumerate(dl)]
    gp = dev(sum(host(a) for a, _ in st2))                   # all-reduce(sum)
    gs = dev(sum(host(b) for _, b in st2))
    for r, x in enumerate(dl):
        loss, grad = ops.softmax_ce_stage3(x, lab, gmax, gp, gs, *bounds[r])
        close(host(loss), loss_ref, 1e-5)
        close(host(grad), grads_ref[r], 1e-5)


def test_adam_bit_exact(golden_layers):
    G = golden_layers
    p = dev(G["adam/p0"].copy())
    m, v = rt.zeros((8, 8)), rt.zeros((8, 8))
    f = lambda x: float(np.float32(x))
    for i in range(3):
        g = dev(G[f"adam/g{i}"])
        table = ops.AdamTable(1)
        table.fill([(p, g, m, v, None)])
        ops.adam_step(table, f(1e-3), f(0.9), f(1 - 0.9), f(0.99), f(1 - 0.99), f(1 - 0.9), f(1 - 0.99), f(1e-7))
        np.testing.assert_array_equal(host(p), G[f"adam/p{i + 1}"])
    np.testing.assert_array_equal(host(m), G["adam/m"])
    np.testing.assert_array_equal(host(v), G["adam/v"])


def test_adam_multi_tensor_bit_exact_vs_oracle():
    rng = np.random.default_rng(11)
    shapes = [(384, 1152), (1536,), (7,), (384, 65), (100003,)]
    f = lambda x: float(np.float32(x))
    ps = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    gs = [(rng.standard_normal(s) * 10.0 ** rng.integers(-9, 1)).astype(np.float32) for s in shapes]
    ms = [(rng.standard_normal(s) * 0.01).astype(np.float32) for s in shapes]
    vs = [(rng.random(s) * 0.01).astype(np.float32) for s in shapes]
    entries = []
    for p, g, m, v in zip(ps, gs, ms, vs):
        pt = dev(p)
        sh = ops.cast_bf16(pt, pad_cols=True) if p.ndim == 2 else None
        entries.append((pt, dev(g), dev(m), dev(v), sh))
    table = ops.AdamTable(len(entries))
    table.fill(entries)
    ops.adam_step(table, f(1e-4), f(0.9), f(1 - 0.9), f(0.99), f(1 - 0.99), f(1 - 0.9), f(1 - 0.99), f(1e-7))
    for (pt, _, mt, vt, sh), p, g, m, v in zip(entries, ps, gs, ms, vs):
        p2, m2, v2 = O.adam_update(p, g, m, v, 1e-4, 0.9, 0.99)
        np.testing.assert_array_equal(host(pt), p2)
        np.testing.assert_array_equal(host(mt), m2)
        np.testing.assert_array_equal(host(vt), v2)
        if sh is not None:
            got = host(sh)
            np.testing.assert_array_equal(got[:, : p.shape[1]], bf16_round(p2))
            assert not got[:, p.shape[1]:].any()


@pytest.mark.parametrize("d_in,d_out", [(384, 1152), (384, 384), (1536, 384), (104, 72)])
def test_adam_consumes_splitk_partials_bit_exact(d_in, d_out):
    """Weight gradient dW = x^T dy of a Linear (nn/linear.py:55) as a split-K tcgen05 GEMM: with
    rt.deferred_wgrad the partials are not reduced; Adam adds them in split order while it reads.  The
    parameters, moments and the bf16 shadow must come out bit-identical to reducing first."""
    from numpitron_b200.nn import functional as F

    T = 8192
    g = torch.Generator(device="cuda").manual_seed(d_in + d_out)
    x = torch.randn((1, T, d_in), device="cuda", generator=g).bfloat16()
    dy = (torch.randn((1, T, d_out), device="cuda", generator=g) * 0.01).bfloat16()
    rt.set_precision("bf16")
    try:
        dW = F.linear_backward_weight(x, dy)
        with rt.deferred_wgrad():
            parts = F.linear_backward_weight(x, dy, may_defer=True)
        assert getattr(parts, "_npt_splits", 1) > 1 and tuple(parts.shape) == (parts._npt_splits, d_in, d_out)
        np.testing.assert_array_equal(host(parts).sum(0, dtype=np.float64).astype(np.float32).shape, host(dW).shape)
        f = lambda v: float(np.float32(v))
        rng = np.random.default_rng(3)
        out = []
        for grad in (dW, parts):
            p = dev(rng.standard_normal((d_in, d_out)).astype(np.float32)) if not out else dev(p0)
            if not out:
                p0, m0, v0 = host(p).copy(), (rng.standard_normal((d_in, d_out)) * 0.01).astype(np.float32), \
                    (rng.random((d_in, d_out)) * 0.01).astype(np.float32)
            m, v = dev(m0), dev(v0)
            sh = ops.cast_bf16(p, pad_cols=True)
            table = ops.AdamTable(1)
            table.fill([(p, grad, m, v, sh)])
            ops.adam_step(table, f(1e-4), f(0.9), f(1 - 0.9), f(0.99), f(1 - 0.99), f(1 - 0.9), f(1 - 0.99), f(1e-7))
            out.append((host(p), host(m), host(v), host(sh)))
        for a, b in zip(out[0], out[1]):
            np.testing.assert_array_equal(a, b)
    finally:
        rt.set_precision("fp32")


# ---------------------------------------------------------------------------------------------
# edge cases: empty inputs, the uint16 token range, error reporting through the C ABI
# ---------------------------------------------------------------------------------------------
def test_empty_inputs_are_noops():
    lib = __import__("numpitron_b200._lib", fromlist=["load"]).load()
    st = rt.stream()
    assert lib.npt_layernorm_fwd(None, None, None, None, None, None, None, None, 0, 384, 1e-5, st) == 0
    assert lib.npt_relu_fwd(None, None, None, 0, st) == 0
    assert lib.npt_embedding_gather(None, None, 65, None, None, None, 0, 256, 384, 0, 64, 1, st) == 0
    assert lib.npt_cast_f32_bf16(None, 8, None, 8, 0, 8, st) == 0


def test_full_uint16_vocabulary_and_single_token():
    """V = 65536 (every uint16 value is a valid token, incl. 65535), T = 1 and T = 3."""
    rng = np.random.default_rng(1)
    V, d = 65536, 16
    E = rng.standard_normal((d, V)).astype(np.float32)
    for toks in ([65535], [0, 65535, 65535]):
        tokens = np.array([toks], dtype=np.uint16)
        T = tokens.size
        d_out = rng.standard_normal((1, T, d)).astype(np.float32)
        y_ref, ctx = O.embedding_forward(tokens.copy(), E, 0, V, True)
        g_ref = O.embedding_backward(ctx, np.zeros_like(E), d_out)
        np.testing.assert_array_equal(host(ops.embedding_gather(dev(tokens), dev(E), None, T, 0, V, True)), y_ref)
        grad = rt.zeros((d, V))
        ops.embedding_scatter_add(dev(tokens), dev(d_out), grad, 0, V, True)
        np.testing.assert_array_equal(host(grad), g_ref)


def test_errors_are_reported_not_swallowed():
    from numpitron_b200._lib import NptError

    a = rt.empty((128, 64), torch.bfloat16)
    with pytest.raises(NptError, match="multiples of 8"):
        ops.gemm(a, a, 128, 60, 64, lda=64, ldb=60, C_out=rt.empty((128, 60)), ldc=60, mode=1)  # ldb % 8 != 0
    with pytest.raises(TypeError):
        ops.gemm(a, a, 128, 64, 64, lda=64, ldb=64, C_out=rt.empty((128, 64)), ldc=64, mode=0)  # bf16 into the fp32 kernel
    with pytest.raises(NptError, match="head_dim 64 and 96"):
        ops.attention_fwd(rt.empty((1, 128, 3 * 80), torch.bfloat16), 1, 128, 1, 80, 10.0)


# ---------------------------------------------------------------------------------------------
# tensor parallelism over two ranks, kernel level, on ONE GPU: the "peer" buffers are local tensors.
# (The two-GPU end-to-end parity is tests/test_multigpu.py; these run wherever a single B200 is available.)
# ---------------------------------------------------------------------------------------------
def test_gemm_split_peer_store():
    """npt_gemm_args.peer_rows_lo/hi: tiles of the rows in [lo, hi) go to C_bf16_peer ONLY, the others to C_bf16 only
    (the reduce-scatter by push of distributed/peer.py); without the bounds every tile goes to both."""
    rng = np.random.default_rng(7)
    M, N, K = 512, 384, 192
    a, b = bf16_round(rng.standard_normal((M, K))), bf16_round(rng.standard_normal((K, N)))
    want = a.astype(np.float64) @ b.astype(np.float64)
    at, bt = dev(a).bfloat16(), dev(b).bfloat16()
    own = torch.full((M, N), 7.0, dtype=torch.bfloat16, device="cuda")
    peer = torch.full((M, N), 7.0, dtype=torch.bfloat16, device="cuda")
    ops.gemm(at, bt, M, N, K, lda=K, ldb=N, Cb_out=own, ldcb=N, Cb_peer=(peer.data_ptr(), 256, 512), mode=1)
    o, p = host(own), host(peer)
    close(o[:256], want[:256], 1e-2)
    close(p[256:], want[256:], 1e-2)
    assert (o[256:] == 7.0).all() and (p[:256] == 7.0).all()          # untouched halves
    ops.gemm(at, bt, M, N, K, lda=K, ldb=N, Cb_out=own, ldcb=N, Cb_peer=peer.data_ptr(), mode=1)
    np.testing.assert_array_equal(host(own), host(peer))               # double store
    close(host(own), want, 1e-2)


class _FakeExchange:
    """Stands in for distributed.peer.PeerExchange on one GPU: `bufs[r]` plays rank r's arena."""

    def __init__(self, rank, pools):
        self.rank, self.pools, self.i = rank, pools, 0

    def arena(self, shape, dtype):
        key = self.i
        self.i += 1
        for r in (0, 1):
            if key not in self.pools[r]:
                self.pools[r][key] = torch.zeros(shape, dtype=dtype, device="cuda")
        return self.pools[self.rank][key], self.pools[1 - self.rank][key].data_ptr()


@pytest.mark.parametrize("T,d", [(256, 384), (512, 768), (256, 1024)])
def test_layernorm_row_sharded_two_ranks_on_one_gpu(T, d):
    """npt_layernorm_fwd_v3 / _bwd_v4 as distributed/peer.py drives them: two ranks hold partial sums x0 + x1 of the
    LayerNorm input (and dy0 + dy1 of its output gradient); rank r normalises rows [r T/2, (r+1) T/2) and pushes
    its bf16 results and per-block gradient partials into the other rank's buffers.  Afterwards both ranks must hold
    the complete tensors, equal to the LayerNorm of round_bf16(x0 + x1) (nn/normalization.py:22-52 after the
    reference's all-reduce, nn/attention.py:57-58)."""
    rng = np.random.default_rng(T + d)
    x0 = bf16_round(0.75 + 0.02 * rng.standard_normal((1, T, d)))
    x1 = bf16_round(0.75 + 0.02 * rng.standard_normal((1, T, d)))
    dy0, dy1 = bf16_round(rng.standard_normal((1, T, d))), bf16_round(rng.standard_normal((1, T, d)))
    g = rng.random(d).astype(np.float32) * 0.2
    b = (rng.standard_normal(d) * 0.1).astype(np.float32)
    res = rng.standard_normal((1, T, d)).astype(np.float32)
    x, dy = bf16_round(x0 + x1), bf16_round(dy0 + dy1)
    y_ref, ctx = O.layernorm_forward(x, g, b)
    dx_ref, dg_ref, db_ref = O.layernorm_backward(ctx, g, dy)
    H = T // 2
    xs, dys = [dev(x0).bfloat16(), dev(x1).bfloat16()], [dev(dy0).bfloat16(), dev(dy1).bfloat16()]
    gt, bt, rt_ = dev(g), dev(b), dev(res)
    pools = ({}, {})
    state = []
    for r in (0, 1):       # forward of both ranks (rank r: own partial xs[r], received partial xs[1 - r])
        px = _FakeExchange(r, pools)
        y, mean, var, x_sum = ops.layernorm_fwd_rows(xs[r], xs[1 - r].data_ptr(), gt, bt, 1e-5, rt_, (r * H, H), px)
        state.append((px, y, mean, var, x_sum))
    for r in (0, 1):
        px, y, mean, var, x_sum = state[r]
        rows = slice(r * H, (r + 1) * H)
        close(host(y)[0, rows], (y_ref + res)[0, rows], 2e-5)                       # fp32 result: own rows
        close(host(x_sum)[0, rows], x[0, rows], 1e-6)
        close(host(y._npt_bf16)[0], (y_ref + res)[0], 1e-2)                         # bf16 result: EVERY row, on both ranks
    np.testing.assert_array_equal(host(state[0][1]._npt_bf16), host(state[1][1]._npt_bf16))
    finishes = []
    for r in (0, 1):       # backward of both ranks
        px, y, mean, var, x_sum = state[r]
        dx, finish = ops.layernorm_bwd_rows(x_sum, gt, mean, var, dys[r], dys[1 - r].data_ptr(), 1e-5, (r * H, H), px, True)
        finishes.append((dx, finish))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(host(finishes[0][0]), host(finishes[1][0]))        # dx complete and identical on both
    close(host(finishes[0][0]), dx_ref, 1e-2)
    grads = [f() for _, f in finishes]
    for k, want in enumerate((dg_ref, db_ref, dx_ref.reshape(T, d).astype(np.float64).sum(0))):
        np.testing.assert_array_equal(host(grads[0][k]), host(grads[1][k]))          # replicated parameters stay bit-identical
        close(host(grads[0][k]), want, 2e-2 if k == 2 else 1e-3)
    with rt.deferred_wgrad():   # inside the graphed step the two stacks go to Adam un-reduced: same sums
        stacks = finishes[0][1]()
    for k in range(3):
        assert stacks[k]._npt_splits == stacks[k].shape[0] and stacks[k]._npt_split_stride == 3 * d
        close(host(stacks[k]).astype(np.float64).sum(0), host(grads[0][k]), 1e-5)


def test_memcpy2d_reassembles_vocabulary_shards():
    """npt_memcpy2d_async: the (d, V) embedding table put back together from two pitched shards (nn/embedding.py:19-37)."""
    lib = __import__("numpitron_b200._lib", fromlist=["load"]).load()
    rng = np.random.default_rng(11)
    d, w0, w1 = 48, 33, 32
    E = rng.standard_normal((d, w0 + w1)).astype(np.float32)
    shards = [dev(np.ascontiguousarray(E[:, :w0])), dev(np.ascontiguousarray(E[:, w0:]))]
    full = rt.zeros((d, w0 + w1))
    for sh, off, w in ((shards[0], 0, w0), (shards[1], w0, w1)):
        ops.check(lib.npt_memcpy2d_async(full.data_ptr() + off * 4, (w0 + w1) * 4, sh.data_ptr(), w * 4, w * 4, d, rt.stream()), "memcpy2d")
    np.testing.assert_array_equal(host(full), E)
