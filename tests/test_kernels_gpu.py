"""GPU parity tests of the C-ABI kernels against the numpy oracle (oracle/intexact.py) on identical seeded inputs.

Bars (north star): int32 accumulators and requantized int8 tensors bit-exact; fp32 epilogues bit-exact (same op
order, no FMA); float reductions (LayerNorm, softmax, P.V) within 1e-3 relative (observed ~1e-6).
"""
import numpy as np
import pytest
import torch

from oracle import intexact as ox

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    from onnx_transformer_b200 import kernels
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return kernels


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rand_i8(rng, shape):
    return rng.integers(-127, 128, size=shape, dtype=np.int8)


@pytest.mark.parametrize("M,N,K_", [(64, 512, 512), (128, 512, 512), (200, 1536, 512), (64, 2048, 512), (64, 512, 2048),
                                    (1000, 512, 2048), (4096, 2048, 512), (4096, 512, 2048), (1, 512, 512), (300, 96, 144)])
def test_gemm_int32_bit_exact(K, M, N, K_):
    rng = np.random.default_rng(M * 7 + N * 3 + K_)
    a, w = rand_i8(rng, (M, K_)), rand_i8(rng, (N, K_))
    out = K.linear_w8a8(dev(a), dev(w), out_kind=K.OUT_I32).cpu().numpy()
    ref = ox.int_matmul(a, w)
    assert out.dtype == np.int32 and np.array_equal(out, ref)


@pytest.mark.parametrize("M,N,K_,relu,res", [(64, 512, 512, False, True), (257, 2048, 512, True, False), (4096, 512, 2048, False, True),
                                             (64, 1536, 512, False, False)])
def test_gemm_fp32_epilogue_bit_exact(K, M, N, K_, relu, res):
    rng = np.random.default_rng(11 + M + N)
    a, w = rand_i8(rng, (M, K_)), rand_i8(rng, (N, K_))
    sx = rng.uniform(1e-3, 5e-2, size=M).astype(np.float32)
    sw = rng.uniform(1e-4, 1e-2, size=N).astype(np.float32)
    b = rng.normal(size=N).astype(np.float32)
    r = rng.normal(size=(M, N)).astype(np.float32) if res else None
    out = K.linear_w8a8(dev(a), dev(w), row_scale=dev(sx), col_scale=dev(sw), bias=dev(b), residual=dev(r) if res else None, relu=relu,
                        out_kind=K.OUT_F32).cpu().numpy()
    ref = ox.linear_w8a8(a, sx, w, sw, b, relu, r)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("M,N,K_,group,relu", [(64, 1536, 512, 512, False), (64, 2048, 512, 2048, True), (4096, 1536, 512, 512, False),
                                               (4096, 2048, 512, 2048, True), (130, 512, 2048, 512, False), (64, 1024, 512, 512, False)])
def test_gemm_fused_requant_bit_exact(K, M, N, K_, group, relu):
    rng = np.random.default_rng(5 + M + N + group)
    a, w = rand_i8(rng, (M, K_)), rand_i8(rng, (N, K_))
    sx = rng.uniform(1e-3, 5e-2, size=M).astype(np.float32)
    sw = rng.uniform(1e-4, 1e-2, size=N).astype(np.float32)
    b = rng.normal(size=N).astype(np.float32)
    q, s = K.linear_w8a8(dev(a), dev(w), row_scale=dev(sx), col_scale=dev(sw), bias=dev(b), relu=relu, out_kind=K.OUT_Q8,
                         quant_group=group)
    y = ox.linear_w8a8(a, sx, w, sw, b, relu)
    qr, sr = ox.group_quant(y, group)
    assert np.array_equal(s.cpu().numpy().view(np.uint32), sr.view(np.uint32))
    assert np.array_equal(q.cpu().numpy(), qr)


@pytest.mark.parametrize("M,N,K_,group,relu,res", [(2100, 1536, 512, 512, False, False), (8192, 2048, 512, 2048, True, False), (3000, 512, 512, 0, False, True),
                                                   (5000, 512, 2048, 0, False, True), (2048, 6144, 512, 512, False, False), (4096, 1024, 256, 256, True, False),
                                                   # fp32 output through per-warp TMA boxes (K <= 512), without / with a residual, ragged rows
                                                   (2500, 512, 512, 0, False, False), (2049, 768, 256, 0, True, True),
                                                   # CTA pairs (cta_group::2, K > 512): an odd number of 128-row tiles (the peer's last tile is empty),
                                                   # one column tile, ReLU without a residual
                                                   (2176, 1024, 1024, 0, False, True), (2048, 256, 2048, 0, True, False)])
def test_gemm_stream_kernel_equals_tile_kernel_and_oracle(K, M, N, K_, group, relu, res):
    """The persistent double-accumulator kernel (ot_gemm_stream.cu, M >= 2048) against the one-tile-per-CTA kernel bit for bit, and both
    against the oracle on a row sample (ragged last tiles, clusters of 1 / 2 / 8, the 12-projection cross-K/V shape, the TMA epilogue of
    the K <= 512 fp32 GEMMs, the CTA-pair form of the K > 512 ones)."""
    import os
    rng = np.random.default_rng(31 + M + N)
    a, w = rand_i8(rng, (M, K_)), rand_i8(rng, (N, K_))
    sx = rng.uniform(1e-3, 5e-2, size=M).astype(np.float32)
    sw = rng.uniform(1e-4, 1e-2, size=N).astype(np.float32)
    b = rng.normal(size=N).astype(np.float32)
    r = rng.normal(size=(M, N)).astype(np.float32) if res else None
    ad, wd_, sxd, swd, bd, rd = dev(a), dev(w), dev(sx), dev(sw), dev(b), (dev(r) if res else None)

    def run():
        if group:
            return K.linear_w8a8(ad, wd_, row_scale=sxd, col_scale=swd, bias=bd, relu=relu, out_kind=K.OUT_Q8, quant_group=group)
        return (K.linear_w8a8(ad, wd_, row_scale=sxd, col_scale=swd, bias=bd, residual=rd, relu=relu, out_kind=K.OUT_F32),)
    n0 = K._lib.launch_count()
    new = run()
    os.environ["OT_GEMM_STREAM"] = "0"
    try:
        old = run()
    finally:
        del os.environ["OT_GEMM_STREAM"]
    assert K._lib.launch_count() == n0 + 2
    for x, y in zip(new, old):
        assert torch.equal(x, y)
    rows = np.unique(np.concatenate([np.arange(0, 130), np.arange(M - 140, M), rng.integers(0, M, size=300)]))
    yo = ox.linear_w8a8(a[rows], sx[rows], w, sw, b, relu, r[rows] if res else None)
    if group:
        qr, sr = ox.group_quant(yo, group)
        assert np.array_equal(new[0].cpu().numpy()[rows], qr) and np.array_equal(new[1].cpu().numpy()[rows].view(np.uint32), sr.view(np.uint32))
    else:
        assert np.array_equal(new[0].cpu().numpy()[rows].view(np.uint32), yo.view(np.uint32))


def test_gemm_stream_kernel_non_finite_inputs_take_the_exact_path(K):
    """Rows with an infinite / NaN scale and a NaN bias column: same bytes as the tile kernel (whose chunk-wise slow path is exact)."""
    import os
    rng = np.random.default_rng(5)
    M, N, K_ = 2048, 512, 512
    a, w = rand_i8(rng, (M, K_)), rand_i8(rng, (N, K_))
    sx = rng.uniform(1e-3, 5e-2, size=M).astype(np.float32)
    sx[7], sx[300], sx[2047] = np.inf, np.nan, 3e38
    sw = rng.uniform(1e-4, 1e-2, size=N).astype(np.float32)
    b = rng.normal(size=N).astype(np.float32)
    args = dict(row_scale=dev(sx), col_scale=dev(sw), bias=dev(b), out_kind=K.OUT_Q8, quant_group=512)
    new = K.linear_w8a8(dev(a), dev(w), **args)
    os.environ["OT_GEMM_STREAM"] = "0"
    try:
        old = K.linear_w8a8(dev(a), dev(w), **args)
    finally:
        del os.environ["OT_GEMM_STREAM"]
    assert torch.equal(new[0], old[0]) and torch.equal(new[1].view(torch.int32), old[1].view(torch.int32))
    b2 = b.copy(); b2[100] = np.nan
    args["bias"] = dev(b2)
    new = K.linear_w8a8(dev(a), dev(w), **args)
    os.environ["OT_GEMM_STREAM"] = "0"
    try:
        old = K.linear_w8a8(dev(a), dev(w), **args)
    finally:
        del os.environ["OT_GEMM_STREAM"]
    assert torch.equal(new[0], old[0]) and torch.equal(new[1].view(torch.int32), old[1].view(torch.int32))


@pytest.mark.parametrize("M,N,group,relu", [(64, 1536, 512, False), (64, 512, 512, False), (64, 2048, 2048, True), (200, 1536, 512, False), (1, 512, 512, False)])
def test_gemm_layernorm_prologue_equals_unfused(K, M, N, group, relu):
    """ot_ln_linear_w8a8 == ot_layernorm_quant followed by ot_linear_w8a8, bit for bit (same op order), and vs the oracle."""
    rng = np.random.default_rng(77 + M + N)
    x = (rng.normal(size=(M, 512)) * 2 + 0.5).astype(np.float32)
    ga = (1 + 0.1 * rng.normal(size=512)).astype(np.float32)
    be = (0.1 * rng.normal(size=512)).astype(np.float32)
    w = rand_i8(rng, (N, 512))
    sw = rng.uniform(1e-4, 1e-2, size=N).astype(np.float32)
    b = rng.normal(size=N).astype(np.float32)
    xd, gd, bd, wd, swd, bbd = dev(x), dev(ga), dev(be), dev(w), dev(sw), dev(b)
    q, s = K.ln_linear_w8a8(xd, gd, bd, wd, col_scale=swd, bias=bbd, relu=relu, out_kind=K.OUT_Q8, quant_group=group)
    _, xq, sx = K.layernorm_quant(xd, gd, bd, want_q=True)
    q2, s2 = K.linear_w8a8(xq, wd, row_scale=sx, col_scale=swd, bias=bbd, relu=relu, out_kind=K.OUT_Q8, quant_group=group)
    assert torch.equal(q, q2) and torch.equal(s, s2)
    f = K.ln_linear_w8a8(xd, gd, bd, wd, col_scale=swd, bias=bbd, residual=xd if N == 512 else None, out_kind=K.OUT_F32)
    f2 = K.linear_w8a8(xq, wd, row_scale=sx, col_scale=swd, bias=bbd, residual=xd if N == 512 else None, out_kind=K.OUT_F32)
    assert torch.equal(f, f2)
    # and against the oracle, given the kernel's own LayerNorm rounding (float tolerance class): integer tensors from identical xq
    y = ox.linear_w8a8(xq.cpu().numpy(), sx.cpu().numpy(), w, sw, b, relu)
    qr, sr = ox.group_quant(y, group)
    assert np.array_equal(q.cpu().numpy(), qr) and np.array_equal(s.cpu().numpy().view(np.uint32), sr.view(np.uint32))


def test_gemm_w4_bit_exact(K):
    rng = np.random.default_rng(3)
    M, N, K_ = 192, 512, 512
    a = rand_i8(rng, (M, K_))
    w = rng.integers(-8, 8, size=(N, K_), dtype=np.int8)
    packed = ((w[:, 0::2].astype(np.uint8) & 0xF) | ((w[:, 1::2].astype(np.uint8) & 0xF) << 4)).astype(np.uint8)
    assert np.array_equal(K.unpack_int4(dev(packed)).cpu().numpy(), w)
    out = K.linear_w8a8(dev(a), dev(packed), out_kind=K.OUT_I32, w4=True).cpu().numpy()
    assert np.array_equal(out, ox.int_matmul(a, w))


@pytest.mark.parametrize("M,N,K_,group,relu", [(4096, 1536, 512, 512, False), (2100, 2048, 512, 2048, True), (2048, 512, 256, 512, False)])
def test_gemm_w4_weight_stationary_requant_bit_exact(K, M, N, K_, group, relu):
    """cfg4 at encoder sizes: packed int4 weights through the weight-stationary requant kernel (gemm_wres_kernel<.., W4>: the tile is
    unpacked once per launch into the resident swizzled int8 layout) == the same GEMM on the unpacked int8 copy == the oracle, bit for bit."""
    rng = np.random.default_rng(M + N)
    a = rand_i8(rng, (M, K_))
    w = rng.integers(-8, 8, size=(N, K_), dtype=np.int8)
    packed = ((w[:, 0::2].astype(np.uint8) & 0xF) | ((w[:, 1::2].astype(np.uint8) & 0xF) << 4)).astype(np.uint8)
    sx = rng.uniform(1e-3, 5e-2, size=M).astype(np.float32)
    sw = rng.uniform(1e-3, 1e-1, size=N).astype(np.float32)
    b = rng.normal(size=N).astype(np.float32)
    n0 = K._lib.launch_count()
    q4, s4 = K.linear_w8a8(dev(a), dev(packed), row_scale=dev(sx), col_scale=dev(sw), bias=dev(b), relu=relu, out_kind=K.OUT_Q8, quant_group=group, w4=True)
    assert K._lib.launch_count() == n0 + 1
    q8, s8 = K.linear_w8a8(dev(a), dev(w), row_scale=dev(sx), col_scale=dev(sw), bias=dev(b), relu=relu, out_kind=K.OUT_Q8, quant_group=group)
    assert torch.equal(q4, q8) and torch.equal(s4, s8)
    rows = np.unique(np.concatenate([np.arange(0, 130), np.arange(M - 140, M), rng.integers(0, M, size=200)]))
    qr, sr = ox.group_quant(ox.linear_w8a8(a[rows], sx[rows], w, sw, b, relu, None), group)
    assert np.array_equal(q4.cpu().numpy()[rows], qr) and np.array_equal(s4.cpu().numpy()[rows].view(np.uint32), sr.view(np.uint32))


def test_gemm_faults(K):
    rng = np.random.default_rng(9)
    M, N, K_ = 96, 512, 512
    a, w = rand_i8(rng, (M, K_)), rand_i8(rng, (N, K_))
    golden = ox.int_matmul(a, w)
    # INPUT fault at A[i,k], bit 6, 16-column window
    i, k, bit = 37, 211, 6
    f = K.make_fault(K.FAULT_INPUT, flat_index=i * K_ + k, bit=bit, window_start=32, window_len=16)
    out = K.linear_w8a8(dev(a), dev(w), out_kind=K.OUT_I32, fault=f).cpu().numpy()
    ref = golden.copy()
    delta = ox.flip_int8_bit(int(a[i, k]), bit) - int(a[i, k])
    ref[i, 32:48] += delta * w[32:48, k].astype(np.int32)
    assert np.array_equal(out, ref)
    # WEIGHT fault at W[n,k], whole column
    n, k, bit = 300, 5, 7
    f = K.make_fault(K.FAULT_WEIGHT, flat_index=n * K_ + k, bit=bit)
    out = K.linear_w8a8(dev(a), dev(w), out_kind=K.OUT_I32, fault=f).cpu().numpy()
    ref = golden.copy()
    delta = ox.flip_int8_bit(int(w[n, k]), bit) - int(w[n, k])
    ref[:, n] += a[:, k].astype(np.int32) * delta
    assert np.array_equal(out, ref)
    # RANDOM_BITFLIP on the fp32 MatMul output (before bias)
    sx = rng.uniform(1e-3, 5e-2, size=M).astype(np.float32)
    sw = rng.uniform(1e-4, 1e-2, size=N).astype(np.float32)
    b = rng.normal(size=N).astype(np.float32)
    r_, c_, bit = 5, 77, 30
    f = K.make_fault(K.FAULT_RANDOM_BITFLIP, flat_index=r_ * N + c_, bit=bit)
    out = K.linear_w8a8(dev(a), dev(w), row_scale=dev(sx), col_scale=dev(sw), bias=dev(b), out_kind=K.OUT_F32, fault=f).cpu().numpy()
    mm = ox.linear_epilogue(golden, sx, sw)
    mm[r_, c_] = ox.float32_bit_flip(mm[r_, c_], bit)
    ref = (mm + b.reshape(1, -1)).astype(np.float32)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("rows,n,group", [(64, 512, 512), (4096, 2048, 2048), (100, 1536, 512), (7, 512, 64)])
def test_rowquant_bit_exact(K, rows, n, group):
    rng = np.random.default_rng(rows + n)
    x = (rng.normal(size=(rows, n)) * rng.uniform(0.01, 10, size=(rows, 1))).astype(np.float32)
    x[0, :] = 0.0  # clamp path
    q, s = K.rowquant(dev(x), group)
    qr, sr = ox.group_quant(x, group)
    assert np.array_equal(s.cpu().numpy().reshape(rows, -1).view(np.uint32), sr.view(np.uint32))
    assert np.array_equal(q.cpu().numpy(), qr)


def test_layernorm_quant(K):
    rng = np.random.default_rng(0)
    rows, n = 777, 512
    x = rng.normal(size=(rows, n)).astype(np.float32) * 3 + 1
    g = rng.normal(size=n).astype(np.float32)
    b = rng.normal(size=n).astype(np.float32)
    y, q, s = K.layernorm_quant(dev(x), dev(g), dev(b), want_y=True, want_q=True)
    y = y.cpu().numpy()
    ref = ox.layer_norm(x, g, b)
    np.testing.assert_allclose(y, ref, rtol=1e-3, atol=1e-5)   # float tolerance class of the north star
    assert np.max(np.abs(y - ref)) < 1e-5
    # the fused quantization must be the exact RowQuant of the kernel's own y
    qr, sr = ox.row_quant(y)
    assert np.array_equal(q.cpu().numpy(), qr)
    assert np.array_equal(s.cpu().numpy().view(np.uint32), sr.reshape(-1).view(np.uint32))


def test_residual_embed(K):
    rng = np.random.default_rng(1)
    a = rng.normal(size=(64, 512)).astype(np.float32)
    b = rng.normal(size=(64, 512)).astype(np.float32)
    assert np.array_equal(K.residual_add(dev(a), dev(b)).cpu().numpy(), a + b)
    lut = rng.normal(size=(100, 512)).astype(np.float32)
    pe = ox.positional_encoding(80)
    ids = rng.integers(0, 100, size=(4, 9))
    out = K.embed_pe(dev(ids.reshape(-1)), dev(lut), dev(pe), seq_len=9).cpu().numpy().reshape(4, 9, 512)
    assert np.array_equal(out, ox.embed(ids, lut, pe))


def _attn_inputs(rng, B, Tq, Tk):
    qq, kq, vq = rand_i8(rng, (B, Tq, 512)), rand_i8(rng, (B, Tk, 512)), rand_i8(rng, (B, Tk, 512))
    sq = rng.uniform(0.005, 0.02, size=(B, Tq)).astype(np.float32)
    sk = rng.uniform(0.005, 0.02, size=(B, Tk)).astype(np.float32)
    sv = rng.uniform(0.005, 0.02, size=(B, Tk)).astype(np.float32)
    return qq, sq, kq, sk, vq, sv


def _check_attn(ctx, pq, ref_ctx, ref_pq, ref_p):
    diff = pq.astype(np.int32) - ref_pq.astype(np.int32)
    assert np.max(np.abs(diff)) <= 1
    bad = diff != 0
    # any disagreement must sit on a rounding boundary of 127 p (float softmax within tolerance)
    if bad.any():
        frac = (ref_p * 127.0)[bad] % 1.0
        assert np.all(np.abs(frac - 0.5) < 1e-3)
        assert bad.mean() < 1e-3
    np.testing.assert_allclose(ctx, ref_ctx, rtol=1e-3, atol=2e-3 if bad.any() else 1e-5)


@pytest.mark.parametrize("B,Tq,Tk,mask", [(3, 64, 64, "pad"), (2, 37, 37, "causal"), (2, 1, 50, "none"), (2, 128, 128, "pad"), (5, 1, 72, "pad")])
def test_attention(K, B, Tq, Tk, mask):
    rng = np.random.default_rng(B * 100 + Tq + Tk)
    qq, sq, kq, sk, vq, sv = _attn_inputs(rng, B, Tq, Tk)
    key_mask = None
    if mask == "pad":
        lens = rng.integers(max(1, Tk // 3), Tk + 1, size=B)
        key_mask = (np.arange(Tk)[None, :] < lens[:, None]).astype(np.uint8)
    ctx, cq, cs, probs = K.attention_q8(dev(qq), dev(sq), dev(kq), dev(vq), dev(sk), dev(sv), B=B, Tq=Tq, Tk=Tk,
                                        mask_kind={"none": 0, "pad": 1, "causal": 2}[mask],
                                        key_mask=dev(key_mask) if key_mask is not None else None,
                                        want_ctx=True, want_q=True, want_probs=True)
    ctx = ctx.cpu().numpy().reshape(B, Tq, 512)
    probs = probs.cpu().numpy()
    for b in range(B):
        rc, rpq, rp = ox.attention(qq[b], sq[b], kq[b], sk[b], vq[b], sv[b], key_mask[b] if key_mask is not None else None,
                                   causal=(mask == "causal"), return_all=True)
        _check_attn(ctx[b], probs[b], rc, rpq, rp)
    # fused RowQuant of the context == oracle RowQuant of the kernel's own fp32 context (bit-exact)
    qr, sr = ox.row_quant(ctx.reshape(B * Tq, 512))
    assert np.array_equal(cq.cpu().numpy(), qr)
    assert np.array_equal(cs.cpu().numpy().view(np.uint32), sr.reshape(-1).view(np.uint32))


@pytest.mark.parametrize("Tq,Tk", [(64, 64), (100, 128), (128, 40)])
def test_attention_tensor_core_kernel_vs_cuda_core_kernel_and_oracle(K, Tq, Tk):
    """ot_attention_tc.cu (tcgen05 scores + fp16 hi/lo context) against the dp4a / fp32-FMA kernel it replaces: the quantized
    probabilities may differ only at verified rounding boundaries, the context within the float class; measured error vs the float64
    oracle reported through the assertion bounds (P.V: hi/lo split of s_v*vq = 22 significant bits)."""
    import os
    rng = np.random.default_rng(Tq * 7 + Tk)
    B = 3
    qq, sq, kq, sk, vq, sv = _attn_inputs(rng, B, Tq, Tk)
    sv[:, 3] *= 40.0                                   # a wide spread of V scales exercises the power-of-two pre-scaling of the split
    sv[:, 5] *= 1e-3
    lens = rng.integers(max(1, Tk // 3), Tk + 1, size=B)
    key_mask = (np.arange(Tk)[None, :] < lens[:, None]).astype(np.uint8)
    args = (dev(qq), dev(sq), dev(kq), dev(vq), dev(sk), dev(sv))
    kw = dict(B=B, Tq=Tq, Tk=Tk, mask_kind=1, key_mask=dev(key_mask), want_ctx=True, want_q=True, want_probs=True)
    ctx, cq, cs, probs = K.attention_q8(*args, **kw)
    os.environ["OT_ATTN_TC"] = "0"
    try:
        ctx0, cq0, cs0, probs0 = K.attention_q8(*args, **kw)
    finally:
        del os.environ["OT_ATTN_TC"]
    ctx, ctx0 = ctx.cpu().numpy().reshape(B, Tq, 512), ctx0.cpu().numpy().reshape(B, Tq, 512)
    probs, probs0 = probs.cpu().numpy(), probs0.cpu().numpy()
    worst = 0.0
    for b in range(B):
        rc, rpq, rp = ox.attention(qq[b], sq[b], kq[b], sk[b], vq[b], sv[b], key_mask[b], return_all=True)
        _check_attn(ctx[b], probs[b], rc, rpq, rp)
        _check_attn(ctx0[b], probs0[b], rc, rpq, rp)
        same = (probs[b] == rpq).all(axis=2)                      # [8, Tq]: rows whose probabilities equal the oracle's
        vmax = np.abs(vq[b].astype(np.float32) * sv[b].reshape(-1, 1)).reshape(Tk, 8, 64).max(axis=(0, 2))
        err = np.abs(ctx[b] - rc).reshape(Tq, 8, 64) / vmax[None, :, None]
        worst = max(worst, float(err[same.T].max()))
    assert worst < 2e-6, worst                                     # relative to the head's largest |vhat|: fp32-class accuracy
    assert (probs != probs0).mean() < 1e-4
    qr, sr = ox.row_quant(ctx.reshape(B * Tq, 512))
    assert np.array_equal(cq.cpu().numpy(), qr) and np.array_equal(cs.cpu().numpy().view(np.uint32), sr.reshape(-1).view(np.uint32))


@pytest.mark.parametrize("case", ["q_input16", "k_weight", "scores_bitflip", "scores_random", "p_input16", "v_weight16", "ctx_bitflip"])
def test_attention_tensor_core_kernel_faults(K, case):
    """Fault hooks of ot_attention_tc.cu vs the oracle's integer-domain restatement; every element the fault does not
    reach is bit-identical to the golden launch (same kernel, same data)."""
    rng = np.random.default_rng(99)
    B, T = 2, 64
    qq, sq, kq, sk, vq, sv = _attn_inputs(rng, B, T, T)
    args = (dev(qq), dev(sq), dev(kq), dev(vq), dev(sk), dev(sv))
    kw = dict(B=B, Tq=T, Tk=T, mask_kind=0, want_ctx=True, want_q=False, want_probs=True)
    g_ctx, _, _, g_probs = K.attention_q8(*args, **kw)
    b, h, i, j, d, t = 1, 5, 17, 40, 33, 22
    c = h * 64 + d
    spec = {
        "q_input16": (K.FAULT_INPUT, K.OPERAND_Q, (b * T + t) * 512 + c, 6, 32, 16, 0, dict(target="qk", type="INPUT16", index=(0, t, c))),
        "k_weight": (K.FAULT_WEIGHT, K.OPERAND_K, (b * T + t) * 512 + c, 7, 0, 0, 0, dict(target="qk", type="WEIGHT", index=(0, t, c))),
        "scores_bitflip": (K.FAULT_RANDOM_BITFLIP, K.OPERAND_SCORES, ((b * 8 + h) * T + i) * T + j, 30, 0, 0, 0, dict(target="qk", type="RANDOM_BITFLIP", index=(0, h, i, j))),
        "scores_random": (K.FAULT_RANDOM, K.OPERAND_SCORES, ((b * 8 + h) * T + i) * T + j, 0, 0, 0, 0x42f00000, dict(target="qk", type="RANDOM", index=(0, h, i, j))),
        "p_input16": (K.FAULT_INPUT, K.OPERAND_P, ((b * 8 + h) * T + i) * T + j, 6, 16, 16, 0, dict(target="pv", type="INPUT16", index=(0, h, i, j))),
        "v_weight16": (K.FAULT_WEIGHT, K.OPERAND_V, (b * T + t) * 512 + c, 7, 16, 9, 0, dict(target="pv", type="WEIGHT16", index=(0, t, c))),
        "ctx_bitflip": (K.FAULT_RANDOM_BITFLIP, K.OPERAND_CTX, ((b * 8 + h) * T + i) * 64 + d, 29, 0, 0, 0, dict(target="pv", type="RANDOM_BITFLIP", index=(0, h, i, d))),
    }[case]
    mode, operand, flat, bit, ws, wl, vbits, of = spec
    fault = K.make_fault(mode, flat_index=flat, bit=bit, window_start=ws, window_len=wl, value_bits=vbits, operand=operand)
    f_ctx, _, _, f_probs = K.attention_q8(*args, fault=fault, **kw)
    g, fc = g_ctx.cpu().numpy().reshape(B, T, 512), f_ctx.cpu().numpy().reshape(B, T, 512)
    assert np.array_equal(g[0].view(np.uint32), fc[0].view(np.uint32))                  # the other sentence: untouched
    other_heads = np.ones(512, bool); other_heads[h * 64:(h + 1) * 64] = False
    assert np.array_equal(g[1][:, other_heads].view(np.uint32), fc[1][:, other_heads].view(np.uint32))
    of.update(bit=bit, window_start=ws, window_len=wl, value_bits=vbits)
    rc, rpq, rp = ox.attention(qq[b], sq[b], kq[b], sk[b], vq[b], sv[b], None, return_all=True, fault=of)
    rg = ox.attention(qq[b], sq[b], kq[b], sk[b], vq[b], sv[b], None)
    assert not np.array_equal(rc, rg), "the oracle's fault must change something"
    _check_attn(fc[1], f_probs.cpu().numpy()[1], rc, rpq, rp)
    changed = np.abs(rc - rg) > 1e-4 * (1 + np.abs(rg))
    assert np.all(np.abs(fc[1] - g[1])[changed] > 0), "the fault did not reach the elements the oracle changes"
    # batched-trial entry point (one fault per sentence, indices relative to the sentence): same bits as the single-fault launch
    rel = K.make_fault(mode, flat_index=flat - {K.OPERAND_Q: b * T * 512, K.OPERAND_K: b * T * 512, K.OPERAND_V: b * T * 512,
                                                 K.OPERAND_P: b * 8 * T * T, K.OPERAND_SCORES: b * 8 * T * T, K.OPERAND_CTX: b * 8 * T * 64}[operand],
                       bit=bit, window_start=ws, window_len=wl, value_bits=vbits, operand=operand)
    unit = torch.tensor([-1, 0], dtype=torch.int32, device="cuda")
    m_ctx, _, _, _ = K.attention_q8(*args, mf=(K.pack_faults([rel], "cuda"), unit), B=B, Tq=T, Tk=T, mask_kind=0, want_ctx=True, want_q=False)
    assert torch.equal(m_ctx.view(torch.int32), f_ctx.view(torch.int32))


def test_attention_kv_cache_append(K):
    """Decode-style: Tq=1, cache of capacity 72 holding t keys, the new key/value appended by the kernel."""
    rng = np.random.default_rng(42)
    B, cap, t = 4, 72, 17
    qq, sq, kq, sk, vq, sv = _attn_inputs(rng, B, 1, t + 1)
    kc = np.zeros((B, cap, 512), np.int8); vc = np.zeros((B, cap, 512), np.int8)
    skc = np.zeros((B, cap), np.float32); svc = np.zeros((B, cap), np.float32)
    kc[:, :t], vc[:, :t], skc[:, :t], svc[:, :t] = kq[:, :t], vq[:, :t], sk[:, :t], sv[:, :t]
    kcd, vcd, skd, svd = dev(kc), dev(vc), dev(skc), dev(svc)
    step = torch.tensor([t], dtype=torch.int32, device="cuda")
    ctx, _, _, _ = K.attention_q8(dev(qq), dev(sq), kcd, vcd, skd, svd, B=B, Tq=1, Tk=t + 1, Tk_cap=cap,
                                  k_new=dev(kq[:, t:]), v_new=dev(vq[:, t:]), sk_new=dev(sk[:, t:]), sv_new=dev(sv[:, t:]), ld_new=512,
                                  mask_kind=2, step_dev=step)
    ctx = ctx.cpu().numpy().reshape(B, 1, 512)
    for b in range(B):
        rc = ox.attention(qq[b], sq[b], kq[b], sk[b], vq[b], sv[b], causal=True, q_pos0=t)
        np.testing.assert_allclose(ctx[b], rc, rtol=1e-3, atol=2e-3)
    assert np.array_equal(kcd.cpu().numpy()[:, t], kq[:, t]) and np.array_equal(vcd.cpu().numpy()[:, t], vq[:, t])
    assert np.array_equal(skd.cpu().numpy()[:, t], sk[:, t]) and np.array_equal(svd.cpu().numpy()[:, t], sv[:, t])


def test_generator_argmax(K):
    rng = np.random.default_rng(8)
    rows, d, vocab = 64, 512, 4444
    h = rng.normal(size=(rows, d)).astype(np.float32)
    W = (rng.normal(size=(vocab, d)) * 0.05).astype(np.float32)
    b = rng.normal(size=vocab).astype(np.float32) * 0.1
    ids, logp, margin, logits = K.generator_argmax(dev(h), dev(W), dev(b), want_logp=True, want_margin=True)
    ref_ids, ref_logits = ox.generator(h, W, b)
    np.testing.assert_allclose(logits.cpu().numpy(), ref_logits, rtol=1e-4, atol=1e-4)
    srt = np.sort(ref_logits, axis=-1)
    safe = (srt[:, -1] - srt[:, -2]) > 1e-3
    assert np.array_equal(ids.cpu().numpy()[safe], ref_ids[safe])
    ref_logp = ref_logits - np.log(np.sum(np.exp(ref_logits - ref_logits.max(-1, keepdims=True)), -1, keepdims=True)) - ref_logits.max(-1, keepdims=True)
    np.testing.assert_allclose(logp.cpu().numpy(), ref_logp, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(margin.cpu().numpy(), srt[:, -1] - srt[:, -2], atol=1e-4)


def test_elementwise_family(K):
    rng = np.random.default_rng(2)
    x = rng.normal(size=(2, 8, 5, 7)).astype(np.float32)
    y = rng.normal(size=(1, 8, 1, 7)).astype(np.float32)
    xd, yd = dev(x), dev(y)
    assert np.array_equal(K.unary("Abs", xd).cpu().numpy(), np.abs(x))
    assert np.array_equal(K.unary("Relu", xd).cpu().numpy(), np.maximum(x, 0))
    assert np.array_equal(K.unary("Round", dev(x * 10)).cpu().numpy(), np.rint(x * 10))
    assert np.array_equal(K.unary("Sqrt", dev(np.abs(x))).cpu().numpy(), np.sqrt(np.abs(x)))
    for op, fn in [("Add", np.add), ("Sub", np.subtract), ("Mul", np.multiply), ("Div", np.divide)]:
        assert np.array_equal(K.binary(op, xd, yd).cpu().numpy(), fn(x, y).astype(np.float32)), op
        assert np.array_equal(K.binary(op, yd, xd).cpu().numpy(), fn(y, x).astype(np.float32)), op
    assert np.array_equal(K.clip(xd, 1e-5, 3.4e38).cpu().numpy(), np.clip(x, np.float32(1e-5), np.float32(3.4e38)))
    assert np.array_equal(K.reduce_last("ReduceMax", dev(np.abs(x))).cpu().numpy(), np.abs(x).max(-1, keepdims=True))
    np.testing.assert_allclose(K.reduce_last("ReduceMean", xd).cpu().numpy(), x.mean(-1, keepdims=True), rtol=1e-5, atol=1e-6)
    sm = np.exp(x - x.max(-1, keepdims=True)); sm /= sm.sum(-1, keepdims=True)
    np.testing.assert_allclose(K.softmax_last(xd).cpu().numpy(), sm, rtol=1e-5, atol=1e-7)
    cond = rng.integers(0, 2, size=(2, 1, 1, 7)).astype(bool)
    assert np.array_equal(K.where_scalar(dev(cond), -1e9, xd).cpu().numpy(), np.where(cond, np.float32(-1e9), x))
    m = rng.integers(0, 2, size=(3, 1, 9)).astype(np.int64)
    assert np.array_equal(K.equal_scalar_i64(dev(m), 0).cpu().numpy(), m == 0)
    assert np.array_equal(K.cast(dev(cond), torch.int64).cpu().numpy(), cond.astype(np.int64))
    assert np.array_equal(K.cast(dev(x * 50), torch.int8).cpu().numpy(), (x * 50).astype(np.int8))
    assert np.array_equal(K.transpose(xd, (0, 2, 1, 3)).cpu().numpy(), x.transpose(0, 2, 1, 3))
    assert np.array_equal(K.transpose(xd, (0, 2, 3, 1)).cpu().numpy(), x.transpose(0, 2, 3, 1))
    a = rng.normal(size=(2, 8, 33, 64)).astype(np.float32)
    b = rng.normal(size=(2, 8, 64, 47)).astype(np.float32)
    np.testing.assert_allclose(K.matmul_f32(dev(a), dev(b)).cpu().numpy(), a @ b, rtol=1e-4, atol=1e-4)
    w = rng.normal(size=(64, 20)).astype(np.float32)
    np.testing.assert_allclose(K.matmul_f32(dev(a), dev(w)).cpu().numpy(), a @ w, rtol=1e-4, atol=1e-4)


def test_attention_tc_fused_rowquant_equals_two_launches(K):
    """ot_attention_tc.cu with the RowQuant of the merged rows inside (cluster of the 8 head CTAs; taken when no fp32 context buffer is
    passed) against the default two-launch path (fp32 context + rowquant_kernel): int8 rows and scales bit for bit, ragged lengths."""
    rng = np.random.default_rng(77)
    for B, S in ((3, 128), (2, 41)):
        M = B * S
        qkv = dev(rng.integers(-127, 128, size=(M, 1536), dtype=np.int8))
        sqkv = dev(rng.uniform(1e-3, 2e-2, size=(M, 3)).astype(np.float32))
        mask = np.ones((B, S), dtype=np.uint8)
        mask[:, S - 5:] = 0
        mask = dev(mask)
        kw = dict(B=B, Tq=S, Tk=S, ldq=1536, sq_stride=3, ldk=1536, skv_stride=3, mask_kind=1, key_mask=mask, mask_stride=S)
        args = (qkv, sqkv, qkv[:, 512:], qkv[:, 1024:], sqkv[:, 1:], sqkv[:, 2:])
        _, q2, s2, _ = K.attention_q8(*args, want_ctx=True, want_q=True, **kw)
        _, q1, s1, _ = K.attention_q8(*args, want_ctx=False, want_q=True, **kw)
        assert torch.equal(q1, q2) and torch.equal(s1.view(torch.int32), s2.view(torch.int32))


def test_layernorm_quant512_fast_kernel_equals_general_kernel(K):
    """layernorm_quant512_kernel (rows >= 2048: packed fp32 arithmetic, divisions as exact FMA quotients) against the general kernel
    bit for bit -- int8 rows and scales -- on ordinary rows and on rows the fast path must hand back: all-zero, constant, tiny, huge,
    infinite and NaN rows, zero gamma / beta columns."""
    import os
    rng = np.random.default_rng(12)
    rows, n = 4100, 512
    x = (rng.normal(size=(rows, n)) * rng.uniform(0.01, 30.0, size=(rows, 1)) + rng.normal(size=(rows, 1))).astype(np.float32)
    x[5] = 0.0
    x[6] = 3.25
    x[7] *= 1e-30
    x[8] *= 1e30
    x[9, 100] = np.inf
    x[10, 3] = np.nan
    x[11] = 0.0; x[11, 0] = 1.0; x[11, 1] = -1.0; x[11, 2] = 1e-36
    x[12] *= 1e-20
    g = rng.normal(size=n).astype(np.float32)
    b = rng.normal(size=n).astype(np.float32)
    g[17] = 0.0
    b[17] = 0.0
    b[300:310] = 0.0
    xd, gd, bd = dev(x), dev(g), dev(b)
    n0 = K._lib.launch_count()
    _, q1, s1 = K.layernorm_quant(xd, gd, bd, want_y=False, want_q=True)
    os.environ["OT_LN512_MIN_ROWS"] = str(1 << 40)
    try:
        _, q0, s0 = K.layernorm_quant(xd, gd, bd, want_y=False, want_q=True)
    finally:
        del os.environ["OT_LN512_MIN_ROWS"]
    assert K._lib.launch_count() == n0 + 2
    assert torch.equal(s1.view(torch.int32), s0.view(torch.int32))
    assert torch.equal(q1, q0)
    # and against the oracle's LayerNorm + RowQuant with the boundary accounting of the float-reduction class
    ref = ox.layer_norm(x[:4000], g, b)
    qr, _ = ox.row_quant(ref)
    ok = np.all(np.isfinite(x[:4000]), axis=1) & (np.abs(x[:4000]).max(axis=1) < 1e20) & (np.abs(x[:4000]).max(axis=1) > 1e-10)
    diff = q1.cpu().numpy()[:4000][ok].astype(np.int32) - qr[ok].astype(np.int32)
    assert np.abs(diff).max() <= 1 and np.count_nonzero(diff) < 1e-4 * diff.size
