"""Torch-tensor level wrappers over the C ABI (include/ot_b200.h).

PyTorch is used for device memory and the current CUDA stream only; every computation below is one call into
libot_b200.so.  All functions raise OtError if the library or a CUDA device is missing (no fallback).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import (FAULT_ACC_BITFLIP, FAULT_INPUT, FAULT_NONE, FAULT_OUT_Q8_BITFLIP, FAULT_RANDOM, FAULT_RANDOM_BITFLIP,  # noqa: F401
                   FAULT_WEIGHT, OUT_F32, OUT_I32, OUT_Q8, OtError, OtFault)

OPERAND_Q, OPERAND_K, OPERAND_P, OPERAND_V, OPERAND_SCORES, OPERAND_CTX = range(6)
UNARY = {"Abs": 0, "Relu": 1, "Sqrt": 2, "Round": 3, "Neg": 4, "Exp": 5, "Identity": 6}
BINARY = {"Add": 0, "Sub": 1, "Mul": 2, "Div": 3, "Max": 4, "Min": 5}
CAST_KIND = {torch.float32: 0, torch.int64: 1, torch.uint8: 2, torch.bool: 2, torch.int8: 3, torch.int32: 4}


def _ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise OtError("expected a CUDA tensor (this package has no CPU path)")
    return C.c_void_p(t.data_ptr())


try:      # the raw stream handle of the current device: two C calls instead of torch.cuda.current_stream()'s Python stack (17 us per launch)
    _raw_stream, _cur_device = torch._C._cuda_getCurrentRawStream, torch._C._cuda_getDevice
except AttributeError:    # pragma: no cover  (older / newer torch without these entry points)
    _raw_stream = _cur_device = None


def _stream():
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(_cur_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _broadcast(a, b):
    """numpy-style broadcast of two shapes (torch.broadcast_shapes goes through the symbolic-shape machinery: 35 us per call)."""
    ra, rb = len(a), len(b)
    out = []
    for i in range(max(ra, rb)):
        da = a[ra - 1 - i] if i < ra else 1
        db = b[rb - 1 - i] if i < rb else 1
        if da != db and da != 1 and db != 1:
            raise OtError("shapes %s and %s do not broadcast" % (tuple(a), tuple(b)))
        out.append(db if da == 1 else da)
    return tuple(reversed(out))


def _req(t: torch.Tensor, dtype, name: str):
    if t.dtype != dtype:
        raise OtError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_cuda:
        raise OtError("%s must be a CUDA tensor" % name)


def make_fault(mode=FAULT_NONE, flat_index=0, bit=0, window_start=0, window_len=0, value_bits=0, operand=0) -> OtFault:
    return OtFault(int(mode), int(bit), int(flat_index), int(window_start), int(window_len), int(value_bits) & 0xFFFFFFFF,
                   int(operand))


def _fault_ref(fault: Optional[OtFault]):
    return C.byref(fault) if fault is not None else None


_FAULT_DTYPE = None


def faults_to_numpy(faults):
    """A list of OtFault -> numpy uint8 array holding the C array (32 bytes per entry)."""
    import numpy as np
    global _FAULT_DTYPE
    if _FAULT_DTYPE is None:
        _FAULT_DTYPE = np.dtype([("mode", "<i4"), ("bit", "<i4"), ("flat_index", "<i8"), ("window_start", "<i4"), ("window_len", "<i4"),
                                 ("value_bits", "<u4"), ("reserved", "<i4")])
        assert _FAULT_DTYPE.itemsize == C.sizeof(OtFault)
    arr = np.zeros(len(faults), dtype=_FAULT_DTYPE)
    for i, f in enumerate(faults):
        arr[i] = (f.mode, f.bit, f.flat_index, f.window_start, f.window_len, f.value_bits, f.reserved)
    return arr.view(np.uint8)


def pack_faults(faults, device) -> torch.Tensor:
    """A list of OtFault -> device byte tensor holding the C array (32 bytes per entry)."""
    return torch.from_numpy(faults_to_numpy(faults)).to(device)


# ------------------------------------------------------------------------------------------------ GEMM
def linear_w8a8(a_q: torch.Tensor, w_q: torch.Tensor, *, row_scale=None, col_scale=None, bias=None, residual=None,
                relu=False, out_kind=OUT_F32, quant_group=0, fault: Optional[OtFault] = None, out=None, out_scale=None,
                w4=False, mf=None):
    """a_q int8 [M,K] (row stride multiple of 16), w_q int8 [N,K] (or uint8 [N,K/2] packed int4 when w4).
    Returns out (int32 / fp32 / int8 [M,N]) and, for OUT_Q8, (out, out_scale [M, N/quant_group])."""
    lib = _lib.load()
    _req(a_q, torch.int8, "a_q")
    M, K = a_q.shape
    N = w_q.shape[0]
    if w4:
        _req(w_q, torch.uint8, "w_q (packed int4)")
        assert w_q.shape[1] * 2 == K
    else:
        _req(w_q, torch.int8, "w_q")
        assert w_q.shape[1] == K
    assert a_q.stride(1) == 1 and w_q.stride(1) == 1
    dev = a_q.device
    if out is None:
        dt = {OUT_I32: torch.int32, OUT_F32: torch.float32, OUT_Q8: torch.int8}[out_kind]
        out = torch.empty((M, N), dtype=dt, device=dev)
    if out_kind == OUT_Q8 and out_scale is None:
        out_scale = torch.empty((M, N // quant_group), dtype=torch.float32, device=dev)
    if mf is not None:
        faults_dev, unit_dev, rows_per_unit = mf
        assert fault is None
        fn_mf = lib.ot_linear_w4a8_mf if w4 else lib.ot_linear_w8a8_mf
        rc = fn_mf(_ptr(a_q), a_q.stride(0), _ptr(w_q), w_q.stride(0), M, N, K, _ptr(row_scale), _ptr(col_scale), _ptr(bias),
                                   _ptr(residual), residual.stride(0) if residual is not None else 0, 1 if relu else 0, out_kind, _ptr(out),
                                   out.stride(0), _ptr(out_scale), int(quant_group), _ptr(faults_dev), _ptr(unit_dev), int(rows_per_unit), _stream())
        _lib.check(rc, "ot_linear_w8a8_mf")
        return (out, out_scale) if out_kind == OUT_Q8 else out
    fn = lib.ot_linear_w4a8 if w4 else lib.ot_linear_w8a8
    rc = fn(_ptr(a_q), a_q.stride(0), _ptr(w_q), w_q.stride(0), M, N, K,
            _ptr(row_scale), _ptr(col_scale), _ptr(bias), _ptr(residual), residual.stride(0) if residual is not None else 0,
            1 if relu else 0, out_kind, _ptr(out), out.stride(0), _ptr(out_scale), int(quant_group), _fault_ref(fault), _stream())
    _lib.check(rc, "ot_linear_w4a8" if w4 else "ot_linear_w8a8")
    return (out, out_scale) if out_kind == OUT_Q8 else out


def rowsum_i8(x: torch.Tensor) -> torch.Tensor:
    """sum over the last axis of an int8 matrix -> int32 [rows] (operand sums of the zero-point correction)."""
    _req(x, torch.int8, "x")
    assert x.dim() == 2 and x.stride(1) == 1
    out = torch.empty((x.shape[0],), dtype=torch.int32, device=x.device)
    _lib.check(_lib.load().ot_rowsum_i8(_ptr(x), x.stride(0), x.shape[0], x.shape[1], _ptr(out), _stream()), "ot_rowsum_i8")
    return out


def _zp_vec(zp, n: int, device) -> Optional[torch.Tensor]:
    """Zero point (None, scalar or [n] tensor of any integer type) -> int32 [n] on the device, or None when identically zero is
    KNOWN without a device read (None)."""
    if zp is None:
        return None
    z = zp.to(device=device, dtype=torch.int32).reshape(-1)
    return z.expand(n).contiguous() if z.numel() == 1 else z.contiguous()


def matmul_integer(a_q: torch.Tensor, w_q: torch.Tensor, a_zp=None, b_zp=None, fault: Optional[OtFault] = None, out=None):
    """ONNX MatMulInteger on K-major operands: a_q int8 [M,K], w_q int8 [N,K] (= B transposed); a_zp per row / scalar, b_zp per
    column / scalar (int tensors, None = 0).  The zero-point correction runs in the GEMM epilogue (ot_matmul_integer)."""
    lib = _lib.load()
    _req(a_q, torch.int8, "a_q")
    _req(w_q, torch.int8, "w_q")
    M, K = a_q.shape
    N = w_q.shape[0]
    assert w_q.shape[1] == K and a_q.stride(1) == 1 and w_q.stride(1) == 1
    az, bz = _zp_vec(a_zp, M, a_q.device), _zp_vec(b_zp, N, a_q.device)
    rs = rowsum_i8(a_q) if bz is not None else None
    cs = rowsum_i8(w_q) if az is not None else None
    if out is None:
        out = torch.empty((M, N), dtype=torch.int32, device=a_q.device)
    rc = lib.ot_matmul_integer(_ptr(a_q), a_q.stride(0), _ptr(w_q), w_q.stride(0), M, N, K, _ptr(az), _ptr(bz), _ptr(rs), _ptr(cs), _ptr(out),
                               out.stride(0), _fault_ref(fault), _stream())
    _lib.check(rc, "ot_matmul_integer")
    return out


def qlinear_matmul(a_q, a_scale, a_zp, w_q, b_scale, b_zp, y_scale: float, y_zp: int, out=None):
    """ONNX QLinearMatMul on K-major operands (w_q = B transposed, [N,K]); a_scale fp32 scalar / [M], b_scale scalar / [N]."""
    lib = _lib.load()
    _req(a_q, torch.int8, "a_q")
    _req(w_q, torch.int8, "w_q")
    M, K = a_q.shape
    N = w_q.shape[0]
    dev = a_q.device
    sa = a_scale.to(device=dev, dtype=torch.float32).reshape(-1)
    sa = sa.expand(M).contiguous() if sa.numel() == 1 else sa.contiguous()
    sb = b_scale.to(device=dev, dtype=torch.float32).reshape(-1)
    sb = sb.expand(N).contiguous() if sb.numel() == 1 else sb.contiguous()
    az, bz = _zp_vec(a_zp, M, dev), _zp_vec(b_zp, N, dev)
    rs = rowsum_i8(a_q) if bz is not None else None
    cs = rowsum_i8(w_q) if az is not None else None
    if out is None:
        out = torch.empty((M, N), dtype=torch.int8, device=dev)
    rc = lib.ot_qlinear_matmul(_ptr(a_q), a_q.stride(0), _ptr(w_q), w_q.stride(0), M, N, K, _ptr(sa), _ptr(sb), _ptr(az), _ptr(bz), _ptr(rs), _ptr(cs),
                               float(y_scale), int(y_zp), _ptr(out), out.stride(0), _stream())
    _lib.check(rc, "ot_qlinear_matmul")
    return out


def ln_linear_w8a8(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, w_q: torch.Tensor, *, eps: float = 1e-6, col_scale=None,
                   bias=None, residual=None, relu=False, out_kind=OUT_F32, quant_group=0, out=None, out_scale=None):
    """LayerNorm + RowQuant prologue fused into the int8 GEMM (x fp32 [M,512])."""
    lib = _lib.load()
    _req(x, torch.float32, "x")
    _req(w_q, torch.int8, "w_q")
    M, K = x.shape
    N = w_q.shape[0]
    if out is None:
        dt = {OUT_I32: torch.int32, OUT_F32: torch.float32, OUT_Q8: torch.int8}[out_kind]
        out = torch.empty((M, N), dtype=dt, device=x.device)
    if out_kind == OUT_Q8 and out_scale is None:
        out_scale = torch.empty((M, N // quant_group), dtype=torch.float32, device=x.device)
    rc = lib.ot_ln_linear_w8a8(_ptr(x), x.stride(0), _ptr(gamma), _ptr(beta), eps, _ptr(w_q), w_q.stride(0), M, N, K, _ptr(col_scale),
                               _ptr(bias), _ptr(residual), residual.stride(0) if residual is not None else 0, 1 if relu else 0, out_kind,
                               _ptr(out), out.stride(0), _ptr(out_scale), int(quant_group), _stream())
    _lib.check(rc, "ot_ln_linear_w8a8")
    return (out, out_scale) if out_kind == OUT_Q8 else out


def unpack_int4(w4: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req(w4, torch.uint8, "w4")
    rows, half = w4.shape
    if out is None:
        out = torch.empty((rows, half * 2), dtype=torch.int8, device=w4.device)
    else:
        _req(out, torch.int8, "out")
        assert out.shape == (rows, half * 2)
    _lib.check(_lib.load().ot_unpack_int4(_ptr(w4), _ptr(out), rows, half * 2, _stream()), "ot_unpack_int4")
    return out


def pack_int4(w8: torch.Tensor) -> torch.Tensor:
    _req(w8, torch.int8, "w8")
    rows, cols = w8.shape
    out = torch.empty((rows, cols // 2), dtype=torch.uint8, device=w8.device)
    _lib.check(_lib.load().ot_pack_int4(_ptr(w8.contiguous()), _ptr(out), rows, cols, _stream()), "ot_pack_int4")
    return out


# ------------------------------------------------------------------------------------------------ row ops
def layernorm_quant(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-6, want_y=False, want_q=True,
                    y=None, q=None, s=None):
    _req(x, torch.float32, "x")
    n = x.shape[-1]
    rows = x.numel() // n
    assert x.is_contiguous()
    if want_y and y is None:
        y = torch.empty_like(x)
    if want_q and q is None:
        q = torch.empty(x.shape, dtype=torch.int8, device=x.device)
        s = torch.empty(x.shape[:-1], dtype=torch.float32, device=x.device)
    rc = _lib.load().ot_layernorm_quant(_ptr(x), _ptr(gamma), _ptr(beta), rows, n, eps, _ptr(y), _ptr(q), _ptr(s), _stream())
    _lib.check(rc, "ot_layernorm_quant")
    return y, q, s


def rowquant(x: torch.Tensor, group: Optional[int] = None, want_xhat=False, q=None, s=None):
    """Per-row abs-max quantization of fp32 [..., n] over groups of `group` columns (default: the whole row)."""
    _req(x, torch.float32, "x")
    n = x.shape[-1]
    group = group or n
    rows = x.numel() // n
    assert x.stride(-1) == 1
    x2 = x.reshape(rows, n) if x.is_contiguous() else x
    ldx = x2.stride(0) if x2.dim() == 2 else n
    if q is None:
        q = torch.empty(x.shape, dtype=torch.int8, device=x.device)
    if s is None:
        s = torch.empty(tuple(x.shape[:-1]) + ((n // group,) if group != n else ()), dtype=torch.float32, device=x.device)
    xhat = torch.empty(x.shape, dtype=torch.float32, device=x.device) if want_xhat else None
    rc = _lib.load().ot_rowquant(_ptr(x2), ldx, rows, n, group, _ptr(q), _ptr(s), _ptr(xhat), _stream())
    _lib.check(rc, "ot_rowquant")
    return (q, s, xhat) if want_xhat else (q, s)


def residual_add(a: torch.Tensor, b: torch.Tensor, out=None):
    _req(a, torch.float32, "a")
    _req(b, torch.float32, "b")
    assert a.shape == b.shape and a.is_contiguous() and b.is_contiguous()
    if out is None:
        out = torch.empty_like(a)
    _lib.check(_lib.load().ot_residual_add(_ptr(a), _ptr(b), _ptr(out), a.numel(), _stream()), "ot_residual_add")
    return out


def embed_pe(ids: torch.Tensor, table: torch.Tensor, pe: torch.Tensor, *, seq_len: int, pos0: int = 0, pos_dev=None,
             ids_stride: int = 1, rows: Optional[int] = None, scale: Optional[float] = None, out=None):
    _req(ids, torch.int64, "ids")
    d = table.shape[1]
    rows = ids.numel() if rows is None else rows
    if scale is None:
        scale = float(d) ** 0.5
    if out is None:
        out = torch.empty((rows, d), dtype=torch.float32, device=table.device)
    rc = _lib.load().ot_embed_pe(_ptr(ids), ids_stride, _ptr(table), _ptr(pe), rows, seq_len, d, pos0, _ptr(pos_dev), scale, _ptr(out),
                                 _stream())
    _lib.check(rc, "ot_embed_pe")
    return out


# ------------------------------------------------------------------------------------------------ attention
def attention_q8(q, sq, k, v, sk, sv, *, B, Tq, Tk, Tk_cap=None, ldq=None, sq_stride=1, ldk=None, skv_stride=1,
                 k_new=None, v_new=None, sk_new=None, sv_new=None, ld_new=0, snew_stride=1,
                 mask_kind=0, key_mask=None, mask_stride=0, q_pos0=0, step_dev=None,
                 want_ctx=True, want_q=False, want_probs=False, fault: Optional[OtFault] = None,
                 ctx=None, ctx_q=None, ctx_s=None, mf=None):
    """Pointer-level wrapper; q/k/v may be views into fused QKV buffers (give ldq/ldk/strides explicitly)."""
    Tk_cap = Tk_cap or Tk
    ldq = ldq if ldq is not None else q.stride(-2)
    ldk = ldk if ldk is not None else k.stride(-2)
    dev = q.device
    if want_ctx and ctx is None:
        ctx = torch.empty((B * Tq, 512), dtype=torch.float32, device=dev)
    if not want_ctx:
        ctx = None
    if want_q and ctx_q is None:
        ctx_q = torch.empty((B * Tq, 512), dtype=torch.int8, device=dev)
        ctx_s = torch.empty((B * Tq,), dtype=torch.float32, device=dev)
    probs = torch.empty((B, 8, Tq, Tk), dtype=torch.uint8, device=dev) if want_probs else None
    if key_mask is not None and mask_stride == 0:
        mask_stride = key_mask.stride(0)
    rc = _lib.load().ot_attention_q8_mf(
        _ptr(q), ldq, _ptr(sq), sq_stride, _ptr(k), _ptr(v), ldk, _ptr(sk), _ptr(sv), skv_stride,
        _ptr(k_new), _ptr(v_new), ld_new, _ptr(sk_new), _ptr(sv_new), snew_stride,
        B, 8, Tq, Tk, Tk_cap, mask_kind, _ptr(key_mask), mask_stride, q_pos0, _ptr(step_dev),
        _ptr(ctx), ctx.stride(0) if ctx is not None else 0, _ptr(ctx_q), _ptr(ctx_s), _ptr(probs), _fault_ref(fault),
        _ptr(mf[0]) if mf is not None else None, _ptr(mf[1]) if mf is not None else None, _stream())
    _lib.check(rc, "ot_attention_q8")
    return ctx, ctx_q, ctx_s, probs


# ------------------------------------------------------------------------------------------------ generator / loop glue
def generator_argmax(h: torch.Tensor, Wg: torch.Tensor, bg: torch.Tensor, *, want_logp=False, want_margin=False,
                     next_ids=None, scratch=None, margin=None):
    _req(h, torch.float32, "h")
    rows, d = h.shape
    vocab = Wg.shape[0]
    dev = h.device
    if next_ids is None:
        next_ids = torch.empty((rows,), dtype=torch.int64, device=dev)
    if scratch is None:
        scratch = torch.empty((rows, vocab), dtype=torch.float32, device=dev)
    logp = torch.empty((rows, vocab), dtype=torch.float32, device=dev) if want_logp else None
    if want_margin and margin is None:
        margin = torch.empty((rows,), dtype=torch.float32, device=dev)
    rc = _lib.load().ot_generator_argmax(_ptr(h), h.stride(0), _ptr(Wg), _ptr(bg), rows, d, vocab, _ptr(next_ids), _ptr(scratch),
                                         _ptr(logp), _ptr(margin), _stream())
    _lib.check(rc, "ot_generator_argmax")
    return next_ids, logp, margin, scratch


def append_token(ys: torch.Tensor, next_ids: torch.Tensor, step_dev: torch.Tensor):
    _req(ys, torch.int64, "ys")
    _req(step_dev, torch.int32, "step_dev")
    rc = _lib.load().ot_append_token(_ptr(ys), ys.stride(0), _ptr(next_ids), ys.shape[0], _ptr(step_dev), _stream())
    _lib.check(rc, "ot_append_token")


# ------------------------------------------------------------------------------------------------ persistent decoder
class DecoderPlan:
    """Device-resident plan of the persistent greedy decoder (ot_decoder_plan_build).  Keeps every tensor it points to alive."""

    def __init__(self, layers, ws_tensors, *, n_layers: int, B: int, S: int, cap: int, vocab: int, ys: torch.Tensor, trace: bool = False):
        lib = _lib.load()
        dev = ys.device
        self.keep = (layers, ws_tensors)
        self.buf = torch.zeros(lib.ot_decoder_plan_size() + 256, dtype=torch.uint8, device=dev)
        self.bar = torch.zeros(1, dtype=torch.int32, device=dev)
        self.trace = torch.zeros(256, dtype=torch.int64, device=dev) if trace else None
        flat_l = [t for layer in layers for t in layer]
        assert len(flat_l) == 28 * n_layers
        ws_all = list(ws_tensors) + [ys, self.bar, self.trace]
        assert len(ws_all) == 24
        for t in flat_l + ws_all[:-1]:
            if t is not None and not t.is_contiguous() and t.dim() > 1 and t.stride(-1) != 1:
                raise OtError("decoder plan tensors must be row-major")
        lp = (C.c_void_p * len(flat_l))(*[t.data_ptr() for t in flat_l])
        wp = (C.c_void_p * 24)(*[(t.data_ptr() if t is not None else None) for t in ws_all])
        base = (self.buf.data_ptr() + 255) & ~255
        self.ptr = C.c_void_p(base)
        rc = lib.ot_decoder_plan_build(self.ptr, n_layers, B, S, cap, vocab, ys.stride(0), lp, wp)
        _lib.check(rc, "ot_decoder_plan_build")

    def run(self, t0: int, n_steps: int):
        rc = _lib.load().ot_decoder_run(self.ptr, _ptr(self.bar), int(t0), int(n_steps), _stream())
        _lib.check(rc, "ot_decoder_run")


def cdecoder_max_sentences() -> int:
    """Sentences the cluster-resident decoder advances in ONE wave of thread-block clusters on the current device: 8 per co-resident
    cluster (ot_cdecoder_max_clusters: 15 clusters = 120 sentences on a B200).  The step time does not depend on how many are in use."""
    n = C.c_int(0)
    _lib.check(_lib.load().ot_cdecoder_max_clusters(C.byref(n)), "ot_cdecoder_max_clusters")
    return 8 * max(1, int(n.value))


class ClusterDecoderPlan:
    """Device-resident plan of the cluster-resident greedy decoder (ot_cdecoder_plan_build): groups of `spc` sentences, one
    8-CTA cluster each.  Keeps every tensor it points to alive."""

    PHASES = ["qkv", "self_attn", "o", "ln2", "cq", "cross_attn", "co", "ln3", "ffn1", "ffn2", "ln1"]      # ffn1 includes the RowQuant of the hidden rows

    def __init__(self, layers, ws_tensors, *, n_layers: int, B: int, S: int, cap: int, vocab: int, ys: torch.Tensor, spc: int = 8,
                 trace: bool = False):
        lib = _lib.load()
        dev = ys.device
        self.keep = (layers, ws_tensors)
        self.B, self.spc, self.n_layers = B, spc, n_layers
        self.buf = torch.zeros(lib.ot_cdecoder_plan_size() + 256, dtype=torch.uint8, device=dev)
        self.trace = torch.zeros(256, dtype=torch.int64, device=dev) if trace else None
        flat_l = [t for layer in layers for t in layer]
        assert len(flat_l) == 28 * n_layers
        ws_all = list(ws_tensors[:9]) + [ys, self.trace, ws_tensors[9] if len(ws_tensors) > 9 else None]
        assert len(ws_all) == 12
        for t in flat_l + ws_all:
            if t is not None and not t.is_contiguous() and t.dim() > 1 and t.stride(-1) != 1:
                raise OtError("decoder plan tensors must be row-major")
        lp = (C.c_void_p * len(flat_l))(*[t.data_ptr() for t in flat_l])
        wp = (C.c_void_p * 12)(*[(t.data_ptr() if t is not None else None) for t in ws_all])
        base = (self.buf.data_ptr() + 255) & ~255
        self.ptr = C.c_void_p(base)
        rc = lib.ot_cdecoder_plan_build(self.ptr, n_layers, B, S, cap, vocab, spc, ys.stride(0), lp, wp)
        _lib.check(rc, "ot_cdecoder_plan_build")

    def run(self, t0: int, n_steps: int):
        rc = _lib.load().ot_cdecoder_run(self.ptr, self.B, self.spc, int(t0), int(n_steps), _stream())
        _lib.check(rc, "ot_cdecoder_run")

    def phase_names(self):
        """Names of the barrier.cluster intervals of one greedy step, in trace order."""
        names = ["ln1"]                     # embedding + positional encoding + LayerNorm 1 of layer 0
        for _ in range(self.n_layers):
            names += self.PHASES          # ... the last one is LayerNorm 1 of the NEXT layer
        names[-1] = "final_norm"
        return names + ["generator"]


# ------------------------------------------------------------------------------------------------ elementwise family
def _shape4(shape):
    shape = tuple(int(x) for x in shape)
    if len(shape) > 4:
        raise OtError("rank > 4 is not supported: %s" % (shape,))
    arr = (1,) * (4 - len(shape)) + shape
    return (C.c_int64 * 4)(*arr)


def unary(op: str, x: torch.Tensor, out=None):
    _req(x, torch.float32, "x")
    x = x.contiguous()
    if out is None:
        out = torch.empty_like(x)
    _lib.check(_lib.load().ot_unary_f32(UNARY[op], _ptr(x), _ptr(out), x.numel(), _stream()), "ot_unary_f32(%s)" % op)
    return out


def binary(op: str, a: torch.Tensor, b: torch.Tensor):
    _req(a, torch.float32, "a")
    _req(b, torch.float32, "b")
    a = a.contiguous()
    b = b.contiguous()
    out_shape = _broadcast(tuple(a.shape), tuple(b.shape))
    out = torch.empty(out_shape, dtype=torch.float32, device=a.device)
    rank = len(out_shape)
    ash = (1,) * (rank - a.dim()) + tuple(a.shape)
    bsh = (1,) * (rank - b.dim()) + tuple(b.shape)
    rc = _lib.load().ot_binary_f32(BINARY[op], _ptr(a), _shape4(ash), _ptr(b), _shape4(bsh), _ptr(out), _shape4(out_shape), _stream())
    _lib.check(rc, "ot_binary_f32(%s)" % op)
    return out


def clip(x: torch.Tensor, lo: float, hi: float):
    _req(x, torch.float32, "x")
    x = x.contiguous()
    out = torch.empty_like(x)
    _lib.check(_lib.load().ot_clip_f32(_ptr(x), lo, hi, _ptr(out), x.numel(), _stream()), "ot_clip_f32")
    return out


def reduce_last(op: str, x: torch.Tensor, keepdims=True):
    _req(x, torch.float32, "x")
    x = x.contiguous()
    n = x.shape[-1]
    rows = x.numel() // n
    out = torch.empty(tuple(x.shape[:-1]) + ((1,) if keepdims else ()), dtype=torch.float32, device=x.device)
    code = {"ReduceMax": 0, "ReduceMean": 1}[op]
    _lib.check(_lib.load().ot_reduce_last_f32(code, _ptr(x), rows, n, _ptr(out), _stream()), "ot_reduce_last_f32")
    return out


def softmax_last(x: torch.Tensor):
    _req(x, torch.float32, "x")
    x = x.contiguous()
    n = x.shape[-1]
    out = torch.empty_like(x)
    _lib.check(_lib.load().ot_softmax_f32(_ptr(x), x.numel() // n, n, _ptr(out), _stream()), "ot_softmax_f32")
    return out


def where_scalar(cond: torch.Tensor, a_scalar: float, x: torch.Tensor):
    _req(x, torch.float32, "x")
    cond = cond.contiguous()
    if cond.dtype == torch.bool:
        cond = cond.view(torch.uint8)
    x = x.contiguous()
    out_shape = _broadcast(tuple(cond.shape), tuple(x.shape))
    if tuple(out_shape) != tuple(x.shape):
        raise OtError("Where: x must already have the broadcast shape")
    out = torch.empty_like(x)
    csh = (1,) * (x.dim() - cond.dim()) + tuple(cond.shape)
    rc = _lib.load().ot_where_f32(_ptr(cond), _shape4(csh), float(a_scalar), _ptr(x), _ptr(out), _shape4(x.shape), _stream())
    _lib.check(rc, "ot_where_f32")
    return out


def equal_scalar_i64(x: torch.Tensor, scalar: int):
    _req(x, torch.int64, "x")
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=torch.bool, device=x.device)
    _lib.check(_lib.load().ot_equal_i64(_ptr(x), int(scalar), _ptr(out), x.numel(), _stream()), "ot_equal_i64")
    return out


def cast(x: torch.Tensor, dtype: torch.dtype, numeric_u8: bool = False):
    """ONNX Cast.  torch.uint8 as a destination means `bool` (0/1) unless numeric_u8 (QuantizeLinear's uint8 tensors)."""
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=dtype, device=x.device)
    dk = 5 if (numeric_u8 and dtype == torch.uint8) else CAST_KIND[dtype]
    rc = _lib.load().ot_cast(CAST_KIND[x.dtype], _ptr(x), dk, _ptr(out), x.numel(), _stream())
    _lib.check(rc, "ot_cast")
    return out


def transpose(x: torch.Tensor, perm):
    if x.element_size() != 4:
        raise OtError("transpose supports 4-byte element types")
    x = x.contiguous()
    rank = x.dim()
    pad = 4 - rank
    shape4 = (1,) * pad + tuple(x.shape)
    perm4 = tuple(range(pad)) + tuple(p + pad for p in perm)
    out_shape = tuple(x.shape[p] for p in perm)
    out = torch.empty(out_shape, dtype=x.dtype, device=x.device)
    rc = _lib.load().ot_transpose4_b32(_ptr(x), _shape4(shape4), (C.c_int * 4)(*perm4), _ptr(out), _stream())
    _lib.check(rc, "ot_transpose4_b32")
    return out


def matmul_f32(a: torch.Tensor, b: torch.Tensor):
    """Batched fp32 MatMul with numpy-style broadcasting of the leading (batch) dims (batch of b may be 1)."""
    _req(a, torch.float32, "a")
    _req(b, torch.float32, "b")
    a = a.contiguous()
    b = b.contiguous()
    M, K = a.shape[-2], a.shape[-1]
    K2, N = b.shape[-2], b.shape[-1]
    assert K == K2
    batch_shape = _broadcast(tuple(a.shape[:-2]), tuple(b.shape[:-2]))
    batch = 1
    for s in batch_shape:
        batch *= s
    a_b = a.numel() // (M * K)
    b_b = b.numel() // (K * N)
    if a_b not in (1, batch) or b_b not in (1, batch):
        a = a.expand(batch_shape + (M, K)).contiguous()
        b = b.expand(batch_shape + (K, N)).contiguous()
        a_b = b_b = batch
    out = torch.empty(tuple(batch_shape) + (M, N), dtype=torch.float32, device=a.device)
    rc = _lib.load().ot_matmul_f32(_ptr(a), _ptr(b), _ptr(out), batch, M, N, K, M * K if a_b == batch and batch > 1 else (0 if a_b == 1 and batch > 1 else M * K),
                                   K * N if b_b == batch and batch > 1 else (0 if b_b == 1 and batch > 1 else K * N), M * N, _stream())
    _lib.check(rc, "ot_matmul_f32")
    return out
