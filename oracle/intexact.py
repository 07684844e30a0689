"""ORACLE (test infrastructure, never imported by the product): CPU restatement in numpy of the arithmetic of the
reference's quantized Transformer ("dialect A") in its *int-exact* factorisation

    (s_x[m] * xq[m,k]) * (s_w[n] * wq[n,k])  ==  s_x[m] * s_w[n] * sum_k xq*wq          

Every function cites the reference lines it follows.  fp32 operations are applied one at a time in the order the
exported graph applies them (numpy float32 arithmetic is IEEE round-to-nearest, no FMA), so integer results
(accumulators, requantized int8 tensors) are defined bit-exactly; float reductions are accumulated in float64
and compared with a tolerance by the tests.

Parity status: the reference's executor (qonnx + onnxruntime) is not installable here and the reference ships no
golden vectors, so this oracle is pinned against the reference's own torch modules
(quant_linear.W8A8Linear, attention.MultiHeadedAttention, layer_norm.LayerNorm ...) imported from /root/reference
by tests/golden/make_golden.py, whose outputs are committed under tests/golden/.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
EPS_CLAMP = F32(1e-5)   # quant_linear.py:10,36  scales.clamp(min=1e-5)
QMAX = F32(127.0)       # quant_linear.py:9,35   2**(8-1)-1
LN_EPS = F32(1e-6)      # layer_norm.py:6


def row_quant(x: np.ndarray):
    """quant_linear.py:31-43 (activations, per token) and :6-17 (weights, per output channel):
    s = clamp(max|x|, 1e-5) / 127 ; q = round(x / s) ; returns (q int8, s fp32[..., 1])."""
    x = np.asarray(x, dtype=F32)
    amax = np.max(np.abs(x), axis=-1, keepdims=True)
    s = (np.maximum(amax, EPS_CLAMP) / QMAX).astype(F32)
    q = np.rint((x / s).astype(F32))
    return q.astype(np.int8), s


def dequant(q: np.ndarray, s: np.ndarray) -> np.ndarray:
    """quant_linear.py:16,42  w.mul(scales)."""
    return (q.astype(F32) * s).astype(F32)


def layer_norm(x: np.ndarray, a_2: np.ndarray, b_2: np.ndarray) -> np.ndarray:
    """layer_norm.py:12-15 in the op order of the exported graph:
    mu = mean(x); d = x - mu; v = mean(d*d) * N / (N-1); y = (a*d) / (sqrt(v) + eps) + b."""
    x = np.asarray(x, dtype=F32)
    n = x.shape[-1]
    mu = (np.sum(x.astype(np.float64), axis=-1, keepdims=True) / n).astype(F32)
    d = (x - mu).astype(F32)
    dd = (d * d).astype(F32)
    v = (np.sum(dd.astype(np.float64), axis=-1, keepdims=True)).astype(F32) / F32(n)
    v = (v * F32(n)).astype(F32) / F32(n - 1)
    den = (np.sqrt(v.astype(F32)) + LN_EPS).astype(F32)
    return (((a_2.astype(F32) * d).astype(F32) / den).astype(F32) + b_2.astype(F32)).astype(F32)


def int_matmul(aq: np.ndarray, wq: np.ndarray) -> np.ndarray:
    """sum_k aq[m,k] * wq[n,k] exactly (|acc| <= 127*127*2048 < 2^31; float64 BLAS is exact below 2^53)."""
    acc = aq.astype(np.float64) @ wq.astype(np.float64).T
    return np.rint(acc).astype(np.int64).astype(np.int32)


def linear_epilogue(acc: np.ndarray, sx, sw, bias=None, relu=False, residual=None) -> np.ndarray:
    """Canonical fp32 epilogue: z = fl(fl(float(acc) * sx[m]) * sw[n]) + bias[n]
    (quant_linear.py:117 F.linear), then Relu (position_feed_forward.py:12) and the residual Add
    (sublayer_connection.py:17)."""
    z = acc.astype(F32)
    if sx is not None:
        z = (z * np.asarray(sx, dtype=F32).reshape(-1, 1)).astype(F32)
    if sw is not None:
        z = (z * np.asarray(sw, dtype=F32).reshape(1, -1)).astype(F32)
    if bias is not None:
        z = (z + np.asarray(bias, dtype=F32).reshape(1, -1)).astype(F32)
    if relu:
        z = np.maximum(z, F32(0))
    if residual is not None:
        z = (np.asarray(residual, dtype=F32) + z).astype(F32)
    return z


def linear_w8a8(xq, sx, wq, sw, bias=None, relu=False, residual=None):
    """quant_linear.py:111-119 W8A8Linear.forward in the int-exact factorisation."""
    return linear_epilogue(int_matmul(xq, wq), sx, sw, bias, relu, residual)


def group_quant(y: np.ndarray, group: int):
    """RowQuant over column groups (Q, K, V outputs share one GEMM; get_quantized_model.py:160-168)."""
    m, n = y.shape
    q, s = row_quant(y.reshape(m, n // group, group))
    return q.reshape(m, n), s.reshape(m, n // group)


def apply_output_fault(value, fault):
    """RANDOM_BITFLIP / RANDOM on one fp32 element (inject_utils/layers.py:18-33)."""
    if fault["type"] == "RANDOM_BITFLIP":
        return float32_bit_flip(value, fault["bit"])
    return bits_to_float32(fault["value_bits"])


def attention(qq, sq, kq, sk, vq, sv, key_mask=None, causal=False, q_pos0=0, return_all=False, fault=None):
    """attention.py:23-36 per sentence.  qq int8 [Tq,512], kq/vq int8 [Tk,512], sq [Tq], sk/sv [Tk].
    scores = fl(fl(float(dot) * sq[i]) * sk[j]) / 8 ; masked_fill(mask==0, -1e9) ; softmax ; pq = rint(127 p) ;
    ctx = sum_j (pq/127) * (sv[j] * vq[j, :])  accumulated in float64 (tolerance class).
    `fault` (optional, one sentence): dict(target "qk"|"pv", type, bit, index=(...), window_start, window_len, value_bits) in
    the integer-domain form of SURVEY.md App. D: an operand fault replaces q by q' = flip(q) for the affected output
    window only (INPUT: one query row x key/feature window; WEIGHT: one key column / feature x query window).
    Returns ctx fp32 [Tq,512] (heads merged, attention.py:65-66) and, if return_all, (ctx, pq uint8 [8,Tq,Tk], p)."""
    Tq, Tk = qq.shape[0], kq.shape[0]
    H, dk = 8, 64
    ctx = np.zeros((Tq, H * dk), dtype=F32)
    pq_all = np.zeros((H, Tq, Tk), dtype=np.uint8)
    p_all = np.zeros((H, Tq, Tk), dtype=np.float64)
    sq = np.asarray(sq, dtype=F32).reshape(Tq, 1)
    sk = np.asarray(sk, dtype=F32).reshape(1, Tk)
    sv = np.asarray(sv, dtype=F32).reshape(Tk, 1)
    visible = np.ones((Tq, Tk), dtype=bool)
    if key_mask is not None:
        visible &= np.asarray(key_mask).astype(bool).reshape(1, Tk)
    if causal:
        visible &= (np.arange(Tk)[None, :] <= (q_pos0 + np.arange(Tq))[:, None])
    ft = fault["type"] if fault else None
    for h in range(H):
        sl = slice(h * dk, (h + 1) * dk)
        dot = int_matmul(qq[:, sl], kq[:, sl])
        if fault and fault["target"] == "qk" and ft.startswith(("INPUT", "WEIGHT")):
            _, t, c = fault["index"]                       # Round tensor [1, T, 512]
            if c // dk == h:
                d = c % dk
                if ft.startswith("INPUT"):
                    q0 = int(qq[t, h * dk + d]); delta = flip_int8_bit(q0, fault["bit"]) - q0
                    w0, w1 = (fault["window_start"], min(Tk, fault["window_start"] + fault["window_len"])) if fault["window_len"] > 0 else (0, Tk)
                    dot[t, w0:w1] += delta * kq[w0:w1, h * dk + d].astype(np.int32)
                else:
                    k0 = int(kq[t, h * dk + d]); delta = flip_int8_bit(k0, fault["bit"]) - k0
                    w0, w1 = (fault["window_start"], min(Tq, fault["window_start"] + fault["window_len"])) if fault["window_len"] > 0 else (0, Tq)
                    dot[w0:w1, t] += qq[w0:w1, h * dk + d].astype(np.int32) * delta
        mm = ((dot.astype(F32) * sq).astype(F32) * sk).astype(F32)
        if fault and fault["target"] == "qk" and ft.startswith("RANDOM") and fault["index"][1] == h:
            _, _, i, j = fault["index"]
            mm[i, j] = apply_output_fault(mm[i, j], fault)
        s = mm / F32(8.0)
        s = np.where(visible, s, F32(-1e9)).astype(F32)
        m = np.max(s, axis=-1, keepdims=True)
        e = np.exp((s - m).astype(F32).astype(np.float64))
        p = e / np.sum(e, axis=-1, keepdims=True)
        pq = np.rint((p.astype(F32) * QMAX).astype(F32))
        phat = (pq.astype(F32) / QMAX).astype(F32)
        vhat = (vq[:, sl].astype(F32) * sv).astype(F32)
        c_h = (phat.astype(np.float64) @ vhat.astype(np.float64))
        if fault and fault["target"] == "pv":
            if ft.startswith("INPUT") and fault["index"][1] == h:      # P tensor [1,8,Tq,Tk]
                _, _, i, j = fault["index"]
                pf = F32(flip_int8_bit(int(pq[i, j]), fault["bit"])) / QMAX
                w0, w1 = (fault["window_start"], min(dk, fault["window_start"] + fault["window_len"])) if fault["window_len"] > 0 else (0, dk)
                c_h[i, w0:w1] += (np.float64(pf) - np.float64(phat[i, j])) * vhat[j, w0:w1].astype(np.float64)
            elif ft.startswith("WEIGHT") and fault["index"][2] // dk == h:   # V Round tensor [1,Tk,512]
                _, j, c = fault["index"]
                d = c % dk
                v0 = int(vq[j, h * dk + d]); vf = flip_int8_bit(v0, fault["bit"])
                w0, w1 = (fault["window_start"], min(Tq, fault["window_start"] + fault["window_len"])) if fault["window_len"] > 0 else (0, Tq)
                dv = np.float64(F32(F32(vf) * sv[j, 0])) - np.float64(vhat[j, d])
                c_h[w0:w1, d] += phat[w0:w1, j].astype(np.float64) * dv
        c_h = c_h.astype(F32)
        if fault and fault["target"] == "pv" and ft.startswith("RANDOM") and fault["index"][1] == h:
            _, _, i, d = fault["index"]
            c_h[i, d] = apply_output_fault(c_h[i, d], fault)
        ctx[:, sl] = c_h
        pq_all[h] = pq.astype(np.uint8)
        p_all[h] = p
    if return_all:
        return ctx, pq_all, p_all
    return ctx


def generator(h, Wg, bg):
    """generator.py:14-15 + torch.max (first index on ties).  Returns (next ids, logits float64)."""
    logits = h.astype(np.float64) @ Wg.astype(np.float64).T + bg.astype(np.float64)
    return np.argmax(logits, axis=-1), logits


def positional_encoding(max_len: int, d_model: int = 512) -> np.ndarray:
    """positional_encodings.py:14-21 (fp32 torch ops)."""
    import math
    position = np.arange(0.0, max_len, dtype=F32).reshape(-1, 1)
    div_term = np.exp((np.arange(0.0, d_model, 2, dtype=F32) * F32(-(math.log(10000.0) / d_model))).astype(F32)).astype(F32)
    pe = np.zeros((max_len, d_model), dtype=F32)
    ang = (position * div_term).astype(F32)
    pe[:, 0::2] = np.sin(ang).astype(F32)
    pe[:, 1::2] = np.cos(ang).astype(F32)
    return pe


def embed(ids, lut, pe, pos0=0):
    """embeddings.py:13 lut(x) * sqrt(d_model) ; positional_encodings.py:24 x + pe[:, :T]."""
    ids = np.asarray(ids)
    d = lut.shape[1]
    T = ids.shape[-1]
    x = (lut[ids].astype(F32) * F32(np.sqrt(F32(d)))).astype(F32)
    return (x + pe[pos0:pos0 + T].astype(F32)).astype(F32)


def flip_int8_bit(value: int, bit: int) -> int:
    """inject_utils/layers.py:61-68 with Python ints (np.int8 ^ 128 raises under NumPy >= 2)."""
    flipped = int(value) ^ (1 << bit)
    if flipped > 127:
        flipped -= 256
    if flipped < -128:
        flipped += 256
    return flipped


def flip_int4_bit(value: int, bit: int) -> int:
    """inject_utils/layers.py:48-59."""
    flipped = int(value) ^ (1 << bit)
    if flipped > 7:
        flipped -= 16
    if flipped < -8:
        flipped += 16
    return flipped


def float32_bit_flip(value, bit: int) -> np.float32:
    """inject_utils/layers.py:24-33 + bin2fp32 :10-16 (NaN -> 0); bit 0 = LSB."""
    bits = np.array([value], dtype=F32).view(np.uint32)
    bits ^= np.uint32(1 << bit)
    out = bits.view(F32)[0]
    return F32(0) if np.isnan(out) else out


def bits_to_float32(pattern: int) -> np.float32:
    """inject_utils/layers.py:18-22 delta_init with the 32 random bits given explicitly."""
    out = np.array([pattern & 0xFFFFFFFF], dtype=np.uint32).view(F32)[0]
    return F32(0) if np.isnan(out) else out
