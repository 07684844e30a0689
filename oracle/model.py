"""ORACLE (test infrastructure, never imported by the product): numpy restatement of the reference's quantized
Transformer forward pass and greedy loop, layer by layer.

Follows: get_quantized_model.py:10-36,46-148 (SmoothQuant pre-pass), :150-172 (which linears quantize their
output), quant_linear.py:111-119 (W8A8Linear.forward), attention.py:23-67, layer_norm.py:12-15,
sublayer_connection.py:15-17, encoder.py:14-32, decoder.py:13-33, position_feed_forward.py:11-12,
encoder_decoder.py:54-58, generator.py:14-15, parallelized_inject_onnx_transformer.py:616,632-758 (greedy loop),
batch_output.py:659-672 (batched greedy).

Two numeric modes:
  "ref-float": fp32 MatMul of the de-quantized operands, literally what the exported graph computes;
  "int-exact": exact int32 contraction + canonical fp32 epilogue (oracle/intexact.py), the factorisation the
               CUDA kernels implement.  Integer tensors agree between the modes except for ~1e-5 of elements
               at rounding boundaries.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from . import intexact as ox

F32 = np.float32
D_MODEL, N_HEADS, D_K = 512, 8, 64


# ------------------------------------------------------------------------------------------------ model preparation
def smooth_ln_fcs(w: Dict[str, np.ndarray], ln: str, fcs, act_scales: np.ndarray, alpha: float = 0.5) -> None:
    """get_quantized_model.py:10-36 (in place)."""
    weight_scales = np.stack([np.abs(w[fc + ".weight"]).max(axis=0) for fc in fcs], axis=0).max(axis=0)
    weight_scales = np.maximum(weight_scales, F32(1e-5)).astype(F32)
    scales = (np.power(act_scales.astype(F32), F32(alpha)) / np.power(weight_scales, F32(1 - alpha))).astype(F32)
    scales = np.maximum(scales, F32(1e-5)).astype(F32)
    w[ln + ".a_2"] = (w[ln + ".a_2"] / scales).astype(F32)
    w[ln + ".b_2"] = (w[ln + ".b_2"] / scales).astype(F32)
    for fc in fcs:
        w[fc + ".weight"] = (w[fc + ".weight"] * scales.reshape(1, -1)).astype(F32)


def smooth_lm(w: Dict[str, np.ndarray], scales: Dict[str, np.ndarray], n_layers: int = 6, alpha: float = 0.5) -> None:
    """get_quantized_model.py:46-148: LN -> {q,k,v} and LN -> w_1 smoothing; note that the decoder's src_attn k/v
    linears (which consume `memory`, not the LN output) are scaled as well (:123-131) -- restated as is."""
    for l in range(n_layers):
        p = "encoder.layers.%d" % l
        smooth_ln_fcs(w, p + ".sublayer.0.norm", [p + ".self_attn.linears.%d" % i for i in range(3)],
                      scales[p + ".self_attn.linears.0"], alpha)
        smooth_ln_fcs(w, p + ".sublayer.1.norm", [p + ".feed_forward.w_1"], scales[p + ".feed_forward.w_1"], alpha)
    for l in range(n_layers):
        p = "decoder.layers.%d" % l
        smooth_ln_fcs(w, p + ".sublayer.0.norm", [p + ".self_attn.linears.%d" % i for i in range(3)],
                      scales[p + ".self_attn.linears.0"], alpha)
        smooth_ln_fcs(w, p + ".sublayer.1.norm", [p + ".src_attn.linears.%d" % i for i in range(3)],
                      scales[p + ".src_attn.linears.0"], alpha)
        smooth_ln_fcs(w, p + ".sublayer.2.norm", [p + ".feed_forward.w_1"], scales[p + ".feed_forward.w_1"], alpha)


def fake_quantize_weights(w: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """quantize_transformer -> W8A8Linear.from_float (get_quantized_model.py:150-172, quant_linear.py:122-147):
    every attention / FFN linear weight is replaced by round(W/s)*s, s per output channel."""
    out = dict(w)
    for k, v in w.items():
        if k.endswith(".weight") and v.ndim == 2 and (".linears." in k or ".feed_forward." in k):
            q, s = ox.row_quant(v)
            out[k] = ox.dequant(q, s)
    return out


def quantize_weights_int4(w: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Config #4 (dialect B style, SURVEY.md App. C -- Brevitas semantics from its documentation, parity unpinned): every
    attention / FFN linear weight on the signed 4-bit grid [-8, 7] with per-output-channel scale amax/8.  The result is
    stored de-quantized, with "<name>.int4_scale" beside it so the executor can recover (q, s) exactly."""
    out = dict(w)
    for k, v in w.items():
        if k.endswith(".weight") and v.ndim == 2 and (".linears." in k or ".feed_forward." in k):
            amax = np.max(np.abs(v), axis=-1, keepdims=True).astype(F32)
            s = (np.maximum(amax, F32(1e-5)) / F32(8.0)).astype(F32)
            q = np.clip(np.rint((v / s).astype(F32)), -8, 7).astype(np.int8)
            out[k] = q                       # int8 values in [-8,7]
            out[k[:-7] + ".int4_scale"] = s
    return out


def get_quantized(float_weights: Dict[str, np.ndarray], scales: Optional[Dict[str, np.ndarray]] = None, n_layers: int = 6):
    """get_quantized_model.py:174-178: smooth_lm (when scales are given) then quantize_transformer."""
    w = {k: np.array(v, dtype=F32, copy=True) for k, v in float_weights.items()}
    if scales is not None:
        smooth_lm(w, scales, n_layers)
    return fake_quantize_weights(w)


# ------------------------------------------------------------------------------------------------ building blocks
class Trace(dict):
    """Optional capture of named intermediates (integer Round tensors, MatMul outputs) for parity tests."""


def _linear(w, prefix: str, x: np.ndarray, mode: str, relu=False, residual=None, quantize_output=False, cap=None, cap_name=None,
            xq_sx=None, fault=None):
    """W8A8Linear.forward (quant_linear.py:111-119): act RowQuant, weight RowQuant (recomputed), F.linear, [output quant]."""
    shape = x.shape
    x2 = x.reshape(-1, shape[-1])
    xq, sx = xq_sx if xq_sx is not None else ox.row_quant(x2)
    if prefix + ".int4_scale" in w:          # config #4: static 4-bit weights
        wq, sw = w[prefix + ".weight"], w[prefix + ".int4_scale"]
    else:
        wq, sw = ox.row_quant(w[prefix + ".weight"])
    bias = w[prefix + ".bias"]
    if mode == "int-exact":
        acc = ox.int_matmul(xq, wq)
        if fault is not None and fault["type"].startswith(("INPUT", "WEIGHT")):
            # integer-domain rank-1 update of SURVEY.md App. D
            K_ = xq.shape[1]
            r, k = divmod(int(fault["flat_index"]), K_)
            if fault["type"].startswith("INPUT"):
                q0 = int(xq[r, k]); delta = ox.flip_int8_bit(q0, fault["bit"]) - q0
                acc[r, :] += delta * wq[:, k].astype(np.int32)
            else:
                q0 = int(wq[r, k]); delta = ox.flip_int8_bit(q0, fault["bit"]) - q0
                acc[:, r] += xq[:, k].astype(np.int32) * delta
        mm = ox.linear_epilogue(acc, sx, sw)
        if fault is not None and fault["type"].startswith("RANDOM"):
            r, c = divmod(int(fault["flat_index"]), mm.shape[1])
            mm[r, c] = ox.apply_output_fault(mm[r, c], fault)
    else:
        mm = (ox.dequant(xq, sx) @ ox.dequant(wq, sw).T).astype(F32)
    if cap is not None and cap_name:
        cap[cap_name + ":xq"] = xq
        cap[cap_name + ":mm"] = mm
    y = (mm + bias.reshape(1, -1)).astype(F32)
    if relu:
        y = np.maximum(y, F32(0))
    if residual is not None:
        y = (residual.reshape(-1, y.shape[-1]) + y).astype(F32)
    y = y.reshape(shape[:-1] + (y.shape[-1],))
    if quantize_output:
        q, s = ox.row_quant(y)
        return q, s
    return y


def _attention(qq, sq, kq, sk, vq, sv, mask, mode: str, causal=False, q_pos0=0, fault=None):
    """attention.py:23-36 for a batch.  qq [B,Tq,512] int8, sq [B,Tq,1]; mask: bool [B,1,Tk] or None."""
    B, Tq, _ = qq.shape
    out = np.zeros((B, Tq, D_MODEL), dtype=F32)
    for b in range(B):
        km = mask[b, 0] if mask is not None else None
        if mode == "int-exact":
            fb = None
            if fault is not None and fault["index"][0] == b:
                fb = fault
            out[b] = ox.attention(qq[b], sq[b].reshape(-1), kq[b], sk[b].reshape(-1), vq[b], sv[b].reshape(-1), km, causal, q_pos0, fault=fb)
        else:
            out[b] = _attention_ref_float(qq[b], sq[b], kq[b], sk[b], vq[b], sv[b], km, causal, q_pos0)
    return out


def _attention_ref_float(qq, sq, kq, sk, vq, sv, key_mask, causal, q_pos0):
    Tq, Tk = qq.shape[0], kq.shape[0]
    qh = ox.dequant(qq, sq.reshape(Tq, 1)).reshape(Tq, N_HEADS, D_K).transpose(1, 0, 2)
    kh = ox.dequant(kq, sk.reshape(Tk, 1)).reshape(Tk, N_HEADS, D_K).transpose(1, 0, 2)
    vh = ox.dequant(vq, sv.reshape(Tk, 1)).reshape(Tk, N_HEADS, D_K).transpose(1, 0, 2)
    scores = (np.matmul(qh, kh.transpose(0, 2, 1)).astype(F32) / F32(8.0)).astype(F32)
    visible = np.ones((Tq, Tk), dtype=bool)
    if key_mask is not None:
        visible &= np.asarray(key_mask).astype(bool).reshape(1, Tk)
    if causal:
        visible &= (np.arange(Tk)[None, :] <= (q_pos0 + np.arange(Tq))[:, None])
    scores = np.where(visible[None], scores, F32(-1e9)).astype(F32)
    m = scores.max(-1, keepdims=True)
    e = np.exp((scores - m).astype(F32)).astype(F32)
    p = (e / e.sum(-1, keepdims=True, dtype=F32)).astype(F32)
    p = (np.rint((p * F32(127.0)).astype(F32)) / F32(127.0)).astype(F32)
    ctx = np.matmul(p, vh).astype(F32)
    return ctx.transpose(1, 0, 2).reshape(Tq, D_MODEL)


def _norm(w, prefix, x):
    return ox.layer_norm(x, w[prefix + ".a_2"], w[prefix + ".b_2"])


# ------------------------------------------------------------------------------------------------ encoder / decoder
def encode(w, src_emb: np.ndarray, src_mask: Optional[np.ndarray], mode: str = "int-exact", n_layers: int = 6, cap: Optional[Trace] = None,
           fault: Optional[dict] = None):
    """Encoder.forward (encoder.py:14-18) on embedded input [B,S,512]; returns memory [B,S,512]."""
    x = src_emb.astype(F32)
    for l in range(n_layers):
        p = "encoder.layers.%d" % l
        f = (lambda *t: fault if (fault is not None and fault["module"] == "Encoder" and fault["layer"] == l and fault["target"] in t) else None)  # noqa: E731
        ln = _norm(w, p + ".sublayer.0.norm", x)
        shared = ox.row_quant(ln.reshape(-1, D_MODEL))       # Q,K,V share one quantization of the LN output
        qq, sq = _linear(w, p + ".self_attn.linears.0", ln, mode, quantize_output=True, xq_sx=shared, cap=cap, cap_name="enc%d.q" % l)
        kq, sk = _linear(w, p + ".self_attn.linears.1", ln, mode, quantize_output=True, xq_sx=shared)
        vq, sv = _linear(w, p + ".self_attn.linears.2", ln, mode, quantize_output=True, xq_sx=shared)
        ctx = _attention(qq, sq, kq, sk, vq, sv, src_mask, mode, fault=f("qk", "pv"))
        x = _linear(w, p + ".self_attn.linears.3", ctx, mode, residual=x).reshape(x.shape)
        ln = _norm(w, p + ".sublayer.1.norm", x)
        h = _linear(w, p + ".feed_forward.w_1", ln, mode, relu=True, fault=f("ffn1"))
        x = _linear(w, p + ".feed_forward.w_2", h, mode, residual=x, fault=f("ffn2")).reshape(x.shape)
        if cap is not None:
            cap["enc%d.out" % l] = x.copy()
            cap["enc%d.qq" % l] = qq
    return _norm(w, "encoder.norm", x)


def cross_kv(w, memory: np.ndarray, mode: str = "int-exact", n_layers: int = 6):
    """The 12 hoisted cross-attention K/V projections (decoder MatMul_0..11) sharing one quantization of `memory`."""
    shared = ox.row_quant(memory.reshape(-1, D_MODEL))
    out = []
    for l in range(n_layers):
        p = "decoder.layers.%d.src_attn.linears." % l
        out.append((_linear(w, p + "1", memory, mode, quantize_output=True, xq_sx=shared),
                    _linear(w, p + "2", memory, mode, quantize_output=True, xq_sx=shared)))
    return out


def decode(w, tgt_emb: np.ndarray, memory: np.ndarray, src_mask, mode: str = "int-exact", n_layers: int = 6, ckv=None,
           self_kv=None, pos0: int = 0, fault: Optional[dict] = None):
    """Decoder.forward (decoder.py:13-16) on embedded target prefix [B,T,512] with the causal mask
    (subsequent_mask, utils.py:10-14).  With `self_kv` (list of per-layer dicts) only the new positions are given in
    tgt_emb (starting at pos0) and the keys/values of earlier positions come from / are appended to the cache:
    every op other than attention is row-local, so this equals the reference's full-prefix recompute bit for bit."""
    x = tgt_emb.astype(F32)
    ckv = ckv if ckv is not None else cross_kv(w, memory, mode, n_layers)
    for l in range(n_layers):
        p = "decoder.layers.%d" % l

        def f(*t, _l=l):
            if fault is None or fault["module"] != "Decoder" or fault["layer"] != _l or fault["target"] not in t:
                return None
            # the attention oracle addresses faults by the reference's target names qk / pv
            return dict(fault, target={"cqk": "qk", "cpv": "pv"}.get(fault["target"], fault["target"]))
        ln = _norm(w, p + ".sublayer.0.norm", x)
        shared = ox.row_quant(ln.reshape(-1, D_MODEL))
        qq, sq = _linear(w, p + ".self_attn.linears.0", ln, mode, quantize_output=True, xq_sx=shared)
        kq, sk = _linear(w, p + ".self_attn.linears.1", ln, mode, quantize_output=True, xq_sx=shared)
        vq, sv = _linear(w, p + ".self_attn.linears.2", ln, mode, quantize_output=True, xq_sx=shared)
        if self_kv is not None:
            c = self_kv[l]
            if "k" in c:
                kq = np.concatenate([c["k"], kq], axis=1); sk = np.concatenate([c["sk"], sk], axis=1)
                vq = np.concatenate([c["v"], vq], axis=1); sv = np.concatenate([c["sv"], sv], axis=1)
            c["k"], c["sk"], c["v"], c["sv"] = kq, sk, vq, sv
        ctx = _attention(qq, sq, kq, sk, vq, sv, None, mode, causal=True, q_pos0=pos0, fault=f("qk", "pv"))
        x = _linear(w, p + ".self_attn.linears.3", ctx, mode, residual=x).reshape(x.shape)
        ln = _norm(w, p + ".sublayer.1.norm", x)
        cq, scq = _linear(w, p + ".src_attn.linears.0", ln, mode, quantize_output=True)
        (ckq, sck), (cvq, scv) = ckv[l]
        ctx = _attention(cq, scq, ckq, sck, cvq, scv, src_mask, mode, fault=f("cqk", "cpv"))
        x = _linear(w, p + ".src_attn.linears.3", ctx, mode, residual=x).reshape(x.shape)
        ln = _norm(w, p + ".sublayer.2.norm", x)
        h = _linear(w, p + ".feed_forward.w_1", ln, mode, relu=True, fault=f("ffn1"))
        x = _linear(w, p + ".feed_forward.w_2", h, mode, residual=x, fault=f("ffn2")).reshape(x.shape)
    return _norm(w, "decoder.norm", x)


def greedy_decode(w, src_ids: np.ndarray, src_mask: np.ndarray, max_len: int = 72, start_symbol: int = 0, mode: str = "int-exact",
                  n_layers: int = 6, kv_cache: bool = True, return_margins: bool = False, pe=None, fault: Optional[dict] = None):
    """greedy_decode (parallelized_inject_onnx_transformer.py:536-758, batched as batch_output.py:659-672):
    memory = encode(src); ys = [<s>]; 71 x { out = decode(ys); next = argmax(generator(out[:, -1])); ys = cat }.
    No early stop at </s>.  kv_cache=False re-runs the full prefix every step exactly as the reference does."""
    B = src_ids.shape[0]
    pe = pe if pe is not None else ox.positional_encoding(max(max_len, src_ids.shape[1]) + 1)
    memory = encode(w, ox.embed(src_ids, w["src_embed.0.lut.weight"], pe), src_mask, mode, n_layers, fault=fault)
    ckv = cross_kv(w, memory, mode, n_layers)
    ys = np.full((B, 1), start_symbol, dtype=np.int64)
    caches = [dict() for _ in range(n_layers)] if kv_cache else None
    margins = []
    for i in range(max_len - 1):
        if kv_cache:
            emb = ox.embed(ys[:, i:i + 1], w["tgt_embed.0.lut.weight"], pe, pos0=i)
            out = decode(w, emb, memory, src_mask, mode, n_layers, ckv, caches, pos0=i,
                         fault=fault if (fault is not None and fault.get("step", 0) == i) else None)
        else:
            emb = ox.embed(ys, w["tgt_embed.0.lut.weight"], pe)
            out = decode(w, emb, memory, src_mask, mode, n_layers, ckv)
        nxt, logits = ox.generator(out[:, -1], w["generator.proj.weight"], w["generator.proj.bias"])
        srt = np.sort(logits, axis=-1)
        margins.append(srt[:, -1] - srt[:, -2])
        ys = np.concatenate([ys, nxt.reshape(B, 1).astype(np.int64)], axis=1)
    if return_margins:
        return ys, np.stack(margins, axis=1), memory
    return ys
