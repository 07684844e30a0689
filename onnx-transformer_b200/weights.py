"""Synthetic model parameters with the reference's names and initialisation (BASELINE.json: "random-init weights").

`init_float_weights` reproduces what `make_model(src_vocab, tgt_vocab, N=6)` constructs (model.py:15-37): every
parameter with dim > 1 is `xavier_uniform_`, Linear biases keep torch's default U(-1/sqrt(fan_in), 1/sqrt(fan_in)),
LayerNorm a_2 = 1 / b_2 = 0 (layer_norm.py:8-9).  The random stream is numpy's PCG64 (not torch's Philox), so the
*values* differ from a torch-seeded model but the distribution and shapes are the reference's; the golden
fixtures load these very arrays into the reference torch modules (tests/golden/make_golden.py).

Keys are the reference's `state_dict` names.  This is data generation only -- no model arithmetic happens here.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np

D_MODEL, D_FF, N_HEADS = 512, 2048, 8
SRC_VOCAB, TGT_VOCAB = 5337, 4444   # error.log:3-5
PAD_ID, BOS_ID, EOS_ID = 2, 0, 1    # parallelized_inject_onnx_transformer.py:149-154
MAX_LEN = 72                        # max_padding=72 -> 71 greedy steps


def _xavier(rng, shape) -> np.ndarray:
    fan_out, fan_in = shape[0], shape[1]
    a = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-a, a, size=shape).astype(np.float32)


def _bias(rng, fan_in, n) -> np.ndarray:
    b = 1.0 / math.sqrt(fan_in)
    return rng.uniform(-b, b, size=(n,)).astype(np.float32)


def linear_names(n_layers: int = 6) -> List[Tuple[str, int, int]]:
    """(state_dict prefix, out_features, in_features) of every quantized linear, in module order."""
    out = []
    for l in range(n_layers):
        for i in range(4):
            out.append(("encoder.layers.%d.self_attn.linears.%d" % (l, i), D_MODEL, D_MODEL))
        out.append(("encoder.layers.%d.feed_forward.w_1" % l, D_FF, D_MODEL))
        out.append(("encoder.layers.%d.feed_forward.w_2" % l, D_MODEL, D_FF))
    for l in range(n_layers):
        for attn in ("self_attn", "src_attn"):
            for i in range(4):
                out.append(("decoder.layers.%d.%s.linears.%d" % (l, attn, i), D_MODEL, D_MODEL))
        out.append(("decoder.layers.%d.feed_forward.w_1" % l, D_FF, D_MODEL))
        out.append(("decoder.layers.%d.feed_forward.w_2" % l, D_MODEL, D_FF))
    return out


def norm_names(n_layers: int = 6) -> List[str]:
    out = []
    for l in range(n_layers):
        out += ["encoder.layers.%d.sublayer.%d.norm" % (l, j) for j in range(2)]
    out.append("encoder.norm")
    for l in range(n_layers):
        out += ["decoder.layers.%d.sublayer.%d.norm" % (l, j) for j in range(3)]
    out.append("decoder.norm")
    return out


def init_float_weights(seed: int = 0, src_vocab: int = SRC_VOCAB, tgt_vocab: int = TGT_VOCAB, n_layers: int = 6,
                       randomize_norms: bool = False) -> Dict[str, np.ndarray]:
    """Float (un-quantized) parameters of the Annotated-Transformer model, reference state_dict naming.
    randomize_norms=True perturbs LayerNorm gains/offsets (a trained model's are not 1/0) for stronger tests."""
    rng = np.random.default_rng(seed)
    w: Dict[str, np.ndarray] = {}
    for prefix, n_out, n_in in linear_names(n_layers):
        w[prefix + ".weight"] = _xavier(rng, (n_out, n_in))
        w[prefix + ".bias"] = _bias(rng, n_in, n_out)
    for prefix in norm_names(n_layers):
        if randomize_norms:
            w[prefix + ".a_2"] = (1.0 + 0.1 * rng.standard_normal(D_MODEL)).astype(np.float32)
            w[prefix + ".b_2"] = (0.1 * rng.standard_normal(D_MODEL)).astype(np.float32)
        else:
            w[prefix + ".a_2"] = np.ones(D_MODEL, np.float32)
            w[prefix + ".b_2"] = np.zeros(D_MODEL, np.float32)
    w["src_embed.0.lut.weight"] = _xavier(rng, (src_vocab, D_MODEL))
    w["tgt_embed.0.lut.weight"] = _xavier(rng, (tgt_vocab, D_MODEL))
    w["generator.proj.weight"] = _xavier(rng, (tgt_vocab, D_MODEL))
    w["generator.proj.bias"] = _bias(rng, D_MODEL, tgt_vocab)
    return w


def synthetic_scales(seed: int = 0, n_layers: int = 6) -> Dict[str, np.ndarray]:
    """Stand-in for scales/transformer_scales.pt (per-input-channel activation abs-max, values 1.15-15.65 in the
    reference's file): one vector per quantized linear, keyed like get_quantized_scales.py:125 does."""
    rng = np.random.default_rng(seed + 7919)
    return {prefix: rng.uniform(1.15, 15.65, size=(n_in,)).astype(np.float32) for prefix, _, n_in in linear_names(n_layers)}


def synthetic_tokens(seed: int, batch: int, src_len: int, src_vocab: int = SRC_VOCAB, min_len: int = 0):
    """source ids randint(4, vocab); optional ragged lengths padded with PAD_ID=2.
    Returns (ids int64 [B,S], key mask bool [B,1,S])."""
    rng = np.random.default_rng(seed + 104729)
    ids = rng.integers(4, src_vocab, size=(batch, src_len), dtype=np.int64)
    if min_len and min_len < src_len:
        lens = rng.integers(min_len, src_len + 1, size=batch)
        for b in range(batch):
            ids[b, lens[b]:] = PAD_ID
    mask = (ids != PAD_ID)[:, None, :]
    return ids, mask


# ------------------------------------------------------------------------------------------------ SmoothQuant pre-pass
def smoothing_plan(n_layers: int = 6) -> List[Tuple[str, List[str], str]]:
    """(LayerNorm, linears fed by it, key of the activation-scale vector) for every smoothing step of
    get_quantized_model.smooth_lm (:46-148).  Encoder layer: norm0 -> self-attn q,k,v; norm1 -> w_1.  Decoder layer: norm0 ->
    self-attn q,k,v; norm1 -> src-attn q AND k,v (the reference scales the k/v projections of `memory` too, :123-131: kept);
    norm2 -> w_1.  The scale vector of a q/k/v group is the one recorded for its first linear."""
    plan = []
    for side, attn_groups in (("encoder", [(0, "self_attn")]), ("decoder", [(0, "self_attn"), (1, "src_attn")])):
        for l in range(n_layers):
            p = "%s.layers.%d" % (side, l)
            for sub, attn in attn_groups:
                fcs = ["%s.%s.linears.%d" % (p, attn, i) for i in range(3)]
                plan.append(("%s.sublayer.%d.norm" % (p, sub), fcs, fcs[0]))
            ffn_sub = len(attn_groups)
            plan.append(("%s.sublayer.%d.norm" % (p, ffn_sub), [p + ".feed_forward.w_1"], p + ".feed_forward.w_1"))
    return plan


def smooth_lm(float_weights: Dict[str, np.ndarray], act_scales: Dict[str, np.ndarray], n_layers: int = 6, alpha: float = 0.5) -> Dict[str, np.ndarray]:
    """SmoothQuant as the reference applies it before quantisation (get_quantized_model.py:10-36 smooth_ln_fcs, :46-148 smooth_lm):
    an OFFLINE transform of the fp32 state_dict -- per input channel j of a LayerNorm -> linear(s) pair,
        s_j = clamp(act_j^alpha / clamp(max_rows |W[:, j]|, 1e-5)^(1-alpha), 1e-5);  a_2, b_2 /= s;  W[:, j] *= s_j
    -- so it stays on the host, like in the reference (torch CPU there, numpy fp32 here; same op order).  Returns a new dict; the
    result feeds QuantizedTransformer / graph.build_*_graph."""
    w = dict(float_weights)
    f32 = np.float32
    for ln, fcs, key in smoothing_plan(n_layers):
        act = np.asarray(act_scales[key], dtype=f32)
        col_max = np.max(np.stack([np.max(np.abs(w[fc + ".weight"]), axis=0) for fc in fcs]), axis=0).astype(f32)
        col_max = np.maximum(col_max, f32(1e-5))
        s = np.maximum((np.power(act, f32(alpha)) / np.power(col_max, f32(1.0 - alpha))).astype(f32), f32(1e-5))
        assert s.shape == w[ln + ".a_2"].shape
        w[ln + ".a_2"] = (w[ln + ".a_2"] / s).astype(f32)
        w[ln + ".b_2"] = (w[ln + ".b_2"] / s).astype(f32)
        for fc in fcs:
            w[fc + ".weight"] = (w[fc + ".weight"] * s[None, :]).astype(f32)
    return w


def load_act_scales(path: str) -> Dict[str, np.ndarray]:
    """scales/transformer_scales.pt of the reference (get_quantized_scales.py:125): {linear name: abs-max per input channel}."""
    import torch
    blob = torch.load(path, map_location="cpu")
    return {k: np.asarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v, dtype=np.float32) for k, v in blob.items()}
