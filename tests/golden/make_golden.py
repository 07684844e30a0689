"""Generate golden vectors from the REFERENCE's own torch modules (run in the build container only; needs
/root/reference, which does not exist on the GPU box -- the outputs are committed as tests/golden/*.npz).

What is executed here is reference code, unmodified, imported from /root/reference:
    model.make_model, get_quantized_model.smooth_lm / quantize_transformer, quant_linear.W8A8Linear,
    attention.MultiHeadedAttention, layer_norm.LayerNorm, encoder_decoder.EncoderDecoder.{encode,decode},
    generator.Generator
with the un-installed third-party imports it never calls on this path (brevitas, qonnx) stubbed by empty modules
 and attention.print silenced.  Parameters come from onnx_transformer_b200.weights (seeded numpy)
and are loaded into the reference model through load_state_dict, so the fixtures need not store any weights.

    python tests/golden/make_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def _stub_modules():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m
    mod("brevitas"); mod("brevitas.nn"); mod("brevitas.export", export_onnx_qcdq=None); mod("brevitas.quant")
    mod("brevitas.quant.scaled_int", Int32Bias=None, Uint8ActPerTensorFloat=None, Int8ActPerTensorFloat=None,
        Int8WeightPerChannelFloat=None)
    mod("qonnx"); mod("qonnx.core"); mod("qonnx.core.modelwrapper", ModelWrapper=None)


def build_reference_model(weights, scales, n_layers, src_vocab, tgt_vocab):
    _stub_modules()
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        import attention
        attention.print = lambda *a, **k: None
        from model import make_model
        import get_quantized_model as gq
        import warnings
        warnings.filterwarnings("ignore")
        model = make_model(src_vocab, tgt_vocab, N=n_layers)
        sd = model.state_dict()
        for k in sd:
            if k in weights:
                sd[k] = torch.from_numpy(np.array(weights[k]))
            else:
                assert k.endswith(".pe"), k
        model.load_state_dict(sd)
        model.eval()
        with torch.no_grad():
            if scales is not None:
                gq.smooth_lm(model, {k: torch.from_numpy(v) for k, v in scales.items()})
            model = gq.quantize_transformer(model)
    finally:
        os.chdir(cwd)
    return model


def subsequent_mask(size):
    return torch.from_numpy(np.triu(np.ones((1, size, size)), k=1).astype("uint8")) == 0


def greedy(model, src, src_mask, max_len, start_symbol=0):
    """The batched greedy loop of batch_output.py:659-672 / the reference's greedy_decode, on the torch modules."""
    with torch.no_grad():
        memory = model.encode(src, src_mask)
        ys = torch.zeros(src.shape[0], 1, dtype=src.dtype).fill_(start_symbol)
        outs, margins = [], []
        for _ in range(max_len - 1):
            out = model.decode(memory, src_mask, ys, subsequent_mask(ys.size(1)).type_as(src.data))
            prob = model.generator(out[:, -1])
            top2 = torch.topk(prob, 2, dim=1).values
            margins.append((top2[:, 0] - top2[:, 1]).numpy())
            _, nxt = torch.max(prob, dim=1)
            ys = torch.cat([ys, nxt.reshape(-1, 1)], dim=1)
            outs.append(out[:, -1].numpy().copy())
    return memory.numpy(), ys.numpy(), np.stack(outs, 1), np.stack(margins, 1)


def main():
    from onnx_transformer_b200 import weights as W
    torch.manual_seed(0)
    torch.set_num_threads(8)
    out_dir = os.path.dirname(os.path.abspath(__file__))

    # ---- case A: 2-layer model, small vocab, SmoothQuant with synthetic scales, ragged source lengths
    cfg = dict(seed=1, n_layers=2, src_vocab=211, tgt_vocab=197, batch=3, src_len=11, max_len=9)
    fw = W.init_float_weights(cfg["seed"], cfg["src_vocab"], cfg["tgt_vocab"], cfg["n_layers"], randomize_norms=True)
    sc = W.synthetic_scales(cfg["seed"], cfg["n_layers"])
    model = build_reference_model(fw, sc, cfg["n_layers"], cfg["src_vocab"], cfg["tgt_vocab"])
    ids, mask = W.synthetic_tokens(cfg["seed"], cfg["batch"], cfg["src_len"], cfg["src_vocab"], min_len=5)
    memory, ys, last_h, margins = greedy(model, torch.from_numpy(ids), torch.from_numpy(mask), cfg["max_len"])
    sd = model.state_dict()
    # a few smoothed + fake-quantized parameters, to pin the oracle's get_quantized() restatement
    probe = {k: sd[k].numpy() for k in ["encoder.layers.0.self_attn.linears.0.weight", "encoder.layers.1.feed_forward.w_1.weight",
                                         "decoder.layers.0.src_attn.linears.1.weight", "encoder.layers.0.sublayer.0.norm.a_2",
                                         "decoder.layers.1.sublayer.1.norm.b_2"]}
    src_emb = model.src_embed(torch.from_numpy(ids)).detach().numpy()
    np.savez_compressed(os.path.join(out_dir, "ref_torch_case_a.npz"), cfg=np.array(repr(cfg)), ids=ids, mask=mask, memory=memory, ys=ys,
                        last_h=last_h, margins=margins, src_emb=src_emb,
                        **{"probe:" + k: v[:8] if v.ndim == 2 else v for k, v in probe.items()})

    # ---- case B: per-module pins on one layer's worth of reference modules (no smoothing), explicit intermediates
    cfg_b = dict(seed=2, n_layers=1, src_vocab=97, tgt_vocab=89, batch=2, src_len=7)
    fw = W.init_float_weights(cfg_b["seed"], cfg_b["src_vocab"], cfg_b["tgt_vocab"], 1)
    model = build_reference_model(fw, None, 1, cfg_b["src_vocab"], cfg_b["tgt_vocab"])
    rng = np.random.default_rng(5)
    x = rng.normal(size=(2, 7, 512)).astype(np.float32)
    with torch.no_grad():
        xt = torch.from_numpy(x)
        layer = model.encoder.layers[0]
        ln = layer.sublayer[0].norm(xt).numpy()
        lin_q = layer.self_attn.linears[0]
        q_out = lin_q(torch.from_numpy(ln)).numpy()                         # fake-quant output (quantize_output=True)
        ffn = layer.feed_forward(torch.from_numpy(ln)).numpy()
        msk = torch.ones(2, 1, 7, dtype=torch.bool); msk[1, 0, 5:] = False
        attn = layer.self_attn(torch.from_numpy(ln), torch.from_numpy(ln), torch.from_numpy(ln), msk).numpy()
        p_attn = layer.self_attn.attn.numpy()                               # rint(127 p) (in-place round survives, attention.py:33-35)
        enc_layer = layer(xt, msk).numpy()
        gen = model.generator(torch.from_numpy(x[:, -1])).numpy()
        emb = model.tgt_embed(torch.from_numpy(np.array([[3, 5, 7], [11, 13, 17]]))).numpy()
    np.savez_compressed(os.path.join(out_dir, "ref_torch_case_b.npz"), cfg=np.array(repr(cfg_b)), x=x, ln=ln, q_out=q_out, ffn=ffn,
                        attn=attn, p_attn=p_attn, enc_layer=enc_layer, gen=gen, emb=emb, mask=msk.numpy())
    print("golden fixtures written to", out_dir)


if __name__ == "__main__":
    main()
