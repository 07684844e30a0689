"""SmoothQuant pre-pass of the product (weights.smooth_lm, restating get_quantized_model.py:10-36,46-148) against the values the
reference's own torch code produced (tests/golden/ref_torch_case_a.npz, written by tests/golden/make_golden.py) and against the
oracle's independent restatement; on the GPU, the engine fed by the product's pre-pass reproduces the reference's greedy decode."""
import os

import numpy as np
import pytest

from onnx_transformer_b200 import weights as W
from oracle import intexact as ox
from oracle import model as om

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_torch_case_a.npz")


def _case():
    g = np.load(GOLD, allow_pickle=False)
    cfg = eval(str(g["cfg"]))      # the dict literal make_golden.py wrote
    fw = W.init_float_weights(cfg["seed"], cfg["src_vocab"], cfg["tgt_vocab"], cfg["n_layers"], randomize_norms=True)
    sc = W.synthetic_scales(cfg["seed"], cfg["n_layers"])
    return g, cfg, fw, sc


def test_smoothing_plan_covers_every_layernorm_linear_pair():
    plan = W.smoothing_plan(6)
    assert len(plan) == 6 * 2 + 6 * 3
    assert ("decoder.layers.3.sublayer.1.norm", ["decoder.layers.3.src_attn.linears.%d" % i for i in range(3)],
            "decoder.layers.3.src_attn.linears.0") in plan
    assert all(ln.endswith(".norm") for ln, _, _ in plan)


def test_smooth_lm_matches_reference_torch_values():
    g, cfg, fw, sc = _case()
    sm = W.smooth_lm(fw, sc, cfg["n_layers"])
    # LayerNorm parameters are not quantised afterwards: direct comparison with the reference model's state_dict
    for k in ("encoder.layers.0.sublayer.0.norm.a_2", "decoder.layers.1.sublayer.1.norm.b_2"):
        np.testing.assert_allclose(sm[k], g["probe:" + k], rtol=2e-7, atol=0)
    # linear weights: the reference stores round(W/s)*s (W8A8Linear.from_float); same fake-quant on the product's smoothed weight
    for k in ("encoder.layers.0.self_attn.linears.0.weight", "encoder.layers.1.feed_forward.w_1.weight", "decoder.layers.0.src_attn.linears.1.weight"):
        q, s = ox.row_quant(sm[k][:8])
        ref = g["probe:" + k]
        got = ox.dequant(q, s)
        assert np.mean(np.abs(got - ref) > 1e-6 * np.abs(ref).max()) < 1e-3     # +-1 LSB where pow() differs by an ulp
    # and the oracle's own restatement agrees with the product's
    wo = {k: v.copy() for k, v in fw.items()}
    om.smooth_lm(wo, sc, cfg["n_layers"])
    for k in sm:
        assert np.array_equal(sm[k], wo[k]), k
    assert fw["encoder.layers.0.sublayer.0.norm.a_2"] is not sm["encoder.layers.0.sublayer.0.norm.a_2"]   # input left untouched


@pytest.mark.gpu
def test_engine_on_smoothed_weights_reproduces_the_reference_decode():
    import torch
    from onnx_transformer_b200.engine import QuantizedTransformer
    g, cfg, fw, sc = _case()
    eng = QuantizedTransformer(W.smooth_lm(fw, sc, cfg["n_layers"]), n_layers=cfg["n_layers"], max_len=cfg["max_len"])
    ys = eng.greedy_decode(torch.from_numpy(g["ids"]).cuda(), torch.from_numpy(g["mask"]).cuda(), cfg["max_len"]).cpu().numpy()
    ref, margins = g["ys"], g["margins"]
    for b in range(cfg["batch"]):
        for t in range(cfg["max_len"] - 1):
            if ys[b, t + 1] != ref[b, t + 1]:
                assert margins[b, t] < 0.1, (b, t, float(margins[b, t]))     # identical wherever the top-2 margin allows
                break
    assert (ys[:, 1:] == ref[:, 1:]).mean() > 0.8
