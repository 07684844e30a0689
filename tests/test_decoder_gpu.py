"""The persistent greedy decoders -- csrc/ot_cdecoder.cu ("cluster": 8-CTA clusters exchanging rows / column slices through
distributed shared memory) and csrc/ot_decoder.cu ("grid": phases separated by grid barriers), all steps in ONE kernel -- against
the per-op kernel path (CUDA-graph replay / eager) of the same engine and against the numpy oracle.  The two device paths
execute the same arithmetic instruction for instruction, so tokens AND KV caches must be bit-identical; against the oracle
tokens are identical wherever the top-2 logit margin allows (tolerance class of the float reductions)."""
import numpy as np
import pytest
import torch

from onnx_transformer_b200 import weights as W
from oracle import model as om

pytestmark = pytest.mark.gpu


DECODERS = ["cluster", "grid"]


def _engines(seed, n_layers, src_vocab, tgt_vocab, max_len, decoder="cluster", spc=8):
    from onnx_transformer_b200.engine import QuantizedTransformer
    fw = W.init_float_weights(seed, src_vocab, tgt_vocab, n_layers, randomize_norms=True)
    return fw, QuantizedTransformer(fw, n_layers=n_layers, max_len=max_len, persistent=True, decoder=decoder, sentences_per_cluster=spc), \
        QuantizedTransformer(fw, n_layers=n_layers, max_len=max_len, persistent=False)


def _caches(eng, B, S):
    ws = eng._dec_workspace(B, S)
    return [t.clone() for name in ("kc", "vc", "skc", "svc") for t in ws[name]]


def _assert_same_state(ep, eg, ys_p, ys_g, B, S, steps):
    assert torch.equal(ys_p, ys_g)
    for a, b in zip(_caches(ep, B, S), _caches(eg, B, S)):
        a, b = a[:, :steps], b[:, :steps]
        if a.dtype == torch.float32:
            assert torch.equal(a.view(torch.int32), b.view(torch.int32))
        else:
            assert torch.equal(a, b)


@pytest.mark.parametrize("decoder", DECODERS)
@pytest.mark.parametrize("B,S,min_len", [(3, 11, 5), (1, 4, 0), (7, 33, 9)])
def test_persistent_decoder_bit_identical_to_per_op_path_small(B, S, min_len, decoder):
    fw, ep, eg = _engines(1, 2, 211, 197, 9, decoder)
    ids, mask = W.synthetic_tokens(B + S, B, S, 211, min_len=min_len)
    idt, mt = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
    ys_p = ep.greedy_decode(idt, mt, 9)
    assert ep.persistent_steps == 8 and ep.graph_replays == 0
    ys_g = eg.greedy_decode(idt, mt, 9)
    assert eg.persistent_steps == 0 and eg.graph_replays == 8
    _assert_same_state(ep, eg, ys_p, ys_g, B, S, 8)
    ys_e = eg.greedy_decode(idt, mt, 9, use_graph=False)
    assert torch.equal(ys_p, ys_e)
    # and against the oracle
    ref, margins, _ = om.greedy_decode(om.get_quantized(fw, None, 2), ids, mask, 9, 0, "int-exact", 2, return_margins=True)
    ys = ys_p.cpu().numpy()
    for b in range(B):
        for t in range(8):
            if ys[b, t + 1] != ref[b, t + 1]:
                assert margins[b, t] < 0.1, (b, t, margins[b, t])
                break


@pytest.mark.parametrize("B,spc", [(13, 5), (9, 8), (20, 3)])
def test_cluster_decoder_sentence_groups(B, spc):
    """Several clusters, a ragged last group, fewer than 8 sentences per cluster."""
    fw, ep, eg = _engines(2, 2, 211, 197, 10, "cluster", spc)
    ids, mask = W.synthetic_tokens(B, B, 17, 211, min_len=3)
    idt, mt = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
    ys_p = ep.greedy_decode(idt, mt, 10)
    ys_g = eg.greedy_decode(idt, mt, 10)
    assert ep.persistent_steps == 9
    _assert_same_state(ep, eg, ys_p, ys_g, B, 17, 9)


@pytest.mark.parametrize("decoder", DECODERS)
def test_persistent_decoder_full_size_batch64(decoder):
    """BASELINE config #2 shape: Transformer-base, 64 sentences x 64 source tokens, 71 greedy steps."""
    fw, ep, eg = _engines(0, 6, W.SRC_VOCAB, W.TGT_VOCAB, W.MAX_LEN, decoder)
    ids, mask = W.synthetic_tokens(1000, 64, 64)
    idt, mt = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
    ys_p = ep.greedy_decode(idt, mt)
    ys_g = eg.greedy_decode(idt, mt)
    _assert_same_state(ep, eg, ys_p, ys_g, 64, 64, 71)
    # repeatable (the barrier counter / row maxima are reset correctly between launches)
    assert torch.equal(ep.greedy_decode(idt, mt), ys_p)
    # ragged padding + a second shape on the same engine
    ids2, mask2 = W.synthetic_tokens(5, 17, 40, min_len=3)
    i2, m2 = torch.from_numpy(ids2).cuda(), torch.from_numpy(mask2).cuda()
    assert torch.equal(ep.greedy_decode(i2, m2), eg.greedy_decode(i2, m2))


@pytest.mark.parametrize("decoder", DECODERS)
def test_persistent_decoder_with_an_injected_step(decoder):
    """A fault step runs through the per-op kernels between two persistent launches: same result as the all-per-op run."""
    from onnx_transformer_b200.engine import FaultSpec
    fw, ep, eg = _engines(3, 2, 211, 197, 12, decoder)
    ids, mask = W.synthetic_tokens(3, 4, 10, 211, min_len=4)
    idt, mt = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
    golden = ep.greedy_decode(idt, mt, 12)
    differs = 0
    for step, target, bit in [(0, "ffn1", 6), (4, "q", 7), (10, "o", 6)]:
        f = FaultSpec("Decoder", 1, target, "INPUT", bit=bit, flat_index=2 * 512 + 17, step=step)
        ys_p = ep.greedy_decode(idt, mt, 12, fault=f)
        ys_g = eg.greedy_decode(idt, mt, 12, fault=f)
        assert torch.equal(ys_p, ys_g)
        assert torch.equal(ys_p[:, :step + 1], golden[:, :step + 1])
        differs += int(not torch.equal(ys_p, golden))
    assert ep.persistent_steps > 0


@pytest.mark.parametrize("kind", ["fp16_overflow", "nan_bias", "near_ties", "exact_only"])
def test_screening_generator_corner_cases(kind, monkeypatch):
    """The cluster decoder's generator screens the vocabulary on the tensor cores (fp16) and re-evaluates the entries within the error
    bound of the maximum with the exact fp32 chain (ot_cdecoder.cu phase_generator_tc): tokens must equal the per-op path's
      * when the fp16 copy overflows (huge generator weights: every CTA of a cluster falls back to the exact generator),
      * when a logit is NaN (a NaN bias entry ranks above every number, as torch.max / np.argmax),
      * when MANY entries sit within the bound (duplicated weight rows: exact ties -> lowest index wins; near ties),
      * and with the screening switched off (OT_CD_GEN_TC=0)."""
    from onnx_transformer_b200.engine import QuantizedTransformer
    fw = W.init_float_weights(3, 211, 197, 2, randomize_norms=True)
    gw, gb = fw["generator.proj.weight"].copy(), fw["generator.proj.bias"].copy()
    if kind == "fp16_overflow":
        gw *= 3.0e6
    elif kind == "nan_bias":
        gb[77] = np.nan
    elif kind == "near_ties":
        gw[10:60] = gw[5]                        # 51 identical rows: exact ties
        gb[10:60] = gb[5]
        gw[100:140] = gw[5] * (1.0 + 1e-4)       # and 40 rows within the screening bound
        gb[100:140] = gb[5]
    else:
        monkeypatch.setenv("OT_CD_GEN_TC", "0")
    fw["generator.proj.weight"], fw["generator.proj.bias"] = gw, gb
    ep = QuantizedTransformer(fw, n_layers=2, max_len=12, persistent=True, decoder="cluster")
    eg = QuantizedTransformer(fw, n_layers=2, max_len=12, persistent=False)
    for B, S in ((11, 17), (8, 9)):
        ids, mask = W.synthetic_tokens(B * S, B, S, 211, min_len=4)
        idt, mt = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
        ys_p = ep.greedy_decode(idt, mt, 12)
        ys_g = eg.greedy_decode(idt, mt, 12)
        assert torch.equal(ys_p, ys_g), kind
    if kind == "nan_bias":
        assert bool((ys_p[:, 1:] == 77).all())
