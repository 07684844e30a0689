"""GPU parity of the fused KV-cached engine against the numpy oracle model (int-exact mode) on identical seeded
weights and token batches.  Integer tensors bit-exact up to rounding-boundary flips caused by float reductions
(LayerNorm / softmax / P.V, tolerance class 1e-3); greedy tokens identical wherever the top-2 margin allows."""
import numpy as np
import pytest
import torch

from onnx_transformer_b200 import weights as W
from oracle import intexact as ox
from oracle import model as om

pytestmark = pytest.mark.gpu


def _setup(seed, n_layers, src_vocab, tgt_vocab, B, S, min_len, max_len):
    from onnx_transformer_b200.engine import QuantizedTransformer
    fw = W.init_float_weights(seed, src_vocab, tgt_vocab, n_layers, randomize_norms=True)
    eng = QuantizedTransformer(fw, n_layers=n_layers, max_len=max_len)
    wq = om.get_quantized(fw, None, n_layers)
    ids, mask = W.synthetic_tokens(seed, B, S, src_vocab, min_len=min_len)
    return eng, wq, ids, mask


def test_weight_preparation_bit_exact():
    eng, wq, _, _ = _setup(4, 1, 50, 40, 1, 4, 0, 8)
    q, s = ox.row_quant(wq["encoder.layers.0.feed_forward.w_1.weight"])
    assert np.array_equal(eng.enc[0]["w1"].wq.cpu().numpy(), q)
    assert np.array_equal(eng.enc[0]["w1"].sw.cpu().numpy().view(np.uint32), s.reshape(-1).view(np.uint32))
    qs = [ox.row_quant(wq["decoder.layers.0.self_attn.linears.%d.weight" % i])[0] for i in range(3)]
    assert np.array_equal(eng.dec[0]["qkv"].wq.cpu().numpy(), np.concatenate(qs, 0))


def test_encoder_memory_and_integer_tensors():
    eng, wq, ids, mask = _setup(1, 2, 211, 197, 3, 11, 5, 9)
    cap = {}
    mem = eng.encode(torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda(), capture=cap).cpu().numpy()
    ocap = om.Trace()
    pe = ox.positional_encoding(80)
    ref = om.encode(wq, ox.embed(ids, wq["src_embed.0.lut.weight"], pe), mask, "int-exact", 2, cap=ocap)
    # first-layer Q projection output (Round_37-style int8 tensor): bit-exact except rounding-boundary flips
    q0 = cap["enc0.qkv"].cpu().numpy()[:, :512].reshape(3, 11, 512)
    assert np.mean(q0 != ocap["enc0.qq"]) < 1e-3 and np.max(np.abs(q0.astype(int) - ocap["enc0.qq"].astype(int))) <= 1
    err = np.abs(mem - ref)
    assert err.mean(axis=(1, 2)).min() < 1e-5 or err.mean() < 5e-3
    assert err.mean() < 2e-2 and err.max() < 0.15


def test_greedy_decode_tokens_and_graph_equivalence():
    eng, wq, ids, mask = _setup(1, 2, 211, 197, 3, 11, 5, 9)
    idt, mt = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
    ys_graph = eng.greedy_decode(idt, mt, 9, use_graph=True).cpu().numpy()
    ys_eager = eng.greedy_decode(idt, mt, 9, use_graph=False).cpu().numpy()
    ys_m, margins, _ = eng.greedy_decode(idt, mt, 9, use_graph=False, return_margins=True)
    assert np.array_equal(ys_graph, ys_eager) and np.array_equal(ys_graph, ys_m.cpu().numpy())
    ref, ref_margins, _ = om.greedy_decode(wq, ids, mask, 9, 0, "int-exact", 2, kv_cache=False, return_margins=True)
    assert ys_graph.shape == ref.shape and np.all(ys_graph[:, 0] == 0)
    for b in range(ids.shape[0]):
        for t in range(8):
            if ys_graph[b, t + 1] != ref[b, t + 1]:
                assert ref_margins[b, t] < 0.1, (b, t, ref_margins[b, t])
                break
    np.testing.assert_allclose(margins.cpu().numpy()[0, 0], ref_margins[0, 0], atol=5e-2)


def test_full_size_model_decode_runs_and_matches_oracle_prefix():
    """Transformer-base (6+6 layers, real vocab sizes) at B=2: first greedy steps vs the oracle."""
    eng, wq, ids, mask = _setup(0, 6, W.SRC_VOCAB, W.TGT_VOCAB, 2, 16, 9, W.MAX_LEN)
    idt, mt = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
    ys = eng.greedy_decode(idt, mt).cpu().numpy()
    assert ys.shape == (2, 72)
    ref, ref_margins, _ = om.greedy_decode(wq, ids, mask, 6, 0, "int-exact", 6, kv_cache=True, return_margins=True)
    for b in range(2):
        for t in range(5):
            if ys[b, t + 1] != ref[b, t + 1]:
                assert ref_margins[b, t] < 0.1, (b, t, ref_margins[b, t])
                break
    # decoding is deterministic and batch-invariant: sentence 0 alone gives the same tokens
    ys0 = eng.greedy_decode(idt[:1], mt[:1]).cpu().numpy()
    assert np.array_equal(ys0[0], ys[0])


def test_int4_weight_engine_matches_oracle():
    """BASELINE config #4: packed int4 weights (unpacked to int8 in shared memory by the GEMM), int8 activations."""
    from onnx_transformer_b200.engine import QuantizedTransformer
    fw = W.init_float_weights(9, 101, 89, 2, randomize_norms=True)
    eng = QuantizedTransformer(fw, n_layers=2, max_len=9, weight_bits=4)
    w4 = om.quantize_weights_int4({k: np.array(v) for k, v in fw.items()})
    assert np.array_equal(eng.enc[0]["w1"].wq8.cpu().numpy(), w4["encoder.layers.0.feed_forward.w_1.weight"])
    assert np.array_equal(eng.enc[0]["w1"].sw.cpu().numpy().view(np.uint32), w4["encoder.layers.0.feed_forward.w_1.int4_scale"].reshape(-1).view(np.uint32))
    ids, mask = W.synthetic_tokens(9, 3, 12, 101, min_len=6)
    idt, mt = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
    mem = eng.encode(idt, mt).cpu().numpy()
    pe = ox.positional_encoding(80)
    ref = om.encode(w4, ox.embed(ids, w4["src_embed.0.lut.weight"], pe), mask, "int-exact", 2)
    err = np.abs(mem - ref)
    assert err.mean() < 2e-2 and err.max() < 0.3
    ys_t = eng.greedy_decode(idt, mt, 9)
    assert eng.persistent_steps == 8          # the cluster decoder runs cfg4 too (on the int8 copy of the 4-bit values)
    eng_g = QuantizedTransformer(fw, n_layers=2, max_len=9, weight_bits=4, persistent=False)     # packed-int4 GEMMs, per-op path
    assert torch.equal(ys_t, eng_g.greedy_decode(idt, mt, 9))
    ys = ys_t.cpu().numpy()
    ref_ys, margins, _ = om.greedy_decode(w4, ids, mask, 9, 0, "int-exact", 2, return_margins=True)
    for b in range(3):
        for t in range(8):
            if ys[b, t + 1] != ref_ys[b, t + 1]:
                assert margins[b, t] < 0.1, (b, t, margins[b, t])
                break


def test_int4_weight_encoder_at_encoder_sizes_equals_its_small_batch_path():
    """cfg4 at encoder sizes (>= 2048 rows): the requant GEMMs take the packed nibbles through the weight-stationary kernel, the
    fp32-output GEMMs run the streaming int8 kernels on a per-launch unpacked scratch.  A sentence's rows must not depend on the batch it
    is in: 40 sentences x 64 tokens (2560 rows, the large-M kernels) == the same sentences 8 at a time (512 rows, the packed-int4 tile
    kernel), bit for bit."""
    from onnx_transformer_b200.engine import QuantizedTransformer
    fw = W.init_float_weights(5, 211, 197, 2, randomize_norms=True)
    eng = QuantizedTransformer(fw, n_layers=2, max_len=9, weight_bits=4)
    ids, mask = W.synthetic_tokens(5, 40, 64, 211, min_len=20)
    idt, mt = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
    big = eng.encode(idt, mt).clone()
    for lo in range(0, 40, 8):
        small = eng.encode(idt[lo:lo + 8], mt[lo:lo + 8])
        assert torch.equal(big[lo:lo + 8], small), lo


def test_cfg3_full_size_encoder_is_sentence_shardable():
    """BASELINE config #3 (encoder only, 512 sentences x 128 source tokens): the size-independent property the multi-GPU partition
    rests on -- a sentence's memory rows do not depend on which other sentences share its batch, bit for bit (no cross-sentence op)."""
    from onnx_transformer_b200.engine import QuantizedTransformer
    fw = W.init_float_weights(0)
    eng = QuantizedTransformer(fw)
    ids, mask = W.synthetic_tokens(7, 512, 128, min_len=40)
    idt, mt = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
    full = eng.encode(idt, mt).clone()
    assert full.shape == (512, 128, 512) and bool(torch.isfinite(full).all())
    for lo, hi in ((0, 8), (250, 314), (505, 512)):          # shards of 8 / 64 / 7 sentences, as rank r of N would see them
        part = eng.encode(idt[lo:hi].contiguous(), mt[lo:hi].contiguous())
        assert torch.equal(part.view(torch.int32), full[lo:hi].view(torch.int32)), (lo, hi)


def test_cluster_decoder_any_batch_size():
    """The cluster-resident decoder has no co-residency requirement: 80 sentences = 10 independent clusters, same tokens as the
    per-op path."""
    from onnx_transformer_b200.engine import QuantizedTransformer
    fw = W.init_float_weights(5, 211, 197, 2, randomize_norms=True)
    ep = QuantizedTransformer(fw, n_layers=2, max_len=12, decoder="cluster")
    eg = QuantizedTransformer(fw, n_layers=2, max_len=12, persistent=False)
    ids, mask = W.synthetic_tokens(9, 80, 23, 211, min_len=6)
    idt, mt = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
    assert torch.equal(ep.greedy_decode(idt, mt, 12), eg.greedy_decode(idt, mt, 12))
    assert ep.persistent_steps == 11


def test_front_graph_and_fused_context_quant_are_bit_identical():
    """The CUDA-graph replay of encoder + cross-K/V + state reset (static buffers) against the eager launches, on changing inputs and
    shapes; the in-kernel RowQuant of the attention context against the two-launch form."""
    from onnx_transformer_b200.engine import QuantizedTransformer
    fw = W.init_float_weights(5, 211, 197, 2, randomize_norms=True)
    eg = QuantizedTransformer(fw, n_layers=2, max_len=12)
    ee = QuantizedTransformer(fw, n_layers=2, max_len=12)
    ee.front_graph = False
    ee.fuse_ctx_quant = False
    assert eg.front_graph and eg.fuse_ctx_quant
    shapes = [(9, 40, 1), (9, 40, 2), (5, 33, 3), (9, 40, 4), (9, 40, 1)]
    for B, S, seed in shapes:
        ids, mask = W.synthetic_tokens(seed, B, S, 211, min_len=3)
        idt, mt = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
        a, b = eg.greedy_decode(idt, mt), ee.greedy_decode(idt, mt)
        assert torch.equal(a, b), (B, S, seed)
        assert torch.equal(eg.encode(idt, mt), ee.encode(idt, mt))
    assert eg.front_replays == len(shapes) and ee.front_replays == 0
    # a start symbol other than the captured one rebuilds the graph
    ids, mask = W.synthetic_tokens(7, 9, 40, 211, min_len=3)
    idt, mt = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
    assert torch.equal(eg.greedy_decode(idt, mt, start_symbol=3), ee.greedy_decode(idt, mt, start_symbol=3))
