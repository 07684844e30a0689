"""CPU tests: the dialect-A graph builder reproduces the reference's naming contract (all 60 fault-target files) and op
histogram, and the oracle's node-by-node executor (restatement of onnx_optimized_inference.py) agrees with the oracle's
layer-level model and implements the fault formulas of SURVEY.md App. D."""
import copy
import json
import os

import numpy as np
import pytest

from onnx_transformer_b200 import faults
from onnx_transformer_b200 import graph as G
from onnx_transformer_b200 import weights as W
from oracle import executor as oe
from oracle import intexact as ox
from oracle import model as om

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def small():
    fw = W.init_float_weights(5, 67, 59, 2, randomize_norms=True)
    w = om.get_quantized(fw, None, 2)
    return w, G.build_encoder_graph(w, batch=2, n_layers=2), G.build_decoder_graph(w, batch=2, n_layers=2)


@pytest.fixture(scope="module")
def full_graphs():
    fw = W.init_float_weights(0, 31, 29, 6)
    return G.build_encoder_graph(fw, 1, 6), G.build_decoder_graph(fw, 1, 6)


def test_fault_targets_match_the_reference_files(full_graphs):
    enc, dec = full_graphs
    ref = json.load(open(os.path.join(GOLD, "fault_targets.json")))
    mine_e = {t["target_layer"]: t for t in faults.targets_from_graph(enc, "Encoder")}
    mine_d = {t["target_layer"]: t for t in faults.targets_from_graph(dec, "Decoder")}
    assert len(ref["encoder"]) == 24 and len(ref["decoder"]) == 36
    for r in ref["encoder"]:
        assert mine_e[r["target_layer"]] == r
    for r in ref["decoder"]:
        assert mine_d[r["target_layer"]] == r
    assert len(mine_e) == 24 and len(mine_d) == 36


def test_op_histogram_matches_the_exported_graph(full_graphs):
    """encoder layer 169 / decoder layer 291 non-constant nodes before qonnx folds the static
    Shape/Gather/ReduceProd/Cast/Sub of every norm (5 nodes x 2 resp. 3 norms); final norm 17 - 4... = 13 cleaned."""
    enc, dec = full_graphs
    he, hd = enc.op_histogram(), dec.op_histogram()
    per_layer_e = {"Div": 32, "Mul": 20, "Round": 14, "Abs": 13, "ReduceMax": 13, "Clip": 13, "Add": 12, "Transpose": 10, "MatMul": 8,
                   "ReduceMean": 6, "Sub": 4, "Cast": 3, "Reshape": 4, "Sqrt": 2, "Softmax": 1, "Where": 1, "Equal": 1, "Relu": 1, "Unsqueeze": 1}
    norm = {"ReduceMean": 3, "Sub": 2, "Mul": 3, "Div": 2, "Sqrt": 1, "Add": 2}
    for op, n in per_layer_e.items():
        assert he[op] == 6 * n + norm.get(op, 0), op
    assert sum(per_layer_e.values()) == 169 - 10
    # the exported decoder layer quantizes `memory` once per layer (7 nodes); cleanup keeps a single shared copy (Round_60)
    per_layer_d = {"Div": 54, "Mul": 33, "Round": 24, "Abs": 22, "ReduceMax": 22, "Clip": 22, "Add": 19, "Transpose": 18, "MatMul": 14,
                   "ReduceMean": 9, "Sub": 6, "Cast": 6, "Reshape": 8, "Sqrt": 3, "Softmax": 2, "Where": 2, "Equal": 2, "Unsqueeze": 2, "Relu": 1}
    assert sum(per_layer_d.values()) == 291 - 15 - 7
    extra = {"Abs": 1, "ReduceMax": 1, "Clip": 1, "Div": 2, "Round": 1, "Mul": 1}   # the shared memory quantization (Round_60)
    for op, n in per_layer_d.items():
        assert hd[op] == 6 * n + norm.get(op, 0) + extra.get(op, 0), op
    assert he["MatMul"] == 48 and he["Round"] == 84 and hd["MatMul"] == 84 and hd["Round"] == 145
    assert [v.name for v in enc.input] == ["global_in", "global_in_1"] and enc.output[0].name == "global_out"
    assert [v.name for v in dec.input] == ["global_in", "global_in_1", "global_in_2", "global_in_3"]


def test_traces_have_the_reference_shape(full_graphs):
    """error.log:7-9: input trace ['Cast_10','Div_143','MatMul_28'], weight trace ['Mul_84','Reshape_14','Transpose_49','MatMul_28']."""
    _, dec = full_graphs
    t = [x for x in faults.targets_from_graph(dec, "Decoder") if x["target_layer"] == "MatMul_28"][0]
    (iq, it), (wq, wt), _, (itr, wtr) = faults.get_target_inputs(dec, t["target_layer"], t["input_tensor"], t["weight_tensor"], None, t["output_tensor"])
    assert [n.split("_")[0] for n in itr] == ["Cast", "Div", "MatMul"] and itr[-1] == "MatMul_28" and iq == itr[0]
    assert [n.split("_")[0] for n in wtr] == ["Mul", "Reshape", "Transpose", "MatMul"] and wq == wtr[0]
    assert it == "Round_89_out0" and wt == "Round_88_out0"


def _enc_inputs(w, B=2, S=5, seed=11):
    ids, mask = W.synthetic_tokens(seed, B, S, 67, min_len=3)
    pe = ox.positional_encoding(32)
    return ox.embed(ids, w["src_embed.0.lut.weight"], pe), mask


@pytest.mark.parametrize("mode", ["ref-float", "int-exact"])
def test_oracle_executor_equals_oracle_model(small, mode):
    w, enc, dec = small
    x, mask = _enc_inputs(w)
    wd, g = oe.prepare_inference(enc, {"global_in": x, "global_in_1": mask})
    out, wd = oe.run_module("Encoder", {"global_in": x, "global_in_1": mask}, None, wd, g, None, mode)
    assert list(out.keys()) == ["global_out"]
    ref = om.encode(w, x, mask, mode, 2)
    np.testing.assert_allclose(out["global_out"], ref, rtol=1e-4, atol=2e-4)
    # every intermediate is retained under its ONNX tensor name (drop-in contract)
    assert "Round_36_out0" in wd and "MatMul_3_out0" in wd and wd["Round_36_out0"].shape == (2, 5, 512)
    # decoder graph, 3 target positions
    T = 3
    ys = np.array([[0, 7, 9], [0, 4, 33]])
    temb = ox.embed(ys, w["tgt_embed.0.lut.weight"], ox.positional_encoding(32))
    sub = (np.triu(np.ones((1, T, T)), k=1) == 0).astype(np.int64)
    ins = {"global_in": temb, "global_in_1": ref, "global_in_2": mask, "global_in_3": sub}
    wd, g = oe.prepare_inference(dec, ins)
    out, wd = oe.run_module("Decoder", ins, None, wd, g, None, mode)
    refd = om.decode(w, temb, ref, mask, mode, 2)
    np.testing.assert_allclose(out["global_out"], refd, rtol=1e-4, atol=5e-4)


def _run_fault(enc, w, target_name, fault_model, bit, draws):
    x, mask = _enc_inputs(w)
    ins = {"global_in": x, "global_in_1": mask}
    wd, g = oe.prepare_inference(enc, ins)
    gold, wdg = oe.run_module("Encoder", ins, None, wd, g, None)
    gold_wd = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in wdg.items()}
    t = [d for d in faults.targets_from_graph(enc, "Encoder") if d["target_layer"] == target_name][0]
    p = faults.build_inject_parameters(enc, t, fault_model, bit, rng_draws=draws)
    wd2, g = oe.prepare_inference(enc, ins)
    out, wd2 = oe.run_module("Encoder", ins, None, wd2, g, p)
    return t, p, gold_wd, wd2


def test_input_fault_is_the_rank1_row_update_of_app_d(small):
    w, enc, _ = small
    # FirstFC of layer 0 (MatMul_6): INPUT fault at (b,t,k)
    idx = [1, 2, 77]
    t, p, gold, wd = _run_fault(enc, w, "MatMul_6", "INPUT", 6, {"target_indices": idx})
    assert p["faulty_trace"] == [] and p["rng_draws_used"]["target_indices"] == idx
    q = int(gold[t["input_tensor"]][tuple(idx)])
    s = gold["Div_%s_out0" % "0"] if False else None
    del s
    # dhat = fl(fl(q'*s) - fl(q*s)) with s the row scale the quantizer Mul uses; out[row, :] = golden + fl(dhat * What[k, :])
    quant_node = enc.node_by_name(p["faulty_quantizer_name"])
    scale_name = [i for i in quant_node.input if i != t["input_tensor"]][0]
    s_row = gold[scale_name][idx[0], idx[1], 0]
    qf = ox.flip_int8_bit(q, 6)
    dhat = np.float32(np.float32(qf * s_row) - np.float32(q * s_row))
    mm = enc.node_by_name("MatMul_6")
    what = gold[mm.input[1]]                      # [K, N] de-quantized, transposed weight
    expect = gold["MatMul_6_out0"].copy()
    expect[idx[0], idx[1], :] = (expect[idx[0], idx[1], :] + (dhat * what[idx[2], :]).astype(np.float32)).astype(np.float32)
    np.testing.assert_allclose(wd["MatMul_6_out0"], expect, rtol=0, atol=1e-6)
    assert np.count_nonzero(wd["MatMul_6_out0"] != gold["MatMul_6_out0"]) <= 2048
    assert not np.array_equal(wd["global_out"], gold["global_out"])


def test_weight16_and_random_faults(small):
    w, enc, _ = small
    # SecondMatMul of layer 1 (MatMul_12): WEIGHT16 on the V operand, 4-D output [B,8,S,64]: rows [0, L) of dim 2 (S=5 < 16)
    t, p, gold, wd = _run_fault(enc, w, "MatMul_12", "WEIGHT16", 3, {"target_indices": [0, 3, 130], "window_len": 2})
    diff = wd["MatMul_12_out0"] != gold["MatMul_12_out0"]
    where = np.argwhere(diff)
    assert where.size > 0 and set(where[:, 0]) == {0} and set(where[:, 1]) == {130 // 64} and set(where[:, 3]) == {130 % 64}
    assert set(where[:, 2]) <= {0, 1}
    # RANDOM_BITFLIP on MatMul_3's output element, bit 30
    t, p, gold, wd = _run_fault(enc, w, "MatMul_3", "RANDOM_BITFLIP", None, {"target_indices": [1, 2, 3, 4], "flip_bit": 30})
    g0 = gold["MatMul_3_out0"][1, 2, 3, 4]
    assert wd["MatMul_3_out0"][1, 2, 3, 4] == ox.float32_bit_flip(g0, 30)
    assert np.count_nonzero(wd["MatMul_3_out0"] != gold["MatMul_3_out0"]) == 1
    # RANDOM: replacement by an explicit 32-bit pattern
    t, p, gold, wd = _run_fault(enc, w, "MatMul_7", "RANDOM", None, {"target_indices": [0, 1, 9], "random_bits": 0x41200000})
    assert wd["MatMul_7_out0"][0, 1, 9] == np.float32(10.0)


def test_host_bit_helpers_match_oracle():
    for v in (-128, -5, 0, 5, 127):
        for b in range(8):
            assert faults.flip_int8_bit(v, b) == ox.flip_int8_bit(v, b)
    for v in (1.0, -3.5, 1e-30, 65504.0):
        for b in (0, 7, 22, 23, 30, 31):
            assert np.float32(faults.float32_bit_flip_value(v, b)) == ox.float32_bit_flip(np.float32(v), b)
    assert faults.delta_init_value(0x7FC00001) == 0.0 and faults.delta_init_value(0x3F800000) == 1.0
