"""Real ONNX import: tests/golden/ref_encoder_tiny.onnx.gz is a genuine opset-13 file written by torch's ONNX
exporter from the REFERENCE's own modules (model.make_model + get_quantized_model.quantize_transformer, 1 encoder layer, d_model
128; tests/golden/make_onnx_fixture.py), with the torch model's input / output beside it.  The protobuf wire reader must recover
the graph the survey describes, the oracle executor must reproduce the reference model's output on it (this pins the oracle's
node-by-node restatement against a reference artefact), and on the GPU the product executor must agree with both."""
import os

import numpy as np
import pytest

from onnx_transformer_b200 import onnx_reader as R
from oracle import executor as oex

HERE = os.path.dirname(os.path.abspath(__file__))
ONNX = os.path.join(HERE, "golden", "ref_encoder_tiny.onnx.gz")
IO = os.path.join(HERE, "golden", "ref_encoder_tiny_io.npz")

# non-constant nodes of one exported encoder layer and of the final norm
LAYER = {"Div": 32, "Mul": 20, "Round": 14, "Abs": 13, "ReduceMax": 13, "Clip": 13, "Add": 12, "Transpose": 10, "MatMul": 8, "ReduceMean": 6,
         "Sub": 6, "Cast": 5, "Reshape": 4, "Shape": 2, "Gather": 2, "ReduceProd": 2, "Sqrt": 2, "Softmax": 1, "Where": 1, "Equal": 1, "Relu": 1,
         "Unsqueeze": 1}
FINAL_NORM = {"ReduceMean": 3, "Sub": 3, "Mul": 3, "Div": 2, "Add": 2, "Cast": 1, "Shape": 1, "Gather": 1, "ReduceProd": 1, "Sqrt": 1}


def test_wire_reader_recovers_the_exported_graph():
    g = R.load_onnx(ONNX)
    hist = g.op_histogram()
    identity = hist.pop("Identity", 0)          # the exporter aliases equal (deduplicated) biases through Identity nodes
    assert identity == 3
    want = dict(LAYER)
    for k, v in FINAL_NORM.items():
        want[k] = want.get(k, 0) + v
    assert hist == want, {k: (hist.get(k), want.get(k)) for k in set(hist) | set(want) if hist.get(k) != want.get(k)}
    assert sum(LAYER.values()) == 169 and len(g.node) == 169 + sum(FINAL_NORM.values()) + 3
    assert [(v.name, tuple(v.shape), v.dtype) for v in g.input] == [("global_in", (2, 5, 128), "float32"), ("global_in_1", (2, 1, 5), "bool")]
    assert [v.name for v in g.output] == ["global_out"]
    inits = {i.name: i.array for i in g.initializer}
    assert inits["layers.0.self_attn.linears.1.weight"].shape == (128, 128) and inits["layers.0.feed_forward.w_1.bias"].dtype == np.float32
    names = [n.name for n in g.node]
    assert len(set(names)) == len(names) and "MatMul_7" in names and "Round_13" in names       # unique <OpType>_<k> names
    clip = g.node_by_name("Clip_0")
    assert clip.input[2] == ""                   # the empty third Clip input the reference patches around (onnx_optimized_inference.py:29-31)
    assert all(n.op_type != "Constant" for n in g.node)
    tr = [n for n in g.node if n.op_type == "Transpose"][0]
    assert tr.attr("perm") in ([1, 0], [0, 2, 1, 3], [0, 2, 3, 1])


def test_reader_rejects_garbage():
    with pytest.raises(R.OnnxFormatError):
        R.read_model(b"\x0a\xff\xff\xff\xff\x0f")          # a length-delimited field longer than the buffer
    with pytest.raises(R.OnnxFormatError):
        R.read_model(b"\x08\x01")                           # a ModelProto without a graph


@pytest.mark.parametrize("mode", ["ref-float", "int-exact"])
def test_oracle_executor_reproduces_the_reference_model_on_its_own_export(mode):
    g = R.load_onnx(ONNX)
    io = np.load(IO)
    feeds = {"global_in": io["x"], "global_in_1": io["mask"]}
    wd, graph = oex.prepare_inference(g, feeds)
    out, wd = oex.run_module("Encoder", feeds, None, wd, graph, mode=mode)
    y = out["global_out"]
    assert y.shape == io["y"].shape
    np.testing.assert_allclose(y, io["y"], rtol=0, atol=5e-6)      # |y| <= 3.3: float reductions only
    assert "/layers.0/sublayer.0/self_attn/Softmax_output_0" in wd    # every intermediate is retained under its ONNX tensor name


DEC_ONNX = os.path.join(HERE, "golden", "ref_decoder_tiny.onnx.gz")
DEC_IO = os.path.join(HERE, "golden", "ref_decoder_tiny_io.npz")
# non-constant nodes of one exported decoder layer
DEC_LAYER = {"Div": 56, "Mul": 34, "Round": 25, "Abs": 23, "ReduceMax": 23, "Clip": 23, "Add": 19, "Transpose": 18, "MatMul": 14, "ReduceMean": 9,
             "Sub": 9, "Cast": 9, "Reshape": 8, "Shape": 3, "Gather": 3, "ReduceProd": 3, "Sqrt": 3, "Softmax": 2, "Where": 2, "Equal": 2,
             "Unsqueeze": 2, "Relu": 1}


def _decoder_feeds():
    io = np.load(DEC_IO)
    return io, {"global_in": io["x"], "global_in_1": io["memory"], "global_in_2": io["src_mask"], "global_in_3": io["tgt_mask"]}


def test_decoder_export_matches_the_survey_and_the_reference_model():
    """The decoder of the same reference model (1 layer; inputs as onnx_optimized_custom_inference.py:646-651 feeds them: embedded
    prefix, memory, source key mask, causal mask): op histogram of SURVEY App. B (291 nodes per layer) and the torch output."""
    g = R.load_onnx(DEC_ONNX)
    hist = g.op_histogram()
    hist.pop("Identity", None)
    want = dict(DEC_LAYER)
    for k, v in FINAL_NORM.items():
        want[k] = want.get(k, 0) + v
    assert sum(DEC_LAYER.values()) == 291
    assert hist.get("Cast") in (9, 10)       # this torch version shares one mask Cast between the two attention blocks
    want["Cast"] = hist["Cast"]
    assert hist == want, {k: (hist.get(k), want.get(k)) for k in set(hist) | set(want) if hist.get(k) != want.get(k)}
    assert [(v.name, v.dtype) for v in g.input] == [("global_in", "float32"), ("global_in_1", "float32"), ("global_in_2", "bool"), ("global_in_3", "int64")]
    io, feeds = _decoder_feeds()
    for mode in ("ref-float", "int-exact"):
        wd, graph = oex.prepare_inference(R.load_onnx(DEC_ONNX), feeds)
        out, wd = oex.run_module("Decoder", feeds, None, wd, graph, mode=mode)
        np.testing.assert_allclose(out["global_out"], io["y"], rtol=0, atol=5e-6)


@pytest.mark.gpu
def test_product_executor_runs_the_real_decoder_onnx_file():
    from onnx_transformer_b200 import executor as ex
    io, feeds = _decoder_feeds()
    wd, graph = ex.prepare_inference(DEC_ONNX, feeds)
    out, wd = ex.run_module("Decoder", feeds, DEC_ONNX, wd, graph)
    y = ex.to_numpy(out)["global_out"]
    np.testing.assert_allclose(y, io["y"], rtol=0, atol=3e-3 * float(np.abs(io["y"]).max()))


@pytest.mark.gpu
def test_product_executor_runs_the_real_onnx_file():
    from onnx_transformer_b200 import executor as ex
    io = np.load(IO)
    feeds = {"global_in": io["x"], "global_in_1": io["mask"]}
    wd, graph = ex.prepare_inference(ONNX, feeds)
    out, wd = ex.run_module("Encoder", feeds, ONNX, wd, graph)
    y = ex.to_numpy(out)["global_out"]
    np.testing.assert_allclose(y, io["y"], rtol=0, atol=3e-3 * float(np.abs(io["y"]).max()))   # 1e-3 relative class
    # integer tensors against the oracle walk of the same graph: every Round output, bit for bit up to rounding-boundary flips
    wd_o, graph_o = oex.prepare_inference(R.load_onnx(ONNX), feeds)
    _, wd_o = oex.run_module("Encoder", feeds, None, wd_o, graph_o, mode="int-exact")
    rounds = [n.output[0] for n in graph_o.node if n.op_type == "Round"]
    assert len(rounds) == 14
    for name in rounds:
        got, want = ex.to_numpy({name: wd[name]})[name], wd_o[name]
        assert got.shape == want.shape
        assert np.mean(got != want) < 2e-3 and np.max(np.abs(got - want)) <= 1.0, name


def test_reference_pt_archive_without_the_onnx_package(tmp_path):
    """weights/{encoder,decoder}.pt of the reference = torch.save((weight_dict, GraphProto)) (parallelized_inject_onnx_transformer.py:
    540,621): the pickled protobuf message is (class onnx.onnx_ml_pb2.GraphProto, state {"serialized": bytes}); the loader swaps the
    class for a stub that parses the bytes with the wire reader.  The archive is produced here from the genuine ONNX export fixture by
    a fake `onnx.onnx_ml_pb2` module with exactly that pickle protocol."""
    import gzip
    import pickle
    import sys
    import types

    import torch
    from onnx_transformer_b200 import onnx_reader as R

    data = gzip.open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_encoder_tiny.onnx.gz"), "rb").read()
    graph_bytes = R.model_graph_bytes(data)
    want = R.cleanup(R.read_graph(graph_bytes))

    class GraphProto:                                     # pickles like a protobuf message: copyreg.__newobj__ + {"serialized": ...}
        def __init__(self):
            self.blob = b""

        def __getstate__(self):
            return {"serialized": self.blob}

        def __setstate__(self, state):
            self.blob = state["serialized"]

    mod_onnx, mod_pb = types.ModuleType("onnx"), types.ModuleType("onnx.onnx_ml_pb2")
    GraphProto.__module__, GraphProto.__qualname__ = "onnx.onnx_ml_pb2", "GraphProto"
    mod_pb.GraphProto = GraphProto
    sys.modules["onnx"], sys.modules["onnx.onnx_ml_pb2"] = mod_onnx, mod_pb
    try:
        msg = GraphProto()
        msg.blob = graph_bytes
        weight_dict = {i.name: i.array for i in want.initializer}
        path = str(tmp_path / "encoder.pt")
        torch.save((weight_dict, msg), path)
    finally:
        del sys.modules["onnx"], sys.modules["onnx.onnx_ml_pb2"]
    wd, g = R.load_pt_archive(path)                        # `onnx` is not importable here
    assert [n.name for n in g.node] == [n.name for n in want.node] and [n.op_type for n in g.node] == [n.op_type for n in want.node]
    assert set(wd) == {i.name for i in want.initializer}
    k = want.initializer[3].name
    assert np.array_equal(wd[k], want.initializer[3].array)
