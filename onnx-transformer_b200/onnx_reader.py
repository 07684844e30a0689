"""Real `.onnx` import without the `onnx` package: a protobuf wire-format reader for the subset of ModelProto that the
reference's graphs use, producing the duck-typed graph object of the executor.

What the reference does with the `onnx` package on this path: `onnx.load` + `qonnx ... cleanup` give it a GraphProto whose
`.node / .input / .output / .value_info / .initializer` it walks (`onnx_optimized_inference.py:214-295`), and
`numpy_helper.to_array` turns initializers into the numpy `weight_dict` (`:273-280`).  Here:

  * `read_model(bytes)` decodes ModelProto.graph (field 7) -> GraphProto {node 1, name 2, initializer 5, input 11, output 12,
    value_info 13}; NodeProto {input 1, output 2, name 3, op_type 4, attribute 5}; AttributeProto {name 1, f 2, i 3, s 4, t 5,
    floats 7, ints 8}; TensorProto {dims 1, data_type 2, float_data 4, int32_data 5, int64_data 7, name 8, raw_data 9};
    ValueInfoProto {name 1, type 2 -> tensor_type 1 -> elem_type 1, shape 2 -> dim 1 -> dim_value 1 | dim_param 2}.
  * `cleanup(graph)` does the part of qonnx's cleanup the executor relies on: `Constant` nodes become initializers, nodes get
    unique names `<OpType>_<k>` in graph order when the exporter left them empty, `Identity` of an initializer is folded.

Field numbers are those of onnx.proto3 (IR version 3-9, opset 13 files of the TorchScript exporter).
"""
from __future__ import annotations

import gzip
import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

from .graph import Attribute, Graph, Initializer, Node, ValueInfo

# TensorProto.DataType -> numpy
_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 4: np.uint16, 5: np.int16, 6: np.int32, 7: np.int64, 9: np.bool_, 10: np.float16,
           11: np.float64, 12: np.uint32, 13: np.uint64}
_DTYPE_NAMES = {1: "float32", 2: "uint8", 3: "int8", 6: "int32", 7: "int64", 9: "bool", 10: "float16", 11: "float64"}


class OnnxFormatError(ValueError):
    pass


# ---------------------------------------------------------------------------------------------- wire format
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result, shift = 0, 0
    while True:
        if pos >= len(buf):
            raise OnnxFormatError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise OnnxFormatError("varint too long")


def _fields(buf: bytes):
    """Yields (field number, wire type, value): value is an int (varint / fixed) or a memoryview slice (length-delimited)."""
    pos, n = 0, len(buf)
    view = memoryview(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            if pos + ln > n:
                raise OnnxFormatError("length-delimited field runs past the end of its message")
            val = bytes(view[pos:pos + ln])
            pos += ln
        elif wt == 5:
            val = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise OnnxFormatError("unsupported wire type %d" % wt)
        yield field, wt, val


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_varints(val, wt) -> List[int]:
    if wt == 0:
        return [_signed64(val)]
    out, pos = [], 0
    while pos < len(val):
        v, pos = _varint(val, pos)
        out.append(_signed64(v))
    return out


# ---------------------------------------------------------------------------------------------- messages
def _tensor(buf: bytes) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    dtype, name, raw = 1, "", None
    floats: List[float] = []
    ints: List[int] = []
    for f, wt, v in _fields(buf):
        if f == 1:
            dims += _packed_varints(v, wt)
        elif f == 2:
            dtype = v
        elif f == 4:      # float_data (packed fixed32 or single)
            floats += list(struct.unpack("<%df" % (len(v) // 4), v)) if wt == 2 else [struct.unpack("<f", struct.pack("<I", v))[0]]
        elif f in (5, 7):  # int32_data / int64_data
            ints += _packed_varints(v, wt)
        elif f == 8:
            name = v.decode()
        elif f == 9:
            raw = v
    if dtype not in _DTYPES:
        raise OnnxFormatError("tensor %r: unsupported data_type %d" % (name, dtype))
    np_dtype = _DTYPES[dtype]
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np.dtype(np_dtype).newbyteorder("<")).astype(np_dtype)
    elif floats:
        arr = np.asarray(floats, dtype=np_dtype)
    else:
        arr = np.asarray(ints, dtype=np_dtype)
    n = int(np.prod(dims)) if dims else 1
    if arr.size != n:
        raise OnnxFormatError("tensor %r: %d elements for dims %r" % (name, arr.size, dims))
    return name, arr.reshape(dims).copy()


def _attribute(buf: bytes):
    name, f_, i_, ints, floats, t, s = "", None, None, None, None, None, None
    for f, wt, v in _fields(buf):
        if f == 1:
            name = v.decode()
        elif f == 2:
            f_ = struct.unpack("<f", struct.pack("<I", v))[0]
        elif f == 3:
            i_ = _signed64(v)
        elif f == 4:
            s = v
        elif f == 5:
            t = _tensor(v)[1]
        elif f == 7:
            floats = (floats or []) + (list(struct.unpack("<%df" % (len(v) // 4), v)) if wt == 2 else [struct.unpack("<f", struct.pack("<I", v))[0]])
        elif f == 8:
            ints = (ints or []) + _packed_varints(v, wt)
    return name, f_, i_, ints, floats, t, s


def _value_info(buf: bytes) -> ValueInfo:
    name, shape, dtype = "", [], "float32"
    for f, _, v in _fields(buf):
        if f == 1:
            name = v.decode()
        elif f == 2:                                   # TypeProto
            for f2, _, v2 in _fields(v):
                if f2 == 1:                            # tensor_type
                    for f3, _, v3 in _fields(v2):
                        if f3 == 1:
                            dtype = _DTYPE_NAMES.get(v3, "dtype%d" % v3)
                        elif f3 == 2:                  # TensorShapeProto
                            for f4, _, v4 in _fields(v3):
                                if f4 == 1:            # Dimension
                                    dim: object = None
                                    for f5, _, v5 in _fields(v4):
                                        if f5 == 1:
                                            dim = _signed64(v5)
                                        elif f5 == 2:
                                            dim = v5.decode()
                                    shape.append(dim)
    return ValueInfo(name, tuple(shape), dtype)


def _node(buf: bytes, constants: Dict[str, np.ndarray]) -> Node:
    ins, outs, name, op, attrs = [], [], "", "", []
    tensors: Dict[str, np.ndarray] = {}
    for f, _, v in _fields(buf):
        if f == 1:
            ins.append(v.decode())
        elif f == 2:
            outs.append(v.decode())
        elif f == 3:
            name = v.decode()
        elif f == 4:
            op = v.decode()
        elif f == 5:
            an, af, ai, aints, afloats, at, _ = _attribute(v)
            if at is not None:
                tensors[an] = at
            elif afloats is not None and aints is None and af is None and ai is None:
                tensors[an] = np.asarray(afloats, dtype=np.float32)
            attrs.append(Attribute(an, af, ai, aints))
    node = Node(name, op, ins, outs, attrs)
    if op == "Constant" and outs:
        if "value" in tensors:
            constants[outs[0]] = tensors["value"]
        else:     # value_float / value_int / value_ints / value_floats forms
            for a in attrs:
                if a.name == "value_float":
                    constants[outs[0]] = np.asarray(a.f, dtype=np.float32)
                elif a.name == "value_int":
                    constants[outs[0]] = np.asarray(a.i, dtype=np.int64)
                elif a.name == "value_ints":
                    constants[outs[0]] = np.asarray(a.ints, dtype=np.int64)
            if "value_floats" in tensors:
                constants[outs[0]] = tensors["value_floats"]
    return node


def model_graph_bytes(data: bytes) -> bytes:
    """The serialized GraphProto (ModelProto field 7) of a serialized ModelProto."""
    graph_buf = None
    for f, wt, v in _fields(data):
        if f == 7 and wt == 2:
            graph_buf = v
    if graph_buf is None:
        raise OnnxFormatError("no GraphProto (ModelProto field 7) in the file")
    return graph_buf


def read_model(data: bytes, name: Optional[str] = None) -> Graph:
    """ModelProto bytes -> Graph (nodes in file order, initializers as numpy arrays).  `Constant` nodes are kept in `.node`;
    their tensors are available to `cleanup`."""
    return read_graph(model_graph_bytes(data), name)


def read_graph(graph_buf: bytes, name: Optional[str] = None) -> Graph:
    """Serialized GraphProto -> Graph."""
    g = Graph(name or "onnx")
    constants: Dict[str, np.ndarray] = {}
    for f, _, v in _fields(graph_buf):
        if f == 1:
            g.node.append(_node(v, constants))
        elif f == 2 and name is None:
            g.name = v.decode()
        elif f == 5:
            n, arr = _tensor(v)
            g.initializer.append(Initializer(n, arr))
        elif f == 11:
            g.input.append(_value_info(v))
        elif f == 12:
            g.output.append(_value_info(v))
        elif f == 13:
            g.value_info.append(_value_info(v))
    g.roles["__constants__"] = constants      # consumed (and removed) by cleanup()
    return g


def cleanup(g: Graph) -> Graph:
    """The part of qonnx's cleanup the node-by-node executor relies on (the reference runs it before inference,
    parallelized_inject_onnx_transformer.py:413-444 / qonnx.util.cleanup): Constant nodes -> initializers; graph inputs that are
    initializers are dropped from `.input`; every node gets a unique non-empty name `<OpType>_<k>` (k counts per op type, in graph
    order, as qonnx's GiveUniqueNodeNames does)."""
    constants = g.roles.pop("__constants__", {})
    have = {i.name for i in g.initializer}
    for n, arr in constants.items():
        if n not in have:
            g.initializer.append(Initializer(n, arr))
            have.add(n)
    g.node = [n for n in g.node if n.op_type != "Constant"]
    g.input = [v for v in g.input if v.name not in have]
    counters: Dict[str, int] = {}
    for n in g.node:
        k = counters.get(n.op_type, 0)
        counters[n.op_type] = k + 1
        n.name = "%s_%d" % (n.op_type, k)
    return g


def load_onnx(path: str, clean: bool = True) -> Graph:
    """Read a `.onnx` (or `.onnx.gz`) file into the executor's graph object."""
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rb") as f:
        data = f.read()
    g = read_model(data)
    return cleanup(g) if clean else g


# ---------------------------------------------------------------------------------------------- weights/{encoder,decoder}.pt
class _GraphProtoStub:
    """Stand-in for onnx.onnx_ml_pb2.GraphProto while un-pickling: protobuf messages pickle as (class, state) with
    state = {"serialized": bytes}; the bytes are parsed by the wire reader above, so the `onnx` package is not needed."""

    def __init__(self, *args, **kwargs):
        self.graph = None

    def __setstate__(self, state):
        data = state["serialized"] if isinstance(state, dict) else state
        self.graph = cleanup(read_graph(bytes(data)))


def load_pt_archive(path: str):
    """The reference's `weights/encoder.pt` / `weights/decoder.pt` (parallelized_inject_onnx_transformer.py:540,621,772-774:
    `weight_dict, main_graph = torch.load(...)`, written by inject_operations.py:198 / onnx_optimized_inference.py with
    `torch.save((module_weight_dict, module_graph), ...)`): a pickled (dict name -> numpy array, onnx GraphProto) pair.  Also reads
    the archives this package writes (save_pt_archive).  Returns (weight_dict of numpy arrays, Graph)."""
    import pickle

    import torch

    class _Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            if module.startswith("onnx") and name == "GraphProto":
                return _GraphProtoStub
            return super().find_class(module, name)

    class _PickleModule:
        __name__ = "ot_pickle"
        Unpickler = _Unpickler
        load = staticmethod(lambda f, **kw: _Unpickler(f, **kw).load())

    obj = torch.load(path, map_location="cpu", pickle_module=_PickleModule, weights_only=False)
    if not (isinstance(obj, (tuple, list)) and len(obj) == 2):
        raise OnnxFormatError("%s: expected a (weight_dict, graph) pair" % path)
    a, b = obj
    if isinstance(a, dict):
        weight_dict, graph = a, b
    else:
        weight_dict, graph = b, a          # inject_main.py:405 unpacks `weight_dict, main_graph`, older scripts the other way round
    if isinstance(graph, _GraphProtoStub):
        graph = graph.graph
    elif isinstance(graph, dict) and "__ot_graph__" in graph:
        from .executor import _graph_from_dict
        graph = _graph_from_dict(graph["__ot_graph__"], graph["arrays"])
    if not isinstance(graph, Graph):
        raise OnnxFormatError("%s: the second element is not a graph (%s)" % (path, type(graph).__name__))
    out = {}
    for k, v in weight_dict.items():
        out[k] = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
    return out, graph
