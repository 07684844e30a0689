"""Duck-typed ONNX graph objects and the dialect-A graph builder.

The reference walks `onnx.GraphProto` objects loaded from `./try/{encoder,decoder}_try_cleaned.onnx`
(onnx_optimized_inference.py:273-295); neither the `onnx` package nor those files exist here
(.MISSING_LARGE_BLOBS), so this module provides

  * `Node / Attribute / ValueInfo / Initializer / Graph`: plain Python objects exposing the fields the executor
    reads (`.node[i].name/.op_type/.input/.output/.attribute`, `.input/.output/.value_info/.initializer`), and
  * `build_encoder_graph / build_decoder_graph`: emit the op-for-op fake-quant graph the reference's exporter
    produces for `get_quantized(make_model(...))` (output.py:607-614, quant_linear.py, attention.py,
    layer_norm.py; op chains per SURVEY.md 8a a6-a19 and App. B histogram: 169 nodes per encoder layer, 291 per
    decoder layer, 13 for the final norm), with qonnx-cleanup style names (`<OpType>_<k>`, `<node>_out0`,
    `global_in[_k]`, `global_out`).  `MatMul_k` and `Round_k` follow the numbering contract of the 60 fault-target
    files input/{encoder,decoder}/matmul_*.json (verified by tests/test_graph.py); the numbering of the other op
    types is this builder's own (qonnx's topological sort is not reproducible here) -- traces are always derived
    from the graph itself (faults.get_target_inputs), so they are self-consistent.

Like the reference's export, a graph is built for a fixed batch size (attention.py:54 bakes `nbatches` into the
Reshape constants); unlike it, any batch can be requested.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

D_MODEL, N_HEADS, D_K, D_FF = 512, 8, 64, 2048


@dataclass
class Attribute:
    name: str
    f: Optional[float] = None
    i: Optional[int] = None
    ints: Optional[List[int]] = None


@dataclass
class Node:
    name: str
    op_type: str
    input: List[str]
    output: List[str]
    attribute: List[Attribute] = field(default_factory=list)

    def attr(self, name, default=None):
        for a in self.attribute:
            if a.name == name:
                if a.ints is not None:
                    return a.ints
                if a.i is not None:
                    return a.i
                return a.f
        return default


@dataclass
class ValueInfo:
    name: str
    shape: Sequence[object]
    dtype: str = "float32"


@dataclass
class Initializer:
    name: str
    array: np.ndarray


@dataclass
class Graph:
    name: str
    node: List[Node] = field(default_factory=list)
    input: List[ValueInfo] = field(default_factory=list)
    output: List[ValueInfo] = field(default_factory=list)
    value_info: List[ValueInfo] = field(default_factory=list)
    initializer: List[Initializer] = field(default_factory=list)
    # builder metadata (not part of GraphProto): role of every MatMul / Round, used by the fused engine and the tests
    roles: Dict[str, dict] = field(default_factory=dict)

    def node_by_name(self, name: str) -> Node:
        for n in self.node:
            if n.name == name:
                return n
        raise KeyError(name)

    def op_histogram(self) -> Dict[str, int]:
        h: Dict[str, int] = {}
        for n in self.node:
            h[n.op_type] = h.get(n.op_type, 0) + 1
        return h


# ---------------------------------------------------------------------------------------------- naming contract
def encoder_matmul_name(layer: int, role: str) -> str:
    """input/encoder/matmul_*.json: MatMul_{8l + k}, k = q,k,v,qk,pv,o,ffn1,ffn2."""
    return "MatMul_%d" % (8 * layer + ["q", "k", "v", "qk", "pv", "o", "ffn1", "ffn2"].index(role))


def encoder_round_name(layer: int, role: str) -> str:
    """Weight rounds Round_{6l + (wq,wk,wv,wo,w1,w2)}; activation rounds Round_{36 + 8l + (x,q,k,v,p,o_in,ffn1_in,ffn2_in)}."""
    w = ["wq", "wk", "wv", "wo", "w1", "w2"]
    a = ["x", "q", "k", "v", "p", "o_in", "ffn1_in", "ffn2_in"]
    if role in w:
        return "Round_%d" % (6 * layer + w.index(role))
    return "Round_%d" % (36 + 8 * layer + a.index(role))


def decoder_matmul_name(layer: int, role: str) -> str:
    """input/decoder/matmul_*.json: MatMul_{0..11} = cross K,V of layers 0..5 (hoisted); layer l:
    MatMul_{12 + 12l + k}, k = q,k,v,qk,pv,o, cq,cqk,cpv,co, ffn1,ffn2."""
    if role == "ck":
        return "MatMul_%d" % (2 * layer)
    if role == "cv":
        return "MatMul_%d" % (2 * layer + 1)
    order = ["q", "k", "v", "qk", "pv", "o", "cq", "cqk", "cpv", "co", "ffn1", "ffn2"]
    return "MatMul_%d" % (12 + 12 * layer + order.index(role))


def decoder_round_name(layer: int, role: str) -> str:
    """Weight rounds Round_{10l + (wq,wk,wv,wo,cwq,cwk,cwv,cwo,w1,w2)}; Round_60 = memory; Round_{61+2l}/{62+2l} =
    cross K / V outputs; activation rounds Round_{73 + 12l + ...} = x,q,k,v,p,o_in,x2,cq,cp,co_in,ffn1_in,ffn2_in with
    the (v,p) pair swapped in layer 0 (input/decoder/matmul_16.json vs matmul_28.json)."""
    w = ["wq", "wk", "wv", "wo", "cwq", "cwk", "cwv", "cwo", "w1", "w2"]
    if role in w:
        return "Round_%d" % (10 * layer + w.index(role))
    if role == "memory":
        return "Round_60"
    if role == "ck":
        return "Round_%d" % (61 + 2 * layer)
    if role == "cv":
        return "Round_%d" % (62 + 2 * layer)
    a = ["x", "q", "k", "v", "p", "o_in", "x2", "cq", "cp", "co_in", "ffn1_in", "ffn2_in"]
    if layer == 0:
        a = ["x", "q", "k", "p", "v", "o_in", "x2", "cq", "cp", "co_in", "ffn1_in", "ffn2_in"]
    return "Round_%d" % (73 + 12 * layer + a.index(role))


# ---------------------------------------------------------------------------------------------- builder
class _Builder:
    def __init__(self, name: str, module: str):
        self.g = Graph(name)
        self.module = module
        self.counters: Dict[str, int] = {}
        self.forced: Dict[str, str] = {}

    def const(self, node_name: str, idx: int, value, dtype=np.float32) -> str:
        name = "%s_param%d" % (node_name, idx)
        self.g.initializer.append(Initializer(name, np.asarray(value, dtype=dtype)))
        return name

    def init(self, name: str, array: np.ndarray) -> str:
        self.g.initializer.append(Initializer(name, np.ascontiguousarray(array)))
        return name

    def next_name(self, op_type: str, forced: Optional[str] = None) -> str:
        if forced is not None:
            return forced
        k = self.counters.get(op_type, 0)
        self.counters[op_type] = k + 1
        return "%s_%d" % (op_type, k)

    def op(self, op_type: str, inputs, attrs=None, name: Optional[str] = None, consts=None, role: Optional[dict] = None) -> str:
        """Emit one node; `consts` = {input position: scalar/array} become `<node>_param<i>` initializers."""
        node_name = self.next_name(op_type, name)
        inputs = list(inputs)
        if consts:
            for pos, value in sorted(consts.items()):
                if isinstance(value, np.ndarray):
                    dtype = value.dtype
                elif isinstance(value, (np.integer, int)) or (isinstance(value, (list, tuple)) and all(isinstance(v, int) for v in value)):
                    dtype = np.int64
                else:
                    dtype = np.float32
                cname = self.const(node_name, pos, value, dtype)
                while len(inputs) <= pos:
                    inputs.append("")
                inputs[pos] = cname
        out = node_name + "_out0"
        self.g.node.append(Node(node_name, op_type, inputs, [out], list(attrs or [])))
        self.g.value_info.append(ValueInfo(out, ()))
        if role is not None:
            self.g.roles[node_name] = role
        return out

    # ---- op chains -------------------------------------------------------------------------------------
    def row_quant(self, x: str, round_name: str, role: dict):
        """quant_linear.py:6-17 / :31-43 as exported: Abs, ReduceMax(-1, keepdims), Clip(min 1e-5, max ""),
        Div(127), Div, Round, Mul.  Returns (x_hat, round_out, scale)."""
        a = self.op("Abs", [x])
        m = self.op("ReduceMax", [a], [Attribute("axes", ints=[-1]), Attribute("keepdims", i=1)])
        c = self.op("Clip", [m, "", ""], consts={1: np.float32(1e-5)})   # empty 3rd input, as exported (App. B)
        s = self.op("Div", [c], consts={1: np.float32(127.0)})
        d = self.op("Div", [x, s])
        r = self.op("Round", [d], name=round_name, role=role)
        xhat = self.op("Mul", [r, s])
        return xhat, r, s

    def layer_norm(self, x: str, a_2: np.ndarray, b_2: np.ndarray, n: int = D_MODEL) -> str:
        """layer_norm.py:12-15 as exported and cleaned (static Shape/Gather/ReduceProd/Cast/Sub folded): 13 nodes."""
        red = [Attribute("axes", ints=[-1]), Attribute("keepdims", i=1)]
        mean = self.op("ReduceMean", [x], red)
        mean2 = self.op("ReduceMean", [x], red)
        d2 = self.op("Sub", [x, mean2])
        sq = self.op("Mul", [d2, d2])
        var = self.op("ReduceMean", [sq], red)
        var = self.op("Mul", [var], consts={1: np.float32(n)})
        var = self.op("Div", [var], consts={1: np.float32(n - 1)})
        std = self.op("Sqrt", [var])
        d = self.op("Sub", [x, mean])
        nm = self.next_name("Mul")
        ad = self.op("Mul", [self.init(nm + "_param0", a_2.astype(np.float32)), d], name=nm)
        den = self.op("Add", [std], consts={1: np.float32(1e-6)})
        y = self.op("Div", [ad, den])
        return self.op("Add", [y], consts={1: b_2.astype(np.float32)})

    def weight_chain(self, w: np.ndarray, round_name: str, role: dict) -> str:
        """Weight fake-quant + Transpose(1,0): recomputed at run time in the reference (quant_linear.py:114-116)."""
        nm = self.next_name("Abs")
        wname = self.init(nm + "_param0", w.astype(np.float32))
        a = self.op("Abs", [wname], name=nm)
        m = self.op("ReduceMax", [a], [Attribute("axes", ints=[-1]), Attribute("keepdims", i=1)])
        c = self.op("Clip", [m, "", ""], consts={1: np.float32(1e-5)})
        s = self.op("Div", [c], consts={1: np.float32(127.0)})
        d = self.op("Div", [wname, s])
        r = self.op("Round", [d], name=round_name, role=role)
        what = self.op("Mul", [r, s])
        return self.op("Transpose", [what], [Attribute("perm", ints=[1, 0])])

    def linear(self, xhat: str, wt: str, bias: np.ndarray, matmul_name: str, role: dict) -> str:
        mm = self.op("MatMul", [xhat, wt], name=matmul_name, role=role)
        nm = self.next_name("Add")
        return self.op("Add", [self.init(nm + "_param0", bias.astype(np.float32)), mm], name=nm)

    def split_heads(self, x: str, batch: int, key: bool = False) -> str:
        r = self.op("Reshape", [x], consts={1: [batch, -1, N_HEADS, D_K]})
        return self.op("Transpose", [r], [Attribute("perm", ints=[0, 2, 3, 1] if key else [0, 2, 1, 3])])

    def attention_core(self, q_hat: str, k_hat: str, v_hat: str, mask: str, batch: int, names: dict, roles: dict) -> str:
        """attention.py:23-36 + :49-66 as exported."""
        qh = self.split_heads(q_hat, batch)
        kh = self.split_heads(k_hat, batch, key=True)
        vh = self.split_heads(v_hat, batch)
        scores = self.op("MatMul", [qh, kh], name=names["qk"], role=roles["qk"])
        scores = self.op("Div", [scores], consts={1: np.float32(8.0)})
        mu = self.op("Unsqueeze", [mask], [Attribute("axes", ints=[1])])
        mc = self.op("Cast", [mu], [Attribute("to", i=7)])       # int64
        me = self.op("Equal", [mc], consts={1: np.int64(0)})
        mb = self.op("Cast", [me], [Attribute("to", i=9)])       # bool
        scores = self.op("Where", [mb, "", scores], consts={1: np.float32(-1e9)})
        p = self.op("Softmax", [scores], [Attribute("axis", i=-1)])
        p127 = self.op("Mul", [p], consts={1: np.float32(127.0)})
        pr = self.op("Round", [p127], name=names["p"], role=roles["p"])
        pc = self.op("Cast", [pr], [Attribute("to", i=1)])       # float32
        phat = self.op("Div", [pc], consts={1: np.float32(127.0)})
        ctx = self.op("MatMul", [phat, vh], name=names["pv"], role=roles["pv"])
        ctx = self.op("Transpose", [ctx], [Attribute("perm", ints=[0, 2, 1, 3])])
        return self.op("Reshape", [ctx], consts={1: [batch, -1, D_MODEL]})


def _w(weights, key):
    return np.asarray(weights[key], dtype=np.float32)


def build_encoder_graph(weights: Dict[str, np.ndarray], batch: int = 1, n_layers: int = 6) -> Graph:
    """Encoder: global_in f32 [B,S,512], global_in_1 bool [B,1,S] -> global_out f32 [B,S,512]
    (onnx_optimized_custom_inference.py:628).  `weights` uses the reference's state_dict names
    (encoder.layers.{l}.self_attn.linears.{i}.weight ...), already fake-quantized as W8A8Linear.from_float does."""
    b = _Builder("encoder", "Encoder")
    g = b.g
    g.input = [ValueInfo("global_in", (batch, "S", D_MODEL), "float32"), ValueInfo("global_in_1", (batch, 1, "S"), "bool")]
    rn, mn = encoder_round_name, encoder_matmul_name
    # phase 1: initializer-only chains (weight fake-quant) sort first under qonnx cleanup -> Round_0..35
    wt = {}
    for l in range(n_layers):
        p = "encoder.layers.%d." % l
        for role, key in [("wq", "self_attn.linears.0"), ("wk", "self_attn.linears.1"), ("wv", "self_attn.linears.2"),
                          ("wo", "self_attn.linears.3"), ("w1", "feed_forward.w_1"), ("w2", "feed_forward.w_2")]:
            wt[(l, role)] = b.weight_chain(_w(weights, p + key + ".weight"), rn(l, role), {"kind": "weight", "layer": l, "role": role})
    x = "global_in"
    for l in range(n_layers):
        p = "encoder.layers.%d." % l
        R = lambda role, kind="act": {"kind": kind, "layer": l, "role": role}  # noqa: E731
        ln = b.layer_norm(x, _w(weights, p + "sublayer.0.norm.a_2"), _w(weights, p + "sublayer.0.norm.b_2"))
        xh, _, _ = b.row_quant(ln, rn(l, "x"), R("x"))
        proj = {}
        for role, wrole, i in [("q", "wq", 0), ("k", "wk", 1), ("v", "wv", 2)]:
            y = b.linear(xh, wt[(l, wrole)], _w(weights, p + "self_attn.linears.%d.bias" % i), mn(l, role), R(role, "matmul"))
            proj[role], _, _ = b.row_quant(y, rn(l, role), R(role))
        ctx = b.attention_core(proj["q"], proj["k"], proj["v"], "global_in_1", batch,
                               {"qk": mn(l, "qk"), "p": rn(l, "p"), "pv": mn(l, "pv")},
                               {"qk": R("qk", "matmul"), "p": R("p"), "pv": R("pv", "matmul")})
        ch, _, _ = b.row_quant(ctx, rn(l, "o_in"), R("o_in"))
        o = b.linear(ch, wt[(l, "wo")], _w(weights, p + "self_attn.linears.3.bias"), mn(l, "o"), R("o", "matmul"))
        x = b.op("Add", [x, o])
        ln = b.layer_norm(x, _w(weights, p + "sublayer.1.norm.a_2"), _w(weights, p + "sublayer.1.norm.b_2"))
        fh, _, _ = b.row_quant(ln, rn(l, "ffn1_in"), R("ffn1_in"))
        h1 = b.linear(fh, wt[(l, "w1")], _w(weights, p + "feed_forward.w_1.bias"), mn(l, "ffn1"), R("ffn1", "matmul"))
        h1 = b.op("Relu", [h1])
        hh, _, _ = b.row_quant(h1, rn(l, "ffn2_in"), R("ffn2_in"))
        h2 = b.linear(hh, wt[(l, "w2")], _w(weights, p + "feed_forward.w_2.bias"), mn(l, "ffn2"), R("ffn2", "matmul"))
        x = b.op("Add", [x, h2])
    out = b.layer_norm(x, _w(weights, "encoder.norm.a_2"), _w(weights, "encoder.norm.b_2"))
    _rename_output(g, out, "global_out")
    g.output = [ValueInfo("global_out", (batch, "S", D_MODEL), "float32")]
    return g


def build_decoder_graph(weights: Dict[str, np.ndarray], batch: int = 1, n_layers: int = 6) -> Graph:
    """Decoder: global_in f32 [B,T,512], global_in_1 memory f32 [B,S,512], global_in_2 bool [B,1,S],
    global_in_3 int64 [1,T,T] -> global_out f32 [B,T,512] (onnx_optimized_custom_inference.py:646-651)."""
    b = _Builder("decoder", "Decoder")
    g = b.g
    g.input = [ValueInfo("global_in", (batch, "T", D_MODEL), "float32"), ValueInfo("global_in_1", (batch, "S", D_MODEL), "float32"),
               ValueInfo("global_in_2", (batch, 1, "S"), "bool"), ValueInfo("global_in_3", (1, "T", "T"), "int64")]
    rn, mn = decoder_round_name, decoder_matmul_name
    wt = {}
    keys = [("wq", "self_attn.linears.0"), ("wk", "self_attn.linears.1"), ("wv", "self_attn.linears.2"), ("wo", "self_attn.linears.3"),
            ("cwq", "src_attn.linears.0"), ("cwk", "src_attn.linears.1"), ("cwv", "src_attn.linears.2"), ("cwo", "src_attn.linears.3"),
            ("w1", "feed_forward.w_1"), ("w2", "feed_forward.w_2")]
    for l in range(n_layers):
        p = "decoder.layers.%d." % l
        for role, key in keys:
            wt[(l, role)] = b.weight_chain(_w(weights, p + key + ".weight"), rn(l, role), {"kind": "weight", "layer": l, "role": role})
    # phase 2: memory-only chains (cross-attention K/V of every layer) hoist to the graph front: MatMul_0..11
    mh, _, _ = b.row_quant("global_in_1", rn(0, "memory"), {"kind": "act", "layer": -1, "role": "memory"})
    cross = {}
    for l in range(n_layers):
        p = "decoder.layers.%d." % l
        for role, wrole, i in [("ck", "cwk", 1), ("cv", "cwv", 2)]:
            y = b.linear(mh, wt[(l, wrole)], _w(weights, p + "src_attn.linears.%d.bias" % i), mn(l, role),
                         {"kind": "matmul", "layer": l, "role": role})
            cross[(l, role)], _, _ = b.row_quant(y, rn(l, role), {"kind": "act", "layer": l, "role": role})
    x = "global_in"
    for l in range(n_layers):
        p = "decoder.layers.%d." % l
        R = lambda role, kind="act": {"kind": kind, "layer": l, "role": role}  # noqa: E731
        ln = b.layer_norm(x, _w(weights, p + "sublayer.0.norm.a_2"), _w(weights, p + "sublayer.0.norm.b_2"))
        xh, _, _ = b.row_quant(ln, rn(l, "x"), R("x"))
        proj = {}
        for role, wrole, i in [("q", "wq", 0), ("k", "wk", 1), ("v", "wv", 2)]:
            y = b.linear(xh, wt[(l, wrole)], _w(weights, p + "self_attn.linears.%d.bias" % i), mn(l, role), R(role, "matmul"))
            proj[role], _, _ = b.row_quant(y, rn(l, role), R(role))
        ctx = b.attention_core(proj["q"], proj["k"], proj["v"], "global_in_3", batch,
                               {"qk": mn(l, "qk"), "p": rn(l, "p"), "pv": mn(l, "pv")},
                               {"qk": R("qk", "matmul"), "p": R("p"), "pv": R("pv", "matmul")})
        ch, _, _ = b.row_quant(ctx, rn(l, "o_in"), R("o_in"))
        o = b.linear(ch, wt[(l, "wo")], _w(weights, p + "self_attn.linears.3.bias"), mn(l, "o"), R("o", "matmul"))
        x = b.op("Add", [x, o])
        ln = b.layer_norm(x, _w(weights, p + "sublayer.1.norm.a_2"), _w(weights, p + "sublayer.1.norm.b_2"))
        xh2, _, _ = b.row_quant(ln, rn(l, "x2"), R("x2"))
        y = b.linear(xh2, wt[(l, "cwq")], _w(weights, p + "src_attn.linears.0.bias"), mn(l, "cq"), R("cq", "matmul"))
        cq, _, _ = b.row_quant(y, rn(l, "cq"), R("cq"))
        ctx = b.attention_core(cq, cross[(l, "ck")], cross[(l, "cv")], "global_in_2", batch,
                               {"qk": mn(l, "cqk"), "p": rn(l, "cp"), "pv": mn(l, "cpv")},
                               {"qk": R("cqk", "matmul"), "p": R("cp"), "pv": R("cpv", "matmul")})
        ch, _, _ = b.row_quant(ctx, rn(l, "co_in"), R("co_in"))
        o = b.linear(ch, wt[(l, "cwo")], _w(weights, p + "src_attn.linears.3.bias"), mn(l, "co"), R("co", "matmul"))
        x = b.op("Add", [x, o])
        ln = b.layer_norm(x, _w(weights, p + "sublayer.2.norm.a_2"), _w(weights, p + "sublayer.2.norm.b_2"))
        fh, _, _ = b.row_quant(ln, rn(l, "ffn1_in"), R("ffn1_in"))
        h1 = b.linear(fh, wt[(l, "w1")], _w(weights, p + "feed_forward.w_1.bias"), mn(l, "ffn1"), R("ffn1", "matmul"))
        h1 = b.op("Relu", [h1])
        hh, _, _ = b.row_quant(h1, rn(l, "ffn2_in"), R("ffn2_in"))
        h2 = b.linear(hh, wt[(l, "w2")], _w(weights, p + "feed_forward.w_2.bias"), mn(l, "ffn2"), R("ffn2", "matmul"))
        x = b.op("Add", [x, h2])
    out = b.layer_norm(x, _w(weights, "decoder.norm.a_2"), _w(weights, "decoder.norm.b_2"))
    _rename_output(g, out, "global_out")
    g.output = [ValueInfo("global_out", (batch, "T", D_MODEL), "float32")]
    return g


def _rename_output(g: Graph, old: str, new: str) -> None:
    for n in g.node:
        n.output = [new if o == old else o for o in n.output]
        n.input = [new if i == old else i for i in n.input]
    g.value_info = [v for v in g.value_info if v.name != old]


# ---------------------------------------------------------------------------------------------- dialect B (Brevitas QCDQ)
def build_qcdq_block_graph(w1: np.ndarray, w2: np.ndarray, wk: np.ndarray, n_tokens: int, bit_width: int = 8, seed: int = 0) -> Graph:
    """A dialect-B ("Brevitas QCDQ") block in the shape `export_onnx_qcdq` gives the reference's
    quantized_position_feed_forward.py:23-41 and the score product of quantized_attention.py:52-58: every matmul operand --
    weights included -- passes `QuantizeLinear(x, scale, zp=0:int8, axis) -> Clip(-2^(b-1), 2^(b-1)-1) [b < 8] -> DequantizeLinear`
    and the product is a float `MatMul`:

        x [1,T,d] --qcdq(per token)--> MatMul_0( . , qcdq(W1 [d,f], per column)) -> Relu -> qcdq(per token)
                  --> MatMul_1( . , qcdq(W2 [f,d], per column)) = y [1,T,d]
        q = qcdq(y, per token) ; k = Transpose(qcdq(MatMul_2(qcdq(x), qcdq(Wk)), per token), (0,2,1)) ; MatMul_3(q, k) = scores [1,T,T]

    Weights are bias-free `nn.Parameter(d_in, d_out)` used as `x @ W` (quantized_attention.py:27-30); scales are static (learned in
    the reference; here abs-max statistics of seeded activations) with the per-token shape (1,T,1) / per-channel shape (1,d_out) of
    quantized_attention.py:32-45.  The Transpose sits between K's de-quantizer and the score MatMul, which is the case the
    reference's `transposed_axes` handles (inject_utils/layers.py:176-181).  Node / tensor names follow the cleanup convention
    (`<OpType>_<k>`, `<node>_out0`, `global_in`, `global_out`)."""
    d, f = w1.shape
    assert w2.shape == (f, d) and wk.shape == (d, d)
    b = _Builder("qcdq_block", "encoder")
    g = b.g
    g.input = [ValueInfo("global_in", (1, n_tokens, d), "float32")]
    qmax = float(2 ** (bit_width - 1))
    rng = np.random.default_rng(seed)
    zp = np.zeros((), dtype=np.int8)

    def qcdq(x: str, scale: np.ndarray, axis: int) -> str:
        ql = b.op("QuantizeLinear", [x], [Attribute("axis", i=axis)], consts={1: scale.astype(np.float32), 2: zp})
        if bit_width < 8:
            ql = b.op("Clip", [ql], consts={1: np.array(-2 ** (bit_width - 1), dtype=np.int8), 2: np.array(2 ** (bit_width - 1) - 1, dtype=np.int8)})
        return b.op("DequantizeLinear", [ql], [Attribute("axis", i=axis)], consts={1: scale.astype(np.float32), 2: zp})

    def act_scale(cal: np.ndarray) -> np.ndarray:
        # static per-token scale = threshold / 2^(b-1); the threshold comes from abs-max statistics of a seeded calibration pass
        # (Brevitas learns it from percentile statistics), jittered so that some run-time values saturate
        return (np.maximum(np.abs(cal).reshape(n_tokens, -1).max(axis=1), 1e-5) * rng.uniform(0.7, 1.1, size=n_tokens) / qmax).astype(np.float32)

    def weight_q(w: np.ndarray, name: str) -> str:
        wname = b.init(name, w.astype(np.float32))
        s = (np.maximum(np.abs(w).max(axis=0), 1e-5) / qmax).astype(np.float32)          # per output channel (column of [d_in, d_out])
        return qcdq(wname, s, 1)

    x_cal = rng.normal(size=(1, n_tokens, d)).astype(np.float32)
    h_cal = np.maximum(x_cal @ w1, 0)
    y_cal = h_cal @ w2
    k_cal = x_cal @ wk
    xq = qcdq("global_in", act_scale(x_cal), 1)
    h = b.op("MatMul", [xq, weight_q(w1, "weights_1")], role={"kind": "matmul", "role": "ffn1"})
    h = b.op("Relu", [h])
    hq = qcdq(h, act_scale(h_cal), 1)
    y = b.op("MatMul", [hq, weight_q(w2, "weights_2")], role={"kind": "matmul", "role": "ffn2"})
    yq = qcdq(y, act_scale(y_cal), 1)
    kx = b.op("MatMul", [xq, weight_q(wk, "weights_key")], role={"kind": "matmul", "role": "k"})
    kq = qcdq(kx, act_scale(k_cal), 1)
    kt = b.op("Transpose", [kq], [Attribute("perm", ints=[0, 2, 1])])
    out = b.op("MatMul", [yq, kt], role={"kind": "matmul", "role": "qk"})
    _rename_output(g, out, "global_out")
    g.output = [ValueInfo("global_out", (1, n_tokens, n_tokens), "float32")]
    return g
