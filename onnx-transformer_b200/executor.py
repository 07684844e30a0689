"""Drop-in for the reference's custom node-by-node ONNX executor, served by CUDA handlers.

Same entry points, argument meaning and hook behaviour as onnx_optimized_inference.py:
    run_module:297  inference:214  execute_node:18  expand_node_inputs_outputs:236  get_weight_dict:273  prepare_inference:282
The `inject_operations.py` variant of the same API (dialect B / QCDQ graphs: 4-tuple `execute_node`, `inject_input`, the
DequantizeLinear-keyed operand hook, `perturb_matmul`) lives in onnx-transformer_b200/inject_operations.py and is served by the
handler table below.

Differences, all forced by the B200 design:
  * `weight_dict` values are torch CUDA tensors (numpy inputs are uploaded on entry); every intermediate is still kept
    under its ONNX tensor name, plus the reference's side-channel key "delta_4d".
  * a node is not wrapped into a one-node ModelProto + onnxruntime session (onnx_optimized_inference.py:33-54): its
    op_type selects a handler that launches one kernel of libot_b200.so.  There is no CPU fallback: an op without a
    CUDA handler raises.
  * `MatMul` whose operands are de-quantized int8 tensors (Round -> Mul(scale) [-> Transpose]) runs as the tcgen05 int8
    GEMM with the scales applied in the epilogue (int-exact factorisation); any other MatMul (attention
    products in the un-fused walk, fault deltas) runs as the fp32 kernel, i.e. literally the reference's arithmetic.
  * the random draws of the fault hooks can be supplied in inject_parameters["rng_draws"] (faults.Draws).
  * graph files: `.otg` archives written by save_graph (the reference's .onnx files are not in its snapshot).
"""
from __future__ import annotations

import io
import json
import os
import time
from typing import Dict, List, Optional

import numpy as np
import torch

from . import faults
from . import kernels as K
from .graph import Attribute, Graph, Initializer, Node, ValueInfo

META_KEY = "__ot_int8__"   # side table: tensor name -> int8 provenance (never an ONNX tensor name)
HOST_KEY = "__ot_host__"   # side table: small initializers (shapes, Clip bounds, scalars, zero points) kept on the host as numpy, filled by
                           # get_weight_dict, so the walk does not synchronise the stream to read them (.item() / .tolist())
FLOAT_MAX = 3.4e38          # onnx_optimized_inference.py:252


# ---------------------------------------------------------------------------------------------- graph files
def save_graph(graph: Graph, path: str) -> None:
    nodes = [dict(name=n.name, op_type=n.op_type, input=n.input, output=n.output,
                  attribute=[dict(name=a.name, f=a.f, i=a.i, ints=a.ints) for a in n.attribute]) for n in graph.node]
    meta = dict(name=graph.name, node=nodes, roles=graph.roles,
                input=[dict(name=v.name, shape=list(v.shape), dtype=v.dtype) for v in graph.input],
                output=[dict(name=v.name, shape=list(v.shape), dtype=v.dtype) for v in graph.output],
                value_info=[v.name for v in graph.value_info], initializer=[i.name for i in graph.initializer])
    arrays = {"init_%d" % k: i.array for k, i in enumerate(graph.initializer)}
    with open(path, "wb") as f:
        np.savez(f, __meta__=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)


def save_pt_archive(path: str, weight_dict, graph: Graph) -> None:
    """`torch.save((weight_dict, graph), path)` as the reference writes its weights/{encoder,decoder}.pt
    (parallelized_inject_onnx_transformer.py:540,621): tensors go out as numpy arrays, the graph as a plain dict."""
    nodes = [dict(name=n.name, op_type=n.op_type, input=n.input, output=n.output,
                  attribute=[dict(name=a.name, f=a.f, i=a.i, ints=a.ints) for a in n.attribute]) for n in graph.node]
    meta = dict(name=graph.name, node=nodes, roles=graph.roles,
                input=[dict(name=v.name, shape=list(v.shape), dtype=v.dtype) for v in graph.input],
                output=[dict(name=v.name, shape=list(v.shape), dtype=v.dtype) for v in graph.output],
                value_info=[v.name for v in graph.value_info], initializer=[i.name for i in graph.initializer])
    arrays = {"init_%d" % k: i.array for k, i in enumerate(graph.initializer)}
    wd = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in weight_dict.items()
          if k not in (META_KEY, HOST_KEY)}
    torch.save((wd, {"__ot_graph__": meta, "arrays": arrays}), path)


def load_pt_archive(path: str):
    """(weight_dict on the device, graph) from a `.pt` archive: this package's own or the reference's pickled
    (weight_dict, onnx GraphProto) pair (onnx_reader.load_pt_archive: no `onnx` package needed)."""
    from . import onnx_reader
    wd, graph = onnx_reader.load_pt_archive(path)
    weight_dict = {k: _to_device(v) for k, v in wd.items()}
    weight_dict[HOST_KEY] = {i.name: np.array(i.array) for i in graph.initializer if np.asarray(i.array).size <= 64}
    return weight_dict, graph


def _graph_from_dict(meta: dict, arrays) -> Graph:
    g = Graph(meta["name"])
    g.node = [Node(n["name"], n["op_type"], list(n["input"]), list(n["output"]),
                   [Attribute(a["name"], a["f"], a["i"], a["ints"]) for a in n["attribute"]]) for n in meta["node"]]
    g.roles = meta.get("roles", {})
    g.input = [ValueInfo(v["name"], tuple(v["shape"]), v["dtype"]) for v in meta["input"]]
    g.output = [ValueInfo(v["name"], tuple(v["shape"]), v["dtype"]) for v in meta["output"]]
    g.value_info = [ValueInfo(n, ()) for n in meta["value_info"]]
    g.initializer = [Initializer(n, np.asarray(arrays["init_%d" % k])) for k, n in enumerate(meta["initializer"])]
    return g


def load_graph(path: str) -> Graph:
    with open(path, "rb") as f:
        z = np.load(io.BytesIO(f.read()))
    meta = json.loads(bytes(z["__meta__"]).decode())
    g = Graph(meta["name"])
    g.node = [Node(n["name"], n["op_type"], list(n["input"]), list(n["output"]),
                   [Attribute(a["name"], a["f"], a["i"], a["ints"]) for a in n["attribute"]]) for n in meta["node"]]
    g.roles = meta.get("roles", {})
    g.input = [ValueInfo(v["name"], tuple(v["shape"]), v["dtype"]) for v in meta["input"]]
    g.output = [ValueInfo(v["name"], tuple(v["shape"]), v["dtype"]) for v in meta["output"]]
    g.value_info = [ValueInfo(n, ()) for n in meta["value_info"]]
    g.initializer = [Initializer(n, z["init_%d" % k]) for k, n in enumerate(meta["initializer"])]
    return g


def _as_graph(module_path_or_graph) -> Graph:
    """A graph object, an `.otg` archive (save_graph) or a real `.onnx` / `.onnx.gz` file (onnx_reader: protobuf wire reader +
    the Constant-folding / node-naming part of qonnx's cleanup)."""
    if isinstance(module_path_or_graph, str):
        if module_path_or_graph.endswith((".onnx", ".onnx.gz")):
            from . import onnx_reader
            return onnx_reader.load_onnx(module_path_or_graph)
        if module_path_or_graph.endswith(".pt"):
            from . import onnx_reader
            return onnx_reader.load_pt_archive(module_path_or_graph)[1]
        return load_graph(module_path_or_graph)
    return module_path_or_graph


def _to_device(value):
    if isinstance(value, torch.Tensor):
        return value.cuda() if not value.is_cuda else value
    arr = np.asarray(value)
    if arr.dtype == np.float64:
        arr = arr.astype(np.float32)
    return torch.from_numpy(np.ascontiguousarray(arr)).cuda()


# ---------------------------------------------------------------------------------------------- op handlers
def _meta(weight_dict) -> dict:
    return weight_dict.setdefault(META_KEY, {})


def _scalar(t: torch.Tensor):
    return t.numel() == 1


_FLOAT_MAX_HOST = np.float32(FLOAT_MAX).reshape(1)


def _host(node, pos: int, ins, wd) -> np.ndarray:
    """Host copy of operand `pos` of `node`: from the side table of small initializers when the operand is one (no device
    synchronisation), else read back from the device (a shape computed at run time by a raw, un-cleaned export)."""
    name = node.input[pos] if pos < len(node.input) else ""
    if name == "" and node.op_type == "Clip" and pos == 2:
        return _FLOAT_MAX_HOST                    # the max operand patched in by expand_node_inputs_outputs
    cached = wd.get(HOST_KEY, {}).get(name)
    if cached is not None:
        return cached
    return ins[pos].detach().cpu().numpy()


def _h_unary(op):
    def run(node, ins, wd):
        return K.unary(op, ins[0])
    return run


def _h_binary(op):
    def run(node, ins, wd):
        a, b = ins[0], ins[1]
        if a.dtype != torch.float32 or b.dtype != torch.float32:
            raise K.OtError("%s: only float32 operands have a CUDA handler (got %s, %s)" % (node.name, a.dtype, b.dtype))
        out = K.binary(op, a, b)
        if op == "Mul":
            # de-quantization Mul(Round_out, scale): remember the int8 tensor and its per-row scale
            m = _meta(wd)
            for q_name, s_t in ((node.input[0], b), (node.input[1], a)):
                src = m.get(q_name)
                if src is not None and src[0] == "q" and s_t.dim() >= 1 and s_t.shape[-1] == 1 and s_t.numel() == src[1].numel() // src[1].shape[-1]:
                    m[node.output[0]] = ("qs", src[1], s_t.reshape(-1).contiguous())
        return out
    return run


def _h_round(node, ins, wd):
    out = K.unary("Round", ins[0])
    # integer-valued result: keep an int8 copy for the tensor-core path -- only behind the quantizer pattern of dialect A
    # (Round of a Div / Clip output: x / (max(|x|, 1e-5) / 127) lies in [-127, 127]); a Round of anything else (arbitrary .onnx
    # files: values may leave the int8 range) stays on the fp32 MatMul
    prod = _meta(wd).setdefault("__producers__", {}).get(node.input[0])
    if prod in ("Div", "Clip"):
        _meta(wd)[node.output[0]] = ("q", K.cast(out, torch.int8))
    return out


def _h_clip(node, ins, wd):
    x = ins[0]
    if x.dtype in (torch.int8, torch.uint8):
        # dialect B: QuantizeLinear -> Clip(-2^(b-1), 2^(b-1)-1) on the integer tensor when bit_width < 8
        lo = int(_host(node, 1, ins, wd).reshape(-1)[0]) if len(ins) > 1 and ins[1] is not None else (-128 if x.dtype == torch.int8 else 0)
        hi = int(_host(node, 2, ins, wd).reshape(-1)[0]) if len(ins) > 2 and ins[2] is not None else (127 if x.dtype == torch.int8 else 255)
        out = K.cast(K.clip(K.cast(x, torch.float32, numeric_u8=True), float(lo), float(hi)), x.dtype, numeric_u8=True)
        _meta(wd)[node.output[0]] = ("q", out)
        return out
    lo = float(_host(node, 1, ins, wd).reshape(-1)[0]) if len(ins) > 1 and ins[1] is not None else -FLOAT_MAX
    hi = float(_host(node, 2, ins, wd).reshape(-1)[0]) if len(ins) > 2 and ins[2] is not None else FLOAT_MAX
    return K.clip(x, lo, hi)


def _h_reduce(op):
    def run(node, ins, wd):
        axes = node.attr("axes", [-1])
        if list(axes) not in ([-1], [ins[0].dim() - 1]):
            raise K.OtError("%s: only reductions over the last axis have a CUDA handler" % node.name)
        return K.reduce_last(op, ins[0], keepdims=bool(node.attr("keepdims", 1)))
    return run


def _h_softmax(node, ins, wd):
    axis = node.attr("axis", -1)
    if axis not in (-1, ins[0].dim() - 1):
        raise K.OtError("%s: Softmax over a non-last axis has no CUDA handler" % node.name)
    return K.softmax_last(ins[0])


def _h_where(node, ins, wd):
    cond, a, x = ins
    if not _scalar(a):
        raise K.OtError("%s: Where with a tensor `X` operand has no CUDA handler" % node.name)
    return K.where_scalar(cond, float(_host(node, 1, ins, wd).reshape(-1)[0]), x)


def _h_equal(node, ins, wd):
    x, c = ins
    if x.dtype != torch.int64 or not _scalar(c):
        raise K.OtError("%s: Equal is implemented for int64 tensor == scalar" % node.name)
    return K.equal_scalar_i64(x, int(_host(node, 1, ins, wd).reshape(-1)[0]))


_ONNX_DTYPE = {1: torch.float32, 7: torch.int64, 9: torch.bool, 3: torch.int8, 6: torch.int32, 2: torch.uint8}


def _h_cast(node, ins, wd):
    to = _ONNX_DTYPE[int(node.attr("to"))]
    x = ins[0]
    if x.dtype == to:
        out = x.clone()      # same-type Cast (e.g. the float Cast after Round in attention.py:34): a copy
        m = _meta(wd)
        if node.input[0] in m:
            m[node.output[0]] = m[node.input[0]]
        return out
    return K.cast(x, to)


def _h_transpose(node, ins, wd):
    perm = list(node.attr("perm"))
    x = ins[0]
    m = _meta(wd)
    src = m.get(node.input[0])
    if src is not None and src[0] == "qs" and perm == [1, 0] and x.dim() == 2:
        m[node.output[0]] = ("qsT", src[1], src[2])       # weight: keep the K-major int8 tensor, no data movement needed
    if x.dtype in (torch.bool, torch.int8, torch.uint8):        # the transpose kernel moves 4-byte elements: widen, move, narrow (all CUDA handlers)
        wide = K.cast(x.view(torch.uint8) if x.dtype == torch.bool else x, torch.int32, numeric_u8=True)
        out = K.cast(K.transpose(wide, perm), torch.uint8 if x.dtype == torch.bool else x.dtype, numeric_u8=True)
        if x.dtype == torch.bool:
            return out.view(torch.bool)
        if src is not None and src[0] == "q":
            m[node.output[0]] = ("q", out)
        return out
    return K.transpose(x, perm)


def _h_reshape(node, ins, wd):
    shape = [int(v) for v in _host(node, 1, ins, wd).reshape(-1)]
    x = ins[0]
    shape = [x.shape[i] if s == 0 else s for i, s in enumerate(shape)]
    out = x.reshape(shape)
    m = _meta(wd)
    src = m.get(node.input[0])
    if src is not None and src[0] == "qs" and out.shape[-1] == x.shape[-1]:
        m[node.output[0]] = src                               # leading dims regrouped only: rows unchanged
    return out


def _h_unsqueeze(node, ins, wd):
    axes = node.attr("axes")
    if axes is None and len(ins) > 1:
        axes = [int(v) for v in _host(node, 1, ins, wd).reshape(-1)]
    out = ins[0]
    for ax in sorted(int(a) for a in axes):
        out = out.unsqueeze(ax)
    return out


def _h_matmul(node, ins, wd):
    a, b = ins
    m = _meta(wd)
    ma, mb = m.get(node.input[0]), m.get(node.input[1])
    if ma is not None and mb is not None and ma[0] == "qs" and mb[0] == "qsT" and b.dim() == 2:
        aq, sa = ma[1].reshape(-1, ma[1].shape[-1]), ma[2]
        wq, sw = mb[1], mb[2]
        out = K.linear_w8a8(aq.contiguous(), wq.contiguous(), row_scale=sa, col_scale=sw, out_kind=K.OUT_F32)
        return out.reshape(tuple(a.shape[:-1]) + (wq.shape[0],))
    if ma is not None and mb is not None and ma[0] == "dq" and mb[0] == "dq" and b.dim() == 2:
        # dialect B: DequantizeLinear(A, per-row scale) x DequantizeLinear(W [K,N], per-column scale), zero points 0 -> the same
        # int-exact factorisation; the K-major copy of the (static) weight is made once per weight tensor
        if ma[3] in ("row", "scalar") and mb[3] in ("col", "scalar") and ma[1].dtype == torch.int8 and mb[1].dtype == torch.int8:
            aq = ma[1].reshape(-1, ma[1].shape[-1]).contiguous()
            if aq.shape[1] % 16 == 0 and mb[1].shape[1] % 32 == 0:
                wt = _kmajor_weight(wd, node.input[1], mb[1])
                sw = mb[2].reshape(-1)
                sw = sw.expand(wt.shape[0]).contiguous() if sw.numel() == 1 else sw.contiguous()
                sa = ma[2].reshape(-1)
                sa = sa.expand(aq.shape[0]).contiguous() if sa.numel() == 1 else sa.contiguous()
                out = K.linear_w8a8(aq, wt, row_scale=sa, col_scale=sw, out_kind=K.OUT_F32)
                return out.reshape(tuple(a.shape[:-1]) + (wt.shape[0],))
    return K.matmul_f32(a, b)


def _kmajor_weight(wd, name: str, w_kn: torch.Tensor) -> torch.Tensor:
    """int8 [K,N] -> [N,K] (the tensor-core GEMM takes both operands K-major), cached per tensor name and storage."""
    cache = _meta(wd).setdefault("__kmajor__", {})
    hit = cache.get(name)
    if hit is not None and hit[0] == w_kn.data_ptr():
        return hit[1]
    wt = K.cast(K.transpose(K.cast(w_kn, torch.int32), [1, 0]), torch.int8)
    cache[name] = (w_kn.data_ptr(), wt)
    return wt


def _zero_point(node, pos, ins, wd):
    """(device tensor or None, host numpy or None) of an optional zero-point operand; None when absent or identically zero."""
    if pos >= len(ins) or ins[pos] is None:
        return None, None
    host = _host(node, pos, ins, wd)
    if not np.any(host):
        return None, None
    return ins[pos], host


def _as_s8(q: torch.Tensor, zp_dev, zp_host):
    """uint8 operand -> the int8 tensor x ^ 0x80 (= x - 128) with its zero point lowered by 128 (the tensor cores take signed bytes)."""
    if q.dtype == torch.int8:
        return q, zp_dev
    if q.dtype != torch.uint8:
        raise K.OtError("integer MatMul operands must be int8 or uint8, got %s" % q.dtype)
    shifted = K.cast(K.binary("Sub", K.cast(q, torch.float32, numeric_u8=True), torch.full((1,), 128.0, device=q.device)), torch.int8)
    host = (zp_host.astype(np.int64) if zp_host is not None else np.zeros(1, np.int64)) - 128
    return shifted, torch.from_numpy(host.astype(np.int32)).to(q.device)


def _h_matmul_integer(node, ins, wd):
    """ONNX MatMulInteger(A [.., K], B [K, N], a_zero_point, b_zero_point) -> int32: the int8 tensor-core GEMM with the zero-point
    correction  - a_zp[m]*colsum(B)[n] - b_zp[n]*rowsum(A)[m] + K*a_zp[m]*b_zp[n]  applied to the accumulator in the epilogue
    (ot_matmul_integer).  Spec-level parity only."""
    a, b = ins[0], ins[1]
    if b.dim() != 2:
        raise K.OtError("%s: MatMulInteger expects B of rank 2" % node.name)
    azd, azh = _zero_point(node, 2, ins, wd)
    bzd, bzh = _zero_point(node, 3, ins, wd)
    a8, azd = _as_s8(a, azd, azh)
    b8, bzd = _as_s8(b, bzd, bzh)
    wt = _kmajor_weight(wd, node.input[1], b8) if b.dtype == torch.int8 else K.cast(K.transpose(K.cast(b8, torch.int32), [1, 0]), torch.int8)
    a2 = a8.reshape(-1, a8.shape[-1]).contiguous()
    if azd is not None and azd.numel() not in (1, a2.shape[0]):
        raise K.OtError("%s: a_zero_point must be a scalar or one value per row" % node.name)
    out = K.matmul_integer(a2, wt, a_zp=azd, b_zp=bzd)
    return out.reshape(tuple(a.shape[:-1]) + (wt.shape[0],))


def _h_qlinear_matmul(node, ins, wd):
    """ONNX QLinearMatMul(a, a_scale, a_zp, b, b_scale, b_zp, y_scale, y_zp) -> int8:
    saturate(rint(fl(fl(float(acc') * a_scale) * b_scale) / y_scale) + y_zp), acc' as in MatMulInteger (ot_qlinear_matmul)."""
    a, a_scale, b, b_scale = ins[0], ins[1], ins[3], ins[4]
    if b.dim() != 2:
        raise K.OtError("%s: QLinearMatMul expects b of rank 2" % node.name)
    azd, azh = _zero_point(node, 2, ins, wd)
    bzd, bzh = _zero_point(node, 5, ins, wd)
    y_scale = float(_host(node, 6, ins, wd).reshape(-1)[0])
    yzh = _host(node, 7, ins, wd) if len(ins) > 7 and ins[7] is not None else np.zeros(1, np.int8)
    if yzh.dtype != np.int8:
        raise K.OtError("%s: only int8 outputs (y_zero_point int8) have a CUDA handler" % node.name)
    a8, azd = _as_s8(a, azd, azh)
    b8, bzd = _as_s8(b, bzd, bzh)
    wt = _kmajor_weight(wd, node.input[3], b8) if b.dtype == torch.int8 else K.cast(K.transpose(K.cast(b8, torch.int32), [1, 0]), torch.int8)
    a2 = a8.reshape(-1, a8.shape[-1]).contiguous()
    out = K.qlinear_matmul(a2, a_scale, azd, wt, b_scale, bzd, y_scale, int(yzh.reshape(-1)[0]))
    _meta(wd)[node.output[0]] = ("q", out)
    return out.reshape(tuple(a.shape[:-1]) + (wt.shape[0],))


def _axis_view(t: torch.Tensor, rank: int, axis: int) -> torch.Tensor:
    """A per-axis 1-D scale / zero point as a broadcastable tensor of `rank` dims (ONNX QuantizeLinear `axis`)."""
    if t.numel() == 1 or t.dim() != 1:
        return t
    shape = [1] * rank
    shape[axis if axis >= 0 else rank + axis] = t.numel()
    return t.reshape(shape)


def _scale_kind(s: torch.Tensor, x: torch.Tensor) -> str:
    """How a broadcastable scale varies over x: 'scalar', 'row' (one value per row of the last axis) or 'col' (along the last axis)."""
    if s.numel() == 1:
        return "scalar"
    if s.dim() == x.dim() and s.shape[-1] == 1 and s.numel() == x.numel() // x.shape[-1]:
        return "row"
    if s.dim() == x.dim() and s.numel() == x.shape[-1] and s.shape[-1] == x.shape[-1]:
        return "col"
    if s.dim() == x.dim() and s.shape[-1] == 1:
        return "rowb"        # per row of one inner axis, broadcast over the others (e.g. (1,T,1) on [B,T,d])
    return "other"


def _h_quantize_linear(node, ins, wd):
    """ONNX QuantizeLinear(x, y_scale, y_zero_point, axis=1): saturate(rint(x / y_scale) + y_zero_point); the result type is the zero
    point's (uint8 when it is omitted).  Dialect B: Brevitas QuantIdentity exports, SURVEY.md App. C."""
    x, s = ins[0], _axis_view(ins[1], ins[0].dim(), int(node.attr("axis", 1)))
    zpd, zph = _zero_point(node, 2, ins, wd)
    unsigned = (len(ins) < 3 or ins[2] is None) or ins[2].dtype == torch.uint8
    q = K.unary("Round", K.binary("Div", x, s))
    if zpd is not None:
        q = K.binary("Add", q, K.cast(_axis_view(zpd, x.dim(), int(node.attr("axis", 1))), torch.float32, numeric_u8=True))
    lo, hi = (0.0, 255.0) if unsigned else (-128.0, 127.0)
    out = K.cast(K.clip(q, lo, hi), torch.uint8 if unsigned else torch.int8, numeric_u8=True)
    _meta(wd)[node.output[0]] = ("q", out)
    return out


def _h_dequantize_linear(node, ins, wd):
    """ONNX DequantizeLinear(x, x_scale, x_zero_point, axis=1): (x - x_zero_point) * x_scale in fp32."""
    x = ins[0]
    axis = int(node.attr("axis", 1))
    s = _axis_view(ins[1], x.dim(), axis)
    zpd, zph = _zero_point(node, 2, ins, wd)
    xf = K.cast(x, torch.float32, numeric_u8=True) if x.dtype != torch.float32 else x
    if zpd is not None:
        xf = K.binary("Sub", xf, K.cast(_axis_view(zpd, x.dim(), axis), torch.float32, numeric_u8=True))
    out = K.binary("Mul", xf, s)
    if x.dtype == torch.int8 and zpd is None and x.dim() >= 2:
        # int8 provenance for the tensor-core MatMul: ("dq", integer tensor, scale, how the scale varies).  A per-tensor scale serves
        # as a per-row scale of an activation or a per-column scale of a weight.
        kind = _scale_kind(s, x)
        if kind == "rowb":       # e.g. (1,T,1) on [B,T,d]: materialise one scale per row
            s, kind = s.expand(tuple(x.shape[:-1]) + (1,)), "row"
        if kind in ("scalar", "row", "col"):
            _meta(wd)[node.output[0]] = ("dq", x, s.reshape(-1).contiguous(), kind)
    return out


def _h_shape(node, ins, wd):
    return torch.tensor(list(ins[0].shape), dtype=torch.int64, device=ins[0].device)


def _h_gather(node, ins, wd):
    if ins[0].dim() != 1:
        raise K.OtError("%s: Gather is implemented for shape vectors only" % node.name)
    return ins[0][ins[1].long()]


def _h_reduce_prod(node, ins, wd):
    return ins[0].prod().reshape(1) if node.attr("keepdims", 1) else ins[0].prod()


HANDLERS = {
    "Abs": _h_unary("Abs"), "Relu": _h_unary("Relu"), "Sqrt": _h_unary("Sqrt"), "Round": _h_round,
    "Add": _h_binary("Add"), "Sub": _h_binary("Sub"), "Mul": _h_binary("Mul"), "Div": _h_binary("Div"),
    "Clip": _h_clip, "ReduceMax": _h_reduce("ReduceMax"), "ReduceMean": _h_reduce("ReduceMean"), "Softmax": _h_softmax,
    "Where": _h_where, "Equal": _h_equal, "Cast": _h_cast, "Transpose": _h_transpose, "Reshape": _h_reshape,
    "Unsqueeze": _h_unsqueeze, "MatMul": _h_matmul, "MatMulInteger": _h_matmul_integer, "QLinearMatMul": _h_qlinear_matmul,
    "QuantizeLinear": _h_quantize_linear, "DequantizeLinear": _h_dequantize_linear,
    "Shape": _h_shape, "Gather": _h_gather, "ReduceProd": _h_reduce_prod,
    "Identity": lambda node, ins, wd: ins[0],          # raw exports alias shared initializers (deduplicated biases) this way
}


def run_node(node, input_tensors: List[Optional[torch.Tensor]], weight_dict) -> torch.Tensor:
    """Execute one node on explicit inputs (the equivalent of execute_onnx on the one-node model)."""
    handler = HANDLERS.get(node.op_type)
    if handler is None:
        raise K.OtError("no CUDA handler for op %r (node %s): this executor has no CPU fallback" % (node.op_type, node.name))
    out = handler(node, input_tensors, weight_dict)
    if node.output:
        _meta(weight_dict).setdefault("__producers__", {})[node.output[0]] = node.op_type      # provenance for _h_round
    return out


# ---------------------------------------------------------------------------------------------- reference API
def expand_node_inputs_outputs(graph, node, weight_dict, module):
    """onnx_optimized_inference.py:236-271.  Returns (inputs, outputs, seconds): the value-info records of the node's
    operands / results.  The missing third Clip operand is added as max = 3.4e38 (:248-252); the decoder's dynamic
    dimensions need no patching here because tensors carry their own shapes."""
    start = time.time()
    names_in = [n for n in node.input if n]
    known, inits = _graph_tables(graph)
    added_inputs = [known.get(n) or ValueInfo(n, ()) for n in names_in if n in known or n in inits or n in weight_dict]
    added_outputs = [known.get(n) or ValueInfo(n, ()) for n in node.output]
    if "Clip" in node.name and len([n for n in node.input if n]) < 3:
        extra = ValueInfo(node.input[0][:-1] + "2", ())
        added_inputs.append(extra)
        if extra.name not in weight_dict:
            weight_dict[extra.name] = torch.tensor(FLOAT_MAX, dtype=torch.float32, device="cuda")
            weight_dict.setdefault(HOST_KEY, {})[extra.name] = np.float32(FLOAT_MAX).reshape(1)
    return added_inputs, added_outputs, time.time() - start


def _graph_tables(graph):
    """name -> value-info record and the set of initializer names of a graph, built once per graph object (the reference rebuilds
    both for every node it executes: 45 % of this executor's host time per node before they were cached)."""
    key = (len(graph.input), len(graph.output), len(graph.value_info), len(graph.initializer))
    cached = getattr(graph, "_ot_tables", None)
    if cached is None or cached[0] != key:
        known = {v.name: v for v in list(graph.input) + list(graph.output) + list(graph.value_info)}
        cached = (key, known, {i.name for i in graph.initializer})
        try:
            graph._ot_tables = cached
        except AttributeError:        # a graph type with __slots__: no cache
            pass
    return cached[1], cached[2]


def _gather_inputs(node, weight_dict, added_inputs):
    ins = []
    for pos, name in enumerate(node.input):
        if name == "":
            if node.op_type == "Clip" and pos == 2:
                ins.append(weight_dict[added_inputs[-1].name])   # the patched-in max operand
            else:
                ins.append(None)
            continue
        ins.append(weight_dict[name])
    return ins


def execute_node(node, main_graph, final_output_node, weight_dict, module, inject_parameters=None):
    """onnx_optimized_inference.py:18-212: run the node, store its result under its tensor name, then apply the fault
    hooks: RANDOM / RANDOM_BITFLIP on the target MatMul's output (:59-72); INPUT / WEIGHT / INPUT16 / WEIGHT16 along the
    faulty_trace (:74-204) -- perturb_quantizer on the first trace node (inject_utils/layers.py:87-142), re-execution of the
    following nodes on `delta_4d`, window selection and `out += delta` on the last one."""
    added_inputs, added_outputs, list_operation_time = expand_node_inputs_outputs(main_graph, node, weight_dict, module)
    ins = _gather_inputs(node, weight_dict, added_inputs)
    out = run_node(node, ins, weight_dict)
    tensor_output_name = node.output[0]
    weight_dict[tensor_output_name] = out
    output_tensors = {tensor_output_name: out}
    p = inject_parameters

    if p and ("RANDOM" in p["inject_type"]) and (node.name == p["faulty_operation_name"]):
        draws = faults.Draws(p)
        target_indices = draws.indices("target_indices", out.shape)
        golden_value = float(out[tuple(target_indices)].item())
        if "BITFLIP" in p["inject_type"]:
            faulty_value = faults.float32_bit_flip_value(golden_value, draws.randint("flip_bit", 0, 32))
        else:
            faulty_value = faults.delta_init_value(draws.bits32("random_bits"))
        out[tuple(target_indices)] = faulty_value           # in place, like the reference (:68)
        _meta(weight_dict).pop(tensor_output_name, None)

    if p and (module in p["targetted_module"]) and p["faulty_trace"] and (node.name == p["faulty_trace"][0]) and \
            (p["inject_type"] in ["INPUT", "WEIGHT", "INPUT16", "WEIGHT16"]):
        faulty_operation = p["faulty_trace"][0]
        draws = faults.Draws(p)
        if p["faulty_tensor_name"] in node.input:
            # first node of the trace consumes the integer tensor: build the one-hot perturbation
            assert p["faulty_quantizer_name"] == p["faulty_trace"][0]
            _perturb_quantizer(node, ins, weight_dict, p["faulty_tensor_name"], p["faulty_bit_position"], draws)
            p["intermediate_output_name"] = tensor_output_name
        else:
            pos = None
            for k, name in enumerate(node.input):
                if name == p["intermediate_output_name"]:
                    pos = k
            assert pos is not None
            delta_ins = list(ins)
            delta_ins[pos] = weight_dict["delta_4d"]
            scratch = {META_KEY: {}}                          # the delta carries no int8 provenance: fp32 kernels
            weight_dict["delta_4d"] = run_node(node, delta_ins, scratch)
            p["intermediate_output_name"] = tensor_output_name

        if faulty_operation == p["faulty_operation_name"]:
            assert len(p["faulty_trace"]) == 1
            delta = weight_dict["delta_4d"]
            if p["inject_type"] == "INPUT16":
                delta = _window(delta, axis=3, length=16, start_key="window_start", draws=draws, random_len=False)
            elif p["inject_type"] == "WEIGHT16":
                delta = _window(delta, axis=2, length=16, start_key="window_start", draws=draws, random_len=True)
            weight_dict["delta_4d"] = delta
            faulty = K.binary("Add", weight_dict[tensor_output_name], delta)
            weight_dict[tensor_output_name] = faulty
            output_tensors[tensor_output_name] = faulty
            _meta(weight_dict).pop(tensor_output_name, None)
        p["faulty_trace"] = p["faulty_trace"][1:]
    return output_tensors, weight_dict, list_operation_time


def _perturb_quantizer(node, ins, weight_dict, faulty_tensor_name, faulty_bit_position, draws):
    """inject_utils/layers.py:87-142: flip one bit of one random element of the integer tensor (int_bit_flip :70-84), run
    the de-quantizing node on the one-hot tensor holding the faulty value, and turn the result into a delta by
    subtracting the golden de-quantized value at that index."""
    golden = weight_dict[faulty_tensor_name]
    idx = tuple(draws.indices("target_indices", golden.shape))
    q = int(golden[idx].item())                                # np.int8(tensor)[idx]
    faulty_value = faults.flip_int8_bit(q, faulty_bit_position)
    assert -128 <= faulty_value <= 127
    one_hot = torch.zeros_like(golden)
    one_hot[idx] = float(faulty_value)
    pos = list(node.input).index(faulty_tensor_name)
    pert_ins = list(ins)
    pert_ins[pos] = one_hot
    delta = run_node(node, pert_ins, {META_KEY: {}})
    dequantized = weight_dict[node.output[0]]
    delta[idx] = delta[idx] - dequantized[idx]
    weight_dict["delta_4d"] = delta


def _window(delta: torch.Tensor, axis: int, length: int, start_key: str, draws, random_len: bool) -> torch.Tensor:
    """onnx_optimized_inference.py:111-179: keep only a 16-aligned window of the delta along `axis` (shape[3] for INPUT16,
    shape[2] for WEIGHT16 -- 4-D attention outputs only, as in the reference), at the coordinates of the first non-zero."""
    if delta.dim() < 4:
        raise IndexError("INPUT16/WEIGHT16 index shape[%d]: only 4-D MatMul outputs are supported (as in the reference)" % axis)
    shape = list(delta.shape)
    blocks = shape[axis] // 16
    start = 0 if blocks == 0 else draws.randint(start_key, 0, blocks)
    start *= 16
    n = length
    out = torch.zeros_like(delta)
    nz = torch.nonzero(delta)
    if nz.shape[0] == 0:
        return out
    if random_len:
        n = draws.randint("window_len", 1, 16)
    index = [int(v) for v in nz[0].tolist()]
    index[axis] = start
    for i in range(n):
        if i >= shape[axis]:
            break
        out[tuple(index)] = delta[tuple(index)]
        index[axis] += 1
        if index[axis] >= shape[axis]:
            break
    return out


def inference(main_graph, weight_dict, module, inject_parameters=None):
    """onnx_optimized_inference.py:214-234: walk graph.node in order; returns the LAST node's {name: tensor} and the dict."""
    output_tensors = None
    for node in main_graph.node:
        output_tensors, weight_dict, _ = execute_node(node, main_graph, node.output[0], weight_dict, module, inject_parameters)
    return output_tensors, weight_dict


def get_weight_dict(module_path):
    """onnx_optimized_inference.py:273-280: (graph, {initializer name: tensor})."""
    graph = _as_graph(module_path)
    weight_dict = {i.name: _to_device(i.array) for i in graph.initializer}
    # small initializers (Reshape shapes, Clip bounds, Where / Equal scalars, zero points, y_scale ...) also stay on the host: the
    # handlers read them from this side table instead of synchronising the stream once per node
    weight_dict[HOST_KEY] = {i.name: np.array(i.array) for i in graph.initializer if np.asarray(i.array).size <= 64}
    return graph, weight_dict


def prepare_inference(module_path, module_input_values):
    """onnx_optimized_inference.py:282-295: (weight_dict incl. the graph inputs, graph)."""
    graph, weight_dict = get_weight_dict(module_path)
    init_names = {i.name for i in graph.initializer}
    for v in graph.input:
        if v.name not in init_names:
            weight_dict[v.name] = _to_device(module_input_values[v.name])
            weight_dict[HOST_KEY].pop(v.name, None)
    return weight_dict, graph


# ---------------------------------------------------------------------------------------------- CUDA-graph replay of whole passes
# A fault-free pass over a module is ~1,600 handler calls whose host side (~18 us per node through Python + ctypes) is 10x the GPU
# time of the kernels they launch.  With replay enabled, run_module captures the node walk of a (graph, input shapes) pair into a CUDA
# graph the first time the pair comes back and replays it afterwards: same handlers, same kernels, same launch order, one host call.
# Opt-in (enable_graph_replay / OT_EXEC_REPLAY=1) because a replayed pass returns the final output only: the intermediates of a
# captured walk live in a memory pool shared by all captured passes (0.6 GB for the full-size decoder instead of 0.6 GB per prefix
# length) and are not left in weight_dict.  Passes with inject_parameters always take the node walk.
REPLAY_KEY = "__ot_replay__"
REPLAY_MAX_GRAPHS = 512       # captured passes kept per weight_dict (one per distinct set of input shapes)
_replay_enabled = os.environ.get("OT_EXEC_REPLAY", "0") == "1"
replay_stats = {"eager": 0, "captured": 0, "replayed": 0, "failed": 0}


def enable_graph_replay(on: bool = True) -> None:
    global _replay_enabled
    _replay_enabled = bool(on)


def _replay_pass(module, input_values, wd, graph):
    """Returns (output_tensors, weight_dict) of a replayed / freshly captured pass, or None when this pass has to be walked."""
    cache = wd.get(REPLAY_KEY)
    if cache is None or cache["graph"] is not graph:
        cache = wd[REPLAY_KEY] = {"graph": graph, "records": {}, "pool": None, "bad": set(), "warm": False}
    if not cache["warm"]:
        cache["warm"] = True        # the first walk of a weight_dict patches constants into it (Clip bounds ...) with blocking copies
        return None
    dev_in = {k: _to_device(v) for k, v in input_values.items()}
    sig = tuple(sorted((k, tuple(t.shape), str(t.dtype)) for k, t in dev_in.items()))
    if sig in cache["bad"]:
        return None
    rec = cache["records"].get(sig)
    if rec is None:
        static_in = {k: t.clone() for k, t in dev_in.items()}
        host = wd.get(HOST_KEY, {})
        for k, t in static_in.items():
            wd[k] = t
            host.pop(k, None)
        wd.pop(META_KEY, None)
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        try:
            with torch.cuda.graph(g, pool=cache["pool"]):
                out_tensors, _ = inference(graph, wd, module, None)
        except Exception:
            # a handler that reads a value back (raw exports with run-time shapes) cannot be captured: this shape is walked from now on
            torch.cuda.synchronize()
            cache["bad"].add(sig)
            replay_stats["failed"] += 1
            for node in graph.node:
                for o in node.output:
                    wd.pop(o, None)
            return None
        if cache["pool"] is None:
            cache["pool"] = g.pool()
        name = list(out_tensors.keys())[0]
        if len(cache["records"]) >= REPLAY_MAX_GRAPHS:            # bounded: drop the oldest captured pass (dicts keep insertion order)
            cache["records"].pop(next(iter(cache["records"])))
        rec = cache["records"][sig] = {"graph": g, "inputs": static_in, "out_name": name, "out": out_tensors[name]}
        del out_tensors
        for node in graph.node:                     # the captured intermediates belong to the shared pool: no reference survives
            for o in node.output:
                wd.pop(o, None)
        wd.pop(META_KEY, None)
        replay_stats["captured"] += 1
    else:
        for k, t in dev_in.items():
            rec["inputs"][k].copy_(t)
            wd[k] = rec["inputs"][k]
        replay_stats["replayed"] += 1
    rec["graph"].replay()
    out = rec["out"].clone()                        # the pool is shared: the next captured pass may overwrite rec["out"]
    wd[rec["out_name"]] = out
    return {rec["out_name"]: out}, wd


def run_module(module, input_values, module_filepath, module_weight_dict, module_graph, inject_parameters=None):
    """onnx_optimized_inference.py:297-304."""
    if _replay_enabled and inject_parameters is None:
        done = _replay_pass(module, input_values, module_weight_dict, module_graph)
        if done is not None:
            return done
        replay_stats["eager"] += 1
    for input_name in list(input_values.keys()):
        module_weight_dict[input_name] = _to_device(input_values[input_name])
        module_weight_dict.get(HOST_KEY, {}).pop(input_name, None)     # an overwritten initializer is no longer a known constant
    module_weight_dict.pop(META_KEY, None)
    return inference(module_graph, module_weight_dict, module, inject_parameters)


def to_numpy(tensors: Dict[str, torch.Tensor]) -> Dict[str, np.ndarray]:
    """Convenience for callers that index the reference's numpy outputs."""
    return {k: v.detach().cpu().numpy() for k, v in tensors.items() if isinstance(v, torch.Tensor) and k not in (META_KEY, HOST_KEY)}
