"""ORACLE (test infrastructure, never imported by the product): numpy restatement of the reference's custom
node-by-node ONNX executor and its fault hooks.

Follows onnx_optimized_inference.py:18-212 (execute_node: run one node, store every intermediate, RANDOM /
RANDOM_BITFLIP hook :59-72, INPUT/WEIGHT[16] trace hook :74-204), :214-234 (inference), :236-271
(expand_node_inputs_outputs: the missing Clip max = 3.4e38), :273-304 (get_weight_dict, prepare_inference,
run_module) and inject_utils/layers.py:70-142 (int_bit_flip, perturb_quantizer).  The third-party call
`qonnx.core.onnx_exec.execute_onnx(one-node model)` (-> onnxruntime CPU EP; neither package is installable here,
versions unpinned: SURVEY.md 8c) is restated per op from the ONNX opset-13 operator specification: fp32 IEEE
arithmetic, Round = half-to-even, Softmax over `axis`, MatMul = fp32 matrix product.

Dialect B (Brevitas QCDQ; inject_operations.py) adds the ONNX-spec operators QuantizeLinear, DequantizeLinear, Clip on integer
tensors, MatMulInteger and QLinearMatMul -- restated from the ONNX operator specification only: no reference artefact exists for
them: PARITY UNPINNED for these five handlers.

mode = "ref-float": every MatMul is an fp32 product of the de-quantized operands (the reference as is).
mode = "int-exact": a MatMul whose operands are Round -> Mul(scale) [-> Transpose] chains is evaluated as the exact
integer contraction followed by fl(fl(float(acc)*s_row)*s_col) (the factorisation the CUDA GEMM implements).
"""
from __future__ import annotations

import time
from typing import Dict, List, Optional

import numpy as np

from . import intexact as ox

F32 = np.float32
FLOAT_MAX = 3.4e38
_ONNX_DTYPE = {1: np.float32, 7: np.int64, 9: np.bool_, 3: np.int8, 6: np.int32, 2: np.uint8}


class Draws:
    """Explicit random draws of a trial (same contract as the product's faults.Draws)."""

    def __init__(self, p: dict):
        self.given = dict(p.get("rng_draws") or {})
        self.used = p.setdefault("rng_draws_used", {})

    def indices(self, key, shape):
        idx = [int(i) for i in self.given[key]] if key in self.given else [int(np.random.randint(0, d)) for d in shape]
        self.used[key] = idx
        return idx

    def randint(self, key, lo, hi):
        v = int(self.given[key]) if key in self.given else int(np.random.randint(lo, hi))
        self.used[key] = v
        return v

    def bits32(self, key):
        v = int(self.given[key]) if key in self.given else int("".join(str(np.random.randint(0, 2)) for _ in range(32)), 2)
        self.used[key] = v
        return v


def _attr(node, name, default=None):
    for a in node.attribute:
        if a.name == name:
            if a.ints is not None:
                return list(a.ints)
            return a.i if a.i is not None else a.f
    return default


def _axis_view(t: np.ndarray, rank: int, axis: int) -> np.ndarray:
    """A per-axis 1-D scale / zero point broadcast along `axis` (ONNX QuantizeLinear / DequantizeLinear)."""
    if t.size == 1 or t.ndim != 1:
        return t
    shape = [1] * rank
    shape[axis if axis >= 0 else rank + axis] = t.size
    return t.reshape(shape)


def _scale_kind(s: np.ndarray, x: np.ndarray) -> str:
    if s.size == 1:
        return "scalar"
    if s.ndim == x.ndim and s.shape[-1] == 1 and s.size == x.size // x.shape[-1]:
        return "row"
    if s.ndim == x.ndim and s.size == x.shape[-1] and s.shape[-1] == x.shape[-1]:
        return "col"
    if s.ndim == x.ndim and s.shape[-1] == 1:
        return "rowb"
    return "other"


def run_node(node, ins: List[Optional[np.ndarray]], prov: Optional[dict] = None, mode: str = "ref-float") -> np.ndarray:
    """ONNX opset-13 semantics of one node on numpy arrays.  `prov` tracks int8 provenance for int-exact MatMuls."""
    op = node.op_type
    prov = prov if prov is not None else {}
    x = ins[0]
    if op == "Abs":
        return np.abs(x)
    if op == "Relu":
        return np.maximum(x, F32(0))
    if op == "Sqrt":
        return np.sqrt(x).astype(F32)
    if op == "Round":
        out = np.rint(x).astype(F32)
        prov[node.output[0]] = ("q", out)
        return out
    if op in ("Add", "Sub", "Mul", "Div"):
        a, b = ins[0], ins[1]
        out = {"Add": np.add, "Sub": np.subtract, "Mul": np.multiply, "Div": np.divide}[op](a, b).astype(F32)
        if op == "Mul":
            for qn, s in ((node.input[0], b), (node.input[1], a)):
                src = prov.get(qn)
                if src is not None and src[0] == "q" and s.ndim >= 1 and s.shape[-1] == 1 and s.size == src[1].size // src[1].shape[-1]:
                    prov[node.output[0]] = ("qs", src[1], s.reshape(-1))
        return out
    if op == "Clip" and x.dtype in (np.int8, np.uint8):          # dialect B: Clip on the integer tensor (bit_width < 8)
        info = np.iinfo(x.dtype)
        lo = int(ins[1]) if len(ins) > 1 and ins[1] is not None else info.min
        hi = int(ins[2]) if len(ins) > 2 and ins[2] is not None else info.max
        out = np.clip(x.astype(np.int32), lo, hi).astype(x.dtype)
        prov[node.output[0]] = ("q", out)
        return out
    if op == "Clip":
        lo = F32(ins[1]) if len(ins) > 1 and ins[1] is not None else F32(-FLOAT_MAX)
        hi = F32(ins[2]) if len(ins) > 2 and ins[2] is not None else F32(FLOAT_MAX)
        return np.minimum(np.maximum(x, lo), hi).astype(F32)
    if op == "ReduceMax":
        return np.max(x, axis=tuple(_attr(node, "axes", [-1])), keepdims=bool(_attr(node, "keepdims", 1)))
    if op == "ReduceMean":
        ax = tuple(_attr(node, "axes", [-1]))
        return (np.sum(x.astype(np.float64), axis=ax, keepdims=bool(_attr(node, "keepdims", 1))) / np.prod([x.shape[a] for a in ax])).astype(F32)
    if op == "Softmax":
        ax = _attr(node, "axis", -1)
        m = np.max(x, axis=ax, keepdims=True)
        e = np.exp((x - m).astype(F32).astype(np.float64))
        return (e / np.sum(e, axis=ax, keepdims=True)).astype(F32)
    if op == "Where":
        return np.where(ins[0], ins[1], ins[2]).astype(F32)
    if op == "Equal":
        return np.equal(ins[0], ins[1])
    if op == "Cast":
        out = x.astype(_ONNX_DTYPE[int(_attr(node, "to"))])
        if out.dtype == x.dtype and node.input[0] in prov:
            prov[node.output[0]] = prov[node.input[0]]
        return out
    if op == "Transpose":
        perm = _attr(node, "perm")
        src = prov.get(node.input[0])
        if src is not None and src[0] == "qs" and perm == [1, 0] and x.ndim == 2:
            prov[node.output[0]] = ("qsT", src[1], src[2])
        return np.ascontiguousarray(np.transpose(x, perm))
    if op == "Reshape":
        shape = [int(v) for v in ins[1]]
        shape = [x.shape[i] if s == 0 else s for i, s in enumerate(shape)]
        out = x.reshape(shape)
        src = prov.get(node.input[0])
        if src is not None and src[0] == "qs" and out.shape[-1] == x.shape[-1]:
            prov[node.output[0]] = src
        return out
    if op == "Unsqueeze":
        axes = _attr(node, "axes")
        if axes is None:                       # opset 13: axes is the second input
            axes = [int(a) for a in np.asarray(ins[1]).reshape(-1)]
        out = x
        for ax in sorted(int(a) for a in axes):
            out = np.expand_dims(out, ax)
        return out
    if op == "Identity":
        return x
    if op == "Shape":                          # the raw (un-cleaned) export computes N and N-1 of every LayerNorm on the fly
        return np.asarray(x.shape, dtype=np.int64)
    if op == "Gather":
        return np.take(x, np.asarray(ins[1], dtype=np.int64), axis=int(_attr(node, "axis", 0)))
    if op == "ReduceProd":
        return np.prod(x, axis=None if _attr(node, "axes") is None else tuple(_attr(node, "axes")), keepdims=bool(_attr(node, "keepdims", 1)))
    if op == "QuantizeLinear":
        # ONNX opset 13: saturate(round_half_even(x / y_scale) + y_zero_point); output type = zero point's (uint8 if omitted)
        axis = int(_attr(node, "axis", 1))
        s = _axis_view(np.asarray(ins[1], dtype=F32), x.ndim, axis)
        zp = ins[2] if len(ins) > 2 and ins[2] is not None else None
        dt = zp.dtype if zp is not None else np.dtype(np.uint8)
        q = np.rint((x / s).astype(F32)).astype(F32)
        if zp is not None and np.any(zp):
            q = (q + _axis_view(np.asarray(zp), x.ndim, axis).astype(F32)).astype(F32)
        info = np.iinfo(dt)
        out = np.clip(q, info.min, info.max).astype(dt)
        prov[node.output[0]] = ("q", out)
        return out
    if op == "DequantizeLinear":
        # (x - x_zero_point) * x_scale
        axis = int(_attr(node, "axis", 1))
        s = _axis_view(np.asarray(ins[1], dtype=F32), x.ndim, axis)
        zp = ins[2] if len(ins) > 2 and ins[2] is not None else None
        xf = x.astype(F32)
        has_zp = zp is not None and np.any(zp)
        if has_zp:
            xf = (xf - _axis_view(np.asarray(zp), x.ndim, axis).astype(F32)).astype(F32)
        out = (xf * s).astype(F32)
        if x.dtype == np.int8 and not has_zp and x.ndim >= 2:
            kind = _scale_kind(s, x)
            if kind == "rowb":
                s, kind = np.broadcast_to(s, x.shape[:-1] + (1,)), "row"
            if kind in ("scalar", "row", "col"):
                prov[node.output[0]] = ("dq", x, np.ascontiguousarray(s).reshape(-1), kind)
        return out
    if op == "MatMulInteger":
        # sum_k (A - a_zp)(B - b_zp) in int32 (wrap-around); a_zp scalar or per row, b_zp scalar or per column
        a, b = ins[0].astype(np.int64), ins[1].astype(np.int64)
        if len(ins) > 2 and ins[2] is not None:
            az = np.asarray(ins[2]).astype(np.int64)
            a = a - (az.reshape(-1, 1) if az.size > 1 else az)
        if len(ins) > 3 and ins[3] is not None:
            bz = np.asarray(ins[3]).astype(np.int64)
            b = b - (bz.reshape(1, -1) if bz.size > 1 else bz)
        return np.matmul(a, b).astype(np.int32)
    if op == "QLinearMatMul":
        a, a_s, a_z, b, b_s, b_z, y_s, y_z = (list(ins) + [None] * 8)[:8]
        ai = a.astype(np.int64) - (0 if a_z is None else (np.asarray(a_z).astype(np.int64).reshape(-1, 1) if np.asarray(a_z).size > 1 else np.asarray(a_z).astype(np.int64)))
        bi = b.astype(np.int64) - (0 if b_z is None else (np.asarray(b_z).astype(np.int64).reshape(1, -1) if np.asarray(b_z).size > 1 else np.asarray(b_z).astype(np.int64)))
        acc = np.matmul(ai, bi).astype(np.int32)
        sa = np.asarray(a_s, dtype=F32).reshape(-1, 1) if np.asarray(a_s).size > 1 else F32(a_s)
        sb = np.asarray(b_s, dtype=F32).reshape(1, -1) if np.asarray(b_s).size > 1 else F32(b_s)
        y = ((acc.astype(F32) * sa).astype(F32) * sb).astype(F32)
        yz = np.zeros((), np.int8) if y_z is None else np.asarray(y_z)
        q = (np.rint((y / F32(y_s)).astype(F32)) + yz.astype(F32)).astype(F32)
        info = np.iinfo(yz.dtype)
        return np.clip(q, info.min, info.max).astype(yz.dtype)
    if op == "MatMul":
        a, b = ins[0], ins[1]
        pa, pb = prov.get(node.input[0]), prov.get(node.input[1])
        if mode == "int-exact" and pa is not None and pb is not None and pa[0] == "dq" and pb[0] == "dq" and b.ndim == 2 and \
                pa[3] in ("row", "scalar") and pb[3] in ("col", "scalar") and pa[1].shape[-1] % 16 == 0 and pb[1].shape[1] % 32 == 0:
            aq = pa[1].reshape(-1, pa[1].shape[-1])
            acc = ox.int_matmul(aq, np.ascontiguousarray(pb[1].T))
            sa = np.broadcast_to(pa[2], (aq.shape[0],)) if pa[2].size == 1 else pa[2]
            sw = np.broadcast_to(pb[2], (pb[1].shape[1],)) if pb[2].size == 1 else pb[2]
            return ox.linear_epilogue(acc, sa, sw).reshape(a.shape[:-1] + (pb[1].shape[1],))
        if mode == "int-exact" and pa is not None and pb is not None and pa[0] == "qs" and pb[0] == "qsT" and b.ndim == 2:
            aq = pa[1].reshape(-1, pa[1].shape[-1]).astype(np.int8)
            acc = ox.int_matmul(aq, pb[1].astype(np.int8))
            out = ox.linear_epilogue(acc, pa[2], pb[2])
            return out.reshape(a.shape[:-1] + (pb[1].shape[0],))
        return np.matmul(a, b).astype(F32)
    raise NotImplementedError("oracle: op %s" % op)


# ---------------------------------------------------------------------------------------------- reference API
def expand_node_inputs_outputs(graph, node, weight_dict, module):
    """onnx_optimized_inference.py:236-271."""
    start = time.time()
    added_inputs = [n for n in node.input if n]
    if "Clip" in node.name and len(added_inputs) < 3:
        extra = node.input[0][:-1] + "2"                     # :250  name[:-1] + "2"
        weight_dict[extra] = np.array(FLOAT_MAX, dtype=F32)  # :251
        added_inputs.append(extra)
    return added_inputs, list(node.output), time.time() - start


def _inputs(node, weight_dict, added):
    ins = []
    for pos, name in enumerate(node.input):
        if name == "":
            ins.append(weight_dict[added[-1]] if (node.op_type == "Clip" and pos == 2) else None)
        else:
            ins.append(weight_dict[name])
    return ins


def execute_node(node, main_graph, final_output_node, weight_dict, module, inject_parameters=None, mode="ref-float"):
    """onnx_optimized_inference.py:18-212."""
    prov = weight_dict.setdefault("__prov__", {})
    added, _, op_time = expand_node_inputs_outputs(main_graph, node, weight_dict, module)
    ins = _inputs(node, weight_dict, added)
    out = run_node(node, ins, prov, mode)
    name = node.output[0]
    weight_dict[name] = out
    output_tensors = {name: out}
    p = inject_parameters

    if p and ("RANDOM" in p["inject_type"]) and (node.name == p["faulty_operation_name"]):       # :59-72
        d = Draws(p)
        idx = tuple(d.indices("target_indices", out.shape))
        if "BITFLIP" in p["inject_type"]:
            faulty = ox.float32_bit_flip(out[idx], d.randint("flip_bit", 0, 32))
        else:
            faulty = ox.bits_to_float32(d.bits32("random_bits"))
        out[idx] = faulty
        prov.pop(name, None)

    if p and (module in p["targetted_module"]) and p["faulty_trace"] and (node.name == p["faulty_trace"][0]) and \
            (p["inject_type"] in ["INPUT", "WEIGHT", "INPUT16", "WEIGHT16"]):                     # :74
        faulty_operation = p["faulty_trace"][0]
        d = Draws(p)
        if p["faulty_tensor_name"] in node.input:                                                 # :78-81
            assert p["faulty_quantizer_name"] == p["faulty_trace"][0]
            golden = weight_dict[p["faulty_tensor_name"]]
            idx = tuple(d.indices("target_indices", golden.shape))                                # layers.py:73
            faulty_value = ox.flip_int8_bit(int(np.int8(golden[idx])), p["faulty_bit_position"])  # layers.py:72,77
            assert -128 <= faulty_value <= 127
            one_hot = np.zeros(golden.shape, dtype=golden.dtype)                                  # layers.py:105-107
            one_hot[idx] = faulty_value
            pert = list(ins)
            pert[list(node.input).index(p["faulty_tensor_name"])] = one_hot
            delta = run_node(node, pert, {}, "ref-float").copy()                                  # layers.py:134-135
            delta[idx] = delta[idx] - weight_dict[name][idx]                                      # layers.py:139-140
            weight_dict["delta_4d"] = delta
            p["intermediate_output_name"] = name
        else:                                                                                     # :84-104
            pos = [k for k, n in enumerate(node.input) if n == p["intermediate_output_name"]]
            assert pos
            pert = list(ins)
            pert[pos[-1]] = weight_dict["delta_4d"]
            weight_dict["delta_4d"] = run_node(node, pert, {}, "ref-float")
            p["intermediate_output_name"] = name
        if faulty_operation == p["faulty_operation_name"]:                                        # :107-199
            assert len(p["faulty_trace"]) == 1
            delta = weight_dict["delta_4d"]
            if p["inject_type"] == "INPUT16":
                delta = _window(delta, 3, d, random_len=False)
            elif p["inject_type"] == "WEIGHT16":
                delta = _window(delta, 2, d, random_len=True)
            weight_dict["delta_4d"] = delta
            faulty = np.add(weight_dict[name], delta).astype(F32)                                 # :191
            weight_dict[name] = faulty
            output_tensors[name] = faulty
            prov.pop(name, None)
        p["faulty_trace"] = p["faulty_trace"][1:]                                                 # :204
    return output_tensors, weight_dict, op_time


def _window(delta, axis, d, random_len):
    """onnx_optimized_inference.py:111-139 (INPUT16: shape[3]) / :156-179 (WEIGHT16: shape[2], randint(1,16) rows)."""
    shape = list(delta.shape)
    blocks = shape[axis] // 16
    start = 0 if blocks == 0 else d.randint("window_start", 0, blocks)
    start *= 16
    out = np.zeros(delta.shape, dtype=F32)
    nz = np.nonzero(delta)
    if len(nz[0]) == 0:
        return out
    n = d.randint("window_len", 1, 16) if random_len else 16
    index = [int(a[0]) for a in nz]
    index[axis] = start
    for i in range(n):
        if i >= shape[axis] or index[axis] >= shape[axis]:
            break
        out[tuple(index)] = delta[tuple(index)]
        index[axis] += 1
    return out


def inference(main_graph, weight_dict, module, inject_parameters=None, mode="ref-float"):
    """onnx_optimized_inference.py:214-234."""
    output_tensors = None
    for node in main_graph.node:
        output_tensors, weight_dict, _ = execute_node(node, main_graph, node.output[0], weight_dict, module, inject_parameters, mode)
    return output_tensors, weight_dict


def get_weight_dict(graph):
    """onnx_optimized_inference.py:273-280."""
    return graph, {i.name: np.array(i.array) for i in graph.initializer}


def prepare_inference(graph, module_input_values):
    """onnx_optimized_inference.py:282-295."""
    graph, wd = get_weight_dict(graph)
    for v in graph.input:
        wd[v.name] = np.asarray(module_input_values[v.name])
    return wd, graph


def run_module(module, input_values, module_filepath, module_weight_dict, module_graph, inject_parameters=None, mode="ref-float"):
    """onnx_optimized_inference.py:297-304."""
    for k in list(input_values.keys()):
        module_weight_dict[k] = np.asarray(input_values[k])
    module_weight_dict.pop("__prov__", None)
    return inference(module_graph, module_weight_dict, module, inject_parameters, mode)


def greedy_decode(float_weights, encoder_graph, decoder_graph, src_ids, src_mask, max_len, start_symbol=0, mode="ref-float", timings=None):
    """8-bit_onnx_optimized_custom_inference.py:649-721 on this oracle's node walk (custom_decoder=True): host embedding
    (embeddings.py:13, positional_encodings.py:24), run_module("Encoder"), then max_len - 1 x run_module("Decoder") on the FULL
    prefix with subsequent_mask (utils.py:10-14), generator + arg-max (generator.py:14-15).  Returns (ys, margins)."""
    import time
    t0 = time.perf_counter()
    B = src_ids.shape[0]
    pe = ox.positional_encoding(max(max_len, src_ids.shape[1]) + 1)
    src_float = ox.embed(src_ids, float_weights["src_embed.0.lut.weight"], pe)
    enc_in = {"global_in": src_float, "global_in_1": np.asarray(src_mask)}
    wd, _ = prepare_inference(encoder_graph, enc_in)
    memory, _ = run_module("Encoder", enc_in, None, wd, encoder_graph, None, mode)
    memory = memory[list(memory.keys())[0]]
    ys = np.full((B, 1), start_symbol, dtype=np.int64)
    dwd = None
    margins = []
    if timings is not None:
        timings["encoder_s"] = time.perf_counter() - t0
        timings["decoder_step_s"] = []
    for _ in range(max_len - 1):
        t1 = time.perf_counter()
        T = ys.shape[1]
        dec_in = {"global_in": ox.embed(ys, float_weights["tgt_embed.0.lut.weight"], pe), "global_in_1": memory, "global_in_2": np.asarray(src_mask),
                  "global_in_3": (np.triu(np.ones((1, T, T)), k=1) == 0).astype(np.int64)}
        if dwd is None:
            dwd, _ = prepare_inference(decoder_graph, dec_in)
        out, _ = run_module("Decoder", dec_in, None, dwd, decoder_graph, None, mode)
        out = out[list(out.keys())[0]]
        nxt, logits = ox.generator(out[:, -1], float_weights["generator.proj.weight"], float_weights["generator.proj.bias"])
        srt = np.sort(logits, axis=-1)
        margins.append(srt[:, -1] - srt[:, -2])
        ys = np.concatenate([ys, nxt.reshape(B, 1).astype(np.int64)], axis=1)
        if timings is not None:
            timings["decoder_step_s"].append(time.perf_counter() - t1)
    return ys, np.stack(margins, axis=1)
