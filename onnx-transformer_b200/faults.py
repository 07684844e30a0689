"""Host-side fault-model helpers: the reference's inject_utils restated with explicit, replayable random draws.

Follows inject_utils/layers.py:7-33 (fp32 bit codecs, float32_bit_flip, delta_init), :48-84 (flip_int4_bit,
flip_int8_bit, int_bit_flip) and inject_utils/utils.py:165-246 (get_target_inputs).  Scalar bit arithmetic only --
tensors never come to the host here; the tensor work of a fault (one-hot delta, trace propagation, output patch) runs
through the CUDA handlers of executor.py or inside the fused kernels' epilogues.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Sequence

import numpy as np


# ---------------------------------------------------------------------------------------------- scalar bit flips
def flip_int8_bit(value: int, bit_position: int) -> int:
    """inject_utils/layers.py:61-68, on Python ints (np.int8 ^ 128 raises under NumPy >= 2)."""
    flipped = int(value) ^ (1 << int(bit_position))
    if flipped > 127:
        flipped -= 256
    if flipped < -128:
        flipped += 256
    return flipped


def flip_int4_bit(value: int, bit_position: int) -> int:
    """inject_utils/layers.py:48-59."""
    flipped = int(value) ^ (1 << int(bit_position))
    if flipped > 7:
        flipped -= 16
    if flipped < -8:
        flipped += 16
    return flipped


def fp32tobin(value: float) -> str:
    """inject_utils/layers.py:7-8: big-endian bit string, index 0 = sign bit."""
    return "".join(bin(c).replace("0b", "").rjust(8, "0") for c in struct.pack("!f", float(value)))


def bin2fp32(bin_str: str) -> float:
    """inject_utils/layers.py:10-16: NaN -> 0."""
    assert len(bin_str) == 32
    data = struct.unpack("!f", struct.pack("!I", int(bin_str, 2)))[0]
    return 0.0 if data != data else data


def float32_bit_flip_value(golden_value: float, flip_bit: int) -> float:
    """inject_utils/layers.py:24-33 with the drawn bit explicit (bit 0 = LSB)."""
    s = fp32tobin(golden_value)
    pos = 31 - int(flip_bit)
    s = s[:pos] + ("0" if s[pos] == "1" else "1") + s[pos + 1:]
    return bin2fp32(s)


def delta_init_value(bits: int) -> float:
    """inject_utils/layers.py:18-22: a random 32-bit pattern as fp32 (NaN -> 0), pattern given explicitly."""
    return bin2fp32(bin(int(bits) & 0xFFFFFFFF)[2:].rjust(32, "0"))


# ---------------------------------------------------------------------------------------------- random draws
class Draws:
    """The reference draws from the global np.random at five places (inject_utils/layers.py:73, 21, 27;
    onnx_optimized_inference.py:63, 118, 160, 167).  A trial carries them explicitly in
    inject_parameters["rng_draws"]; anything missing is drawn from np.random exactly as the reference would and
    recorded in inject_parameters["rng_draws_used"] so the trial can be replayed."""

    def __init__(self, inject_parameters: dict):
        self.given = dict(inject_parameters.get("rng_draws") or {})
        self.used = inject_parameters.setdefault("rng_draws_used", {})

    def indices(self, key: str, shape: Sequence[int]) -> List[int]:
        if key in self.given:
            idx = [int(i) for i in self.given[key]]
        else:
            idx = [int(np.random.randint(0, dim)) for dim in shape]
        self.used[key] = idx
        return idx

    def randint(self, key: str, low: int, high: int) -> int:
        v = int(self.given[key]) if key in self.given else int(np.random.randint(low, high))
        self.used[key] = v
        return v

    def bits32(self, key: str) -> int:
        if key in self.given:
            v = int(self.given[key])
        else:
            v = int("".join(str(np.random.randint(0, 2)) for _ in range(32)), 2)
        self.used[key] = v
        return v


# ---------------------------------------------------------------------------------------------- trace discovery
def get_target_inputs(graph, layer_name: str, input_name: str, weight_name: str, bias_name, output_tensor: str):
    """inject_utils/utils.py:165-246: from a fault-target JSON (target_layer, input_tensor, weight_tensor, output_tensor)
    find, for the input and the weight operand, the quantizer node (first consumer of the Round tensor) and the chain
    of node names from it to the target MatMul, e.g. ['Mul_84', 'Reshape_14', 'Transpose_49', 'MatMul_28'].
    Returns ((input_quantizer, input_tensor), (weight_quantizer, weight_tensor), (None, None), (input_trace, weight_trace))."""
    layer_node = None
    for node in graph.node:
        if node.name == layer_name:
            layer_node = node
    if layer_node is None:
        raise SystemExit("get_target_inputs: target layer %r not in graph" % layer_name)

    def chase(tensor_name: str):
        quantizer = None
        for node in graph.node:
            if tensor_name in node.input:
                quantizer = node
                break
        if quantizer is None:
            return None, None, None
        int_tensor = None
        for t in quantizer.input:
            if "out0" in t:
                int_tensor = t
                break
        names = [quantizer.name]
        current = quantizer.output[0]
        for outer in graph.node:
            if names[-1] == layer_name:
                break
            if current in outer.input:
                names.append(outer.name)
                current = outer.output[0]
        return quantizer, int_tensor, names

    q_in, t_in, trace_in = chase(input_name)
    q_w, t_w, trace_w = chase(weight_name)
    check_1 = q_in is not None and q_w is not None and t_in in q_in.input and t_w in q_w.input
    check_2 = output_tensor in layer_node.output
    if not (check_1 and check_2):
        # the reference prints and exit()s (inject_utils/utils.py:241-245)
        raise SystemExit("get_target_inputs: inconsistent target %s / %s / %s" % (input_name, weight_name, output_tensor))
    return (q_in.name, t_in), (q_w.name, t_w), (None, None), (trace_in, trace_w)


def build_inject_parameters(graph, target: Dict[str, str], fault_model: str, bit_position: Optional[int], experiment_output_file: str = "",
                            target_inference_number: int = 1, rng_draws: Optional[dict] = None) -> dict:
    """The inject_parameters dict of parallelized_inject_onnx_transformer.py:803-858 for one (target JSON, fault model,
    bit) combination."""
    (in_q, in_t), (w_q, w_t), _, (in_trace, w_trace) = get_target_inputs(graph, target["target_layer"], target["input_tensor"],
                                                                           target["weight_tensor"], None, target["output_tensor"])
    faulty_trace, quantizer, tensor = None, None, target["output_tensor"]
    if "INPUT" in fault_model:
        faulty_trace, quantizer, tensor = list(in_trace), in_q, in_t
    elif "WEIGHT" in fault_model:
        faulty_trace, quantizer, tensor = list(w_trace), w_q, w_t
    p = {
        "inject_type": fault_model,
        "faulty_tensor_name": tensor,
        "faulty_quantizer_name": quantizer,
        "faulty_trace": faulty_trace,
        "faulty_output_tensor": target["output_tensor"],
        "faulty_operation_name": target["target_layer"],
        "targetted_module": target["module"],
        "target_inference_number": target_inference_number,
        "experiment_output_file": experiment_output_file,
        "faulty_bit_position": None if "RANDOM" in fault_model else bit_position,
    }
    if rng_draws is not None:
        p["rng_draws"] = rng_draws
    return p


# ---------------------------------------------------------------------------------------------- fault-target descriptors
_MODULE_TAG = {"qk": "FirstMatMul", "cqk": "FirstMatMul", "pv": "SecondMatMul", "cpv": "SecondMatMul", "ffn1": "FirstFC", "ffn2": "SecondFC"}


def targets_from_graph(graph, module: str, model_name: Optional[str] = None) -> List[Dict[str, str]]:
    """Re-create the reference's fault-target descriptors (input/encoder/matmul_*.json x24, input/decoder/matmul_*.json
    x36: the attention MatMuls and the two FFN MatMuls of every layer) from the graph itself.  Each dict has the
    reference's keys: target_layer, input_tensor, weight_tensor, bias_tensor, output_tensor, module, model_name."""
    producers = {}
    for n in graph.node:
        for o in n.output:
            producers[o] = n

    def round_behind(tensor: str) -> str:
        # walk back through Mul / Cast / Div / Reshape / Transpose to the Round node that defines the integer tensor
        seen = 0
        while seen < 16:
            n = producers.get(tensor)
            if n is None:
                raise KeyError(tensor)
            if n.op_type == "Round":
                return n.output[0]
            nxt = None
            for i in n.input:
                if i in producers:
                    nxt = i
                    break
            tensor = nxt
            seen += 1
        raise KeyError("no Round behind tensor")

    out = []
    for n in graph.node:
        role = graph.roles.get(n.name)
        if n.op_type != "MatMul" or role is None or role.get("role") not in _MODULE_TAG:
            continue
        out.append({
            "target_layer": n.name,
            "input_tensor": round_behind(n.input[0]),
            "weight_tensor": round_behind(n.input[1]),
            "bias_tensor": "None",
            "output_tensor": n.output[0],
            "module": "%s/%s" % (module, _MODULE_TAG[role["role"]]),
            "model_name": model_name or ("./try/%s_try_cleaned.onnx" % module.lower()),
        })
    return out
