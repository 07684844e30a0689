"""ORACLE (test infrastructure, never imported by the product): numpy restatement of the reference's dialect-B executor
inject_operations.py (Brevitas-QCDQ graphs, driven by inject_main.py:339-385, 403-443) and of the fault helpers it calls.

Follows inject_operations.py:11-108 (execute_node: run the node, then -- in this order -- the DequantizeLinear-keyed operand hook
:59-61, the RANDOM / RANDOM_BITFLIP hook :63-74, the MatMul delta hook :76-104; 4-tuple return :108), :110-118 (inference),
:174-190 (run_module) and inject_utils/layers.py:48-59 (flip_int4_bit), :61-68 (flip_int8_bit), :70-84 (int_bit_flip), :87-142
(perturb_quantizer arithmetic, in the 5-argument / 6-result form inject_operations.py:60 calls) and :174-185 (perturb_matmul with
transposed_axes).  Nodes are evaluated by oracle/executor.run_node (ONNX opset-13 semantics).

PARITY UNPINNED: the reference ships no dialect-B graph, checkpoint or golden vector, and inject_operations.py raises TypeError as
shipped; this restatement follows the call sites and the ONNX operator specification.
"""
from __future__ import annotations

import sys

import numpy as np

from . import executor as oe
from . import intexact as ox

F32 = np.float32


def int_bit_flip(weight_dict, target_tensor, target_bit_position, bit_precision, draws):
    """inject_utils/layers.py:70-84."""
    tensor = weight_dict[target_tensor]
    idx = draws.indices("target_indices", tensor.shape)
    golden = int(np.int8(tensor[tuple(idx)]))                              # :72 np.int8(faulty_tensor)
    if bit_precision == 4:
        faulty = ox.flip_int4_bit(golden, target_bit_position)             # :48-59 (the commented-out branch of :80-83)
        assert -8 <= faulty <= 7
    else:
        faulty = ox.flip_int8_bit(golden, target_bit_position)             # :77
        assert -128 <= faulty <= 127
    return faulty, idx


def perturb_quantizer(node, ins, weight_dict, faulty_tensor_name, faulty_bit_position, inject_input):
    """inject_utils/layers.py:87-142 on the de-quantizing node."""
    faulty_value, idx = int_bit_flip(weight_dict, faulty_tensor_name, faulty_bit_position, int(inject_input.get("bit_width", 8)), oe.Draws(inject_input))
    idx = tuple(idx)
    golden = weight_dict[faulty_tensor_name]
    one_hot = np.zeros(golden.shape, dtype=golden.dtype)                   # :105-107
    one_hot[idx] = faulty_value
    pert = list(ins)
    pert[list(node.input).index(faulty_tensor_name)] = one_hot
    delta = oe.run_node(node, pert, {}, "ref-float").copy()                # :134-135
    name = node.output[0]
    delta[idx] = delta[idx] - weight_dict[name][idx]                       # :139-140
    weight_dict["delta_4d"] = delta
    return weight_dict, name, list(idx), int(golden[idx]), faulty_value, ("Unsigned" if golden.dtype == np.uint8 else "Signed")


def perturb_matmul(node, ins, weight_dict, input_tensor_name, transposed_axes=None):
    """inject_utils/layers.py:174-185."""
    if transposed_axes is not None and transposed_axes.input[0] in input_tensor_name:
        perm = list(oe._attr(transposed_axes, "perm"))
        input_tensor_name = transposed_axes.output[0]
        weight_dict["delta_4d"] = np.transpose(weight_dict["delta_4d"], tuple(perm))
    pert = list(ins)
    pert[list(node.input).index(input_tensor_name)] = weight_dict["delta_4d"]
    return oe.run_node(node, pert, {}, "ref-float")


def execute_node(node, main_graph, final_output_node, weight_dict, module, inject_input, mode="ref-float"):
    """inject_operations.py:11-108."""
    prov = weight_dict.setdefault("__prov__", {})
    added, _, op_time = oe.expand_node_inputs_outputs(main_graph, node, weight_dict, module)
    ins = oe._inputs(node, weight_dict, added)
    original = oe.run_node(node, ins, prov, mode)
    name = node.output[0]
    weight_dict[name] = original
    output_tensors = {name: original}
    p = inject_input
    if p:
        if ("RANDOM" not in p["inject_type"]) and (node.op_type == "DequantizeLinear") and (p["faulty_quantizer_name"] in node.name):   # :59
            weight_dict, dq_name, _, _, _, _ = perturb_quantizer(node, ins, weight_dict, p["faulty_tensor_name"], p["faulty_bit_position"], p)
            p["dequantized_operation_input_name"] = dq_name
        if "RANDOM" in p["inject_type"]:                                                                                                # :63
            if p["faulty_operation_name"] in node.name:
                d = oe.Draws(p)
                target = weight_dict[p["faulty_tensor_name"]]
                idx = tuple(d.indices("target_indices", target.shape))
                if "BITFLIP" in p["inject_type"]:
                    target[idx] = ox.float32_bit_flip(target[idx], d.randint("flip_bit", 0, 32))
                else:
                    target[idx] = ox.bits_to_float32(d.bits32("random_bits"))
                prov.pop(p["faulty_tensor_name"], None)
        if "INPUT" in p["inject_type"] or "WEIGHT" in p["inject_type"]:                                                                 # :76
            if (node.op_type == "MatMul") and (node.name == p["faulty_operation_name"]):
                if not p.get("dequantized_operation_input_name"):
                    print("Error with dequantized value")
                    sys.exit(0)
                delta = perturb_matmul(node, ins, weight_dict, p["dequantized_operation_input_name"], p.get("transposed_axes"))
                result = np.add(original, delta).astype(F32)                                                                            # :100
                output_tensors[name] = result
                weight_dict[name] = result
                prov.pop(name, None)
    return output_tensors, weight_dict, op_time, inject_input


def inference(main_graph, weight_dict, module, inject_input, mode="ref-float"):
    """inject_operations.py:110-118."""
    output_tensors = None
    for node in main_graph.node:
        output_tensors, weight_dict, _, inject_input = execute_node(node, main_graph, node.output[0], weight_dict, module, inject_input, mode)
    return output_tensors, weight_dict


def run_module(module, input_values, module_filepath, module_weight_dict, module_graph, inject_input=None, mode="ref-float"):
    """inject_operations.py:174-190."""
    for k in list(input_values.keys()):
        module_weight_dict[k] = np.asarray(input_values[k])
    module_weight_dict.pop("__prov__", None)
    return inference(module_graph, module_weight_dict, module, inject_input, mode)
