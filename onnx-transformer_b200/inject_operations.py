"""Drop-in for the reference's OTHER executor module, inject_operations.py (the dialect-B / Brevitas-QCDQ variant of the custom
node-by-node ONNX executor, driven by inject_main.py:339-385,403-443), served by the same CUDA handlers as executor.py.

Same entry points and argument meaning as the reference file:
    execute_node:11 (returns the 4-tuple incl. inject_input, :108)   inference:110   expand_node_inputs_outputs:120
    get_weight_dict:152   prepare_inference:161   run_module:174 (last argument `inject_input=None`)
and the fault helpers it calls in inject_utils/layers.py: int_bit_flip:70-84, flip_int4_bit:48-59, perturb_quantizer (the 5-argument
form inject_operations.py:60 calls, 6 results), perturb_matmul:174-185 (with `transposed_axes`).

Hook semantics (inject_operations.py:58-106), in the reference's order after the node has run and its result is stored:
  1. operand faults: at the `DequantizeLinear` node whose name contains `faulty_quantizer_name`, flip one bit of one element of the
     integer tensor `faulty_tensor_name`, de-quantize the one-hot tensor holding the faulty value with that very node and subtract
     the golden de-quantized value at that index -> weight_dict["delta_4d"]; remember the node's output name;
  2. RANDOM / RANDOM_BITFLIP: at the node whose name contains `faulty_operation_name`, overwrite one element of
     weight_dict[`faulty_tensor_name`] in place (random 32-bit pattern / one flipped fp32 bit, NaN -> 0);
  3. at the target `MatMul` (name == `faulty_operation_name`): re-run it with the de-quantized operand replaced by delta_4d
     (transposed first when `transposed_axes` -- the Transpose node between the de-quantizer and the MatMul -- applies) and add the
     result to the golden output.
As shipped, the reference file calls perturb_quantizer with a signature that no longer exists (it raises TypeError);
this module implements the call it makes.  Bit width: the dialect-B campaign uses `bit_width = 4`, `range(4)` (inject_main.py:403,
410): inject_input["bit_width"] == 4 selects flip_int4_bit, anything else flip_int8_bit (the reference's current int_bit_flip).
Random draws can be supplied in inject_input["rng_draws"] (faults.Draws), like in executor.py.

There is no CPU fallback: every node runs through executor.run_node -> libot_b200.so.
"""
from __future__ import annotations

import sys

import torch

from . import executor as X
from . import faults
from . import kernels as K

expand_node_inputs_outputs = X.expand_node_inputs_outputs      # inject_operations.py:120-150 (shape patching is moot: tensors carry shapes)
get_weight_dict = X.get_weight_dict                            # :152-159
prepare_inference = X.prepare_inference                        # :161-172


def int_bit_flip(weight_dict, target_tensor, target_bit_position, bit_precision=8, draws=None):
    """inject_utils/layers.py:70-84: (faulty integer value, random indices) for one element of the integer tensor."""
    tensor = weight_dict[target_tensor]
    draws = draws if draws is not None else faults.Draws({})
    indices = draws.indices("target_indices", tensor.shape)
    golden = int(tensor[tuple(indices)].item())
    if bit_precision == 4:
        faulty = faults.flip_int4_bit(golden, target_bit_position)
        assert -8 <= faulty <= 7
    else:
        faulty = faults.flip_int8_bit(golden if golden < 128 else golden - 256, target_bit_position)     # np.int8(tensor) of the reference
        assert -128 <= faulty <= 127
    return faulty, indices


def perturb_quantizer(node, ins, weight_dict, faulty_tensor_name, faulty_bit_position, inject_input=None):
    """The call of inject_operations.py:60 (perturb_quantizer(model, input_dict, weight_dict, faulty_tensor_name, faulty_bit_position)
    -> 6 results; arithmetic of inject_utils/layers.py:87-142): `node` + `ins` stand for the one-node model and its input_dict."""
    p = inject_input if inject_input is not None else {}
    faulty_value, target_indices = int_bit_flip(weight_dict, faulty_tensor_name, faulty_bit_position, int(p.get("bit_width", 8)), faults.Draws(p))
    idx = tuple(target_indices)
    golden_tensor = weight_dict[faulty_tensor_name]
    golden_value = int(golden_tensor[idx].item())
    is_signed = "Unsigned" if golden_tensor.dtype == torch.uint8 else "Signed"
    one_hot = torch.zeros_like(golden_tensor)
    one_hot[idx] = faulty_value if golden_tensor.dtype != torch.uint8 else (faulty_value & 0xFF)
    pert = list(ins)
    pert[list(node.input).index(faulty_tensor_name)] = one_hot
    delta = X.run_node(node, pert, {X.META_KEY: {}, X.HOST_KEY: weight_dict.get(X.HOST_KEY, {})})
    name = node.output[0]
    delta[idx] = delta[idx] - weight_dict[name][idx]
    weight_dict["delta_4d"] = delta
    return weight_dict, name, target_indices, golden_value, faulty_value, is_signed


def perturb_matmul(node, ins, weight_dict, input_tensor_name, transposed_axes=None):
    """inject_utils/layers.py:174-185: the target MatMul on delta_4d in place of the de-quantized operand."""
    if transposed_axes is not None and transposed_axes.input[0] in input_tensor_name:
        perm = list(transposed_axes.attr("perm") if hasattr(transposed_axes, "attr") else transposed_axes.attribute[0].ints)
        input_tensor_name = transposed_axes.output[0]
        weight_dict["delta_4d"] = K.transpose(weight_dict["delta_4d"], perm)
    pert = list(ins)
    pert[list(node.input).index(input_tensor_name)] = weight_dict["delta_4d"]
    return X.run_node(node, pert, {X.META_KEY: {}, X.HOST_KEY: weight_dict.get(X.HOST_KEY, {})})     # no int8 provenance: the fp32 MatMul kernel


def execute_node(node, main_graph, final_output_node, weight_dict, module, inject_input):
    """inject_operations.py:11-108."""
    added_inputs, added_outputs, list_operation_time = expand_node_inputs_outputs(main_graph, node, weight_dict, module)
    ins = X._gather_inputs(node, weight_dict, added_inputs)
    original_tensor_output = X.run_node(node, ins, weight_dict)
    tensor_output_name = node.output[0]
    weight_dict[tensor_output_name] = original_tensor_output
    output_tensors = {tensor_output_name: original_tensor_output}

    if inject_input:
        p = inject_input
        if ("RANDOM" not in p["inject_type"]) and (node.op_type == "DequantizeLinear") and (p["faulty_quantizer_name"] in node.name):
            weight_dict, dq_name, _, _, _, _ = perturb_quantizer(node, ins, weight_dict, p["faulty_tensor_name"], p["faulty_bit_position"], p)
            p["dequantized_operation_input_name"] = dq_name

        if "RANDOM" in p["inject_type"]:
            if p["faulty_operation_name"] in node.name:
                draws = faults.Draws(p)
                target = weight_dict[p["faulty_tensor_name"]]
                target_indices = tuple(draws.indices("target_indices", target.shape))
                golden_value = float(target[target_indices].item())
                if "BITFLIP" in p["inject_type"]:
                    faulty_value = faults.float32_bit_flip_value(golden_value, draws.randint("flip_bit", 0, 32))
                else:
                    faulty_value = faults.delta_init_value(draws.bits32("random_bits"))
                target[target_indices] = faulty_value                       # in place (:74)
                X._meta(weight_dict).pop(p["faulty_tensor_name"], None)

        if "INPUT" in p["inject_type"] or "WEIGHT" in p["inject_type"]:
            if (node.op_type == "MatMul") and (node.name == p["faulty_operation_name"]):
                if not p.get("dequantized_operation_input_name"):
                    print("Error with dequantized value")
                    sys.exit(0)
                delta_perturb = perturb_matmul(node, ins, weight_dict, p["dequantized_operation_input_name"], p.get("transposed_axes"))
                perturb_result = K.binary("Add", original_tensor_output, delta_perturb)
                output_tensors[tensor_output_name] = perturb_result
                weight_dict[tensor_output_name] = perturb_result
                X._meta(weight_dict).pop(tensor_output_name, None)
    return output_tensors, weight_dict, list_operation_time, inject_input


def inference(main_graph, weight_dict, module, inject_input):
    """inject_operations.py:110-118."""
    output_tensors = None
    for node in main_graph.node:
        output_tensors, weight_dict, _, inject_input = execute_node(node, main_graph, node.output[0], weight_dict, module, inject_input)
    return output_tensors, weight_dict


def run_module(module, input_values, module_filepath, module_weight_dict, module_graph, inject_input=None):
    """inject_operations.py:174-190."""
    for input_name in list(input_values.keys()):
        module_weight_dict[input_name] = X._to_device(input_values[input_name])
        module_weight_dict.get(X.HOST_KEY, {}).pop(input_name, None)
    module_weight_dict.pop(X.META_KEY, None)
    return inference(module_graph, module_weight_dict, module, inject_input)


def get_target_inputs(graph, layer_name, input_name, weight_name, bias_name, output_tensor):
    """The get_target_inputs contract inject_main.py:415 unpacks: ((input quantizer, int input tensor), (weight quantizer, int weight
    tensor), (None, None), transposed_axes) -- the quantizer is the node consuming the integer tensor (a DequantizeLinear in dialect
    B) and transposed_axes the Transpose NODE feeding the target MatMul, if any (inject_utils/utils.py:180-192 still collects it)."""
    (in_q, in_t), (w_q, w_t), _, _ = faults.get_target_inputs(graph, layer_name, input_name, weight_name, bias_name, output_tensor)
    layer = [n for n in graph.node if n.name == layer_name][0]
    transposed = None
    for n in graph.node:
        if n.op_type == "Transpose" and n.output[0] in layer.input:
            transposed = n
    return (in_q, in_t), (w_q, w_t), (None, None), transposed
