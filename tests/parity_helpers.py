"""Shared helpers of the parity tests (test infrastructure: may import oracle/)."""
import numpy as np

from onnx_transformer_b200 import campaign as C
from onnx_transformer_b200 import weights as W


def oracle_fault(tr, S_, T_):
    """campaign.Trial -> the fault dict oracle/model.py understands (index in the reference's tensor layout)."""
    operand = "input" if tr.inject_type.startswith("INPUT") else ("weight" if tr.inject_type.startswith("WEIGHT") else "output")
    shape = C._tensor_shape(tr.module, tr.target, operand, S_, T_)
    idx = tuple(int(i) for i in np.unravel_index(tr.flat_index, shape))
    return dict(module=tr.module, layer=tr.layer, target=tr.target, type=tr.inject_type, bit=tr.bit, flat_index=tr.flat_index, index=idx,
                window_start=tr.window_start, window_len=tr.window_len, value_bits=tr.value_bits, step=0)


CFG5_SEED_TOKENS, CFG5_SEED_TRIALS, CFG5_N_TRIALS, CFG5_ALIAS_BIAS = 11, 5, 200, 0.05


def cfg5_weights(alias_token: int):
    """Full-size random-init model whose generator can emit </s>: the </s> row is an alias of `alias_token`'s row with a slightly larger
    bias, so every step that would emit `alias_token` emits </s> instead (random-init weights never emit </s> on their own, which
    made every full-size trial of round 1 the reference's "no-EOS" row)."""
    fw = W.init_float_weights(0)
    fw["generator.proj.weight"][W.EOS_ID] = fw["generator.proj.weight"][alias_token]
    fw["generator.proj.bias"][W.EOS_ID] = fw["generator.proj.bias"][alias_token] + np.float32(CFG5_ALIAS_BIAS)
    return fw


def first_divergence(a: np.ndarray, b: np.ndarray) -> int:
    """Index of the first greedy STEP (0-based: step t produced column t+1) at which two token rows differ, -1 if identical."""
    d = np.nonzero(a != b)[0]
    return int(d[0]) - 1 if len(d) else -1
