"""Encoder pass (+ cross-attention K/V projection) timing at the headline shape (64 x 64) and at cfg3 (512 x 128)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnx_transformer_b200 import weights as W  # noqa: E402
from onnx_transformer_b200.engine import QuantizedTransformer  # noqa: E402

fw = W.init_float_weights(0)
eng = QuantizedTransformer(fw)
for B, S in [(64, 64), (512, 128)]:
    ids, mask = W.synthetic_tokens(1000, B, S)
    ids, mask = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
    for _ in range(3):
        eng.encode(ids, mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        eng.encode(ids, mask)
    e1.record()
    torch.cuda.synchronize()
    print("encode %d x %d: %.3f ms" % (B, S, e0.elapsed_time(e1) / reps), flush=True)
