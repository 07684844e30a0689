"""Encoder-only forward at BASELINE config #3 (B=512, S=128) -- target for launch lists / ncu captures."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnx_transformer_b200 import weights as W  # noqa: E402
from onnx_transformer_b200.engine import QuantizedTransformer  # noqa: E402

B, S = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (512, 128)
eng = QuantizedTransformer(W.init_float_weights(0))
ids, mask = W.synthetic_tokens(7, B, S)
ids, mask = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
for _ in range(3):
    eng.encode(ids, mask)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    eng.encode(ids, mask)
e1.record()
torch.cuda.synchronize()
print("encode B=%d S=%d: %.2f ms" % (B, S, e0.elapsed_time(e1) / 3))
