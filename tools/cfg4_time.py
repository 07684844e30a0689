"""cfg4 (4-bit weights): batch-64 greedy decode and encoder-only times (64 x 64 and 512 x 128), next to the 8-bit engine."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnx_transformer_b200 import weights as W  # noqa: E402
from onnx_transformer_b200.engine import QuantizedTransformer  # noqa: E402

fw = W.init_float_weights(0)
for bits in (8, 4):
    eng = QuantizedTransformer(fw, n_layers=6, max_len=72, weight_bits=bits)
    ids, mask = W.synthetic_tokens(11, 64, 64)
    ids, mask = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
    for _ in range(2):
        eng.greedy_decode(ids, mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        eng.greedy_decode(ids, mask)
    e1.record()
    torch.cuda.synchronize()
    line = "%d-bit weights: decode 64 x 64 %.3f ms" % (bits, e0.elapsed_time(e1) / 3)
    for B, S in [(64, 64), (512, 128)]:
        i2, m2 = W.synthetic_tokens(7, B, S)
        i2, m2 = torch.from_numpy(i2).cuda(), torch.from_numpy(m2).cuda()
        for _ in range(2):
            eng.encode(i2, m2)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            eng.encode(i2, m2)
        e1.record()
        torch.cuda.synchronize()
        line += "   encode %d x %d %.3f ms" % (B, S, e0.elapsed_time(e1) / 5)
    print(line, flush=True)
    del eng
    torch.cuda.empty_cache()
