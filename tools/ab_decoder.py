"""Development aid: time the batch-64 greedy decode with the library OT_B200_LIB points at and print a digest of the tokens
(A/B runs of kernel variants: `for v in v0 k0 k1; do OT_B200_LIB=.../libot_b200_$v.so python tools/ab_decoder.py $v; done`)."""
import hashlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnx_transformer_b200 import weights as W  # noqa: E402
from onnx_transformer_b200.engine import QuantizedTransformer  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "?"
fw = W.init_float_weights(0)
ids, mask = W.synthetic_tokens(1000, 64, 64)
ids, mask = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
eng = QuantizedTransformer(fw)
for _ in range(3):
    ys = eng.greedy_decode(ids, mask)
ws = eng._dec_workspace(64, 64)
plan = eng._decoder_plan(ws, 64, 64)
best = 1e9
for rep in range(5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    plan.run(0, 71)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("%-6s decoder launch (71 steps): %.3f ms = %.1f us/step   tokens %s" % (tag, best, best * 1e3 / 71, hashlib.md5(ys.cpu().numpy().tobytes()).hexdigest()[:10]), flush=True)
