import sys, torch
sys.path.insert(0, "/root/repo")
from onnx_transformer_b200 import weights as W
from onnx_transformer_b200.engine import QuantizedTransformer
eng = QuantizedTransformer(W.init_float_weights(0))
for B in [64, 72, 80, 88, 96, 104, 112, 120]:
    ids, mask = W.synthetic_tokens(1000, B, 64)
    ids, mask = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
    for _ in range(2): eng.greedy_decode(ids, mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): eng.greedy_decode(ids, mask)
    e1.record(); torch.cuda.synchronize()
    print("B=%3d  %.2f ms per decode  persistent_steps=%d" % (B, e0.elapsed_time(e1) / 3, eng.persistent_steps), flush=True)
