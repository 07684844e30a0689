"""One launch of each encoder GEMM shape at M rows (for ncu): python tools/ncu_gemm.py [M] [which]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnx_transformer_b200 import kernels as K  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
which = sys.argv[2] if len(sys.argv) > 2 else "all"
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
shapes = [("qkv_q8", 1536, 512, dict(out_kind=K.OUT_Q8, quant_group=512), False), ("o_f32", 512, 512, dict(out_kind=K.OUT_F32), True),
          ("ffn1_q8", 2048, 512, dict(out_kind=K.OUT_Q8, quant_group=2048, relu=True), False), ("ffn2_f32", 512, 2048, dict(out_kind=K.OUT_F32), True)]
for name, N, Kd, kw, res in shapes:
    if which not in ("all", name):
        continue
    a = torch.randint(-127, 128, (M, Kd), dtype=torch.int8, device=dev, generator=g)
    w = torch.randint(-127, 128, (N, Kd), dtype=torch.int8, device=dev, generator=g)
    sx = torch.rand(M, device=dev) * 0.05 + 1e-3
    sw = torch.rand(N, device=dev) * 0.01 + 1e-4
    b = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev) if res else None
    for _ in range(2):
        K.linear_w8a8(a, w, row_scale=sx, col_scale=sw, bias=b, residual=r, **kw)
    torch.cuda.synchronize()
