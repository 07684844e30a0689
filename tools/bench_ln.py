"""LayerNorm + RowQuant (512 features) at encoder sizes: time per launch and algorithmic GB/s (2048 B read + 516 B written per row)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnx_transformer_b200 import kernels as K  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
torch.manual_seed(0)
x = torch.randn(rows, 512, device="cuda")
g, b = torch.rand(512, device="cuda") + 0.5, torch.randn(512, device="cuda") * 0.1
q = torch.empty(rows, 512, dtype=torch.int8, device="cuda")
s = torch.empty(rows, device="cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
ts = []
for rep in range(8):
    flush.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        K.layernorm_quant(x, g, b, want_q=True, q=q, s=s)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 4 * 1e3)
us = sorted(ts)[len(ts) // 2]
print("rows %d  ctas/sm %s: %.1f us per launch  %.0f GB/s   checksum %d" % (rows, os.environ.get("OT_LN512_CTAS_PER_SM", "default"), us, rows * 2564 / us / 1e3, int(q.to(torch.int32).sum().item())))
