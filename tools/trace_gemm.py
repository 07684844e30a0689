"""Phase timestamps (globaltimer, ns) of one GEMM launch per CTA: setup | LN prologue | accumulator ready | pass 1 | cluster sync | end."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
buf = torch.zeros(4096 * 8, dtype=torch.int64, device="cuda")
os.environ["OT_GEMM_TRACE"] = hex(buf.data_ptr())
from onnx_transformer_b200 import kernels as K  # noqa: E402

M, N = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 512
group = int(sys.argv[2]) if len(sys.argv) > 2 else 512
ln = (sys.argv[3] == "ln") if len(sys.argv) > 3 else True
rng = np.random.default_rng(0)
w = torch.from_numpy(rng.integers(-127, 128, size=(N, 512), dtype=np.int8)).cuda()
a = torch.from_numpy(rng.integers(-127, 128, size=(M, 512), dtype=np.int8)).cuda()
sx = torch.rand(M, device="cuda") * 0.02 + 0.001
sw = torch.rand(N, device="cuda") * 0.01 + 1e-4
b = torch.randn(N, device="cuda")
x = torch.randn(M, 512, device="cuda")
ga, be = torch.rand(512, device="cuda") + 0.5, torch.randn(512, device="cuda") * 0.1
for it in range(4):
    buf.zero_()
    if ln:
        K.ln_linear_w8a8(x, ga, be, w, col_scale=sw, bias=b, out_kind=K.OUT_Q8, quant_group=group)
    else:
        K.linear_w8a8(a, w, row_scale=sx, col_scale=sw, bias=b, out_kind=K.OUT_Q8, quant_group=group)
    torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(-1, 8)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
print("cta  setup  ln_done  acc_ready  pass1  clsync  end   (us since first CTA start)")
for i, r in enumerate(t[:24]):
    print(i, " ".join("%7.2f" % ((v - t0) / 1e3) if v > 0 else "      -" for v in r[:6]))
