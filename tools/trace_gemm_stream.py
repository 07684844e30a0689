"""Per-tile phase timeline of CTA 0 of the persistent GEMM (OT_GEMM_STREAM_TRACE): python tools/trace_gemm_stream.py [qkv|ffn1] [M]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnx_transformer_b200 import kernels as K  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "qkv"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
N, Kd, kw = {"qkv": (1536, 512, dict(out_kind=K.OUT_Q8, quant_group=512)), "ffn1": (2048, 512, dict(out_kind=K.OUT_Q8, quant_group=2048, relu=True))}[which]
dev = torch.device("cuda")
a = torch.randint(-127, 128, (M, Kd), dtype=torch.int8, device=dev)
w = torch.randint(-127, 128, (N, Kd), dtype=torch.int8, device=dev)
sx = torch.rand(M, device=dev) * 0.05 + 1e-3
sw = torch.rand(N, device=dev) * 0.01 + 1e-4
b = torch.randn(N, device=dev)
K.linear_w8a8(a, w, row_scale=sx, col_scale=sw, bias=b, **kw)
torch.cuda.synchronize()
trace = torch.zeros(32 * 16 + 8 * 16 * 8, dtype=torch.int64, device=dev)
os.environ["OT_GEMM_STREAM_TRACE"] = hex(trace.data_ptr())
K.linear_w8a8(a, w, row_scale=sx, col_scale=sw, bias=b, **kw)
torch.cuda.synchronize()
del os.environ["OT_GEMM_STREAM_TRACE"]
full = trace.cpu().numpy()
t = full[:512].reshape(32, 16)
t0 = t[0][t[0] > 0].min()
print("tile | epi: start acc_ready pass1 xchg pass2 | mma: free issued | tma: first last   (us since first stamp)")
for i in range(14):
    r = [(x - t0) / 1e3 if x > 0 else -1 for x in t[i]]
    print("%3d  | %7.2f %7.2f %7.2f %7.2f %7.2f | %7.2f %7.2f | %7.2f %7.2f" % (i, r[0], r[1], r[2], r[3], r[4], r[8], r[9], r[12], r[13]))

w = full[512:].reshape(8, 16, 8)
for li in (3, 4):
    print("tile %d: per epilogue warp (e: SMSP) start acc pass1 xchg pass2" % li)
    for e in range(16):
        r = [(x - t0) / 1e3 if x > 0 else -1 for x in w[li, e, :5]]
        print("   e=%2d smsp=%d  %7.2f %7.2f %7.2f %7.2f %7.2f" % (e, (e + 2) % 4, *r))
