"""Per-tile phase timeline of CTA 0 of the weight-stationary GEMM (OT_GEMM_WRES_TRACE): python tools/trace_gemm_wres.py [qkv|ffn1] [M]"""
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
trace = torch.zeros(16 * 16 * 8 + 2 * 160, dtype=torch.int64, device=dev)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
os.environ["OT_GEMM_WRES_TRACE"] = hex(trace.data_ptr())
ev0.record()
K.linear_w8a8(a, w, row_scale=sx, col_scale=sw, bias=b, **kw)
ev1.record()
torch.cuda.synchronize()
del os.environ["OT_GEMM_WRES_TRACE"]
full = trace.cpu().numpy()
t = full[:16 * 16 * 8].reshape(16, 16, 8)
cta = full[16 * 16 * 8:].reshape(160, 2)
cta = cta[cta[:, 0] > 0]
print('event time %.1f us; %d CTAs: begin spread %.2f us, end - first begin: min %.1f median %.1f max %.1f us' % (ev0.elapsed_time(ev1) * 1e3, len(cta), (cta[:, 0].max() - cta[:, 0].min()) / 1e3, (cta[:, 1].min() - cta[:, 0].min()) / 1e3, float(__import__('numpy').median(cta[:, 1] - cta[:, 0].min())) / 1e3, (cta[:, 1].max() - cta[:, 0].min()) / 1e3))
t0 = t[0][t[0] > 0].min()
names = "p2start tile_top pass1 qbar xchg p2done sempty p2half"
print("%s M=%d: CTA 0, us since the first stamp; per tile: warp e=0 | slowest warp per phase" % (which, M))
print("tile | " + names + " | (max over the 16 warps) " + names)
for i in range(12):
    r0 = [(x - t0) / 1e3 if x > 0 else -1 for x in t[i, 0, :7]]
    rm = [(t[i, :, s].max() - t0) / 1e3 for s in range(8)]
    print("%3d  | " % i + " ".join("%7.2f" % x for x in r0) + " | " + " ".join("%7.2f" % x for x in rm))
for li in (4, 5):
    print("tile %d: per epilogue warp (e: SMSP) %s" % (li, names))
    for e in range(16):
        r = [(x - t0) / 1e3 if x > 0 else -1 for x in t[li, e, :7]]
        print("   e=%2d smsp=%d cq=%d " % (e, (e + 2) % 4, e // 4) + " ".join("%7.2f" % x for x in r))
