"""Event-timed int8 GEMM shapes of one cfg3 encoder layer (M = 65,536 tokens): q|k|v -> int8 (groups of 512), o -> fp32 + residual,
ffn1 -> ReLU -> int8 (group 2048), ffn2 -> fp32 + residual; weight-stationary kernel (ot_gemm_wres.cu, int8 outputs) vs the persistent streaming kernel
(ot_gemm_stream.cu, OT_GEMM_WRES=0) vs the tile kernel (OT_GEMM_STREAM=0).  L2 is flushed before every timed group of REPS back-to-back launches (the tensors of one launch, 134-400 MB, exceed L2).  Usage: python tools/bench_gemm.py [M]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnx_transformer_b200 import kernels as K  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
REPS = 4
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)


def ri8(*s):
    return torch.randint(-127, 128, s, dtype=torch.int8, device=dev, generator=g)


flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
shapes = [("qkv_q8", 1536, 512, dict(out_kind=K.OUT_Q8, quant_group=512), False), ("o_f32", 512, 512, dict(out_kind=K.OUT_F32), True),
          ("ffn1_q8", 2048, 512, dict(out_kind=K.OUT_Q8, quant_group=2048, relu=True), False), ("ffn2_f32", 512, 2048, dict(out_kind=K.OUT_F32), True)]
out = {}
for name, N, Kd, kw, res in shapes:
    a, w = ri8(M, Kd), ri8(N, Kd)
    sx = torch.rand(M, device=dev) * 0.05 + 1e-3
    sw = torch.rand(N, device=dev) * 0.01 + 1e-4
    b = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev) if res else None
    rec = {}
    for mode in (("wres", "stream", "tile") if kw["out_kind"] == K.OUT_Q8 else ("stream", "tile")):
        os.environ.pop("OT_GEMM_STREAM", None)
        os.environ.pop("OT_GEMM_WRES", None)
        if mode == "tile":
            os.environ["OT_GEMM_STREAM"] = "0"
        elif mode == "stream":
            os.environ["OT_GEMM_WRES"] = "0"
        for _ in range(2):
            K.linear_w8a8(a, w, row_scale=sx, col_scale=sw, bias=b, residual=r, **kw)
        ts = []
        for _ in range(5):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(REPS):       # back to back: a single launch between two events also times the host's launch path (~40 us of idle GPU)
                K.linear_w8a8(a, w, row_scale=sx, col_scale=sw, bias=b, residual=r, **kw)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / REPS)
        us = sorted(ts)[len(ts) // 2]
        ops = 2.0 * M * N * Kd
        nbytes = M * Kd + N * Kd + (M * N * (4 if res else 0)) + (M * N * (4 if kw["out_kind"] == K.OUT_F32 else 1))
        rec[mode] = {"us": us, "TOPs": ops / us / 1e6, "frac_4.5POPS": ops / us / 1e6 / 4500.0, "GBs": nbytes / us / 1e3}
    os.environ.pop("OT_GEMM_STREAM", None)
    os.environ.pop("OT_GEMM_WRES", None)
    out[name] = rec
print(json.dumps({"M": M, "gemm": out}, indent=1))
