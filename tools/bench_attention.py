"""Event-timed encoder attention at cfg3 (B = 512 sentences x S = 128 tokens, 8 heads): the fused tensor-core kernel with int8 + scale
output (RowQuant of the merged rows inside, cluster of 8 head CTAs) and with the fp32 context only.  python tools/bench_attention.py [B] [S]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnx_transformer_b200 import kernels as K  # noqa: E402

B, S = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (512, 128)
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
M = B * S
qkv = torch.randint(-127, 128, (M, 1536), dtype=torch.int8, device=dev, generator=g)
sqkv = torch.rand(M, 3, device=dev, generator=g) * 0.02 + 1e-3
mask = torch.ones(B, S, dtype=torch.uint8, device=dev)
mask[:, S - 9:] = 0
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
ctx = torch.empty(M, 512, device=dev)
cq = torch.empty(M, 512, dtype=torch.int8, device=dev)
cs = torch.empty(M, device=dev)


def run(want_ctx, want_q):
    K.attention_q8(qkv, sqkv, qkv[:, 512:], qkv[:, 1024:], sqkv[:, 1:], sqkv[:, 2:], B=B, Tq=S, Tk=S, ldq=1536, sq_stride=3, ldk=1536, skv_stride=3,
                   mask_kind=1, key_mask=mask, mask_stride=S, want_ctx=want_ctx, ctx=ctx if want_ctx else None, want_q=want_q,
                   ctx_q=cq if want_q else None, ctx_s=cs if want_q else None)


for name, wc, wq in (("int8 + scale (fused RowQuant)", False, True), ("fp32 context", True, False), ("both", True, True)):
    for _ in range(2):
        run(wc, wq)
    ts = []
    for _ in range(5):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            run(wc, wq)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / 4)
    us = sorted(ts)[2]
    macs = B * 8 * S * S * 64 * 2
    print("%-32s %8.1f us   %6.1f T MAC/s (QK^T + PV)" % (name, us, macs / us / 1e6))
