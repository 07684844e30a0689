"""Warm per-kernel timing of the GEMM variants used by a greedy step (CUDA-graph replay of 50 launches each), and a
target for `ncu --set full` captures.  usage: python tools/probe_gemm.py [variant ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnx_transformer_b200 import kernels as K  # noqa: E402

VARIANTS = {
    # name: (M, N, K, out_kind, quant_group, relu, residual)
    "qkv_q8": (64, 1536, 512, K.OUT_Q8, 512, False, False),
    "cq_q8": (64, 512, 512, K.OUT_Q8, 512, False, False),
    "ffn1_q8": (64, 2048, 512, K.OUT_Q8, 2048, True, False),
    "o_f32": (64, 512, 512, K.OUT_F32, 0, False, True),
    "ffn2_f32": (64, 512, 2048, K.OUT_F32, 0, False, True),
    "enc_qkv_q8": (4096, 1536, 512, K.OUT_Q8, 512, False, False),
    "enc_ffn1_q8": (4096, 2048, 512, K.OUT_Q8, 2048, True, False),
    "enc_ffn2_f32": (4096, 512, 2048, K.OUT_F32, 0, False, True),
    "big_ffn1_q8": (65536, 2048, 512, K.OUT_Q8, 2048, True, False),
    "big_ffn2_f32": (65536, 512, 2048, K.OUT_F32, 0, False, True),
    "big_i32": (65536, 2048, 512, K.OUT_I32, 0, False, False),
    "ln_qkv_q8": (64, 1536, 512, K.OUT_Q8, 512, False, False),
    "ln_cq_q8": (64, 512, 512, K.OUT_Q8, 512, False, False),
    "ln_ffn1_q8": (64, 2048, 512, K.OUT_Q8, 2048, True, False),
}


def main():
    names = sys.argv[1:] or list(VARIANTS)
    rng = np.random.default_rng(0)
    for name in names:
        M, N, Kd, kind, group, relu, res = VARIANTS[name]
        a = torch.from_numpy(rng.integers(-127, 128, size=(M, Kd), dtype=np.int8)).cuda()
        w = torch.from_numpy(rng.integers(-127, 128, size=(N, Kd), dtype=np.int8)).cuda()
        sx = torch.rand(M, device="cuda") * 0.02 + 0.001
        sw = torch.rand(N, device="cuda") * 0.01 + 0.0001
        b = torch.randn(N, device="cuda")
        r = torch.randn(M, N, device="cuda") if res else None
        out = torch.empty((M, N), dtype={K.OUT_I32: torch.int32, K.OUT_F32: torch.float32, K.OUT_Q8: torch.int8}[kind], device="cuda")
        osc = torch.empty((M, max(1, N // group if group else 1)), dtype=torch.float32, device="cuda")

        xf = torch.randn(M, Kd, device="cuda")
        ga = torch.rand(Kd, device="cuda") + 0.5
        be = torch.randn(Kd, device="cuda") * 0.1

        def run_ln():
            K.ln_linear_w8a8(xf, ga, be, w, col_scale=sw, bias=b, relu=relu, out_kind=kind, quant_group=group, out=out, out_scale=osc)

        def run():
            if name.startswith("ln_"):
                return run_ln()
            K.linear_w8a8(a, w, row_scale=sx, col_scale=sw, bias=b, residual=r, relu=relu, out_kind=kind, quant_group=group, out=out,
                          out_scale=osc if kind == K.OUT_Q8 else None)

        for _ in range(3):
            run()
        torch.cuda.synchronize()
        reps = 50 if M <= 4096 else 10
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                run()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        tops = 2.0 * M * N * Kd / (us * 1e-6) / 1e12
        print("%-14s M=%-6d N=%-5d K=%-5d  %9.2f us/launch  %8.2f TOP/s" % (name, M, N, Kd, us, tops), flush=True)


if __name__ == "__main__":
    main()
