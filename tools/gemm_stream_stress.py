"""Randomised stress of the streaming fp32 GEMMs (TMA epilogue at K <= 512, CTA pairs at K > 512, LSU epilogue) against the
one-tile-per-CTA kernel, bit for bit, every configuration launched three times (races would show as run-to-run differences).
python tools/gemm_stream_stress.py [n_configs] [seed]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnx_transformer_b200 import kernels as K  # noqa: E402

n_cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
dev = torch.device("cuda")
bad = 0
for i in range(n_cfg):
    M = int(rng.integers(2048, 9000))
    N = int(rng.choice([256, 512, 768, 1024, 1536]))
    Kd = int(rng.choice([128, 256, 384, 512, 640, 1024, 1536, 2048]))
    relu, res, bias = bool(rng.integers(0, 2)), bool(rng.integers(0, 2)), bool(rng.integers(0, 4))
    a = torch.from_numpy(rng.integers(-127, 128, size=(M, Kd), dtype=np.int8)).to(dev)
    w = torch.from_numpy(rng.integers(-127, 128, size=(N, Kd), dtype=np.int8)).to(dev)
    sx = torch.from_numpy(rng.uniform(1e-3, 5e-2, size=M).astype(np.float32)).to(dev)
    sw = torch.from_numpy(rng.uniform(1e-4, 1e-2, size=N).astype(np.float32)).to(dev)
    b = torch.from_numpy(rng.normal(size=N).astype(np.float32)).to(dev) if bias else None
    r = torch.from_numpy(rng.normal(size=(M, N)).astype(np.float32)).to(dev) if res else None

    def run():
        return K.linear_w8a8(a, w, row_scale=sx, col_scale=sw, bias=b, residual=r, relu=relu, out_kind=K.OUT_F32)
    outs = [run() for _ in range(3)]
    os.environ["OT_GEMM_STREAM"] = "0"
    ref = run()
    del os.environ["OT_GEMM_STREAM"]
    ok = all(torch.equal(o, ref) for o in outs)
    if not ok:
        bad += 1
        print("MISMATCH M=%d N=%d K=%d relu=%s res=%s bias=%s" % (M, N, Kd, relu, res, bias), flush=True)
print("%d configurations x 3 launches, %d mismatches" % (n_cfg, bad))
sys.exit(1 if bad else 0)
