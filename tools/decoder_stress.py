"""Randomised stress of the cluster-resident decoder against the per-op kernel path: random batch sizes (ragged last group),
sentences per cluster, source lengths and decode lengths on a 2-layer model; every configuration is decoded three times (the runs
must agree with each other and with the per-op path, tokens and KV caches)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnx_transformer_b200 import weights as W  # noqa: E402
from onnx_transformer_b200.engine import QuantizedTransformer  # noqa: E402

budget_s = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(123)
t_end = time.time() + budget_s
n_cfg = n_bad = 0
while time.time() < t_end:
    B = int(rng.integers(1, 81))
    spc = int(rng.integers(1, 9))
    S = int(rng.integers(4, 65))
    max_len = int(rng.integers(4, 24))
    seed = int(rng.integers(0, 1 << 30))
    fw = W.init_float_weights(seed % 1000, 211, 197, 2, randomize_norms=True)
    ep = QuantizedTransformer(fw, n_layers=2, max_len=max_len, persistent=True, decoder="cluster", sentences_per_cluster=spc)
    eg = QuantizedTransformer(fw, n_layers=2, max_len=max_len, persistent=False)
    ids, mask = W.synthetic_tokens(seed, B, S, 211, min_len=min(3, S))
    idt, mt = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
    ref = eg.greedy_decode(idt, mt, max_len)
    ok = True
    for rep in range(3):
        ys = ep.greedy_decode(idt, mt, max_len)
        ok = ok and bool(torch.equal(ys, ref))
    ws_p, ws_g = ep._dec_workspace(B, S), eg._dec_workspace(B, S)
    for name in ("kc", "vc", "skc", "svc"):
        for a, b in zip(ws_p[name], ws_g[name]):
            a, b = a[:, :max_len - 1], b[:, :max_len - 1]
            ok = ok and bool(torch.equal(a.view(torch.int32) if a.dtype == torch.float32 else a, b.view(torch.int32) if b.dtype == torch.float32 else b))
    n_cfg += 1
    if not ok:
        n_bad += 1
        print("MISMATCH B=%d spc=%d S=%d max_len=%d seed=%d" % (B, spc, S, max_len, seed), flush=True)
print("%d configurations, %d mismatches" % (n_cfg, n_bad))
