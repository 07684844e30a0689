"""Greedy-decode time of the cluster-resident decoder vs sentences per cluster (B = 64 x S = 64, 71 steps) + per-phase trace."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from onnx_transformer_b200 import weights as W  # noqa: E402
from onnx_transformer_b200.engine import QuantizedTransformer  # noqa: E402

fw = W.init_float_weights(0)
ids_np, mask_np = W.synthetic_tokens(1000, 64, 64)
ids, mask = torch.from_numpy(ids_np).cuda(), torch.from_numpy(mask_np).cuda()
ref = None
for spc in [int(x) for x in (sys.argv[1:] or ["8", "6", "5", "4"])]:
    eng = QuantizedTransformer(fw, sentences_per_cluster=spc)
    for _ in range(3):
        ys = eng.greedy_decode(ids, mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ys = eng.greedy_decode(ids, mask)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    if ref is None:
        ref = ys.clone()
    ws = eng._dec_workspace(64, 64)
    phases, step_us = bench.persistent_phase_trace(eng, ws, 64, 64)
    print("spc=%d clusters=%d  %.3f ms per decode  %.1f us/step (traced %.1f)  tokens_equal=%s" % (spc, (64 + spc - 1) // spc, ms, ms * 1e3 / 71, step_us, bool(torch.equal(ys, ref))))
    print("   ", phases, flush=True)
    tr = eng._last_trace
    if tr[240] > 0:      # gen marks: operands ready, MMA + epilogue done, maxima gathered, decision made, phase end (us since the step began)
        print("    generator marks (us):", [round((int(tr[k]) - int(tr[255])) / 1e3, 2) for k in (240, 241, 242, 243, 254)])
    del eng
