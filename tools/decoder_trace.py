"""Persistent decoder: decode time (persistent vs per-op graph path) and the per-phase timeline of one greedy step
(%globaltimer of CTA 0 after every grid barrier)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnx_transformer_b200 import weights as W  # noqa: E402
from onnx_transformer_b200.engine import QuantizedTransformer  # noqa: E402

B, S = int(os.environ.get("OT_B", "64")), 64
fw = W.init_float_weights(0)
ids, mask = W.synthetic_tokens(1000, B, S)
ids, mask = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()


def timeit(eng, reps=5):
    for _ in range(2):
        ys = eng.greedy_decode(ids, mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ys = eng.greedy_decode(ids, mask)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, ys


DEC = "grid" if "--grid" in sys.argv else "cluster"
SPC = int(os.environ.get("OT_SPC", "8"))
ep = QuantizedTransformer(fw, persistent=True, decoder=DEC, sentences_per_cluster=SPC)
ms_p, ys_p = timeit(ep)
print("persistent: %.2f ms/decode  %.0f tok/s" % (ms_p, B * 71 / ms_p * 1e3), flush=True)
if "--no-graph" not in sys.argv:
    eg = QuantizedTransformer(fw, persistent=False)
    ms_g, ys_g = timeit(eg)
    print("per-op graph: %.2f ms/decode  %.0f tok/s   identical=%s" % (ms_g, B * 71 / ms_g * 1e3, bool(torch.equal(ys_p, ys_g))), flush=True)

ws = ep._dec_workspace(B, S)
plan = ep._decoder_plan(ws, B, S, trace=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
plan.run(35, 1)
torch.cuda.synchronize()
e0.record()
plan.run(30, 10)
e1.record()
torch.cuda.synchronize()
print("10 steps in one launch: %.1f us/step" % (e0.elapsed_time(e1) * 100))
t = plan.trace.cpu().numpy()
if DEC == "cluster":
    names = plan.phase_names()
else:
    names = []
    for l in range(6):
        names += ["ln1", "qkv", "sattn", "o", "ln2", "cq", "cattn", "co", "ln3", "ffn1a", "ffn1b", "ffn2"]
    names += ["lnf", "gen"]
work, wait = {}, {}
if DEC == "cluster":
    # cluster decoder: t[2i] = phase i begins waiting for its input bytes, t[2i+1] = they have all arrived; t[254] = step end
    for i, n in enumerate(names):
        end = t[2 * (i + 1)] if i + 1 < len(names) else t[254]
        wait.setdefault(n, []).append((t[2 * i + 1] - t[2 * i]) / 1e3)
        work.setdefault(n, []).append((end - t[2 * i + 1]) / 1e3)
    print("phase   n   input_wait_us  work_us(CTA0)   sum_us")
    for n in work:
        print("%-10s x%d   %7.2f   %7.2f   %7.1f" % (n, len(work[n]), np.mean(wait[n]), np.mean(work[n]), np.sum(work[n]) + np.sum(wait[n])))
    print("step total %.1f us" % ((t[254] - t[255]) / 1e3))
else:
    prev = t[255]
    for i, n in enumerate(names):
        work.setdefault(n, []).append((t[2 * i] - prev) / 1e3)          # CTA 0: phase start -> its arrival at the barrier
        wait.setdefault(n, []).append((t[2 * i + 1] - t[2 * i]) / 1e3)  # CTA 0: arrival -> release
        prev = t[2 * i + 1]
    print("phase   n   work_us(CTA0)  barrier_wait_us   sum_us")
    for n in work:
        print("%-6s x%d   %7.2f   %7.2f   %7.1f" % (n, len(work[n]), np.mean(work[n]), np.mean(wait[n]), np.sum(work[n]) + np.sum(wait[n])))
    print("step total %.1f us" % ((t[2 * len(names) - 1] - t[255]) / 1e3))
if DEC == "cluster" and "--marks" in sys.argv:
    print("fine marks of layer 2 (id: us since step start, delta):")
    prev = None
    for i in range(150, 250):
        v = int(t[i])
        if v == 0:
            break
        mid, ns = v >> 32, v & 0xffffffff
        print("  %3d  %9.2f  %+7.2f" % (mid, ns / 1e3, 0.0 if prev is None else (ns - prev) / 1e3))
        prev = ns
