"""Device-side timeline of ONE replayed greedy step (B=64, S=64): per kernel start / dependency-wait end / end, gaps."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnx_transformer_b200 import _lib, weights as W  # noqa: E402
from onnx_transformer_b200.engine import QuantizedTransformer  # noqa: E402

trace_n = None
for a in sys.argv:
    if a.startswith("--trace-n="):
        trace_n = int(a.split("=")[1])
tbuf = torch.zeros(4096 * 8, dtype=torch.int64, device="cuda")
if trace_n is not None:      # per-CTA phase stamps of the GEMMs with N == trace_n (the last such launch of the step wins)
    os.environ["OT_GEMM_TRACE"] = hex(tbuf.data_ptr())
    os.environ["OT_GEMM_TRACE_N"] = str(trace_n)
pdl = "--no-pdl" not in sys.argv
fused = "--fused-ln" in sys.argv
eng = QuantizedTransformer(W.init_float_weights(0), pdl=pdl, fused_ln=fused)
ids, mask = W.synthetic_tokens(1000, 64, 64)
ids, mask = torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda()
eng.greedy_decode(ids, mask)
ws = eng._dec_workspace(64, 64)
cap = 4096
buf = torch.zeros(1 + 4 * cap, dtype=torch.int64, device="cuda")
lib = _lib.load()
torch.cuda.synchronize()
ws["step"].fill_(30)
lib.ot_set_timeline(C.c_void_p(buf.data_ptr()), cap)
for _ in range(3):
    ws["graph"].replay()
torch.cuda.synchronize()
lib.ot_set_timeline(None, 0)
t = buf.cpu().numpy()
n = int(t[0])
rec = t[1:1 + 4 * n].reshape(n, 4)
per = n // 3
rec = rec[2 * per:]                      # third replay
names = {1: "gemm", 2: "attn", 3: "attn_dec", 4: "ln", 5: "rowq", 6: "embed", 7: "gen_logits", 8: "gen_reduce", 9: "append"}
order = np.argsort(rec[:, 1])
rec = rec[order]
t0 = rec[0, 1]
print("pdl=%s fused_ln=%s kernels/step=%d  step span = %.1f us" % (pdl, fused, per, (rec[:, 3].max() - t0) / 1e3))
print("%-10s %9s %9s %9s | %8s %8s %8s" % ("kernel", "start", "ready", "end", "wait", "body", "gap_prev"))
prev_end = t0
tot = {}
for r in rec:
    k = names.get(int(r[0]), "?")
    wait, body, gap = (r[2] - r[1]) / 1e3, (r[3] - r[2]) / 1e3, (r[2] - prev_end) / 1e3
    tot.setdefault(k, [0, 0.0, 0.0])
    tot[k][0] += 1; tot[k][1] += body; tot[k][2] += max(gap, 0)
    if "-v" in sys.argv:
        print("%-10s %9.2f %9.2f %9.2f | %8.2f %8.2f %8.2f" % (k, (r[1] - t0) / 1e3, (r[2] - t0) / 1e3, (r[3] - t0) / 1e3, wait, body, gap))
    prev_end = max(prev_end, r[3])
print("family      n   body_us  idle_before_us")
for k, (c, b, g) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-10s %3d %9.1f %9.1f" % (k, c, b, g))

if trace_n is not None:
    tt = tbuf.cpu().numpy().reshape(-1, 8)
    tt = tt[tt[:, 0] > 0]
    t0 = tt[:, 0].min()
    print("GEMM N=%d phase stamps per CTA (us): setup ln_done acc_ready pass1 clsync end" % trace_n)
    for i, r in enumerate(tt[:20]):
        print(i, " ".join("%7.2f" % ((v - t0) / 1e3) if v > 0 else "      -" for v in r[:6]))
