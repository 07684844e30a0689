"""Where a pipelined batch of 64 fault trials spends the host's time (campaign.run_trials_batched's loop with timers): enqueue of the
faulty decode, wait for the previous batch's event, classification.  A wait near zero means the loop is host-bound."""
import os
import sys
import time
from dataclasses import asdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnx_transformer_b200 import campaign as C, weights as W  # noqa: E402
from onnx_transformer_b200.engine import FaultSpec, QuantizedTransformer  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3200
eng = QuantizedTransformer(W.init_float_weights(0))
ids_np, mask_np = W.synthetic_tokens(11, 64, 64)
trials = C.make_trials(n, 0, 64, 64)
dev = eng.dev
ids, mask = torch.from_numpy(ids_np).to(dev), torch.from_numpy(mask_np).to(dev)
golden = eng.greedy_decode(ids, mask).cpu().numpy()
C.run_trials_batched(eng, ids_np, mask_np, trials[:128], 64)
torch.cuda.synchronize()
rows_pin = [torch.zeros(64, dtype=torch.int64).pin_memory() for _ in range(2)]
ys_pin = [torch.zeros((64, eng.max_len), dtype=torch.int64).pin_memory() for _ in range(2)]
t_spec = t_enq = t_wait = t_cls = 0.0
pending = None
T0 = time.perf_counter()
for i, c0 in enumerate(range(0, n, 64)):
    chunk = trials[c0:c0 + 64]
    a = time.perf_counter()
    specs = [FaultSpec(t.module, t.layer, t.target, t.inject_type, t.bit, t.flat_index, t.window_start, t.window_len, t.value_bits, step=0) for t in chunk]
    rp = rows_pin[i % 2]
    rp[:len(chunk)] = torch.tensor([t.sentence for t in chunk], dtype=torch.int64)
    rows = rp.to(dev, non_blocking=True)
    b = time.perf_counter()
    ys = eng.greedy_decode(ids[rows].contiguous(), mask[rows].contiguous(), fault=specs)
    buf = ys_pin[i % 2]
    buf[:, :ys.shape[1]].copy_(ys, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    c = time.perf_counter()
    t_spec += b - a
    t_enq += c - b
    if pending is not None:
        pch, pbuf, pev = pending
        d = time.perf_counter()
        pev.synchronize()
        e = time.perf_counter()
        f = pbuf.numpy()
        for k, tr in enumerate(pch):
            res = C.classify(golden[tr.sentence], f[k])
            res.update(asdict(tr))
        g = time.perf_counter()
        t_wait += e - d
        t_cls += g - e
    pending = (chunk, buf, ev)
torch.cuda.synchronize()
dt = time.perf_counter() - T0
nb = n / 64
print("%d trials: %.2f ms per batch | specs %.2f  enqueue %.2f  wait-for-GPU %.2f  classify %.2f (ms per batch)" % (n, dt / nb * 1e3, t_spec / nb * 1e3, t_enq / nb * 1e3, t_wait / nb * 1e3, t_cls / nb * 1e3))
