import os, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from onnx_transformer_b200 import campaign as C, weights as W
from onnx_transformer_b200.engine import QuantizedTransformer, FaultSpec
eng = QuantizedTransformer(W.init_float_weights(0))
ids_np, mask_np = W.synthetic_tokens(11, 64, 64)
trials = C.make_trials(640, 0, 64, 64)
dev = eng.dev
ids = torch.from_numpy(ids_np).to(dev); mask = torch.from_numpy(mask_np).to(dev)
golden = eng.greedy_decode(ids, mask).cpu().numpy()
C.run_trials_batched(eng, ids_np, mask_np, trials[:64], 64)
torch.cuda.synchronize()
t_prep = t_gpu = t_cls = 0.0
for c0 in range(0, 640, 64):
    chunk = trials[c0:c0 + 64]
    t0 = time.perf_counter()
    rows = torch.tensor([t.sentence for t in chunk], dtype=torch.int64, device=dev)
    specs = [FaultSpec(t.module, t.layer, t.target, t.inject_type, t.bit, t.flat_index, t.window_start, t.window_len, t.value_bits, step=0) for t in chunk]
    a, b = ids[rows].contiguous(), mask[rows].contiguous()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    ys = eng.greedy_decode(a, b, fault=specs)
    t1b = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    faulty = ys.cpu().numpy()
    for k, trial in enumerate(chunk):
        res = C.classify(golden[trial.sentence], faulty[k])
    t3 = time.perf_counter()
    t_prep += t1 - t0; t_gpu += t2 - t1; t_cls += t3 - t2
    if c0 == 0: print("launch (host) %.2f ms of %.2f ms" % ((t1b - t1) * 1e3, (t2 - t1) * 1e3))
print("per batch of 64: prep %.2f ms, decode (host launch + GPU) %.2f ms, classify %.2f ms" % (t_prep * 100, t_gpu * 100, t_cls * 100))
