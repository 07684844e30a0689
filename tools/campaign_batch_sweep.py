"""Fault-trial throughput vs trials per batched decode: the cluster decoder's step time does not depend on how many of the (at most 15
co-resident) 8-CTA clusters are in use, so a batch of 120 trials decodes in the time of 64."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnx_transformer_b200 import campaign as C, weights as W  # noqa: E402
from onnx_transformer_b200.engine import QuantizedTransformer  # noqa: E402

eng = QuantizedTransformer(W.init_float_weights(0))
ids, mask = W.synthetic_tokens(11, 64, 64)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 7680
trials = C.make_trials(n, 0, 64, 64)
ref = None
for batch in [int(x) for x in (sys.argv[2:] or ["64", "96", "112", "120", "128"])]:
    C.run_trials_batched(eng, ids, mask, trials[:2 * batch], batch)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = C.run_trials_batched(eng, ids, mask, trials, batch, return_tokens=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    toks = {r["trial_id"]: r["faulty_ys"].tobytes() for r in res}
    if ref is None:
        ref = toks
    same = sum(1 for k in ref if toks[k] == ref[k])
    print("batch %3d: %d trials in %.2f s = %.0f trials/s   (%.2f ms per batch)   tokens equal to batch-64 run: %d / %d" % (batch, n, dt, n / dt, dt / ((n + batch - 1) // batch) * 1e3, same, len(ref)), flush=True)
