"""Batched fault-injection trials at full model size (debug / timing aid)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnx_transformer_b200 import campaign as C, weights as W  # noqa: E402
from onnx_transformer_b200.engine import QuantizedTransformer  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng = QuantizedTransformer(W.init_float_weights(0))
ids, mask = W.synthetic_tokens(11, 64, 64)
trials = C.make_trials(n, 0, 64, 64)
if len(sys.argv) > 2:   # bisect: one trial at a time, batched path
    for t in trials:
        print(t, flush=True)
        C.run_trials_batched(eng, ids, mask, [t], 64)
        torch.cuda.synchronize()
    sys.exit(0)
C.run_trials_batched(eng, ids, mask, trials[:64], 64)
torch.cuda.synchronize()
t0 = time.perf_counter()
res = C.run_trials_batched(eng, ids, mask, trials, 64)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
from collections import Counter
print("%d trials in %.2f s = %.1f trials/s" % (n, dt, n / dt), Counter(r["outcome"] for r in res), Counter(r["tokens_equal"] for r in res))
