"""Batched fault-injection trials: (a) two runs of the pipelined loop give the same tokens trial for trial, (b) the persistent
cluster decoder and the per-op kernel path give the same tokens for every trial (the fault step runs through the per-op kernels in
both; the remaining 70 steps differ in path only)."""
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnx_transformer_b200 import campaign as C  # noqa: E402
from onnx_transformer_b200 import weights as W  # noqa: E402
from onnx_transformer_b200.engine import QuantizedTransformer  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6400
fw = W.init_float_weights(0)
ids, mask = W.synthetic_tokens(11, 64, 64)
trials = C.make_trials(n, 0, 64, 64)


def run(eng):
    toks = {}
    orig = C.classify

    def spy(golden, faulty):
        r = orig(golden, faulty)
        r["hash"] = hashlib.md5(np.ascontiguousarray(faulty).tobytes()).hexdigest()
        return r
    C.classify = spy
    try:
        res = C.run_trials_batched(eng, ids, mask, trials, 64)
    finally:
        C.classify = orig
    for r in res:
        toks[r["trial_id"]] = r["hash"]
    return toks


ep = QuantizedTransformer(fw, persistent=True)
a = run(ep)
b = run(ep)
print("pipelined persistent run twice: %d / %d trials differ" % (sum(a[k] != b[k] for k in a), len(a)))
eg = QuantizedTransformer(fw, persistent=False)
g = run(eg)
bad = [k for k in a if a[k] != g[k]]
print("persistent vs per-op path: %d / %d trials differ" % (len(bad), len(a)))
for k in bad[:10]:
    t = trials[k]
    print("  trial", k, t.module, t.layer, t.target, t.inject_type, t.bit, t.flat_index)
