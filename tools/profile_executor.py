"""cProfile of the drop-in executor's Python node walk on cfg1 (B = 1, S = 72, full-size graphs): where the per-node host time goes."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from onnx_transformer_b200 import decode as D  # noqa: E402
from onnx_transformer_b200 import executor as X  # noqa: E402
from onnx_transformer_b200 import weights as W  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
fw, enc, dec = bench.cfg1_graphs()
model = D.HostModel(fw)
ids, mask = W.synthetic_tokens(1000, 1, 72)
D.greedy_decode(model, ids, mask, 4, 0, enc, dec, executor=X)
torch.cuda.synchronize()
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
D.greedy_decode(model, ids, mask, steps + 1, 0, enc, dec, executor=X)
torch.cuda.synchronize()
pr.disable()
dt = time.perf_counter() - t0
print("%d decoder passes + encoder: %.3f s (%.1f ms per pass, %d nodes per pass)" % (steps, dt, dt / steps * 1e3, len(dec.node)))
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
pstats.Stats(pr).sort_stats("tottime").print_stats(30)
