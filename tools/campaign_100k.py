"""SURVEY 8d cfg5 at its full size: 100,000 fault-injection trials drawn from np.random.default_rng(0) over 64 sentences x 64 source
tokens (Encoder + Decoder targets), 64 trials per batched faulty decode, on one GPU (or sharded trial-wise under torchrun).
Prints one JSON line: wall time, trials/s and the outcome histogram (random-init weights never emit </s>, so every trial lands in the
reference's "no-EOS" row; `tokens_changed` counts the trials whose faulty token sequence differs from the golden one)."""
import json
import os
import sys
import time
from collections import Counter

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnx_transformer_b200 import campaign as C  # noqa: E402
from onnx_transformer_b200 import weights as W  # noqa: E402
from onnx_transformer_b200.engine import QuantizedTransformer  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    dist.init_process_group("nccl")
eng = QuantizedTransformer(W.init_float_weights(0))
ids, mask = W.synthetic_tokens(11, 64, 64)
trials = C.make_trials(n, 0, 64, 64)
tpd = int(os.environ.get("OT_TRIALS_PER_DECODE", "0")) or C.trials_per_decode(eng)
C.run_trials_batched(eng, ids, mask, trials[:2 * tpd], tpd)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
res = C.run_trials_batched(eng, ids, mask, trials, tpd, None, rank, world)
torch.cuda.synchronize()
hist = Counter(r["outcome"] for r in res)
hist["tokens_changed"] = sum(1 for r in res if not r["tokens_equal"])
mods = Counter((r["module"], r["inject_type"]) for r in res)
if world > 1:
    parts = [None] * world
    dist.all_gather_object(parts, dict(hist))
    hist = Counter()
    for p in parts:
        hist.update(p)
    dist.barrier()
dt = time.perf_counter() - t0
if rank == 0:
    print(json.dumps({"trials": n, "n_gpus": world, "trials_per_decode": tpd, "wall_s": dt, "trials_per_s": n / dt, "outcomes": dict(hist),
                      "rank0_targets": {"%s/%s" % k: v for k, v in mods.items()}}))
if world > 1:
    dist.destroy_process_group()
