"""GPU time of the sections of one batched faulty decode (64 trials), measured with CUDA events while the launches queue up behind a
long kernel (so the host's launch path does not show): encoder with faults, cross-K/V, the per-op fault step, the persistent decoder."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from onnx_transformer_b200 import campaign as C, weights as W  # noqa: E402
from onnx_transformer_b200.engine import FaultSpec, QuantizedTransformer, _FaultBatch  # noqa: E402

eng = QuantizedTransformer(W.init_float_weights(0))
ids_np, mask_np = W.synthetic_tokens(11, 64, 64)
trials = C.make_trials(1280, 0, 64, 64)
B_ = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = eng.dev
ids, mask = torch.from_numpy(ids_np).to(dev), torch.from_numpy(mask_np).to(dev)
eng.greedy_decode(ids, mask)
C.run_trials_batched(eng, ids_np, mask_np, trials[:2 * B_], B_)
torch.cuda.synchronize()
B, S = (int(sys.argv[1]) if len(sys.argv) > 1 else 64), 64
ws = eng._dec_workspace(B, S)
plan = eng._decoder_plan(ws, B, S)
acc = {}
for c0 in range(0, 10 * B_, B_):
    chunk = trials[c0:c0 + B_]
    fb = _FaultBatch([FaultSpec(t.module, t.layer, t.target, t.inject_type, t.bit, t.flat_index, t.window_start, t.window_len, t.value_bits, step=0) for t in chunk])
    rows = torch.tensor([t.sentence for t in chunk], dtype=torch.int64, device=dev)
    a, b = ids[rows].contiguous(), mask[rows].contiguous()
    torch.cuda.synchronize()
    plan.run(0, 71)                      # the plug: ~10 ms during which the host enqueues everything below
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    ev[0].record()
    memory = eng.encode(a, b, fault=fb)
    ev[1].record()
    ws["mask"].copy_(b.reshape(B, S).to(torch.uint8))
    eng._prepare_cross_kv(ws, memory, fb)
    ws["ys"].zero_(); ws["ys"][:, 0] = 0; ws["step"].zero_()
    ev[2].record()
    eng._decode_step(ws, B, S, fault=fb)
    ev[3].record()
    plan.run(1, 70)
    ev[4].record()
    torch.cuda.synchronize()
    for k, name in enumerate(["encoder", "cross_kv+reset", "fault_step", "decoder_70"]):
        acc[name] = acc.get(name, 0.0) + ev[k].elapsed_time(ev[k + 1])
print("GPU ms per faulty batch (OT_MF_PATCH=%s):" % os.environ.get("OT_MF_PATCH", "1"), {k: round(v / 10, 3) for k, v in acc.items()}, "sum %.2f" % (sum(acc.values()) / 10))

