"""Full-size golden vectors for BASELINE.json cfg2 / cfg3 (run in the build container only: imports /root/reference).

cfg2 -- the benchmarked workload itself: Transformer-base 6+6, weights `init_float_weights(0)`, source batch
`synthetic_tokens(1000, 64, 64)` (bench.py rank 0), 71 greedy steps.  Two decodes are recorded:
  * `ref_*`:    the REFERENCE's own torch modules (model.make_model + get_quantized_model.quantize_transformer imported from
                /root/reference, unmodified; the batched greedy loop of batch_output.py:659-672 with full-prefix recompute) --
                i.e. the reference's fp32 fake-quant arithmetic ("ref-float");
  * `oracle_*`: oracle/model.py in "int-exact" mode with the KV cache (the factorisation the CUDA kernels implement).
For both: token ids [64,72], the top-2 logit margin of every step [64,71] and the arg-max / runner-up ids.
cfg3 -- encoder only, sentences 0..7 of `synthetic_tokens(7, 512, 128, min_len=40)` (shard invariance of the CUDA encoder is
proven bit for bit by tests/test_engine_gpu.py, so one shard pins the whole batch): the reference torch encoder's memory rows.

    python tests/golden/make_fullsize_golden.py            # ~10 min on 8 cores
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402


def greedy_top2(model, src, src_mask, max_len, start_symbol=0):
    with torch.no_grad():
        memory = model.encode(src, src_mask)
        ys = torch.zeros(src.shape[0], 1, dtype=src.dtype).fill_(start_symbol)
        margins, second = [], []
        for _ in range(max_len - 1):
            out = model.decode(memory, src_mask, ys, mg.subsequent_mask(ys.size(1)).type_as(src.data))
            prob = model.generator(out[:, -1])
            top2 = torch.topk(prob, 2, dim=1)
            margins.append((top2.values[:, 0] - top2.values[:, 1]).numpy())
            second.append(top2.indices[:, 1].numpy())
            _, nxt = torch.max(prob, dim=1)
            ys = torch.cat([ys, nxt.reshape(-1, 1)], dim=1)
    return memory.numpy(), ys.numpy(), np.stack(margins, 1), np.stack(second, 1)


def main():
    from onnx_transformer_b200 import weights as W
    from oracle import model as om
    torch.manual_seed(0)
    torch.set_num_threads(8)
    fw = W.init_float_weights(0)
    model = mg.build_reference_model(fw, None, 6, W.SRC_VOCAB, W.TGT_VOCAB)

    # ---- cfg2
    ids, mask = W.synthetic_tokens(1000, 64, 64)
    t0 = time.time()
    memory, ys, margins, second = greedy_top2(model, torch.from_numpy(ids), torch.from_numpy(mask), W.MAX_LEN)
    print("cfg2 reference torch modules: %.1f s" % (time.time() - t0), flush=True)
    wq = om.get_quantized(fw, None, 6)
    t0 = time.time()
    o_ys, o_margins, o_memory = om.greedy_decode(wq, ids, mask, W.MAX_LEN, 0, "int-exact", 6, kv_cache=True, return_margins=True)
    print("cfg2 oracle int-exact: %.1f s" % (time.time() - t0), flush=True)
    np.savez_compressed(os.path.join(HERE, "cfg2_fullsize.npz"), ids=ids, ref_ys=ys.astype(np.int16), ref_margins=margins.astype(np.float32),
                        ref_second=second.astype(np.int16), oracle_ys=o_ys.astype(np.int16), oracle_margins=o_margins.astype(np.float32),
                        ref_memory_s0=memory[0], oracle_memory_s0=o_memory[0])
    same = (ys == o_ys)
    first = [int(np.argmin(r)) if not r.all() else -1 for r in same]
    print("cfg2: reference vs oracle: %d / 64 sentences identical; first divergences %s" % (sum(f < 0 for f in first), first), flush=True)

    # ---- cfg3 (one 8-sentence shard)
    ids3, mask3 = W.synthetic_tokens(7, 512, 128, min_len=40)
    with torch.no_grad():
        mem3 = model.encode(torch.from_numpy(ids3[:8]), torch.from_numpy(mask3[:8])).numpy()
    np.savez_compressed(os.path.join(HERE, "cfg3_shard0.npz"), ids=ids3[:8], mask=mask3[:8], ref_memory=mem3.astype(np.float32))
    print("written")


if __name__ == "__main__":
    main()
