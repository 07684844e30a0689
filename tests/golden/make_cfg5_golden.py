"""cfg5 at full size, oracle side (run in the build container; ~10 min on 8 cores): 200 fault-injection trials on the 6+6-layer
Transformer-base with a generator that can emit </s> (tests/parity_helpers.cfg5_weights), sentences `synthetic_tokens(11, 64, 64)`,
trials `campaign.make_trials(200, 5, 64, 64)`.  For every trial the ORACLE's golden and faulty greedy decodes (oracle/model.py,
int-exact, batch 1 like the reference's trials) are recorded with their smallest top-2 margin, so tests/test_fullsize_parity_gpu.py
can compare the CUDA campaign trial for trial (tokens_equal, masked / changed / no-EOS) without spending GPU-box time on CPU work.

    python tests/golden/make_cfg5_golden.py
"""
import os
import sys
from collections import Counter
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from onnx_transformer_b200 import campaign as C  # noqa: E402
from onnx_transformer_b200 import weights as W  # noqa: E402
from oracle import model as om  # noqa: E402
import parity_helpers as ph  # noqa: E402

_STATE = {}


def _init(alias):
    os.environ["OMP_NUM_THREADS"] = "1"
    _STATE["w"] = om.get_quantized(ph.cfg5_weights(alias), None, 6)
    _STATE["tok"] = W.synthetic_tokens(ph.CFG5_SEED_TOKENS, 64, 64)


def _golden(b):
    ids, mask = _STATE["tok"]
    ys, m, _ = om.greedy_decode(_STATE["w"], ids[b:b + 1], mask[b:b + 1], W.MAX_LEN, 0, "int-exact", 6, return_margins=True)
    return b, ys[0], m[0]


def _faulty(tr):
    ids, mask = _STATE["tok"]
    b = tr.sentence
    ys, m, _ = om.greedy_decode(_STATE["w"], ids[b:b + 1], mask[b:b + 1], W.MAX_LEN, 0, "int-exact", 6, return_margins=True,
                                fault=ph.oracle_fault(tr, 64, 1))
    return tr.trial_id, ys[0], m[0]


def main():
    alias = ph.CFG5_ALIAS_TOKEN
    trials = C.make_trials(ph.CFG5_N_TRIALS, ph.CFG5_SEED_TRIALS, 64, 64)
    with Pool(8, initializer=_init, initargs=(alias,)) as pool:
        gold = {b: (ys, m) for b, ys, m in pool.map(_golden, sorted({t.sentence for t in trials}))}
        print("golden decodes done", flush=True)
        faulty = {tid: (ys, m) for tid, ys, m in pool.map(_faulty, trials, chunksize=4)}
    g_ys = np.stack([gold[t.sentence][0] for t in trials]).astype(np.int16)
    g_m = np.stack([gold[t.sentence][1] for t in trials]).astype(np.float32)
    f_ys = np.stack([faulty[t.trial_id][0] for t in trials]).astype(np.int16)
    f_m = np.stack([faulty[t.trial_id][1] for t in trials]).astype(np.float32)
    outcomes = Counter(C.classify(g, f)["outcome"] for g, f in zip(g_ys.astype(np.int64), f_ys.astype(np.int64)))
    print("oracle outcomes", outcomes, flush=True)
    np.savez_compressed(os.path.join(HERE, "cfg5_fullsize.npz"), alias=np.int64(alias), golden_ys=g_ys, golden_margins=g_m, faulty_ys=f_ys,
                        faulty_margins=f_m)


if __name__ == "__main__":
    main()
