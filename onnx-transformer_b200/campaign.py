"""Fault-injection campaign driver on the fused engine: the B200 counterpart of
parallelized_inject_onnx_transformer.py:789-861 (load_trained_model: target files x fault models x bit positions),
:446-493 (run_model_example: sentence draw + worker pool), :305-409 (check_outputs: golden and faulty greedy decode,
sentence BLEU, one CSV row per trial) -- SURVEY.md 8f item 1.

Differences by design:
  * a trial is an explicit record (sentence id, target, fault model, bit, element index, window, bit pattern) drawn
    from np.random.default_rng(seed) -- the reference draws from the unseeded global np.random inside the hooks -- so
    every trial can be replayed and compared trial-for-trial with the oracle;
  * the golden decode of a sentence is computed once and cached (the reference decodes golden + faulty per trial);
    for a Decoder target the encoder memory is reused as well, and only the injection step runs outside the CUDA graph;
  * there are no ./separated/*.onnx temp files (the reference's three-segment graph cut and its file race,
    :411-444, :583): the fault is applied inside the epilogue of the target kernel;
  * trials are sharded over ranks (rank r takes trials r::N); the only collective is the gather of the result records.

CSV schema (results_fault_injection/results.csv, 5-column rows): op,golden_bleu,faulty_bleu,bit,type ; "op,0,0,bit,type"
when the output has no </s> (:379-383).  With synthetic data there is no reference translation, so the golden output
itself serves as the BLEU reference (golden BLEU = 1.0; faulty BLEU < 1.0 iff the decoded sentence changed).
"""
from __future__ import annotations

import argparse
import math
import os
from collections import Counter
from dataclasses import asdict, dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import weights as W

FAULT_MODELS = ["INPUT", "WEIGHT", "INPUT16", "WEIGHT16", "RANDOM", "RANDOM_BITFLIP"]   # :805
ENC_TARGETS = ["qk", "pv", "ffn1", "ffn2"]                                              # input/encoder/*.json
DEC_TARGETS = ["qk", "pv", "cqk", "cpv", "ffn1", "ffn2"]                                # input/decoder/*.json
D, FF, H, DK = 512, 2048, 8, 64


@dataclass
class Trial:
    trial_id: int
    sentence: int
    module: str          # "Encoder" | "Decoder"
    layer: int
    target: str          # role name (graph.py)
    op_name: str         # reference node name, e.g. "MatMul_28"
    inject_type: str
    bit: int             # 0..7 for operand faults; fp32 bit 0..31 for RANDOM_BITFLIP
    flat_index: int
    window_start: int = 0
    window_len: int = 0
    value_bits: int = 0


def _tensor_shape(module: str, target: str, operand: str, S: int, T: int):
    """Shape of the faulty tensor in the reference's layout for one sentence (B = 1, like the reference's trials)."""
    Tq = S if module == "Encoder" else T
    Tk = S if (module == "Encoder" or target in ("cqk", "cpv")) else T
    if target in ("qk", "cqk"):
        return {"input": (1, Tq, D), "weight": (1, Tk, D), "output": (1, H, Tq, Tk)}[operand]
    if target in ("pv", "cpv"):
        return {"input": (1, H, Tq, Tk), "weight": (1, Tk, D), "output": (1, H, Tq, DK)}[operand]
    if target == "ffn1":
        return {"input": (1, Tq, D), "weight": (FF, D), "output": (1, Tq, FF)}[operand]
    if target == "ffn2":
        return {"input": (1, Tq, FF), "weight": (D, FF), "output": (1, Tq, D)}[operand]
    raise ValueError(target)


def make_trials(n: int, seed: int, n_sentences: int, src_len: int, modules: Sequence[str] = ("Encoder", "Decoder"), n_layers: int = 6,
                weight_bits: int = 8) -> List[Trial]:
    """cfg5: trials = (sentence id, target, inject_type, bit, element index, window) from default_rng(seed).
    Decoder trials inject at greedy step 0 (target_inference_number = 1, :832), i.e. with T = 1."""
    from .graph import decoder_matmul_name, encoder_matmul_name
    rng = np.random.default_rng(seed)
    trials = []
    for tid in range(n):
        module = modules[int(rng.integers(0, len(modules)))]
        layer = int(rng.integers(0, n_layers))
        targets = ENC_TARGETS if module == "Encoder" else DEC_TARGETS
        target = targets[int(rng.integers(0, len(targets)))]
        ftype = FAULT_MODELS[int(rng.integers(0, len(FAULT_MODELS)))]
        if ftype in ("INPUT16", "WEIGHT16") and target in ("ffn1", "ffn2"):
            ftype = ftype[:-2]          # the *16 windows index shape[3]/shape[2] of a 4-D output: attention MatMuls only
        operand = "input" if ftype.startswith("INPUT") else ("weight" if ftype.startswith("WEIGHT") else "output")
        shape = _tensor_shape(module, target, operand, src_len, 1)
        idx = [int(rng.integers(0, d)) for d in shape]
        flat = int(np.ravel_multi_index(idx, shape))
        out_shape = _tensor_shape(module, target, "output", src_len, 1)
        bit = int(rng.integers(0, 32)) if ftype == "RANDOM_BITFLIP" else int(rng.integers(0, 8))
        if weight_bits == 4 and ftype.startswith("WEIGHT") and target in ("ffn1", "ffn2"):
            bit &= 3                    # cfg4: the linear weights are 4-bit (inject_main.py:410 `range(4)`, flip_int4_bit)
        ws, wl = 0, 0
        if ftype == "INPUT16":
            blocks = out_shape[3] // 16
            ws, wl = (16 * int(rng.integers(0, blocks)) if blocks else 0), 16
        elif ftype == "WEIGHT16":
            blocks = out_shape[2] // 16
            ws, wl = (16 * int(rng.integers(0, blocks)) if blocks else 0), int(rng.integers(1, 16))
        name = (encoder_matmul_name if module == "Encoder" else decoder_matmul_name)(layer, target)
        trials.append(Trial(tid, int(rng.integers(0, n_sentences)), module, layer, target, name, ftype, bit, flat, ws, wl,
                            int(rng.integers(0, 2 ** 32)) if ftype == "RANDOM" else 0))
    return trials


# ---------------------------------------------------------------------------------------------- target-directory campaigns
def load_targets(directory: str) -> List[dict]:
    """The fault-target descriptors of `--directory_name` (input/encoder/*.json x24, input/decoder/*.json x36; keys target_layer,
    input_tensor, weight_tensor, bias_tensor, output_tensor, module, model_name), in os.listdir order made deterministic (sorted)."""
    import json
    out = []
    for name in sorted(os.listdir(directory)):
        if name.endswith(".json"):
            with open(os.path.join(directory, name)) as f:
                out.append(json.load(f))
    return out


def target_site(target: dict):
    """(module, layer, role) of a descriptor, from the node-numbering contract of the 60 reference files: encoder
    MatMul_{8l+k}, k = q,k,v,qk,pv,o,ffn1,ffn2; decoder MatMul_{0..11} = cross k,v of layers 0..5, MatMul_{12+12l+k}, k = q,k,v,qk,pv,o,
    cq,cqk,cpv,co,ffn1,ffn2."""
    module = target["module"].split("/")[0]
    k = int(target["target_layer"].split("_")[1])
    if module == "Encoder":
        return module, k // 8, ["q", "k", "v", "qk", "pv", "o", "ffn1", "ffn2"][k % 8]
    if k < 12:
        return module, k // 2, ("ck", "cv")[k % 2]
    k -= 12
    return module, k // 12, ["q", "k", "v", "qk", "pv", "o", "cq", "cqk", "cpv", "co", "ffn1", "ffn2"][k % 12]


def trials_from_targets(targets: Sequence[dict], seed: int, n_sentences: int, src_len: int, experiments: int = 5, n_bits: int = 8,
                        fault_models: Sequence[str] = tuple(FAULT_MODELS)) -> List[Trial]:
    """The campaign loop of parallelized_inject_onnx_transformer.py:794-861: for every target file x fault model (:805) x bit position
    (:852, `range(8)`; RANDOM models run the loop too, with faulty_bit_position None) -> `total_experiments` = 5 (:830) sentences,
    each with its own element / window / bit-pattern draws (explicit here, from default_rng(seed)).  The *16 models index shape[3] /
    shape[2] of a 4-D output, so they exist for the attention MatMuls only (the reference raises IndexError on the FFN targets)."""
    rng = np.random.default_rng(seed)
    trials: List[Trial] = []
    for target in targets:
        module, layer, role = target_site(target)
        if role not in (ENC_TARGETS if module == "Encoder" else DEC_TARGETS):
            continue                       # projection MatMuls (q, k, v, o ...) are not in the reference's target directories
        for ftype in fault_models:
            if ftype in ("INPUT16", "WEIGHT16") and role in ("ffn1", "ffn2"):
                continue
            operand = "input" if ftype.startswith("INPUT") else ("weight" if ftype.startswith("WEIGHT") else "output")
            shape = _tensor_shape(module, role, operand, src_len, 1)
            out_shape = _tensor_shape(module, role, "output", src_len, 1)
            for bit_position in range(n_bits):
                for _ in range(experiments):
                    idx = [int(rng.integers(0, d)) for d in shape]
                    ws, wl = 0, 0
                    if ftype == "INPUT16":
                        blocks = out_shape[3] // 16
                        ws, wl = (16 * int(rng.integers(0, blocks)) if blocks else 0), 16
                    elif ftype == "WEIGHT16":
                        blocks = out_shape[2] // 16
                        ws, wl = (16 * int(rng.integers(0, blocks)) if blocks else 0), int(rng.integers(1, 16))
                    bit = int(rng.integers(0, 32)) if ftype == "RANDOM_BITFLIP" else bit_position
                    trials.append(Trial(len(trials), int(rng.integers(0, n_sentences)), module, layer, role, target["target_layer"], ftype, bit,
                                        int(np.ravel_multi_index(idx, shape)), ws, wl, int(rng.integers(0, 2 ** 32)) if ftype == "RANDOM" else 0))
    return trials


# ---------------------------------------------------------------------------------------------- BLEU (nltk method4)
def sentence_bleu_method4(reference: Sequence[int], hypothesis: Sequence[int]) -> float:
    """nltk.translate.bleu_score.sentence_bleu([reference], hypothesis, smoothing_function=SmoothingFunction().method4)
    (the call at parallelized_inject_onnx_transformer.py:393-397) on token sequences: uniform 4-gram weights, brevity
    penalty, method4 smoothing (zero n-gram matches are replaced by 1 / (2^k * K / ln(len(hyp)))), K = 5.  Restated from nltk 3.8's
    corpus_bleu / modified_precision / brevity_penalty / SmoothingFunction.method4 (nltk is not installable here: the pins are the
    hand-evaluated closed forms of tests/test_bleu_cpu.py)."""
    hyp_len, ref_len = len(hypothesis), len(reference)
    if hyp_len == 0:
        return 0.0
    p_num, p_den = [], []
    for n in range(1, 5):
        h = Counter(tuple(hypothesis[i:i + n]) for i in range(hyp_len - n + 1))
        r = Counter(tuple(reference[i:i + n]) for i in range(ref_len - n + 1))
        p_num.append(sum(min(c, r[g]) for g, c in h.items()))
        p_den.append(max(1, sum(h.values())))
    if p_num[0] == 0:
        return 0.0
    bp = 1.0 if hyp_len > ref_len else math.exp(1 - ref_len / hyp_len)
    incvnt = 1
    p = []
    for num, den in zip(p_num, p_den):
        if num == 0 and hyp_len > 1:
            numerator = 1 / (2 ** incvnt * 5 / math.log(hyp_len))
            p.append(numerator / den)
            incvnt += 1
        else:
            p.append(num / den)
    # nltk sums w_i * log(p_i) over the p_i > 0 only (corpus_bleu): a 1-token hypothesis (method4 does not smooth when
    # len(hyp) == 1) scores bp * p_1 ** 0.25, not 0
    return bp * math.exp(math.fsum(0.25 * math.log(x) for x in p if x > 0))


def _sentence(ys: np.ndarray) -> Optional[List[int]]:
    """Tokens between <s> and the first </s> (:376-386); None when there is no </s>."""
    toks = [int(t) for t in ys if int(t) != W.PAD_ID]
    if W.EOS_ID not in toks:
        return None
    return toks[1:toks.index(W.EOS_ID)]


def classify(golden: np.ndarray, faulty: np.ndarray) -> Dict[str, object]:
    """Outcome of one trial: masked (faulty BLEU == golden BLEU), changed, or no-EOS (the "op,0,0,bit,type" row)."""
    g, f = _sentence(golden), _sentence(faulty)
    if g is None or f is None:
        return {"outcome": "no-EOS", "golden_bleu": 0, "faulty_bleu": 0, "tokens_equal": bool(np.array_equal(golden, faulty))}
    gb, fb = sentence_bleu_method4(g, g), sentence_bleu_method4(g, f)
    return {"outcome": "masked" if gb == fb else "changed", "golden_bleu": gb, "faulty_bleu": fb,
            "tokens_equal": bool(np.array_equal(golden, faulty))}


def csv_row(trial: Trial, result: Dict[str, object]) -> str:
    bit = None if "RANDOM" in trial.inject_type else trial.bit
    return "%s,%s,%s,%s,%s\n" % (trial.op_name, result["golden_bleu"], result["faulty_bleu"], bit, trial.inject_type)


# ---------------------------------------------------------------------------------------------- runner
def run_trials(engine, src_ids: np.ndarray, src_mask: np.ndarray, trials: List[Trial], csv_path: Optional[str] = None,
               rank: int = 0, world: int = 1) -> List[Dict[str, object]]:
    """Golden decode of all sentences in one batch, then one faulty decode per trial of this rank's shard."""
    import torch
    from .engine import FaultSpec
    dev = engine.dev
    ids = torch.from_numpy(src_ids).to(dev)
    mask = torch.from_numpy(src_mask).to(dev)
    golden = engine.greedy_decode(ids, mask).cpu().numpy()
    memory_cache: Dict[int, "torch.Tensor"] = {}
    out = []
    done = set()
    if csv_path and os.path.exists(csv_path + ".ids"):
        done = {int(x) for x in open(csv_path + ".ids").read().split()}     # resume: skip trial ids already written
    for trial in trials[rank::world]:
        if trial.trial_id in done:
            continue
        b = trial.sentence
        spec = FaultSpec(trial.module, trial.layer, trial.target, trial.inject_type, trial.bit, trial.flat_index, trial.window_start,
                         trial.window_len, trial.value_bits, step=0)
        memory = None
        if trial.module == "Decoder":
            if b not in memory_cache:
                memory_cache[b] = engine.encode(ids[b:b + 1], mask[b:b + 1]).clone()
            memory = memory_cache[b]
        faulty = engine.greedy_decode(ids[b:b + 1], mask[b:b + 1], fault=spec, memory=memory).cpu().numpy()[0]
        res = classify(golden[b], faulty)
        res.update(asdict(trial))
        out.append(res)
        if csv_path:
            with open(csv_path, "a") as f:
                f.write(csv_row(trial, res))
            with open(csv_path + ".ids", "a") as f:
                f.write("%d\n" % trial.trial_id)
    return out


def trials_per_decode(engine) -> int:
    """The batch that fills ONE wave of the cluster-resident decoder on this GPU: 8 sentences per co-resident 8-CTA cluster (15
    clusters = 120 trials on a B200).  A greedy step takes the same time for 1 cluster as for 15 (the chain of exchanges sets it, not
    the work), so 120 trials decode in the time of 64: 5.1 k -> 8.4 k trials/s on one GPU (tools/campaign_batch_sweep.py); at 128 the
    launch needs a second wave and the rate halves."""
    from . import kernels as K
    if getattr(engine, "decoder", "cluster") != "cluster" or not getattr(engine, "persistent", True):
        return 64
    return K.cdecoder_max_sentences()


def run_trials_batched(engine, src_ids: np.ndarray, src_mask: np.ndarray, trials: List[Trial], batch: Optional[int] = None,
                       csv_path: Optional[str] = None, rank: int = 0, world: int = 1, return_tokens: bool = False) -> List[Dict[str, object]]:
    """Same outcomes as run_trials, `batch` trials per greedy decode (default: trials_per_decode(engine)): row b of a batch re-decodes
    trial b's sentence with trial b's fault (one fault per batch unit: ot_linear_w8a8_mf / ot_attention_q8_mf).  The kernels are
    batch-invariant, so every row equals the batch-1 decode the reference would run."""
    import torch
    from .engine import FaultSpec
    dev = engine.dev
    if batch is None:
        batch = trials_per_decode(engine)
    ids = torch.from_numpy(src_ids).to(dev)
    mask = torch.from_numpy(src_mask).to(dev)
    golden = engine.greedy_decode(ids, mask).cpu().numpy()
    from .parallel import shard
    mine = shard(trials, rank, world)
    if csv_path and os.path.exists(csv_path + ".ids"):
        done = {int(x) for x in open(csv_path + ".ids").read().split()}     # resume: skip the trial ids whose rows are already written
        mine = [t for t in mine if t.trial_id not in done]
    out = []
    # Two batches in flight: while the GPU decodes batch i (one long persistent-decoder launch), the host classifies batch i-1 and
    # launches the encoder / fault step of batch i+1.  Nothing in the loop synchronises the stream: index and result buffers are
    # pinned, copies are asynchronous, a batch's tokens are read after its own event.
    rows_pin = [torch.zeros(batch, dtype=torch.int64).pin_memory() for _ in range(2)]
    ys_pin = [torch.zeros((batch, engine.max_len), dtype=torch.int64).pin_memory() for _ in range(2)]

    def finish(item):
        chunk, buf, ev = item
        ev.synchronize()
        faulty = buf.numpy()
        for k, trial in enumerate(chunk):
            res = classify(golden[trial.sentence], faulty[k])
            res.update(asdict(trial))
            if return_tokens:
                res["golden_ys"], res["faulty_ys"] = golden[trial.sentence].copy(), faulty[k].copy()
            out.append(res)
            if csv_path:
                with open(csv_path, "a") as f:
                    f.write(csv_row(trial, res))
                with open(csv_path + ".ids", "a") as f:
                    f.write("%d\n" % trial.trial_id)

    # Batches of ONE kind (when the shard mixes modules; a directory-driven campaign has one module anyway).  A batch of Encoder-target
    # trials has no injected decoder step: all of its greedy steps run inside the persistent decoder.  A batch of Decoder-target trials
    # needs no encoder pass: the encoder output of its rows is the golden one, computed once for the source sentences (as run_trials
    # does per sentence).  Rows are appended to the CSV as their batch completes; the returned list keeps the order of `trials`.
    position = {t.trial_id: k for k, t in enumerate(mine)}
    enc_part = [t for t in mine if t.module != "Decoder"]
    dec_part = [t for t in mine if t.module == "Decoder"]
    golden_memory = engine.encode(ids, mask).clone() if dec_part else None
    chunks = [(enc_part[c0:c0 + batch], False) for c0 in range(0, len(enc_part), batch)]
    chunks += [(dec_part[c0:c0 + batch], True) for c0 in range(0, len(dec_part), batch)]
    pending = None
    for i, (chunk, dec_only) in enumerate(chunks):
        specs = [FaultSpec(t.module, t.layer, t.target, t.inject_type, t.bit, t.flat_index, t.window_start, t.window_len, t.value_bits, step=0)
                 for t in chunk]
        rp = rows_pin[i % 2]
        rp.zero_()                                  # keep one workspace / CUDA graph shape: pad with fault-free rows of sentence 0
        rp[:len(chunk)] = torch.tensor([t.sentence for t in chunk], dtype=torch.int64)
        specs = specs + [None] * (batch - len(chunk))
        rows = rp.to(dev, non_blocking=True)
        ys = engine.greedy_decode(ids[rows].contiguous(), mask[rows].contiguous(), fault=specs,
                                  memory=golden_memory[rows] if dec_only else None)
        buf = ys_pin[i % 2]
        buf[:, :ys.shape[1]].copy_(ys, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        if pending is not None:
            finish(pending)
        pending = (chunk, buf, ev)
    if pending is not None:
        finish(pending)
    out.sort(key=lambda r: position[r["trial_id"]])
    return out


def main(argv=None):
    """Same three flags as the reference (parallelized_inject_onnx_transformer.py:47-52): `--directory_name` is the directory of
    fault-target JSON files (input/encoder, input/decoder), iterated x 6 fault models x 8 bit positions x 5 experiments exactly like
    :794-861; `--module` Encoder | Decoder; `--experiment_output_name` the CSV.  When the directory does not exist (the reference's
    input/ tree is not shipped with this package) the same descriptors are re-created from the graph (faults.targets_from_graph:
    equal to the 60 reference files, tests/test_graph_executor.py).  Extra flags: --batch/--src-len/--seed, --trials N (cfg5: N random
    trials from make_trials instead of the directory loop), one process per GPU under torchrun (trials sharded by rank)."""
    ap = argparse.ArgumentParser()
    ap.add_argument("--directory_name", default="input/encoder")
    ap.add_argument("--module", default="Encoder")
    ap.add_argument("--experiment_output_name", default="results_fault_injection/results.csv")
    ap.add_argument("--batch", type=int, default=64, help="source sentences the trials draw from")
    ap.add_argument("--trials-per-decode", type=int, default=0, help="trials per batched faulty decode (0: one wave of decoder clusters, 120 on a B200)")
    ap.add_argument("--src-len", type=int, default=64)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--trials", type=int, default=0)
    ap.add_argument("--experiments", type=int, default=5)
    args = ap.parse_args(argv)
    import torch
    from . import parallel as P
    from .engine import QuantizedTransformer
    rank, world, local_rank = P.rank_world()
    torch.cuda.set_device(local_rank)
    P.init_process_group()
    fw = W.init_float_weights(args.seed)
    eng = QuantizedTransformer(fw)
    ids, mask = W.synthetic_tokens(args.seed, args.batch, args.src_len)
    if args.trials > 0:
        trials = make_trials(args.trials, args.seed, args.batch, args.src_len, modules=(args.module,))
    else:
        if os.path.isdir(args.directory_name):
            targets = load_targets(args.directory_name)
        else:
            from . import faults
            from . import graph as G
            from .engine import _Linear  # noqa: F401  (the engine's weights are the graph's: same state_dict)
            builder = G.build_encoder_graph if args.module == "Encoder" else G.build_decoder_graph
            targets = faults.targets_from_graph(builder(_fake_quantized(fw), batch=1), args.module)
        targets = [t for t in targets if t["module"].split("/")[0] == args.module]
        trials = trials_from_targets(targets, args.seed, args.batch, args.src_len, experiments=args.experiments)
    os.makedirs(os.path.dirname(args.experiment_output_name) or ".", exist_ok=True)
    path = args.experiment_output_name if world == 1 else "%s.rank%d" % (args.experiment_output_name, rank)
    res = run_trials_batched(eng, ids, mask, trials, args.trials_per_decode or None, path, rank, world)
    res_all = P.gather_records([(r["trial_id"], r["outcome"]) for r in res], world)
    P.destroy_process_group()
    if rank == 0:
        print(len(trials), "trials", Counter(o for _, o in res_all))


def _fake_quantized(float_weights):
    """W8A8Linear.from_float on the host (quant_linear.py:122-147: stored weight = round(W/s)*s per output channel), the state the
    reference's exported graphs hold -- numpy fp32 in the op order of the exporter, for the graph builder only."""
    out = dict(float_weights)
    f32 = np.float32
    for k, v in float_weights.items():
        if k.endswith(".weight") and v.ndim == 2 and (".linears." in k or ".feed_forward." in k):
            s = (np.maximum(np.max(np.abs(v), axis=-1, keepdims=True), f32(1e-5)) / f32(127.0)).astype(f32)
            out[k] = (np.rint((v / s).astype(f32)) * s).astype(f32)
    return out


if __name__ == "__main__":
    main()
