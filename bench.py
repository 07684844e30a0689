#!/usr/bin/env python
"""bench.py -- decoded tokens/s of the int8 Transformer-base greedy-decode hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm (oracle port) on host cores

Workload = BASELINE.json configs[1]: 8-bit Transformer-base (6+6 layers, d512, h8, ff2048, vocab 5337/4444),
random-init weights (seed 0), synthetic source batch of 64 sentences x 64 tokens per GPU, 71 greedy steps.
A "step" is one pass of the hot path over one batch: encoder + cross-K/V projection + 71 greedy decoder steps ->
64*71 decoded tokens per GPU.  N > 1: one process per GPU (torchrun), sentences sharded by rank (weak scaling),
the only collective is the all_gather of the token ids.

One JSON line on rank 0: value (inputs resident in HBM), e2e (host buffers, H2D/D2H inside the timed region),
roofline of the dominant kernel family (measured live with CUDA events), cpu_baseline (oracle port on the host).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "decoded tokens/s (int8 Transformer-base, bs64)"
UNIT = "tokens/s"
B_DEFAULT, S_DEFAULT, MAX_LEN = 64, 64, 72


# ----------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        def reader():
            for line in self.proc.stdout:
                self.samples.append((time.time(), line.strip()))
        self.thread = threading.Thread(target=reader, daemon=True)
        self.thread.start()

    def mark_begin(self):
        """Samples from here on belong to the timed region (nvidia-smi itself is started earlier: it needs ~100 ms to come up)."""
        self.t_begin = time.time()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t_begin = getattr(self, "t_begin", 0.0)
        timed = [line for (t, line) in self.samples if t >= t_begin]
        window = "timed region"
        if len(timed) < 2:            # a very short timed region: fall back to the samples taken under the same load just before it
            timed = [line for (_, line) in self.samples][-5:]
            window = "warm-up + timed region"
        if not timed:                 # nothing at all: one synchronous query, flagged as such
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=10).stdout.strip().splitlines()
                timed, window = out[:1], "single query after the timed region"
            except Exception:
                pass
        for line in timed:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


# ----------------------------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_reference_sample(batch: int, src_len: int, greedy_steps: int, seed: int = 0):
    """Time the oracle port of the reference algorithm on the host cores: fp32 MatMul of de-quantized operands
    ("ref-float"), weight fake-quant recomputed every call, decoder re-run on the FULL prefix each step -- exactly
    the work the reference's executor does, minus its per-node session overhead.  Returns (tokens/s, seconds)."""
    from onnx_transformer_b200 import weights as W
    from oracle import model as om
    fw = W.init_float_weights(seed)
    w = om.get_quantized(fw, None, 6)
    ids, mask = W.synthetic_tokens(seed, batch, src_len)
    t0 = time.perf_counter()
    om.greedy_decode(w, ids, mask, greedy_steps + 1, 0, "ref-float", 6, kv_cache=False)
    dt = time.perf_counter() - t0
    return batch * greedy_steps / dt, dt


def blas_threads(want: int) -> int:
    """Give the numpy BLAS pool (and torch's intra-op pool) `want` threads -- torchrun exports OMP_NUM_THREADS=1 to its workers -- and
    return the pool size actually in effect (threadpoolctl), which is what the cpu_baseline `cores` field reports."""
    import torch
    torch.set_num_threads(want)
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=want)
        sizes = [int(p.get("num_threads", 1)) for p in threadpool_info() if p.get("user_api") in ("blas", "openmp")]
        return max(sizes) if sizes else 1
    except Exception:
        return 1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = blas_threads(os.cpu_count() or 1)
    sample_b, sample_steps = 16, 16          # ~3 s of CPU work per timed step on 16 cores
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_sample(2, args.src_len, 2)
    vals, secs = [], []
    for _ in range(args.steps):
        v, dt = cpu_reference_sample(sample_b, args.src_len, sample_steps)
        vals.append(v); secs.append(dt)
    value = float(np.mean(vals))
    sample = ("oracle port (numpy, ref-float, full-prefix recompute, %d BLAS threads in effect): %d sentences x %d src tokens, first %d of 71 "
              "greedy steps per timed step -- a SAMPLE of cfg2, not the same config: early steps have the shortest prefixes, so this over-states "
              "the CPU's full-run rate (the GPU/CPU ratio is conservative)" % (cores, sample_b, args.src_len, sample_steps))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(np.mean(secs) * 1e3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": "cfg2: greedy decode, batch 64 x src 64, 71 steps (bounded CPU sample)", "batch": args.batch,
                                            "src_len": args.src_len, "greedy_steps": MAX_LEN - 1},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------- kernel-family timing
FAMILY = {1: "gemm_i8", 2: "attention_q8", 3: "attention_q8", 4: "layernorm_quant", 5: "rowquant", 6: "embed_pe", 7: "generator", 8: "generator",
          9: "append_token"}


def decode_family_timeline(eng, ws, B, S, reps=8, step=35):
    """In-situ kernel durations of the greedy step: the captured step graph is replayed `reps` times (at a mid-run prefix
    length) with the library's device timeline on -- block 0 of every kernel stamps %globaltimer when its dependency wait
    ends and when it finishes -- and bracketed by CUDA events on the launching stream.  Returns per family
    {launches_per_step, us_per_launch, us_per_step} and the event-timed step duration."""
    import ctypes as C
    import torch
    from onnx_transformer_b200 import _lib
    lib = _lib.load()
    cap = 128 * reps
    buf = torch.zeros(1 + 4 * cap, dtype=torch.int64, device=eng.dev)
    ws["step"].fill_(step)
    ws["graph"].replay()
    torch.cuda.synchronize()
    ws["step"].fill_(step)
    lib.ot_set_timeline(C.c_void_p(buf.data_ptr()), cap)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ws["graph"].replay()
    e1.record()
    torch.cuda.synchronize()
    lib.ot_set_timeline(None, 0)
    t = buf.cpu().numpy()
    n = min(int(t[0]), cap)
    rec = t[1:1 + 4 * n].reshape(n, 4)
    fam = {}
    for kid, _, ready, end in rec:
        if end <= ready:      # a kernel whose block 0 does not stamp its end
            continue
        name = FAMILY.get(int(kid), "other")
        f = fam.setdefault(name, [0, 0.0])
        f[0] += 1
        f[1] += (end - ready) / 1e3
    out = {k: {"launches_per_step": c / reps, "us_per_launch": tot / c, "us_per_step": tot / reps} for k, (c, tot) in fam.items()}
    ws["step"].zero_()
    return out, e0.elapsed_time(e1) * 1e3 / reps


def decode_family_bytes(eng, B, S, T=36):
    """Algorithmic bytes per launch of each family for one greedy step of the per-op path (DESIGN.md section 4 / SURVEY.md 8d)."""
    D, FF, nl, V = 512, 2048, eng.n_layers, eng.vocab
    w_bytes = 3 * D * D + 3 * D * D + 2 * D * FF                       # int8 weights of the 6 GEMMs of a decoder layer: 3,670,016 B
    io = B * (D + 3 * D + 12) + 3 * B * (D + 4 + 2 * D * 4) + B * (D + D + 4) + B * (D + FF + 4) + B * (FF + 4 + 2 * D * 4)
    par = 8 * (3 * D + 3 * D + FF + D)                                 # fp32 scale + bias per output column
    gemm = (w_bytes + io + par) / 6.0
    attn = (B * (T * (2 * D + 8) + 3 * D + D + 16) + B * (S * (2 * D + 8 + 1) + D + D + 8)) / 2.0
    ln = B * (D * 4 + D + 4) + 2 * D * 4
    gen = (V * D * 4 + B * D * 4 + 3 * B * V * 4) / 2.0
    return {"gemm_i8": gemm, "attention_q8": attn, "layernorm_quant": ln, "generator": gen}


def persistent_decoder_bytes(eng, B, S, n_steps):
    """Algorithmic bytes of ONE launch of decoder_steps_kernel (all `n_steps` greedy steps): only what any decoder must move per
    step -- every weight once (int8 GEMM weights + fp32 scales/biases/LayerNorm, fp32 generator), the self-attention KV cache
    prefix read + one appended row, the cached cross-attention K/V + scales + mask, the embedding rows and the token ids.  The
    int32 accumulator planes / activations exchanged between phases through L2 are NOT counted."""
    D, FF, nl, V = 512, 2048, eng.n_layers, eng.vocab
    w_int8 = nl * (3 * D * D + 3 * D * D + 2 * D * FF)               # 6 x 3,670,016 B
    w_par = nl * (8 * (3 * D + 3 * D + FF + D) + 6 * D * 4)          # per-column scale + bias, LayerNorm gamma/beta
    gen = V * D * 4 + V * 4 + 2 * D * 4                              # generator weight + bias, final norm
    cross = nl * B * S * (2 * D + 8) + B * S
    per_step_const = w_int8 + w_par + gen + cross + B * (D * 4 + 8) + D * 4
    total = 0
    for t in range(n_steps):
        total += per_step_const + nl * B * (t * (2 * D + 8) + (2 * D + 8))
    return total, {"int8_weights": w_int8, "scales_bias_ln": w_par, "generator_fp32": gen, "cross_kv": cross,
                   "self_kv_read_mean": nl * B * ((n_steps - 1) / 2.0) * (2 * D + 8)}


def persistent_phase_trace(eng, ws, B, S, t_mid=35):
    """Per-phase timeline of one greedy step inside the persistent decoder kernel, stamped by CTA 0 with %globaltimer.
    cluster decoder (ot_cdecoder.cu xwait): t[2i] = phase i starts waiting for its input bytes, t[2i+1] = they have arrived,
    t[254] = step end; grid decoder (ot_decoder.cu grid_sync): t[2i] / t[2i+1] = arrival at / release from barrier i.
    Returns {phase: us per greedy step} and the step total."""
    import torch
    from onnx_transformer_b200 import kernels as K
    plan = eng._decoder_plan(ws, B, S, trace=True)
    plan.run(t_mid, 2)
    torch.cuda.synchronize()
    t = plan.trace.cpu().numpy()
    out = {}
    if isinstance(plan, K.ClusterDecoderPlan):
        names = plan.phase_names()
        for i, n in enumerate(names):
            end = t[2 * (i + 1)] if i + 1 < len(names) else t[254]
            out[n] = out.get(n, 0.0) + (end - t[2 * i]) / 1e3
        total = (t[254] - t[255]) / 1e3
    else:
        names = ["ln1", "qkv", "self_attn", "o", "ln2", "cq", "cross_attn", "co", "ln3", "ffn1_mma", "ffn1_quant", "ffn2"] * eng.n_layers
        names += ["final_norm", "generator"]
        prev = t[255]
        for i, n in enumerate(names):
            out[n] = out.get(n, 0.0) + (t[2 * i + 1] - prev) / 1e3
            prev = t[2 * i + 1]
        total = (t[2 * len(names) - 1] - t[255]) / 1e3
    eng._last_trace = t          # raw stamps (tools/spc_sweep.py prints the generator's internal marks)
    ws.pop("plan", None)      # drop the tracing plan: the next decode rebuilds the plain one
    return {k: round(float(v), 2) for k, v in out.items()}, float(total)


def gemm_roofline(dev, M=65536):
    """The metric's second half, "int8 GEMM % peak": the four linear-layer GEMMs of one cfg3 encoder layer (M = 512 x 128 tokens), each
    event-timed on the launching stream as REPS back-to-back launches after an L2 flush (one launch between two events would also time
    the host's launch path while the GPU idles).  achieved = 2 M N K / t; peak = 4.5 POPS nominal dense int8 (tcgen05 kind::i8) -- the
    issue-rate microbenchmark of this repo measures 128 x 256 x 32 MACs per 132.5 clk per SM = 4.6 POPS at 1965 MHz
    (profiles/r2_mma_issue_microbench.txt); MEASURED_PEAKS.json has no int8 figure."""
    import torch
    from onnx_transformer_b200 import kernels as K
    g = torch.Generator(device=dev).manual_seed(0)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    shapes = [("qkv_to_q8", 1536, 512, dict(out_kind=K.OUT_Q8, quant_group=512), False, "ot_gemm_wres.cu"),
              ("o_to_f32_residual", 512, 512, dict(out_kind=K.OUT_F32), True, "ot_gemm_stream.cu"),
              ("ffn1_relu_to_q8", 2048, 512, dict(out_kind=K.OUT_Q8, quant_group=2048, relu=True), False, "ot_gemm_wres.cu"),
              ("ffn2_to_f32_residual", 512, 2048, dict(out_kind=K.OUT_F32), True, "ot_gemm_stream.cu")]
    out, reps = {}, 4
    for name, N, Kd, kw, res, src in shapes:
        a = torch.randint(-127, 128, (M, Kd), dtype=torch.int8, device=dev, generator=g)
        w = torch.randint(-127, 128, (N, Kd), dtype=torch.int8, device=dev, generator=g)
        sx = torch.rand(M, device=dev, generator=g) * 0.05 + 1e-3
        sw = torch.rand(N, device=dev, generator=g) * 0.01 + 1e-4
        b = torch.randn(N, device=dev, generator=g)
        r = torch.randn(M, N, device=dev, generator=g) if res else None
        for _ in range(2):
            K.linear_w8a8(a, w, row_scale=sx, col_scale=sw, bias=b, residual=r, **kw)
        ts = []
        for _ in range(5):
            flush.zero_()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                K.linear_w8a8(a, w, row_scale=sx, col_scale=sw, bias=b, residual=r, **kw)
            e1.record()
            torch.cuda.synchronize(dev)
            ts.append(e0.elapsed_time(e1) * 1e3 / reps)
        us = float(np.median(ts))
        ops = 2.0 * M * N * Kd
        nbytes = M * Kd + N * Kd + (M * N * (4 if res else 0)) + (M * N * (4 if kw["out_kind"] == K.OUT_F32 else 1))
        out[name] = {"M": M, "N": N, "K": Kd, "us": us, "achieved_top_per_s": ops / us / 1e6, "frac_of_4.5_POPS": ops / us / 1e6 / 4500.0,
                     "algorithmic_gb_per_s": nbytes / us / 1e3, "kernel": src}
        del a, w, r
    del flush
    torch.cuda.empty_cache()
    return {"peak_top_per_s": 4500.0, "peak_source": "nominal dense int8; measured MMA issue rate 4.6 POPS (profiles/r2_mma_issue_microbench.txt)",
            "timing": "median of 5 groups of %d back-to-back launches, CUDA events, 512 MB L2 flush before each group" % reps, "gemm": out}


def sharded_measurements(eng, dev, rank, world):
    """The other sharded BASELINE.json configs at this N (every rank runs its shard; times are the MAX over ranks):
    cfg3 -- encoder-only forward, 512 x 128 in total, sentences block-partitioned over the ranks (STRONG scaling);
    cfg5 -- fault-injection trials/s, 2048 single-fault trials sharded by rank (64 trials per faulty batch decode)."""
    import torch
    from onnx_transformer_b200 import campaign as C
    from onnx_transformer_b200 import parallel as P
    from onnx_transformer_b200 import weights as W
    out = {}
    ids_np, mask_np = W.synthetic_tokens(7, 512, 128)
    lo, hi = P.shard_rows(512, rank, world)
    ids, mask = torch.from_numpy(ids_np[lo:hi]).to(dev), torch.from_numpy(mask_np[lo:hi]).to(dev)
    for _ in range(2):
        eng.encode(ids, mask)
    torch.cuda.synchronize(dev)
    P.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
        eng.encode(ids, mask)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = P.max_over_ranks([e0.elapsed_time(e1) / reps], device=dev)[0]
    tokens = 512 * 128
    ops = tokens * 37.75e6 + 6 * 2 * (512 * 8 * 128 * 128 * 64)          # linears + QK^T, 2 ops per MAC
    out["cfg3_encoder_only"] = {"batch_total": 512, "batch_per_gpu": hi - lo, "src_len": 128, "n_gpus": world, "scaling": "strong", "ms": ms,
                                "tokens_per_s": tokens / (ms * 1e-3), "int8_top_per_s": ops / (ms * 1e-3) / 1e12,
                                "frac_of_4.5_POPS_per_gpu": ops / (ms * 1e-3) / 4.5e15 / world}
    del ids, mask
    eng._enc_ws = {}
    torch.cuda.empty_cache()
    ids_np, mask_np = W.synthetic_tokens(11, 64, 64)
    tpd = C.trials_per_decode(eng)          # one wave of decoder clusters: 120 trials per faulty decode on a B200
    trials = C.make_trials(64 * tpd, 0, 64, 64)
    C.run_trials_batched(eng, ids_np, mask_np, trials[:tpd * world], tpd, None, rank, world)            # warm-up
    torch.cuda.synchronize(dev)
    P.barrier()
    t0 = time.perf_counter()
    res = C.run_trials_batched(eng, ids_np, mask_np, trials, tpd, None, rank, world)
    torch.cuda.synchronize(dev)
    dt = P.max_over_ranks([time.perf_counter() - t0], device=dev)[0]
    from collections import Counter
    counts = Counter(r["outcome"] for r in res)
    names = sorted(set(counts) | {"masked", "changed", "no-EOS"})
    tot = P.sum_over_ranks([counts.get(k, 0) for k in names], device=dev)
    out["cfg5_fault_injection"] = {"trials": len(trials), "n_gpus": world, "trials_per_s": len(trials) / dt, "trials_per_decode": tpd,
                                   "outcomes": {k: int(v) for k, v in zip(names, tot) if v},
                                   "note": "wall clock, max over ranks, incl. one golden batch decode per rank; trials_per_decode trials per faulty greedy "
                                           "decode (one fault per batch row; one wave of decoder clusters: the step time does not depend on how many of the "
                                           "15 co-resident clusters are in use); random-init weights never emit </s> (tests/test_fullsize_parity_gpu.py "
                                           "pins the three outcome classes on an EOS-capable model)"}
    return out


def rank0_measurements(eng, dev):
    """Single-GPU side measurements (rank 0): cfg4 (4-bit weights) batch decode and cfg1 (the drop-in executor driven like the
    reference's greedy_decode: B = 1, S = 72, 71 full-prefix decoder passes)."""
    import torch
    from onnx_transformer_b200 import weights as W
    out = {}
    ids_np, mask_np = W.synthetic_tokens(11, 64, 64)
    try:
        eng4 = type(eng)(W.init_float_weights(0), n_layers=6, max_len=MAX_LEN, weight_bits=4)
        i4, m4 = torch.from_numpy(ids_np).to(dev), torch.from_numpy(mask_np).to(dev)
        for _ in range(2):
            eng4.greedy_decode(i4, m4)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            eng4.greedy_decode(i4, m4)
        e1.record()
        torch.cuda.synchronize(dev)
        ms4 = e0.elapsed_time(e1) / 3
        out["cfg4_int4_weights_decode"] = {"batch": 64, "src_len": 64, "ms": ms4, "tokens_per_s": 64 * (MAX_LEN - 1) / (ms4 * 1e-3),
                                           "note": "encoder / cross-K/V: the requant GEMMs take the packed int4 tile and unpack it once per launch in shared memory (weight-stationary kernel), "
                                                   "the fp32-output GEMMs run on a per-launch unpacked int8 scratch; the cluster decoder reads the int8 copy of the 4-bit values"}
        del eng4
        torch.cuda.empty_cache()
    except Exception as exc:      # side measurement: never fail the headline line
        out["cfg4_int4_weights_decode"] = {"error": str(exc)[:200]}
    try:
        # NOT the headline configuration (BASELINE names batch 64): the same greedy decode with one full wave of decoder clusters in use
        from onnx_transformer_b200 import kernels as K
        bw = K.cdecoder_max_sentences()
        iw_np, mw_np = W.synthetic_tokens(1000, bw, 64)
        iw, mw = torch.from_numpy(iw_np).to(dev), torch.from_numpy(mw_np).to(dev)
        for _ in range(2):
            eng.greedy_decode(iw, mw)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            eng.greedy_decode(iw, mw)
        e1.record()
        torch.cuda.synchronize(dev)
        msw = e0.elapsed_time(e1) / 3
        out["decode_one_cluster_wave"] = {"batch": bw, "src_len": 64, "ms": msw, "tokens_per_s": bw * (MAX_LEN - 1) / (msw * 1e-3),
                                          "note": "not the headline config: batch = 8 sentences x the decoder clusters one GPU keeps resident (15 on a B200); a greedy step "
                                                  "takes the same time for 8 clusters (batch 64) as for 15"}
    except Exception as exc:
        out["decode_one_cluster_wave"] = {"error": str(exc)[:200]}
    try:
        out["cfg1_executor_greedy_decode"] = cfg1_executor_decode(dev)
    except Exception as exc:
        out["cfg1_executor_greedy_decode"] = {"error": str(exc)[:300]}
    return out


def cfg1_graphs(seed=0):
    from onnx_transformer_b200 import campaign as C
    from onnx_transformer_b200 import graph as G
    from onnx_transformer_b200 import weights as W
    fw = W.init_float_weights(seed)
    w = C._fake_quantized(fw)
    return fw, G.build_encoder_graph(w, batch=1, n_layers=6), G.build_decoder_graph(w, batch=1, n_layers=6)


def cfg1_executor_decode(dev):
    """BASELINE.json configs[0] on the GPU: greedy_decode of 8-bit_onnx_optimized_custom_inference.py:649-721 (prepare_inference ->
    run_module("Encoder") -> 71 x run_module("Decoder") on the full prefix) through this package's drop-in executor, B = 1, S = 72."""
    import torch
    from onnx_transformer_b200 import decode as D
    from onnx_transformer_b200 import executor as X
    from onnx_transformer_b200 import weights as W
    fw, enc, dec = cfg1_graphs()
    model = D.HostModel(fw, device=dev)
    ids, mask = W.synthetic_tokens(1000, 1, 72)
    D.greedy_decode(model, ids, mask, 4, 0, enc, dec, executor=X)          # warm-up (3 decoder passes)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    timings = {}
    ys = D.greedy_decode(model, ids, mask, MAX_LEN, 0, enc, dec, executor=X, timings=timings)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    steps = timings["decoder_step_s"]
    # the same decode with whole-pass CUDA-graph replay in run_module (executor.enable_graph_replay): the first sweep over the 71 prefix
    # lengths captures one graph per length, the following sweeps replay them -- the steady state of a caller that decodes many sentences
    replay = None
    try:
        X.enable_graph_replay(True)
        replay = {"note": "the reference's greedy loop with the two prepared weight_dicts kept across sentences (a captured pass belongs to "
                          "its weight_dict); sweep 0 walks / captures one graph per prefix length, sweep 3 (timed: tokens_per_s) replays"}
        replay.update(cfg1_replay_loop(model, ids, mask, enc, dec, dev))
        replay["tokens_equal_to_node_walk"] = bool(torch.equal(replay.pop("ys"), ys))
    except Exception as exc:       # noqa: BLE001
        replay = {"error": str(exc)[:300]}
    finally:
        X.enable_graph_replay(False)
    return {"batch": 1, "src_len": 72, "greedy_steps": MAX_LEN - 1, "seconds": dt, "tokens_per_s": (MAX_LEN - 1) / dt, "graph_replay": replay,
            "encoder_s": timings["encoder_s"], "decoder_pass_ms_first_last": [steps[0] * 1e3, steps[-1] * 1e3], "nodes_per_decoder_pass": len(dec.node),
            "tokens_head": [int(t) for t in ys[0, :8].tolist()],
            "note": "wall clock of the Python node walk (one libot_b200.so handler per ONNX node, full-prefix recompute as in the reference); "
                    "the fused engine (value / e2e) is the fast path"}


def cfg1_replay_loop(model, ids, mask, enc, dec, dev):
    """The reference's greedy loop (8-bit_onnx_optimized_custom_inference.py:649-721) with the two prepared weight_dicts kept across
    sentences: sweep 0 walks / captures, sweeps 1-2 replay; the last sweep is timed."""
    import torch
    from onnx_transformer_b200 import decode as D
    from onnx_transformer_b200 import executor as X
    src = torch.from_numpy(ids).to(dev)
    m = torch.from_numpy(mask).to(dev)
    enc_wd = dec_wd = None
    out = {}
    for sweep in range(4):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        enc_in = {"global_in": model.get_src_embed(src), "global_in_1": m}
        if enc_wd is None:
            enc_wd, enc_g = X.prepare_inference(enc, enc_in)
        memory, _ = X.run_module("Encoder", enc_in, enc, enc_wd, enc_g)
        memory = memory[list(memory.keys())[0]]
        ys = torch.full((1, 1), 0, dtype=torch.int64, device=dev)
        for i in range(MAX_LEN - 1):
            dec_in = {"global_in": model.get_tgt_embed(ys), "global_in_1": memory, "global_in_2": m,
                      "global_in_3": torch.from_numpy(D.subsequent_mask(ys.shape[1])).to(dev)}
            if dec_wd is None:
                dec_wd, dec_g = X.prepare_inference(dec, dec_in)
            o, _ = X.run_module("Decoder", dec_in, dec, dec_wd, dec_g)
            o = o[list(o.keys())[0]]
            ys = torch.cat([ys, model.next_word(o[:, -1]).reshape(1, 1)], dim=1)
        torch.cuda.synchronize(dev)
        out["sweep%d_s" % sweep] = time.perf_counter() - t0
    out["tokens_per_s"] = (MAX_LEN - 1) / out["sweep3_s"]
    out["stats"] = dict(X.replay_stats)
    out["ys"] = ys
    return out


def cfg1_cpu_walk(steps=6):
    """The reference's CPU executor restated (oracle node walk, ref-float = the reference's own float MatMuls) on cfg1: B = 1, S = 72,
    first `steps` of the 71 full-prefix decoder passes (the early passes have the shortest prefixes: this over-states the CPU rate)."""
    from onnx_transformer_b200 import weights as W
    from oracle import executor as oe
    fw, enc, dec = cfg1_graphs()
    ids, mask = W.synthetic_tokens(1000, 1, 72)
    timings = {}
    t0 = time.perf_counter()
    oe.greedy_decode(fw, enc, dec, ids, mask, steps + 1, 0, "ref-float", timings=timings)
    dt = time.perf_counter() - t0
    return {"batch": 1, "src_len": 72, "greedy_steps_timed": steps, "seconds": dt, "tokens_per_s": steps / dt, "encoder_s": timings.get("encoder_s"),
            "sample": "oracle node walk (numpy, ref-float), first %d of 71 decoder passes" % steps}


# ----------------------------------------------------------------------------------------------- main (GPU arm)
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=B_DEFAULT)
    ap.add_argument("--src-len", type=int, default=S_DEFAULT)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the cfg3 encoder-only and cfg5 fault-injection side measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    # rank 0 must print ONE JSON line on stdout: library chatter (the NCCL version banner ...) goes to stderr until then
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    from onnx_transformer_b200 import kernels as K
    from onnx_transformer_b200 import parallel as P
    from onnx_transformer_b200 import weights as W
    from onnx_transformer_b200.engine import QuantizedTransformer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    args.warmup = max(args.warmup, 3)
    rank, world, local_rank = P.rank_world()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    P.init_process_group("nccl")
    B, S = args.batch, args.src_len
    eng = QuantizedTransformer(W.init_float_weights(0), n_layers=6, max_len=MAX_LEN)
    ids_np, mask_np = W.synthetic_tokens(1000 + rank, B, S)          # each rank: its own shard of sentences
    ids, mask = torch.from_numpy(ids_np).to(dev), torch.from_numpy(mask_np).to(dev)
    ids_pin, mask_pin = torch.from_numpy(ids_np).pin_memory(), torch.from_numpy(mask_np).pin_memory()
    ys_pin = torch.empty((B, MAX_LEN), dtype=torch.int64).pin_memory()
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    # The path's only collective is the gather of the decoded ids.  The K batches of a run are independent, so they are gathered ONCE,
    # after the last decode and inside its timed bracket (parallel.gather_token_ids): a gather after every decode made every rank wait
    # for the slowest one K times (sum of per-step maxima instead of the maximum of the per-rank sums: 2 % at 8 GPUs in round 1).
    ys_all = torch.empty((max(args.steps, args.warmup, 2) * B, MAX_LEN), dtype=torch.int64, device=dev)

    def one_step(i, last):
        ys = eng.greedy_decode(ids, mask)
        ys_all[i * B:(i + 1) * B].copy_(ys)
        if last and world > 1:
            P.gather_token_ids(ys_all[:(i + 1) * B], world)
        return ys

    def one_step_e2e(i, last):
        d_ids = ids_pin.to(dev, non_blocking=True)
        d_mask = mask_pin.to(dev, non_blocking=True)
        ys = eng.greedy_decode(d_ids, d_mask)
        ys_pin.copy_(ys, non_blocking=True)
        ys_all[i * B:(i + 1) * B].copy_(ys)
        if last and world > 1:
            P.gather_token_ids(ys_all[:(i + 1) * B], world)
        return ys

    def sync_all():
        P.barrier()
        torch.cuda.synchronize(dev)

    def timed_loop(fn, steps):
        per_step = []
        for i in range(steps):
            flush.zero_()                                # evict L2 between timed iterations (untimed)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(i, i == steps - 1)
            e1.record()
            torch.cuda.synchronize(dev)
            per_step.append(e0.elapsed_time(e1))
        return per_step

    sampler = ClockSampler(local_rank)
    sampler.start()                                      # before the warm-up: nvidia-smi needs ~100 ms to deliver its first sample
    for i in range(args.warmup):
        one_step(i, i == args.warmup - 1)
    sync_all()
    ws = eng._dec_workspace(B, S)
    persistent = ws.get("plan") is not None
    eng.decoder_events = [] if persistent else None     # CUDA events around every decoder_steps_kernel launch of the timed region
    launches_per_graph = ws.get("graph_launches", 0)
    l0, r0, f0 = K._lib.launch_count(), eng.graph_replays, eng.front_replays
    front_launches = (ws.get("front") or {}).get("launches", 0)
    sync_all()
    sampler.mark_begin()
    step_ms = timed_loop(one_step, args.steps)
    total_ms = float(sum(step_ms))
    sync_all()
    # kernels launched directly + kernels replayed by the CUDA graphs (per-op greedy step; encoder + cross-K/V front graph)
    launches = (K._lib.launch_count() - l0) + (eng.graph_replays - r0) * launches_per_graph + (eng.front_replays - f0) * front_launches
    clocks = sampler.stop()
    dec_events, eng.decoder_events = eng.decoder_events, None
    for i in range(2):                                   # the host-buffer path has its own first-call costs (allocations, pinned copies)
        one_step_e2e(i, i == 1)
    sync_all()
    e2e_step_ms = timed_loop(one_step_e2e, args.steps)
    e2e_ms = float(sum(e2e_step_ms))
    sync_all()
    total_ms, e2e_ms = P.max_over_ranks([total_ms, e2e_ms], device=dev)      # a multi-GPU number is the slowest rank's
    sharded = {} if args.no_extra else sharded_measurements(eng, dev, rank, world)
    tokens = world * B * (MAX_LEN - 1) * args.steps
    value = tokens / (total_ms * 1e-3)
    e2e_value = tokens / (e2e_ms * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s"
        n_dec = MAX_LEN - 1
        if persistent:
            # dominant kernel: the persistent decoder, ONE launch per batch decode (all 71 greedy steps); its duration is the mean of
            # the CUDA-event pairs recorded around each launch INSIDE the timed region, on the launching stream
            kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in dec_events]))
            nbytes, parts = persistent_decoder_bytes(eng, B, S, n_dec)
            achieved = nbytes / (kernel_ms * 1e-3) / 1e9
            phases, step_us = persistent_phase_trace(eng, ws, B, S)
            traffic = None
            try:     # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full capture
                traffic = json.load(open(os.path.join(ROOT, "profiles", "decoder_traffic.json")))["dram_bytes_per_launch"]
            except Exception:
                pass
            kname = "cdecoder_kernel (cluster-resident greedy decoder" if eng.decoder == "cluster" else "decoder_steps_kernel (grid-barrier greedy decoder"
            roofline = {"kernel": "%s, %d steps per launch)" % (kname, n_dec), "bound": "hbm",
                        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                        "us_per_launch": kernel_ms * 1e3, "launches_per_step": 1, "algorithmic_bytes_per_launch": nbytes,
                        "algorithmic_bytes_per_greedy_step": parts, "share_of_step": kernel_ms / (total_ms / args.steps),
                        "greedy_step_us": kernel_ms * 1e3 / n_dec, "traced_greedy_step_us": step_us, "phase_us_per_greedy_step": phases,
                        "note": "working set (weights 31 MB + K/V caches 54 MB) is L2-resident; a greedy step at batch 64 is a chain of "
                                "~68 dependent phases, each bound by exchange + instruction latency, not by bandwidth (DESIGN.md 4/7); "
                                "algorithmic bytes keep round 1's definition (every weight once per step incl. the fp32 generator matrix) "
                                "although the screening generator now reads an fp16 copy (4.7 MB) + a few fp32 rows per step"}
        else:
            fam, step_us = decode_family_timeline(eng, ws, B, S)
            nbytes = decode_family_bytes(eng, B, S)
            share = {k: round(v["us_per_step"], 1) for k, v in fam.items()}
            dom = max((k for k in share if k in nbytes), key=lambda k: share[k])
            achieved = nbytes[dom] / (fam[dom]["us_per_launch"] * 1e-6) / 1e9
            roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": None, "peak_source": peak_src,
                        "us_per_launch": fam[dom]["us_per_launch"], "launches_per_step": fam[dom]["launches_per_step"],
                        "algorithmic_bytes_per_launch": nbytes[dom], "greedy_step_us": step_us, "families_us_per_greedy_step": share,
                        "note": "M=64 decode GEMMs are latency-bound (weights L2-resident); see DESIGN.md 4/7"}
        extra = dict(sharded)
        gemm = None
        if not args.no_extra:
            extra.update(rank0_measurements(eng, dev))
            if world == 1:
                gemm = gemm_roofline(dev)
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cores = blas_threads(os.cpu_count() or 1)
            v, dt = cpu_reference_sample(16, S, 32)            # ~10-15 s of CPU work
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "oracle port (numpy ref-float, full-prefix recompute, %d BLAS threads in effect): 16 sentences x %d src tokens, first 32 "
                             "of 71 greedy steps, %.1f s -- a SAMPLE of cfg2 (early steps have the shortest prefixes: over-states the CPU rate)" % (cores, S, dt)}
            if not args.no_extra:
                try:
                    cpu["cfg1_node_walk"] = cfg1_cpu_walk()
                except Exception as exc:
                    cpu["cfg1_node_walk"] = {"error": str(exc)[:200]}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "ms_per_step_min_median_max": [float(np.min(step_ms)), float(np.median(step_ms)), float(np.max(step_ms))],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8",
                "data": "synthetic",
                "config": {"workload": "cfg2: 8-bit Transformer-base greedy decode, batch 64 x src len 64 per GPU, 71 steps, KV cache",
                           "batch_per_gpu": B, "src_len": S, "greedy_steps": MAX_LEN - 1, "parallelism": "sentence-sharded x%d" % world,
                           "l2": "512 MB flush between timed iterations",
                           "result_gather": "one NCCL all_gather of the K decoded batches inside the last step's timed bracket" if world > 1 else "none (1 GPU)"},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(ids_np.nbytes + mask_np.nbytes), "d2h_bytes_per_step": int(B * MAX_LEN * 8),
                        "ms_per_step_min_median_max": [float(np.min(e2e_step_ms)), float(np.median(e2e_step_ms)), float(np.max(e2e_step_ms))]},
                "gpu_launches": int(launches), "roofline": roofline, "roofline_gemm": gemm, "cpu_baseline": cpu, "extra": extra}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    P.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
