"""Greedy decoding THROUGH the drop-in executor: the caller of run_module the reference ships in
8-bit_onnx_optimized_custom_inference.py:649-721 (`greedy_decode(model, src, src_mask, max_len, start_symbol, custom_decoder=True)`),
i.e. BASELINE.json configs[0]:

    src_float = model.get_src_embed(src)                                   # host step (embeddings.py:13, positional_encodings.py:24)
    encoder_weight_dict, encoder_graph = prepare_inference(encoder, {...})
    memory, _ = run_module("Encoder", {...}, encoder, encoder_weight_dict, encoder_graph)
    decoder_weight_dict, decoder_graph = prepare_inference(decoder, {...})
    for i in range(max_len - 1):                                           # FULL prefix every step, no early stop at </s>
        out, _ = run_module("Decoder", {"global_in": tgt_embed(ys), "global_in_1": memory, "global_in_2": src_mask,
                                        "global_in_3": subsequent_mask(len(ys))}, decoder, decoder_weight_dict, decoder_graph)
        next_word = argmax(model.generator(out[:, -1]));  ys = cat(ys, next_word)

Here the graph walk and the loop stay in Python, every node is one CUDA handler of libot_b200.so (executor.py), and the host steps
either side of the graphs (embedding + positional encoding, generator + arg-max: SURVEY.md 8f item 4) are the ot_embed_pe /
ot_generator_argmax kernels.  Tensors stay on the device between the steps; only the final token ids come back.  This path exists
for drop-in fidelity and for fault trials that need the node-by-node walk -- the fused engine (engine.py) is the fast path.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from . import executor as X
from . import kernels as K

D_MODEL = 512


def subsequent_mask(size: int) -> np.ndarray:
    """utils.py:10-14 (`torch.triu(ones, diagonal=1) == 0`), as the int64 tensor the reference feeds (`.type_as(src.data)`)."""
    return (np.triu(np.ones((1, size, size)), k=1) == 0).astype(np.int64)


class HostModel:
    """The three host-side pieces of the reference's torch `model` the decode loop calls: get_src_embed, get_tgt_embed
    (encoder_decoder.py:54-58) and generator + torch.max (generator.py:14-15), on the device."""

    def __init__(self, float_weights: Dict[str, np.ndarray], max_len: int = 512, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise K.OtError("HostModel needs a CUDA device: this package has no CPU fallback")
        self.dev = device or torch.device("cuda", torch.cuda.current_device())
        t = lambda name: torch.from_numpy(np.ascontiguousarray(float_weights[name], dtype=np.float32)).to(self.dev)  # noqa: E731
        self.src_lut, self.tgt_lut = t("src_embed.0.lut.weight"), t("tgt_embed.0.lut.weight")
        self.gen_w, self.gen_b = t("generator.proj.weight"), t("generator.proj.bias")
        d = self.src_lut.shape[1]
        pe = torch.zeros(max_len, d)
        position = torch.arange(0.0, max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0.0, d, 2) * -(math.log(10000.0) / d))           # positional_encodings.py:14-21
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.pe = pe.to(self.dev)

    def _embed(self, ids: torch.Tensor, lut: torch.Tensor) -> torch.Tensor:
        B, T = ids.shape
        return K.embed_pe(ids.reshape(-1).contiguous(), lut, self.pe, seq_len=T).reshape(B, T, lut.shape[1])

    def get_src_embed(self, src: torch.Tensor) -> torch.Tensor:
        return self._embed(src, self.src_lut)

    def get_tgt_embed(self, ys: torch.Tensor) -> torch.Tensor:
        return self._embed(ys, self.tgt_lut)

    def next_word(self, h_last: torch.Tensor) -> torch.Tensor:
        """arg-max of log_softmax(W h + b) == arg-max of the logits; first index on ties (torch.max)."""
        ids, _, _, _ = K.generator_argmax(h_last.contiguous(), self.gen_w, self.gen_b)
        return ids


def greedy_decode(model: HostModel, src, src_mask, max_len: int, start_symbol: int, encoder, decoder, executor=X, inject_parameters=None,
                  inject_module: Optional[str] = None, target_inference_number: int = 1, timings: Optional[dict] = None) -> torch.Tensor:
    """8-bit_onnx_optimized_custom_inference.py:649-721 (custom_decoder=True).  `encoder` / `decoder`: graph objects or graph file paths
    (.onnx / .onnx.gz / .otg), passed to executor.prepare_inference / run_module exactly where the reference passes the file path.
    `executor` is executor.py (run_module(..., inject_parameters)) or inject_operations.py (run_module(..., inject_input)).  With
    inject_parameters, the Encoder pass or the decoder pass number `target_inference_number` carries the fault
    (parallelized_inject_onnx_transformer.py:639, 832).  Returns ys int64 [B, max_len] on the device."""
    import time
    src = src if isinstance(src, torch.Tensor) else torch.from_numpy(np.asarray(src))
    src = src.to(model.dev)
    mask = src_mask if isinstance(src_mask, torch.Tensor) else torch.from_numpy(np.asarray(src_mask))
    mask = mask.to(model.dev)
    t0 = time.perf_counter()
    src_float = model.get_src_embed(src)
    enc_in = {"global_in": src_float, "global_in_1": mask}
    encoder_weight_dict, encoder_graph = executor.prepare_inference(encoder, enc_in)
    p_enc = inject_parameters if (inject_parameters is not None and inject_module == "Encoder") else None
    memory, _ = executor.run_module("Encoder", enc_in, encoder, encoder_weight_dict, encoder_graph, p_enc)
    memory = memory[list(memory.keys())[0]]
    B = src.shape[0]
    ys = torch.full((B, 1), start_symbol, dtype=torch.int64, device=model.dev)
    dec_in = {"global_in": model.get_tgt_embed(ys), "global_in_1": memory, "global_in_2": mask,
              "global_in_3": torch.from_numpy(subsequent_mask(1)).to(model.dev)}
    decoder_weight_dict, decoder_graph = executor.prepare_inference(decoder, dec_in)
    if timings is not None:
        torch.cuda.synchronize()
        timings["encoder_s"] = time.perf_counter() - t0
        timings["decoder_step_s"] = []
    for i in range(max_len - 1):
        t1 = time.perf_counter()
        dec_in = {"global_in": model.get_tgt_embed(ys), "global_in_1": memory, "global_in_2": mask,
                  "global_in_3": torch.from_numpy(subsequent_mask(ys.shape[1])).to(model.dev)}
        p_dec = inject_parameters if (inject_parameters is not None and inject_module == "Decoder" and i == target_inference_number - 1) else None
        out, _ = executor.run_module("Decoder", dec_in, decoder, decoder_weight_dict, decoder_graph, p_dec)
        out = out[list(out.keys())[0]]
        next_word = model.next_word(out[:, -1])
        ys = torch.cat([ys, next_word.reshape(B, 1)], dim=1)
        if timings is not None:
            torch.cuda.synchronize()
            timings["decoder_step_s"].append(time.perf_counter() - t1)
    return ys
