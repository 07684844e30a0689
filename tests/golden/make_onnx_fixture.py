"""Writes tests/golden/ref_encoder_tiny.onnx(.gz) + ref_encoder_tiny_io.npz: a REAL ONNX file exported by torch from the
reference's own modules (model.make_model + get_quantized_model.quantize_transformer: 1 encoder layer, d_model 128, d_ff 128,
8 heads, opset 13, TorchScript exporter), and the input / output tensors of the torch model on the same weights.  Recipe:
SURVEY.md Appendix B (the reference's export needs brevitas / qonnx, which are absent: their unused imports are stubbed).

Run here (needs /root/reference):  python tests/golden/make_onnx_fixture.py
"""
import gzip
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"


def stub_modules():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
    mod("brevitas"); mod("brevitas.nn"); mod("brevitas.export", export_onnx_qcdq=None); mod("brevitas.quant")
    mod("brevitas.quant.scaled_int", Int32Bias=None, Uint8ActPerTensorFloat=None, Int8ActPerTensorFloat=None, Int8WeightPerChannelFloat=None)
    mod("qonnx"); mod("qonnx.core"); mod("qonnx.core.modelwrapper", ModelWrapper=None)


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    stub_modules()
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)
    import attention
    attention.print = lambda *a, **k: None
    import get_quantized_model as gq
    from model import make_model
    torch.manual_seed(3)
    model = make_model(37, 41, N=1, d_model=128, d_ff=128, h=8)
    for p in model.parameters():
        if p.dim() > 1:
            torch.nn.init.xavier_uniform_(p)
    with torch.no_grad():           # non-trivial LayerNorm parameters
        for n, p in model.named_parameters():
            if n.endswith("a_2"):
                p.copy_(1.0 + 0.1 * torch.randn_like(p))
            if n.endswith("b_2"):
                p.copy_(0.1 * torch.randn_like(p))
    model = gq.quantize_transformer(model).eval()
    os.chdir(cwd)
    B, S = 2, 5
    x = torch.randn(B, S, 128)
    mask = torch.ones(B, 1, S, dtype=torch.bool)
    mask[1, 0, 3:] = False
    with torch.no_grad():
        y = model.encoder(x, mask)
    import torch.onnx._internal.torchscript_exporter.onnx_proto_utils as opu
    opu._add_onnxscript_fn = lambda b, o: b           # the real one needs the `onnx` package
    path = os.path.join(out_dir, "ref_encoder_tiny.onnx")
    torch.onnx.export(model.encoder, (x, mask), path, opset_version=13, dynamo=False, input_names=["global_in", "global_in_1"],
                      output_names=["global_out"])
    raw = open(path, "rb").read()
    with gzip.open(path + ".gz", "wb", compresslevel=9) as f:
        f.write(raw)
    os.remove(path)
    np.savez_compressed(os.path.join(out_dir, "ref_encoder_tiny_io.npz"), x=x.numpy(), mask=mask.numpy(), y=y.numpy())
    print("wrote", path + ".gz", len(raw), "bytes raw")

    # ---- the decoder of the same model (1 layer): inputs as onnx_optimized_custom_inference.py:646-651 feeds them
    T = 4
    xt = torch.randn(B, T, 128)
    tgt_mask = torch.tril(torch.ones(1, T, T, dtype=torch.int64))          # subsequent_mask(T)
    with torch.no_grad():
        yd = model.decoder(xt, y, mask, tgt_mask)
    path = os.path.join(out_dir, "ref_decoder_tiny.onnx")
    torch.onnx.export(model.decoder, (xt, y, mask, tgt_mask), path, opset_version=13, dynamo=False,
                      input_names=["global_in", "global_in_1", "global_in_2", "global_in_3"], output_names=["global_out"])
    raw = open(path, "rb").read()
    with gzip.open(path + ".gz", "wb", compresslevel=9) as f:
        f.write(raw)
    os.remove(path)
    np.savez_compressed(os.path.join(out_dir, "ref_decoder_tiny_io.npz"), x=xt.numpy(), memory=y.numpy(), src_mask=mask.numpy(),
                        tgt_mask=tgt_mask.numpy(), y=yd.numpy())
    print("wrote", path + ".gz", len(raw), "bytes raw")


if __name__ == "__main__":
    main()
