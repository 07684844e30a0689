"""B200-native (sm_100a) hot path of gebegebegebe/onnx-transformer: the node-by-node custom ONNX executor
(`run_module` / `execute_node`, with its bit-flip fault hooks) for the quantized Transformer-base model, served by
hand-written CUDA behind a C ABI (include/ot_b200.h), plus the fused KV-cached greedy-decode engine built on the
same kernels.  See DESIGN.md."""
__version__ = "0.1.0"
