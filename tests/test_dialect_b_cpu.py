"""Dialect B (Brevitas QCDQ) on the CPU: the oracle's ONNX-spec handlers against hand-evaluated cases, and the oracle restatement of
inject_operations.py against the closed-form fault deltas of SURVEY.md App. D.  (The CUDA side is tests/test_dialect_b_gpu.py.)"""
import numpy as np
import pytest

from onnx_transformer_b200 import graph as G
from onnx_transformer_b200 import inject_operations as PIO     # host-side helpers only (no kernel is launched in this file)
from onnx_transformer_b200.graph import Attribute, Node
from oracle import executor as oe
from oracle import inject_ops as oio
from oracle import intexact as ox

F32 = np.float32


def _node(op, n_in, **attrs):
    return Node(op + "_t", op, ["in%d" % i for i in range(n_in)], ["out"], [Attribute(k, i=v) for k, v in attrs.items()])


def test_quantize_dequantize_linear_spec():
    x = np.array([[0.5, 1.5, 2.5, -0.5, -1.5, 300.0, -300.0, 0.26]], dtype=F32)
    s = np.array(1.0, dtype=F32)
    zp = np.array(0, dtype=np.int8)
    q = oe.run_node(_node("QuantizeLinear", 3), [x, s, zp])
    assert q.dtype == np.int8 and q.tolist() == [[0, 2, 2, 0, -2, 127, -128, 0]]          # half to even, saturation
    qu = oe.run_node(_node("QuantizeLinear", 2), [x, s])
    assert qu.dtype == np.uint8 and qu.tolist() == [[0, 2, 2, 0, 0, 255, 0, 0]]           # zero point omitted -> uint8
    q3 = oe.run_node(_node("QuantizeLinear", 3), [x, s, np.array(3, dtype=np.int8)])
    assert q3.tolist() == [[3, 5, 5, 3, 1, 127, -128, 3]]
    # per-axis scale along axis 1 of [2, 3]
    y = np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]], dtype=F32)
    sc = np.array([1.0, 0.5, 0.25], dtype=F32)
    qa = oe.run_node(_node("QuantizeLinear", 3, axis=1), [y, sc, zp])
    assert qa.tolist() == [[1, 4, 12], [4, 10, 24]]
    dq = oe.run_node(_node("DequantizeLinear", 3, axis=1), [qa, sc, zp])
    assert np.array_equal(dq, y)
    dz = oe.run_node(_node("DequantizeLinear", 3), [np.array([[5, -3]], dtype=np.int8), np.array(0.5, dtype=F32), np.array(1, dtype=np.int8)])
    assert dz.tolist() == [[2.0, -2.0]]
    clipped = oe.run_node(_node("Clip", 3), [np.array([-128, -9, -8, 7, 8, 127], dtype=np.int8), np.array(-8, np.int8), np.array(7, np.int8)])
    assert clipped.dtype == np.int8 and clipped.tolist() == [-8, -8, -8, 7, 7, 7]


def test_matmul_integer_and_qlinear_matmul_spec():
    rng = np.random.default_rng(0)
    a = rng.integers(-128, 128, size=(5, 16), dtype=np.int8)
    b = rng.integers(-128, 128, size=(16, 7), dtype=np.int8)
    az = np.array(3, dtype=np.int8)
    bz = rng.integers(-5, 6, size=7, dtype=np.int8)
    out = oe.run_node(_node("MatMulInteger", 4), [a, b, az, bz])
    want = (a.astype(np.int64) - 3) @ (b.astype(np.int64) - bz.astype(np.int64)[None, :])
    assert out.dtype == np.int32 and np.array_equal(out, want)
    # the zero-point identity the CUDA epilogue uses: acc - a_zp*colsum(B) - b_zp*rowsum(A) + K*a_zp*b_zp
    acc = a.astype(np.int64) @ b.astype(np.int64)
    ident = acc - 3 * b.astype(np.int64).sum(0)[None, :] - bz.astype(np.int64)[None, :] * a.astype(np.int64).sum(1)[:, None] + 16 * 3 * bz.astype(np.int64)[None, :]
    assert np.array_equal(ident, want)
    # uint8 operands
    au = rng.integers(0, 256, size=(5, 16), dtype=np.uint8)
    outu = oe.run_node(_node("MatMulInteger", 4), [au, b, np.array(128, dtype=np.uint8), None])
    assert np.array_equal(outu, (au.astype(np.int64) - 128) @ b.astype(np.int64))
    ys, yz = F32(0.37), np.array(-4, dtype=np.int8)
    sa, sb = F32(0.02), rng.uniform(0.01, 0.03, size=7).astype(F32)
    q = oe.run_node(_node("QLinearMatMul", 8), [a, sa, az, b, sb, bz, ys, yz])
    y = ((want.astype(F32) * sa).astype(F32) * sb[None, :]).astype(F32)
    assert q.dtype == np.int8 and np.array_equal(q, np.clip(np.rint((y / ys).astype(F32)) - 4, -128, 127).astype(np.int8))


def _block(bits, T=12, d=64, f=128, seed=0):
    rng = np.random.default_rng(seed)
    g = G.build_qcdq_block_graph((rng.normal(size=(d, f)) * 0.1).astype(F32), (rng.normal(size=(f, d)) * 0.1).astype(F32),
                                 (rng.normal(size=(d, d)) * 0.1).astype(F32), T, bits, seed)
    x = rng.normal(size=(1, T, d)).astype(F32)
    return g, x


@pytest.mark.parametrize("bits", [8, 4])
def test_qcdq_block_integer_ranges_and_int_exact_mode(bits):
    g, x = _block(bits)
    wd, _ = oe.prepare_inference(g, {"global_in": x})
    out, wd = oio.run_module("encoder", {"global_in": x}, None, wd, g, None)
    hi = 2 ** (bits - 1)
    for n in g.node:
        if n.op_type == "DequantizeLinear":
            q = wd[n.input[0]]
            assert q.dtype == np.int8 and q.min() >= -hi and q.max() <= hi - 1
            assert len(np.unique(q)) > 3
    wd2, _ = oe.prepare_inference(g, {"global_in": x})
    out2, wd2 = oio.run_module("encoder", {"global_in": x}, None, wd2, g, None, mode="int-exact")
    # the two numeric modes agree up to fp32 summation order; integer tensors may flip by one step at a rounding boundary
    for n in g.node:
        if n.op_type == "MatMul" and n.name in ("MatMul_0",):
            np.testing.assert_allclose(wd2[n.output[0]], wd[n.output[0]], rtol=1e-4, atol=1e-5)
    assert list(out.keys()) == ["global_out"] and out["global_out"].shape == (1, 12, 12)


@pytest.mark.parametrize("bits,fault_model,target,operand", [(8, "INPUT", "MatMul_0", "input"), (4, "WEIGHT", "MatMul_1", "weight"),
                                                             (4, "INPUT", "MatMul_3", "input"), (8, "WEIGHT", "MatMul_3", "weight")])
def test_oracle_operand_fault_equals_closed_form(bits, fault_model, target, operand):
    """out' - out at the target MatMul == the rank-1 update of SURVEY.md App. D: (dequant(q') - dequant(q)) x the other operand."""
    g, x = _block(bits)
    node = g.node_by_name(target)
    dq_in, dq_w = node.input[0], node.input[1]
    transposed = None
    for n in g.node:
        if n.op_type == "Transpose" and n.output[0] in node.input:
            transposed = n
    w_src = transposed.input[0] if (transposed is not None and dq_w == transposed.output[0]) else dq_w
    int_in = [n for n in g.node if n.output[0] == dq_in][0].input[0]
    int_w = [n for n in g.node if n.output[0] == w_src][0].input[0]
    (in_q, in_t), (w_q, w_t), _, taxes = PIO.get_target_inputs(g, target, int_in, int_w, None, node.output[0])
    assert (in_t, w_t) == (int_in, int_w) and (taxes is None) == (transposed is None)
    wd, _ = oe.prepare_inference(g, {"global_in": x})
    _, gold = oio.run_module("encoder", {"global_in": x}, None, dict(wd), g, None)
    tname, qname = (in_t, in_q) if operand == "input" else (w_t, w_q)
    shape = gold[tname].shape
    rng = np.random.default_rng(3)
    idx = [int(rng.integers(0, s)) for s in shape]
    bit = bits - 1 if operand == "weight" else 1
    p = {"inject_type": fault_model, "faulty_tensor_name": tname, "faulty_quantizer_name": qname, "faulty_bit_position": bit,
         "faulty_operation_name": target, "targetted_module": "encoder", "transposed_axes": taxes, "bit_width": bits,
         "rng_draws": {"target_indices": idx}}
    _, bad = oio.run_module("encoder", {"global_in": x}, None, dict(wd), g, p)
    q0 = int(gold[tname][tuple(idx)])
    q1 = (ox.flip_int4_bit if bits == 4 else ox.flip_int8_bit)(q0, bit)
    dq_node = g.node_by_name(qname)
    scale = np.asarray(gold[dq_node.input[1]], dtype=F32)
    s_at = scale.reshape(-1)[idx[1]] if scale.size > 1 else F32(scale)
    dhat = F32(F32(q1 * s_at) - F32(q0 * s_at))
    a_hat, b_hat = gold[node.input[0]], gold[node.input[1]]
    want = gold[target + "_out0" if target != "MatMul_3" else "global_out"].copy()
    if operand == "input":
        _, t, k = idx
        want[0, t, :] = (want[0, t, :] + (dhat * b_hat[..., k, :].reshape(-1)).astype(F32)).astype(F32)
    elif transposed is not None:                      # fault in K [1,T,d] before Transpose(0,2,1): column t of the scores
        _, t, k = idx
        want[0, :, t] = (want[0, :, t] + (a_hat[0, :, k] * dhat).astype(F32)).astype(F32)
    else:                                              # fault in W [K,N] at (k, n): column n
        k, n_ = idx
        want[0, :, n_] = (want[0, :, n_] + (a_hat[0, :, k] * dhat).astype(F32)).astype(F32)
    got = bad[node.output[0]]
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-6)
    assert np.count_nonzero(got != gold[node.output[0]]) > 0 or dhat == 0
