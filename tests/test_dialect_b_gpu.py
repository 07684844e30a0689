"""Dialect B (Brevitas QCDQ) through the product's inject_operations.py drop-in on the GPU, against the oracle restatement
(oracle/inject_ops.py) on a small QCDQ graph (QuantizeLinear -> Clip -> DequantizeLinear -> MatMul, 8- and 4-bit): every
intermediate tensor of the golden walk, then INPUT / WEIGHT (incl. the `transposed_axes` case, flip_int4_bit at 4 bits), RANDOM and
RANDOM_BITFLIP trials with explicit draws; plus the ONNX-spec MatMulInteger / QLinearMatMul handlers with non-zero zero points and
the north-star accumulator / int8-output bit flips of the GEMM epilogue.  PARITY UNPINNED at the reference boundary (SURVEY.md
0.4, 8c): the oracle is the ONNX operator specification + the reference's call sites."""
import numpy as np
import pytest
import torch

from onnx_transformer_b200 import graph as G
from onnx_transformer_b200.graph import Attribute, Node
from oracle import executor as oe
from oracle import inject_ops as oio
from oracle import intexact as ox

pytestmark = pytest.mark.gpu
F32 = np.float32


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _block(bits, T=16, d=64, f=128, seed=0):
    rng = np.random.default_rng(seed)
    g = G.build_qcdq_block_graph((rng.normal(size=(d, f)) * 0.1).astype(F32), (rng.normal(size=(f, d)) * 0.1).astype(F32),
                                 (rng.normal(size=(d, d)) * 0.1).astype(F32), T, bits, seed)
    x = rng.normal(size=(1, T, d)).astype(F32)
    return g, x


def _compare_walks(g, wd_gpu, wd_ref, fp32_matmul_outputs):
    """Integer tensors bit-exact; elementwise float tensors bit-exact; fp32 MatMul results within summation-order tolerance (and
    whatever lies downstream of one: integers +-1 at a rounding boundary)."""
    downstream = set()
    for n in g.node:
        name = n.output[0]
        a = wd_gpu[name].detach().cpu().numpy()
        r = wd_ref[name]
        assert a.shape == r.shape and a.dtype == r.dtype, name
        tainted = name in fp32_matmul_outputs or any(i in downstream for i in n.input)
        if tainted:
            downstream.add(name)
        if not tainted:
            assert np.array_equal(a, r), name
        elif a.dtype == np.int8:
            diff = np.abs(a.astype(np.int32) - r.astype(np.int32))
            assert diff.max() <= 1 and (diff != 0).mean() < 5e-3, name
        else:
            np.testing.assert_allclose(a, r, rtol=2e-3, atol=2e-3, err_msg=name)


@pytest.mark.parametrize("bits", [8, 4])
def test_qcdq_walk_matches_oracle(bits):
    from onnx_transformer_b200 import inject_operations as PIO
    g, x = _block(bits)
    wd, graph = PIO.prepare_inference(g, {"global_in": x})
    out, wd = PIO.run_module("encoder", {"global_in": x}, None, wd, graph, None)
    wr, _ = oe.prepare_inference(g, {"global_in": x})
    out_r, wr = oio.run_module("encoder", {"global_in": x}, None, wr, g, None, mode="int-exact")
    assert list(out.keys()) == ["global_out"]
    # MatMul_0..2 run as int8 tensor-core GEMMs (int-exact: bit-exact vs the oracle's int-exact mode); MatMul_3 (3-D x 3-D) as fp32
    _compare_walks(g, wd, wr, {"global_out"})
    # 4-tuple contract of execute_node
    wd2, graph = PIO.prepare_inference(g, {"global_in": x})
    ret = PIO.execute_node(graph.node[0], graph, graph.node[0].output[0], wd2, "encoder", None)
    assert len(ret) == 4 and ret[3] is None and list(ret[0].keys()) == [graph.node[0].output[0]]


CASES = [(8, "INPUT", "MatMul_0", "input", 6), (8, "WEIGHT", "MatMul_0", "weight", 7), (4, "INPUT", "MatMul_1", "input", 3),
         (4, "WEIGHT", "MatMul_1", "weight", 2), (4, "WEIGHT", "MatMul_3", "weight", 3), (8, "INPUT", "MatMul_3", "input", 0),
         (8, "RANDOM", "MatMul_1", "output", None), (4, "RANDOM_BITFLIP", "MatMul_0", "output", None)]


@pytest.mark.parametrize("bits,fault_model,target,operand,bit", CASES)
def test_qcdq_fault_trials_match_oracle(bits, fault_model, target, operand, bit):
    from onnx_transformer_b200 import inject_operations as PIO
    g, x = _block(bits, seed=1)
    node = g.node_by_name(target)
    transposed = [n for n in g.node if n.op_type == "Transpose" and n.output[0] in node.input]
    w_src = transposed[0].input[0] if (transposed and node.input[1] == transposed[0].output[0]) else node.input[1]
    int_in = [n for n in g.node if n.output[0] == node.input[0]][0].input[0]
    int_w = [n for n in g.node if n.output[0] == w_src][0].input[0]
    (in_q, in_t), (w_q, w_t), _, taxes = PIO.get_target_inputs(g, target, int_in, int_w, None, node.output[0])
    rng = np.random.default_rng(17)

    def params():
        if operand == "output":
            shape = (1, 16, {"MatMul_0": 128, "MatMul_1": 64}[target])
            draws = {"target_indices": [int(rng.integers(0, s)) for s in shape], "flip_bit": 30, "random_bits": 0x41234567}
            return {"inject_type": fault_model, "faulty_tensor_name": node.output[0], "faulty_quantizer_name": None, "faulty_bit_position": None,
                    "faulty_operation_name": target, "targetted_module": "encoder", "transposed_axes": taxes, "bit_width": bits, "rng_draws": draws}
        tname, qname = (in_t, in_q) if operand == "input" else (w_t, w_q)
        return {"inject_type": fault_model, "faulty_tensor_name": tname, "faulty_quantizer_name": qname, "faulty_bit_position": bit,
                "faulty_operation_name": target, "targetted_module": "encoder", "transposed_axes": taxes, "bit_width": bits,
                "rng_draws": {"target_indices": idx}}

    idx = None
    if operand != "output":
        wr0, _ = oe.prepare_inference(g, {"global_in": x})
        _, gold = oio.run_module("encoder", {"global_in": x}, None, wr0, g, None)
        tname = in_t if operand == "input" else w_t
        idx = [int(rng.integers(0, s)) for s in gold[tname].shape]
    p_gpu = params()
    p_ref = dict(p_gpu, rng_draws=dict(p_gpu["rng_draws"]))
    wd, graph = PIO.prepare_inference(g, {"global_in": x})
    out, wd = PIO.run_module("encoder", {"global_in": x}, None, wd, graph, p_gpu)
    wr, _ = oe.prepare_inference(g, {"global_in": x})
    out_r, wr = oio.run_module("encoder", {"global_in": x}, None, wr, g, p_ref, mode="int-exact")
    # the injected tensor itself: the golden part is bit-exact (int-exact GEMM), the delta is an fp32 product of one-hot operands --
    # a single non-zero term per element, hence exact too
    got, want = wd[node.output[0]].detach().cpu().numpy(), wr[node.output[0]]
    if target != "MatMul_3":
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), "faulty MatMul output"
    else:
        np.testing.assert_allclose(got, want, rtol=2e-3, atol=2e-3)
    if operand != "output":
        assert p_gpu["dequantized_operation_input_name"] == p_ref["dequantized_operation_input_name"]
        d_gpu, d_ref = wd["delta_4d"].detach().cpu().numpy(), wr["delta_4d"]
        assert d_gpu.shape == d_ref.shape and np.array_equal(d_gpu.view(np.uint32), d_ref.view(np.uint32))
        assert np.count_nonzero(d_ref) <= 1
    _compare_walks(g, wd, wr, {"global_out"})


# ------------------------------------------------------------------------------------------------ ONNX-spec integer MatMuls
def _n(op, n_in):
    return Node(op + "_t", op, ["in%d" % i for i in range(n_in)], ["out"], [])


@pytest.mark.parametrize("M,N,K_,azp,bzp", [(64, 512, 512, "scalar", "vector"), (200, 96, 144, "vector", "scalar"), (33, 64, 64, "none", "vector"),
                                            (128, 256, 2048, "scalar", "none")])
def test_matmul_integer_with_zero_points(M, N, K_, azp, bzp):
    from onnx_transformer_b200 import kernels as K
    rng = np.random.default_rng(M + N)
    a = rng.integers(-128, 128, size=(M, K_), dtype=np.int8)
    b = rng.integers(-128, 128, size=(K_, N), dtype=np.int8)
    az = {"none": None, "scalar": np.array(-7, dtype=np.int8), "vector": rng.integers(-20, 21, size=M, dtype=np.int8)}[azp]
    bz = {"none": None, "scalar": np.array(5, dtype=np.int8), "vector": rng.integers(-20, 21, size=N, dtype=np.int8)}[bzp]
    ref = oe.run_node(_n("MatMulInteger", 4), [a, b, az, bz])
    wt = dev(np.ascontiguousarray(b.T))
    out = K.matmul_integer(dev(a), wt, a_zp=None if az is None else dev(az), b_zp=None if bz is None else dev(bz)).cpu().numpy()
    assert out.dtype == np.int32 and np.array_equal(out, ref)


def test_matmul_integer_and_qlinear_handlers_through_the_executor():
    from onnx_transformer_b200 import executor as X
    rng = np.random.default_rng(4)
    a_u8 = rng.integers(0, 256, size=(2, 24, 64), dtype=np.uint8)
    b = rng.integers(-128, 128, size=(64, 96), dtype=np.int8)
    az = np.array(131, dtype=np.uint8)
    bz = rng.integers(-9, 10, size=96, dtype=np.int8)
    node = Node("MatMulInteger_0", "MatMulInteger", ["a", "b", "az", "bz"], ["y"], [])
    wd = {k: dev(v) for k, v in dict(a=a_u8, b=b, az=az, bz=bz).items()}
    out = X.run_node(node, [wd[n] for n in node.input], wd).cpu().numpy()
    ref = oe.run_node(node, [a_u8, b, az, bz]).reshape(2, 24, 96)
    assert np.array_equal(out, ref)
    a = rng.integers(-128, 128, size=(40, 64), dtype=np.int8)
    sa, sb, ys = np.array(0.02, dtype=F32), rng.uniform(0.01, 0.03, size=96).astype(F32), np.array(0.41, dtype=F32)
    a_zp, yz = np.array(2, dtype=np.int8), np.array(-3, dtype=np.int8)
    qn = Node("QLinearMatMul_0", "QLinearMatMul", ["a", "sa", "az", "b", "sb", "bz", "ys", "yz"], ["y"], [])
    vals = dict(a=a, sa=sa, az=a_zp, b=b, sb=sb, bz=bz, ys=ys, yz=yz)
    wd = {k: dev(v) for k, v in vals.items()}
    out = X.run_node(qn, [wd[n] for n in qn.input], wd).cpu().numpy()
    ref = oe.run_node(qn, [vals[n] for n in qn.input])
    assert out.dtype == np.int8 and np.array_equal(out, ref)
    assert (ref == 127).any() or (ref == -128).any() or np.abs(ref).max() > 40      # the case exercises a wide output range


def test_accumulator_and_int8_output_bit_flips():
    """north star: "the bit-flip is applied to the int32 accumulator or int8 output inside the epilogue"."""
    from onnx_transformer_b200 import kernels as K
    rng = np.random.default_rng(12)
    M, N, K_ = 96, 512, 512
    a = rng.integers(-127, 128, size=(M, K_), dtype=np.int8)
    w = rng.integers(-127, 128, size=(N, K_), dtype=np.int8)
    sx = rng.uniform(1e-3, 5e-2, size=M).astype(F32)
    sw = rng.uniform(1e-4, 1e-2, size=N).astype(F32)
    b = rng.normal(size=N).astype(F32)
    acc = ox.int_matmul(a, w)
    for (r, c, bit) in [(5, 77, 30), (95, 511, 0), (0, 0, 31), (40, 300, 17)]:
        f = K.make_fault(K.FAULT_ACC_BITFLIP, flat_index=r * N + c, bit=bit)
        out = K.linear_w8a8(dev(a), dev(w), out_kind=K.OUT_I32, fault=f).cpu().numpy()
        ref = acc.copy()
        ref[r, c] = np.int32(np.uint32(ref[r, c]) ^ np.uint32(1 << bit))
        assert np.array_equal(out, ref), (r, c, bit)
        # through the fp32 epilogue and the fused requant: the flipped accumulator feeds scale / bias / RowQuant like any other
        q, s = K.linear_w8a8(dev(a), dev(w), row_scale=dev(sx), col_scale=dev(sw), bias=dev(b), out_kind=K.OUT_Q8, quant_group=512, fault=f)
        qr, sr = ox.group_quant(ox.linear_epilogue(ref, sx, sw, b), 512)
        assert np.array_equal(q.cpu().numpy(), qr) and np.array_equal(s.cpu().numpy().view(np.uint32), sr.view(np.uint32))
    qg, sg = ox.group_quant(ox.linear_epilogue(acc, sx, sw, b), 512)
    for (r, c, bit) in [(7, 130, 7), (95, 0, 0), (33, 511, 6), (1, 257, 3)]:
        f = K.make_fault(K.FAULT_OUT_Q8_BITFLIP, flat_index=r * N + c, bit=bit)
        q, s = K.linear_w8a8(dev(a), dev(w), row_scale=dev(sx), col_scale=dev(sw), bias=dev(b), out_kind=K.OUT_Q8, quant_group=512, fault=f)
        ref = qg.copy()
        ref[r, c] = ox.flip_int8_bit(int(ref[r, c]), bit)
        assert np.array_equal(q.cpu().numpy(), ref) and np.array_equal(s.cpu().numpy().view(np.uint32), sg.view(np.uint32)), (r, c, bit)


def test_int4_weight_faults_use_flip_int4_bit():
    """cfg4 fault hooks (inject_utils/layers.py:48-59): WEIGHT faults on packed int4 weights wrap to [-8, 7]; single and batched."""
    from onnx_transformer_b200 import kernels as K
    rng = np.random.default_rng(6)
    M, N, K_ = 64, 512, 512
    a = rng.integers(-127, 128, size=(M, K_), dtype=np.int8)
    w = rng.integers(-8, 8, size=(N, K_), dtype=np.int8)
    packed = ((w[:, 0::2].astype(np.uint8) & 0xF) | ((w[:, 1::2].astype(np.uint8) & 0xF) << 4)).astype(np.uint8)
    golden = ox.int_matmul(a, w)
    for n, k, bit in [(300, 5, 3), (0, 511, 0), (511, 256, 2)]:
        f = K.make_fault(K.FAULT_WEIGHT, flat_index=n * K_ + k, bit=bit)
        out = K.linear_w8a8(dev(a), dev(packed), out_kind=K.OUT_I32, w4=True, fault=f).cpu().numpy()
        ref = golden.copy()
        ref[:, n] += a[:, k].astype(np.int32) * (ox.flip_int4_bit(int(w[n, k]), bit) - int(w[n, k]))
        assert np.array_equal(out, ref), (n, k, bit)
    # batched trials: 4 units of 16 rows, unit 1 carries an INPUT fault (int8 activations: flip_int8_bit), unit 3 a WEIGHT fault
    f_in = K.make_fault(K.FAULT_INPUT, flat_index=3 * K_ + 9, bit=6)
    f_w = K.make_fault(K.FAULT_WEIGHT, flat_index=77 * K_ + 100, bit=3)
    faults_dev = K.pack_faults([f_in, f_w], "cuda")
    unit = torch.tensor([-1, 0, -1, 1], dtype=torch.int32, device="cuda")
    out = K.linear_w8a8(dev(a), dev(packed), out_kind=K.OUT_I32, w4=True, mf=(faults_dev, unit, 16)).cpu().numpy()
    ref = golden.copy()
    r = 16 + 3
    ref[r, :] += (ox.flip_int8_bit(int(a[r, 9]), 6) - int(a[r, 9])) * w[:, 9].astype(np.int32)
    ref[48:64, 77] += a[48:64, 100].astype(np.int32) * (ox.flip_int4_bit(int(w[77, 100]), 3) - int(w[77, 100]))
    assert np.array_equal(out, ref)
    with pytest.raises(K.OtError):
        K.linear_w8a8(dev(a), dev(packed), out_kind=K.OUT_I32, w4=True, fault=K.make_fault(K.FAULT_WEIGHT, flat_index=0, bit=5))
