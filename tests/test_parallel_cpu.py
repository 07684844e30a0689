"""world_size-2 gloo test of the N>1 host logic: rank sharding of sentences / trials and the single result gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from onnx_transformer_b200 import campaign as C
from onnx_transformer_b200 import parallel as P


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    trials = C.make_trials(37, 0, 8, 16)
    mine = P.shard(trials, rank, world)
    recs = P.gather_records([(t.trial_id, "masked" if t.trial_id % 3 else "changed") for t in mine], world)
    lo, hi = P.shard_rows(10, rank, world)
    ys = torch.arange(lo, hi).reshape(-1, 1).repeat(1, 4)
    if hi - lo < 5:
        ys = torch.cat([ys, torch.full((5 - (hi - lo), 4), -1)])
    allys = P.gather_token_ids(ys, world)
    q.put((rank, [r[0] for r in recs], len(mine), allys[:, 0].tolist()))
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1] == list(range(37)) and got[1][1] == list(range(37))      # every trial exactly once, same view on all ranks
    assert got[0][2] + got[1][2] == 37 and abs(got[0][2] - got[1][2]) <= 1
    assert got[0][3] == list(range(10)) and got[1][3] == list(range(10))       # sentence order preserved by the gather


def test_trial_list_is_deterministic_and_well_formed():
    a, b = C.make_trials(200, 0, 64, 64), C.make_trials(200, 0, 64, 64)
    assert a == b and a != C.make_trials(200, 1, 64, 64)
    for t in a:
        assert t.inject_type in C.FAULT_MODELS and 0 <= t.sentence < 64 and 0 <= t.layer < 6
        assert t.op_name.startswith("MatMul_")
        if t.inject_type in ("INPUT16", "WEIGHT16"):
            assert t.target in ("qk", "pv", "cqk", "cpv") and t.window_start % 16 == 0 and 1 <= t.window_len <= 16
    assert {t.module for t in a} == {"Encoder", "Decoder"}


def test_directory_campaign_trial_list(tmp_path):
    """--directory_name: the 24 + 36 target files x 6 fault models x 8 bits x 5 experiments of parallelized_inject_onnx_transformer.py:
    794-861 (the *16 models only on the attention MatMuls), from JSON files on disk."""
    import json
    targets = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fault_targets.json")))
    for module, n_files in (("encoder", 24), ("decoder", 36)):
        d = tmp_path / module
        d.mkdir()
        for k, t in enumerate(targets[module]):
            (d / ("matmul_%d.json" % k)).write_text(json.dumps(t))
        loaded = C.load_targets(str(d))
        assert len(loaded) == n_files
        trials = C.trials_from_targets(loaded, 0, 64, 64)
        n_attn = sum(1 for t in loaded if C.target_site(t)[2] in ("qk", "pv", "cqk", "cpv"))
        assert len(trials) == (n_attn * 6 + (n_files - n_attn) * 4) * 8 * 5
        assert [t.trial_id for t in trials] == list(range(len(trials)))
        for t in trials:
            assert t.module == module.capitalize() and t.op_name.startswith("MatMul_")
            assert (t.bit < 8) or t.inject_type == "RANDOM_BITFLIP"
        # the site decoded from the node name is the site the builder gives that name to
        from onnx_transformer_b200 import graph as G
        for t in trials[::97]:
            name = (G.encoder_matmul_name if t.module == "Encoder" else G.decoder_matmul_name)(t.layer, t.target)
            assert name == t.op_name
    assert C.target_site({"module": "Decoder/FirstMatMul", "target_layer": "MatMul_3"}) == ("Decoder", 1, "cv")


def _overlap_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    P.init_process_group("gloo")
    vals = P.max_over_ranks([1.0 + rank, 5.0 - rank], device="cpu")
    sums = P.sum_over_ranks([1.0 + rank], device="cpu")
    q.put((rank, vals, sums))
    P.destroy_process_group()


def test_two_rank_timing_reductions():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_overlap_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1] == [2.0, 5.0] and got[1][1] == [2.0, 5.0] and got[0][2] == [3.0]
