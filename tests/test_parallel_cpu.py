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
