"""CPU suite (gloo, world_size 2): the host-side multi-GPU logic of qsi_b200.dist -- tensor sharding of the compression
pass is a partition (no exchange needed) and timing aggregates as the max over ranks."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import qsi_b200  # noqa: F401
    from qsi_b200 import dist as qd
    r, lr, w = qd.init("gloo")
    assert (r, w) == (rank, world)
    tensors = qd.model_tensors("llama-65b")
    mine = qd.shard_by_layer(tensors, r, w)
    elems = sum(s[0] * s[1] for _, _, s in mine)
    qd.barrier()
    t_max = qd.max_over_ranks(10.0 + 5.0 * rank)          # rank 1 is slower
    total = qd.sum_over_ranks(elems)
    gathered = [None] * w
    dist.all_gather_object(gathered, [(l, i) for l, i, _ in mine])
    q.put((rank, len(mine), elems, t_max, total, gathered))
    dist.destroy_process_group()


def test_compression_pass_sharding_and_max_over_ranks():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    from qsi_b200 import dist as qd
    all_t = qd.model_tensors("llama-65b")
    assert len(all_t) == 560 and sum(s[0] * s[1] for _, _, s in all_t) == 64_760_053_760     # 64.76 G elements
    (r0, n0, e0, tmax0, tot0, g0), (r1, n1, e1, tmax1, tot1, g1) = res
    assert n0 + n1 == 560 and n0 == n1 == 280
    assert tmax0 == tmax1 == 15.0                                   # max over ranks, seen by every rank
    assert tot0 == tot1 == 64_760_053_760
    a, b = set(g0[0]), set(g0[1])
    assert not (a & b) and len(a | b) == 560                        # a partition: disjoint and covering


def test_single_process_defaults():
    from qsi_b200 import dist as qd
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        os.environ.pop(k, None)
    assert qd.env_world() == (0, 0, 1)
    assert qd.max_over_ranks(3.5) == 3.5 and qd.sum_over_ranks(2) == 2.0
    assert qd.shard_by_layer(qd.model_tensors("llama-7b"), 0, 1) == qd.model_tensors("llama-7b")
    assert len(qd.model_tensors("opt-66b")) == 384


def _cp_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import qsi_b200  # noqa: F401
    from qsi_b200 import dist as qd
    qd.init("gloo")
    torch.manual_seed(0)                                   # same full weight / input on every rank
    w, b, x = torch.randn(10, 16), torch.randn(10), torch.randn(3, 4, 16)
    # num_format='fp32' keeps the quantiser out (no GPU here): this checks sharding + gather against the full linear
    cp = qd.ColumnParallelBFPLinear(16, 10, bias=True, num_format="fp32").load_full(w, b)
    y = cp(x)
    q.put((rank, (cp.lo, cp.hi), torch.allclose(y, torch.nn.functional.linear(x, w, b), atol=1e-6), tuple(y.shape)))
    dist.destroy_process_group()


def test_column_parallel_linear_gathers_the_full_output():
    world, port = 3, _free_port()                          # 10 rows over 3 ranks: uneven shards 4/3/3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_cp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [(0, 4), (4, 7), (7, 10)]
    assert all(r[2] and r[3] == (3, 4, 10) for r in res)
