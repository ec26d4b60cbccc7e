"""CPU, world_size 2, gloo: the N > 1 host logic (task partition, gather of scalars, broadcast of
means down the chain).  The per-task evaluator is a stand-in quadratic: no GPU here."""
import os
import sys
import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import gpar_at_scale_b200 as gp
    from gpar_at_scale_b200 import parallel, neldermead
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tasks = [(o, r) for o in range(4) for r in range(3)]            # 4 outputs x 3 restarts
    costs = [0.1 if o == 0 else 1.0 + 0.1 * o for o, r in tasks]

    def run_task(task):
        o, r = task
        target = np.array([o, -r, 0.5])
        res = neldermead.optimize(lambda x: float(np.sum((x - target) ** 2) + o), np.array([0.1 * r, 0.2, 0.3]))
        return res.minimum, res.minimizer

    vals, thetas = parallel.fit_tasks(tasks, costs, run_task, 3)
    means = parallel.broadcast_means(np.arange(5.0) * (rank + 1), src=0)
    mine = parallel.shard_tasks(costs, world, rank)
    q.put((rank, vals, thetas, means, mine))
    dist.destroy_process_group()


def test_two_rank_partition_gather_broadcast():
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, v0, t0, m0, mine0), (r1, v1, t1, m1, mine1) = outs
    assert sorted(mine0 + mine1) == list(range(12)) and not set(mine0) & set(mine1)     # a partition
    assert abs(len(mine0) - len(mine1)) <= 2
    assert np.allclose(v0, v1) and np.allclose(t0, t1)                                   # every rank has every result
    tasks = [(o, r) for o in range(4) for r in range(3)]
    for (o, r), v, th in zip(tasks, v0, t0):
        assert abs(v - o) < 1e-6 and np.allclose(th, [o, -r, 0.5], atol=1e-3)
    assert np.allclose(m0, np.arange(5.0)) and np.allclose(m1, np.arange(5.0))          # rank 0's means everywhere
    sys.path.insert(0, ROOT)
    from gpar_at_scale_b200 import parallel
    best = parallel.best_per_output(tasks, v0, t0)
    assert set(best) == {0, 1, 2, 3}


def test_lpt_sharding_balances_cheap_first_output():
    sys.path.insert(0, ROOT)
    from gpar_at_scale_b200 import parallel
    tasks = [(o, r) for o in range(8) for r in range(8)]
    costs = [0.05 if o == 0 else 1.0 + 0.05 * o for o, r in tasks]
    for world in (1, 2, 4, 8):
        assign = parallel.shard_tasks(costs, world)
        loads = [sum(costs[t] for t in a) for a in assign]
        assert sorted(sum(assign, [])) == list(range(64))
        assert max(loads) <= 1.1 * sum(costs) / world + 1.5
