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


# ---- the row-sharded scaled objective, one process per rank (parallel.scaled_dtc_row_sharded) ---------------------------
class _SliceEngine:
    """CPU stand-in for Context.scaled_slice_*: the same protocol on a toy problem with the same structure — M columns
    whitened by an affine recurrence along the rows (x' = a x + K_k, beta_k = K_k - h x), alpha from the full y on every
    rank, statistics (beta'beta, beta'alpha), value from the summed statistics."""

    def __init__(self, K, y, lo, hi, a=0.9, h=0.3):
        self.K, self.y, self.lo, self.hi, self.a, self.h = K, y, lo, hi, a, h
        self.M = K.shape[1]

    def _walk(self, x, want_beta):
        beta = np.zeros((self.hi - self.lo, self.M))
        for i, k in enumerate(range(self.lo, self.hi)):
            if want_beta:
                beta[i] = self.K[k] - self.h * x
            x = self.a * x + self.K[k]
        return x, beta

    def scaled_slice_begin(self, k_time, k_out, theta, row_lo, grad=False):
        assert row_lo == self.lo
        self.grad = grad
        return 1 + self.M, self.M * self.M + self.M

    def scaled_slice_summary(self, out):
        r, _ = self._walk(np.zeros(self.M), False)
        out.numpy()[:] = np.concatenate([[self.a ** (self.hi - self.lo)], r])

    def _entering(self, gathered, member, stride, off=0):
        x = np.zeros(self.M)
        g = gathered.numpy()
        for hh in range(member):
            blk = g[hh * stride + off: hh * stride + off + 1 + self.M]
            x = blk[0] * x + blk[1:]
        return x

    def scaled_slice_stats(self, gathered, member, out):
        x = self._entering(gathered, member, 1 + self.M)
        _, beta = self._walk(x, True)
        xa = 0.0; alpha = np.zeros(len(self.y))                 # the "1 x N filter" every rank runs on the full y
        for k in range(len(self.y)):
            alpha[k] = self.y[k] - self.h * xa; xa = self.a * xa + self.y[k]
        self.beta = beta
        out.numpy()[:] = np.concatenate([(beta.T @ beta).ravel(), beta.T @ alpha[self.lo:self.hi]])

    @staticmethod
    def value_of(stats, M):
        G = stats[:M * M].reshape(M, M); g = stats[M * M:]
        L = np.linalg.cholesky(np.eye(M) + G)
        c = np.linalg.solve(L, g)
        return float(2 * np.log(np.diag(L)).sum() - c @ c)

    def scaled_slice_value(self, stats):
        return self.value_of(stats.numpy(), self.M)

    def scaled_slice_tangent_summary(self, stats, out):
        self.val = self.value_of(stats.numpy(), self.M)
        r, _ = self._walk(np.zeros(self.M), False)
        out.numpy()[:] = np.concatenate([np.concatenate([[self.a ** (self.hi - self.lo)], (q + 1) * r]) for q in range(3)])

    def scaled_slice_grad_partial(self, gathered2, member):
        s5 = np.zeros(5)
        for q in range(3):
            s5[q] = self._entering(gathered2, member, 3 * (1 + self.M), q * (1 + self.M)).sum() + (self.hi - self.lo)
        s5[3] = self.y[self.lo:self.hi].sum(); s5[4] = (self.y[self.lo:self.hi] ** 2).sum()
        return s5

    def scaled_slice_grad_finish(self, s5):
        return self.val, np.asarray(s5, dtype=np.float64).copy()


def _toy(n=203, m=5):
    rng = np.random.default_rng(8)
    return rng.normal(size=(n, m)), rng.normal(size=n)


def _slice_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from gpar_at_scale_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    K, y = _toy()
    b = parallel.row_slice_bounds(len(y), world, align=4)
    eng = _SliceEngine(K, y, b[rank], b[rank + 1])
    v = parallel.scaled_dtc_row_sharded(eng, 3, 3, np.zeros(5), b[rank])
    vg, g5 = parallel.scaled_dtc_row_sharded(eng, 3, 3, np.zeros(5), b[rank], grad=True)
    q.put((rank, v, vg, g5, b))
    dist.destroy_process_group()


def test_row_sharded_objective_protocol_two_and_three_ranks():
    """world_size 2 and 3 (gloo): summaries gathered rank-major, entering states composed in rank order, statistics summed,
    every rank returns the value of the whole sequence; the gradient steps gather the tangent summaries and sum the partials."""
    import socket
    K, y = _toy()
    whole = _SliceEngine(K, y, 0, len(y))
    _, beta = whole._walk(np.zeros(K.shape[1]), True)
    xa = 0.0; alpha = np.zeros(len(y))
    for k in range(len(y)):
        alpha[k] = y[k] - whole.h * xa; xa = whole.a * xa + y[k]
    v_ref = _SliceEngine.value_of(np.concatenate([(beta.T @ beta).ravel(), beta.T @ alpha]), K.shape[1])
    for world in (2, 3):
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=_slice_worker, args=(r, world, port, q)) for r in range(world)]
        for p in procs:
            p.start()
        outs = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        b = outs[0][4]
        assert b[0] == 0 and b[-1] == len(y) and all(x % 4 == 0 for x in b[:-1])
        # expected partial sums of the stand-in gradient steps, composed serially
        exp = np.zeros(5)
        engines = [_SliceEngine(K, y, b[r], b[r + 1]) for r in range(world)]
        for q_ in range(3):
            x = np.zeros(K.shape[1])
            for r, e in enumerate(engines):
                exp[q_] += x.sum() + (b[r + 1] - b[r])
                rr, _ = e._walk(np.zeros(K.shape[1]), False)
                x = e.a ** (b[r + 1] - b[r]) * x + (q_ + 1) * rr
        exp[3] = y.sum(); exp[4] = (y ** 2).sum()
        for rank, v, vg, g5, _ in outs:
            assert abs(v - v_ref) <= 1e-10 * abs(v_ref) and abs(vg - v_ref) <= 1e-10 * abs(v_ref), (world, rank, v, v_ref)
            assert np.allclose(g5, exp, rtol=1e-12, atol=1e-9), (world, rank, g5, exp)
