"""CPU, world_size 2 over gloo: the N>1 host logic -- env sharding by rank, Philox streams keyed
by GLOBAL env id, mean-over-GLOBAL-envs gradient scaling and the one all-reduce(sum) per cycle
-- reproduces the single-process result on the union batch.  The per-rank compute is the
float64 oracle (CUDA is not available here); what is under test is the sharding arithmetic
that src/agent.py applies around the kernels."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import a3c, philox

A, T, B_GLOBAL = 4, 2, 6


def _inputs():
    rng = np.random.default_rng(42)
    params = a3c.init_params(A, 7)
    screens = rng.integers(0, 256, (T + 4, B_GLOBAL, 84, 84), dtype=np.uint8)
    rew = rng.choice([-1.0, 0.0, 1.0], (T, B_GLOBAL))
    term = rng.random((T, B_GLOBAL)) < 0.2
    probs = rng.dirichlet(np.ones(A), (T, B_GLOBAL)).astype(np.float32)
    return params, screens, rew, term, probs


def _shard_grads(params, screens, rew, term, acts, lo, hi):
    """One rank's share of the cycle: its envs only, scaled by 1/GLOBAL envs (agent.py here)."""
    stacks = a3c.stacks_from_screens(screens[:, lo:hi], T)
    with torch.no_grad():
        _, vb = a3c.forward(a3c.to_torch(params), stacks[T])
    R = a3c.nstep_returns(a3c.clip_rewards(rew[:, lo:hi]), term[:, lo:hi], vb.numpy(), 0.99)
    n = (hi - lo) * T
    grads, _ = a3c.gradients(params, stacks[:T].reshape((n,) + stacks.shape[2:]),
                             acts[:, lo:hi].reshape(-1), R.reshape(-1), 0.01, B_GLOBAL)
    return a3c.flatten_params(grads)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    params, screens, rew, term, probs = _inputs()
    per = B_GLOBAL // world
    lo, hi = rank * per, (rank + 1) * per
    # action sampling keyed by global env id = env_id_base + local index (src/agent.py)
    acts_local = np.stack([philox.sample_actions(probs[t, lo:hi], np.arange(lo, hi), step=t,
                                                 seed=123) for t in range(T)])
    gathered = [torch.zeros(T, per, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(gathered, torch.as_tensor(acts_local))
    acts = torch.cat(gathered, dim=1).numpy()
    g = torch.as_tensor(_shard_grads(params, screens, rew, term, acts, lo, hi))
    dist.all_reduce(g, op=dist.ReduceOp.SUM)                   # the one exchange per cycle
    if rank == 0:
        out.put((acts, g.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_cycle_equals_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    acts2, g2 = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    params, screens, rew, term, probs = _inputs()
    acts1 = np.stack([philox.sample_actions(probs[t], np.arange(B_GLOBAL), step=t, seed=123)
                      for t in range(T)])
    assert np.array_equal(acts1, acts2)                        # sharding-independent RNG
    g1 = _shard_grads(params, screens, rew, term, acts1, 0, B_GLOBAL)
    assert np.abs(g1 - g2).max() <= 1e-12 * max(1.0, np.abs(g1).max())
    # identical update on every rank follows from identical summed gradients
    rms = {k: np.ones_like(v) for k, v in params.items()}
    p1, _ = a3c.update(params, rms, a3c.unflatten_params(g1, A), 7e-4)
    p2, _ = a3c.update(params, rms, a3c.unflatten_params(g2, A), 7e-4)
    assert all(np.allclose(p1[k], p2[k], rtol=0, atol=1e-15) for k in p1)
