"""GPU: the loop replayed from CUDA graphs is the eager loop, bit for bit -- same
sampled actions (the Philox step now comes from a device counter), same learning-rate schedule
(agent.py:393-395 evaluated on the device), same parameters and RMSProp slot after every update."""
import random

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(pkg, cuda, graphs, cycles, B=24, T=5, A=6, host=False, start_step=0):
    random.seed(123)
    cfg = pkg.config.get_config({"model": "m1", "num_envs": B, "t_max": T, "cuda_graphs": graphs})
    env = pkg.GymEnvironment(cfg, env=pkg.SyntheticAtari(B, A, seed=5, pool=T, device=cuda, host=host,
                                                         p_terminal=0.1), device=cuda)
    agent = pkg.Agent(cfg, env, device=cuda)
    agent.step_op = start_step
    agent.before_train()
    acts, snaps = [], []
    for c in range(cycles):
        for _ in range(T):
            a = agent.predict()
            acts.append(a.clone())
            scr, rew, term = env.act(a, is_training=True, fused=True)
            agent.observe(scr, rew, a, term)
            agent.step += 1
        snaps.append((agent.network.params.clone(), agent.network.rms.clone()))
    torch.cuda.synchronize()
    return agent, torch.stack(acts), snaps


@pytest.mark.parametrize("host", [False, True])
def test_graph_replay_equals_eager_loop(pkg, cuda, host):
    cycles = 9                                                   # 2 eager, 2 capturing (period 2), 5 replaying
    eager, acts_e, snaps_e = _run(pkg, cuda, False, cycles, host=host, start_step=79999000)
    graph, acts_g, snaps_g = _run(pkg, cuda, True, cycles, host=host, start_step=79999000)
    assert eager.graph_replays == 0 and not eager._graphs
    assert graph.graph_replays > 0 and graph.history.ring_slots == 10
    # ring period 2 x buffer parity 2 -> a few dozen graphs, all captured by the 4th cycle
    assert len(graph._graphs) <= 2 * 2 * 2 * 5, len(graph._graphs)
    assert torch.equal(acts_e, acts_g)
    for (pe, re_), (pg, rg) in zip(snaps_e, snaps_g):
        assert torch.equal(pe, pg) and torch.equal(re_, rg)
    # the schedule: close to max_step the learning rate changes visibly from update to update
    assert not torch.equal(snaps_g[-1][0], snaps_g[-2][0])
    assert graph.update_count == cycles and graph.step_dev.item() == graph.step


def test_graph_loop_follows_a_manual_step_change(pkg, cuda):
    """Callers that move ``agent.step`` by hand (the parity tests do) are followed: the device
    counter is re-synchronised before the next launch."""
    agent, _, _ = _run(pkg, cuda, True, 5)
    agent.step += 1000
    a1 = agent.predict().clone()
    assert agent.step_dev.item() == agent.step
    ref = pkg.ops.batch_sample(agent.network.policy[agent.network._rows(agent.t)], step=agent.step,
                               seed=agent.seed, env_id_base=0)
    assert torch.equal(a1, ref)
