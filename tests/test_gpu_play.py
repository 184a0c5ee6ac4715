"""GPU: the callers on either side of the hot path -- real gym-style
emulators feeding K1 through GymVectorAdapter (environment.py:14-65) and the evaluation loop
Agent.play (agent.py:351-391)."""
import random

import numpy as np
import pytest
import torch

from oracle import preprocess
from util import StubGymEnv

pytestmark = pytest.mark.gpu


def _agent(pkg, cuda, B, mode, episode_len=4, lives=1):
    cfg = pkg.config.get_config({"model": "m1", "num_envs": B, "t_max": 5, "loss_mode": mode,
                                 "random_start": 3})
    envs = [StubGymEnv(100 + b, n_actions=4, episode_len=episode_len + b % 3, lives=lives) for b in range(B)]
    adapter = pkg.environment.GymVectorAdapter(envs, device=cuda)
    env = pkg.GymEnvironment(cfg, env=adapter, device=cuda)
    return pkg.Agent(cfg, env, device=cuda), env, envs


def test_adapter_frames_reach_k1_bit_exact(pkg, cuda):
    """Host emulator frames -> pinned buffer -> arl_upload_frames (only the rows K1 reads) ->
    Environment.screen == the reference's expression on the very frame the emulator produced."""
    random.seed(5)
    agent, env, envs = _agent(pkg, cuda, 5, "a3c")
    env.new_random_game()
    scr = env.screen.cpu().numpy()
    for b, e in enumerate(envs):
        assert np.array_equal(scr[b], preprocess.screen(e.frames[-1])), b
    acts = torch.tensor([0, 1, 2, 3, 1], dtype=torch.int32, device=cuda)
    s, r, t = env.act(acts, is_training=True)
    for b, e in enumerate(envs):
        assert np.array_equal(s[b].cpu().numpy(), preprocess.screen(e.frames[-1])), b
    assert r.cpu().tolist() == [0.0, 1.0, 2.0, 3.0, 1.0]


@pytest.mark.parametrize("mode", ["a3c", "async_q"])
def test_play_runs_episodes_to_the_end(pkg, cuda, mode):
    """agent.py:351-391: every env plays until its own terminal; rewards are counted up to it."""
    random.seed(7)
    B = 6
    agent, env, envs = _agent(pkg, cuda, B, mode, episode_len=4, lives=1)
    before = agent.network.params.clone()
    best, best_idx, means = agent.play(n_step=50, n_episode=3, test_ep=0.0 if mode == "async_q" else None)
    assert len(means) == 3 and 0 <= best_idx < 3
    # reward = the action index (0..3), at most episode_len+2 steps per episode
    assert 0.0 <= best <= 3.0 * 6 and all(0.0 <= m <= 18.0 for m in means)
    assert bool((agent.network.params == before).all())           # play never trains
    # the rollout slots are untouched too: a training cycle after play still works
    agent.train(num_steps=5)
    torch.cuda.synchronize()
    assert agent.update_count == 1 and bool(torch.isfinite(agent.network.params).all())
