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
    agent, env, envs = _agent(pkg, cuda, 5, "a3c", episode_len=20)   # no life lost in the random start
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


def test_pil_branch_with_host_fed_backends(pkg, cuda):
    """resize='pil' (scipy.misc.imresize, environment.py:5-12) reads every source row,
    so host-fed backends must upload the WHOLE frame for it (arl_upload_frames_full), not only the
    168 rows the cv2 kernel reads.  Real-emulator adapter and the pinned-host synthetic pool."""
    random.seed(11)
    B = 4
    cfg = pkg.config.get_config({"model": "m1", "num_envs": B, "t_max": 5, "resize": "pil",
                                 "random_start": 2})
    envs = [StubGymEnv(40 + b, n_actions=4, episode_len=50, lives=1) for b in range(B)]
    adapter = pkg.environment.GymVectorAdapter(envs, device=cuda)
    env = pkg.GymEnvironment(cfg, env=adapter, device=cuda)
    assert adapter.resize == "pil" and adapter.h2d_bytes_per_step == B * 100800
    env.new_random_game()
    for _ in range(3):                                            # both staging buffers get used
        s, r, t = env.act(torch.zeros(B, dtype=torch.int32, device=cuda), is_training=True)
        for b, e in enumerate(envs):
            assert np.array_equal(s[b].cpu().numpy(), preprocess.screen(e.frames[-1], resize="pil")), b
    syn = pkg.SyntheticAtari(B, 4, seed=3, pool=4, device=cuda, host=True)
    env2 = pkg.GymEnvironment(cfg, env=syn, device=cuda)
    assert syn.resize == "pil"
    for _ in range(3):
        s, r, t = env2.act(torch.zeros(B, dtype=torch.int32, device=cuda), is_training=True)
        ref = preprocess.screen(syn._frames[(syn._i - 1) % syn.pool].numpy(), resize="pil")
        assert np.array_equal(s.cpu().numpy(), ref)


def test_per_env_random_start_and_restart(pkg, cuda):
    """environment.py:35-40 + agent.py:66-67, one reference worker per env: every env draws its OWN
    randint(0, random_start - 1) no-op count in new_random_game, and in the training loop only the
    envs whose act() was terminal restart (reset iff their game is over), each with a fresh draw;
    the History keeps the terminal screen, not the restart screen."""
    B = 5
    cfg = pkg.config.get_config({"model": "m1", "num_envs": B, "t_max": 5, "random_start": 7})
    envs = [StubGymEnv(70 + b, n_actions=4, episode_len=9 + b, lives=1) for b in range(B)]   # > 1 + max draw
    adapter = pkg.environment.GymVectorAdapter(envs, device=cuda, auto_reset=False)
    env = pkg.GymEnvironment(cfg, env=adapter, device=cuda)
    agent = pkg.Agent(cfg, env, device=cuda)
    random.seed(123)
    draws = [random.randint(0, 6) for _ in range(B)]             # the draws new_random_game will make
    random.seed(123)
    agent.before_train()
    assert len(set(draws)) > 1                                    # (not one shared draw)
    assert [len(e.frames) for e in envs] == [2 + d for d in draws]   # reset + 1 + k_b no-op steps
    # 4 copies of each env's own first screen (agent.py:37-38)
    first = agent.history.first_slot(0)
    ring = agent.history.planes().cpu().numpy()
    for b, e in enumerate(envs):
        want = preprocess.screen(e.frames[-1])
        for k in range(4):
            assert np.array_equal(ring[b, (first + k) % agent.history.ring_slots], want), (b, k)
    # one training step at a time: after the step in which env b terminates, ONLY env b restarts
    for step in range(6):
        n0 = [len(e.frames) for e in envs]
        r0 = [e.resets for e in envs]
        _one_more_step(agent)
        for b, e in enumerate(envs):
            grew = len(e.frames) - n0[b]
            if agent.batch_terminal_last[b]:
                assert e.resets == r0[b] + 1 and grew >= 3       # step + reset frame + >= 1 no-op
            else:
                assert e.resets == r0[b] and grew == 1           # just its own step
        # the newest history plane is the screen of the env's step frame (terminal screen included)
        newest = agent.history.planes(agent.history.head).cpu().numpy()
        for b, e in enumerate(envs):
            assert np.array_equal(newest[b], preprocess.screen(e.frames[n0[b]])), (step, b)


def _one_more_step(agent):
    """The body of Agent.train's loop for one more step (no before_train)."""
    action = agent.predict()
    screen, reward, terminal = agent.env.act(action, is_training=True, fused=True)
    agent.observe(screen, reward, action, terminal)
    agent.batch_terminal_last = terminal.cpu().tolist()
    if agent.env.per_env_restart:
        agent.env.new_random_game(mask=terminal)
    agent.step += 1
