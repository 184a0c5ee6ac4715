"""CPU: host-side mirror of the reference interfaces (no kernels)."""
import argparse


def test_config_defaults_match_reference(pkg):
    c = pkg.config.M1
    assert (c.max_step, c.learning_rate, c.discount, c.beta) == (80000000, 0.0007, 0.99, 0.01)
    assert (c.decay, c.epsilon, c.momentum) == (0.99, 0.1, 0.0)
    assert (c.history_length, c.batch_size, c.train_frequency, c.learn_start) == (4, 32, 32, 32)
    assert (c.screen_width, c.screen_height, c.min_reward, c.max_reward) == (84, 84, -1.0, 1.0)
    assert (c.env_name, c.action_repeat, c.random_start) == ('Breakout-v0', 1, 30)
    assert (c.t_max, c.seed, c.clip_norm) == (5, 123, 40.0)


def test_get_config_copies_matching_flags(pkg):
    ns = argparse.Namespace(model='m1', env_name='Pong-v0', t_max=20, num_envs=64, bogus=1,
                            action_repeat=None)
    c = pkg.config.get_config(ns)
    assert (c.env_name, c.t_max, c.num_envs, c.action_repeat) == ('Pong-v0', 20, 64, 1)
    assert not hasattr(c, 'bogus')
    assert pkg.config.M1.env_name == 'Breakout-v0'              # the base class is not mutated
    c2 = pkg.config.get_config({'model': 'm1', 'gpu': False})
    assert c2.cnn_format == 'NHWC'


def test_base_model_flattens_and_strips_underscore(pkg):
    m = pkg.base.BaseModel(pkg.config.M1)
    assert m.test_step == 5000.0 and m.t_max == 5 and m.learning_rate == 0.0007
    assert m.model_dir.startswith('Breakout-v0/') and 'display' not in m.model_dir
    assert m.checkpoint_dir.startswith('checkpoints/Breakout-v0/')


def test_initial_weights_distribution(pkg):
    w = pkg.network.initial_weights(6, seed=123)
    assert list(w) == list(pkg.network.PARAM_NAMES)
    assert w['l1_w'].shape == (8, 8, 4, 16) and float(w['l1_w'].abs().max()) <= 0.04
    assert abs(float(w['l4_w'].std()) - 0.02) < 1e-3 and float(w['l4_b'].abs().max()) == 0.0
    w2 = pkg.network.initial_weights(6, seed=123)
    assert all(bool((w[k] == w2[k]).all()) for k in w)


def test_unbuilt_resize_branch_fails_loudly(pkg):
    """environment.py:5-12 has two resize branches ('cv2', 'pil'); anything else is an error, not a
    silent substitution."""
    import pytest
    cfg = pkg.config.get_config({"model": "m1", "resize": "lanczos"})
    with pytest.raises((NotImplementedError, pkg._cabi.ArlError)):
        pkg.GymEnvironment(cfg, env=object(), device="cuda:0")


def test_gym_vector_adapter_host_logic(pkg):
    """GymVectorAdapter (the real-emulator seam of environment.py:14-65) on the CPU: frames of B
    gym-style emulators are gathered unchanged, rewards / terminals / lives are batched, both the
    gym 4-tuple and the gymnasium 5-tuple APIs are accepted, a finished game is reset in place."""
    import numpy as np
    import torch
    from util import StubGymEnv
    envs = [StubGymEnv(1), StubGymEnv(2, gymnasium=True), StubGymEnv(3, episode_len=2, lives=1)]
    ad = pkg.environment.GymVectorAdapter(envs, device='cpu')
    assert ad.num_envs == 3 and ad.action_space.n == 4
    f0 = ad.reset()
    assert f0.shape == (3, 210, 160, 3) and f0.dtype == torch.uint8
    for b, e in enumerate(envs):
        assert np.array_equal(f0[b].numpy(), e.frames[-1]) and e.resets == 1
    assert ad.ale.lives().tolist() == [2, 2, 1]
    f1, r1, t1, _ = ad.step(torch.tensor([1, 2, 3], dtype=torch.int32))
    assert r1.tolist() == [1.0, 2.0, 3.0] and t1.tolist() == [False, False, False]
    assert all(np.array_equal(f1[b].numpy(), e.frames[-1]) for b, e in enumerate(envs))
    assert not np.array_equal(f1.numpy(), f0.numpy())             # double buffer: f0 is still intact
    assert np.array_equal(f0[0].numpy(), envs[0].frames[0])
    f2, r2, t2, _ = ad.step(torch.tensor([0, 0, 0], dtype=torch.int32))
    assert t2.tolist() == [False, False, True]                    # env 2: episode_len 2, its only life
    # the terminal step hands out the TERMINAL frame (the reference pushes it into the history,
    # agent.py:62-64, before it restarts the env); the emulator is not reset yet
    assert envs[2].resets == 1 and ad.ale.lives().tolist() == [2, 2, 0]
    assert np.array_equal(f2[2].numpy(), envs[2].frames[-1])
    # auto_reset: a finished emulator is reset in place on its NEXT step (reward 0, not terminal)
    f2b, r2b, t2b, _ = ad.step(torch.tensor([1, 1, 1], dtype=torch.int32))
    assert envs[2].resets == 2 and ad.ale.lives().tolist()[2] == 1
    assert r2b.tolist() == [1.0, 1.0, 0.0] and t2b.tolist() == [True, True, False]   # envs 0, 1: third step of a 3-step life
    assert np.array_equal(f2b[2].numpy(), envs[2].frames[-1])
    # masked reset: only env 0 restarts, the others keep their last frame
    f3 = ad.reset(torch.tensor([True, False, False]))
    assert envs[0].resets == 2 and envs[1].resets == 1
    assert np.array_equal(f3[1].numpy(), f2b[1].numpy()) and np.array_equal(f3[0].numpy(), envs[0].frames[-1])
    assert len(ad.action_space.sample()) == 3


def test_gym_vector_adapter_per_env_restart(pkg):
    """agent.py:66-67 restarts the env that died, environment.py:28-40: reset only when the game
    is over (lives == 0), one no-op step, then that env's OWN number of random-start no-op steps.
    The other envs are not touched."""
    import numpy as np
    import torch
    from util import StubGymEnv
    envs = [StubGymEnv(10 + b, episode_len=3, lives=2) for b in range(4)]
    ad = pkg.environment.GymVectorAdapter(envs, device='cpu', auto_reset=False)
    ad.restart([True] * 4, [0, 1, 2, 3])                          # first call: every env resets
    assert [e.resets for e in envs] == [1, 1, 1, 1]
    assert [len(e.frames) for e in envs] == [2, 3, 4, 5]          # reset + (1 + k_b) no-op steps
    n_before = [len(e.frames) for e in envs]
    f, r, t = ad.restart([False, True, False, True], [9, 2, 9, 0])
    assert [len(e.frames) - n for e, n in zip(envs, n_before)] == [0, 3, 0, 1]
    assert np.array_equal(f[1].numpy(), envs[1].frames[-1]) and np.array_equal(f[0].numpy(), envs[0].frames[-1])
    # env 3 (5 steps so far with episode_len 3) lost a life on the way but its game is not over:
    # no reset for it (environment.py:29-30)
    assert envs[3].resets == 1 and envs[3]._lives == 1
    # a game-over emulator without auto_reset idles on its last frame until it is restarted
    while envs[3]._lives > 0:
        f, r, t, _ = ad.step(torch.zeros(4, dtype=torch.int32))
    n3 = len(envs[3].frames)
    f2, r2, t2, _ = ad.step(torch.zeros(4, dtype=torch.int32))
    assert len(envs[3].frames) == n3 and t2.tolist()[3] is True and r2.tolist()[3] == 0.0
    ad.restart([False, False, False, True], [0, 0, 0, 1])
    assert envs[3].resets == 2 and len(envs[3].frames) == n3 + 3   # reset frame + 2 no-op steps


def test_adapter_without_ale_resets_on_terminal(pkg):
    """an emulator without .ale has no lives -- every terminal is a game over."""
    import torch
    from util import StubGymEnv
    e = StubGymEnv(3, episode_len=2, lives=5)
    del e.ale
    ad = pkg.environment.GymVectorAdapter([e], device='cpu')
    ad.reset()
    ad.step(torch.zeros(1, dtype=torch.int32))
    _, _, t, _ = ad.step(torch.zeros(1, dtype=torch.int32))
    assert t.tolist() == [True] and e.resets == 1
    _, r, t, _ = ad.step(torch.zeros(1, dtype=torch.int32))      # reset in place on the next step
    assert e.resets == 2 and t.tolist() == [False]


def test_split_block_encode_decode_round_trip(pkg):
    """The bf16 hi+lo block layout of the fc256 operands (include/asyncrl_b200.h): decode(encode(x))
    reproduces x to 2^-16 relative, and the block is [hi|lo][chunk of 8 columns][row][8]."""
    import torch
    net = pkg.network
    g = torch.Generator().manual_seed(0)
    x = torch.randn(37, 64, generator=g) * 3.0
    blk = net.encode_split(x)
    assert blk.shape == x.shape and blk.dtype == torch.float32
    y = net.decode_split(blk, 37, 64)
    assert float((y - x).abs().max() / x.abs().max()) < 2.0 ** -16
    raw = blk.reshape(-1).view(torch.bfloat16).reshape(2, 8, 37, 8)
    assert bool((raw[0, 3, 5].float() == x[5, 24:32].to(torch.bfloat16).float()).all())
    assert float((raw[0].float() + raw[1].float() - x.reshape(37, 8, 8).permute(1, 0, 2)).abs().max()) < 1e-4
