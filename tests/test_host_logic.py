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
