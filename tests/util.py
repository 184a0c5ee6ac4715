"""Shared helpers for the parity tests."""
import numpy as np

REL_TOL = 1e-3      # BASELINE.json north_star: rel-err <= 1e-3 for logits, values, returns, grads


def rel_err(x, ref):
    """max |x - ref| / max |ref|  (scale-relative max error; 0/0 -> 0)."""
    x = np.asarray(x, np.float64)
    ref = np.asarray(ref, np.float64)
    assert x.shape == ref.shape, (x.shape, ref.shape)
    d = float(np.abs(x - ref).max()) if x.size else 0.0
    s = float(np.abs(ref).max()) if ref.size else 0.0
    return d / s if s > 0 else d


def norm_err(x, ref):
    x = np.asarray(x, np.float64).ravel()
    ref = np.asarray(ref, np.float64).ravel()
    n = float(np.linalg.norm(ref))
    return float(np.linalg.norm(x - ref)) / n if n > 0 else float(np.linalg.norm(x - ref))


def golden_seq_frames(golden):
    """The 10 sequence frames of the fixture (rolls of frame random0, see oracle/make_golden.py)."""
    base = golden["frame_random0"]
    return [np.roll(base, tuple(int(v) for v in s), axis=(0, 1)) for s in golden["seq_shifts"]]


def unblock(ring):
    """Ring planes are stored in 4x4 blocks (space-to-depth, see src/history.py); returns the
    row-major [..., 84, 84] screens.  numpy or torch."""
    lead = tuple(ring.shape[:-2])
    x = ring.reshape(lead + (21, 21, 4, 4))
    n = len(lead)
    order = tuple(range(n)) + (n, n + 2, n + 1, n + 3)
    x = x.transpose(order) if isinstance(x, np.ndarray) else x.permute(*order)
    return x.reshape(lead + (84, 84))


def block(screens):
    """Inverse of unblock."""
    lead = tuple(screens.shape[:-2])
    x = screens.reshape(lead + (21, 4, 21, 4))
    n = len(lead)
    order = tuple(range(n)) + (n, n + 2, n + 1, n + 3)
    x = x.transpose(order) if isinstance(x, np.ndarray) else x.permute(*order)
    return x.reshape(lead + (84, 84))


class StubGymEnv(object):
    """A gym-style emulator for the adapter tests (what gym.make(...) returns in the reference,
    environment.py:16): deterministic 210x160x3 frames, ``episode_len`` steps per life, ``lives``
    lives, reward = action.  ``gymnasium=True`` switches to the 5-tuple / (obs, info) API."""

    class _Space(object):
        def __init__(self, n, rng):
            self.n, self._rng = n, rng

        def sample(self):
            return int(self._rng.integers(0, self.n))

    class _ALE(object):
        def __init__(self, env):
            self._env = env

        def lives(self):
            return self._env._lives

    def __init__(self, seed, n_actions=4, episode_len=3, lives=2, gymnasium=False):
        import numpy as np
        self._np = np
        self._rng = np.random.default_rng(seed)
        self._seed, self._len, self._max_lives, self._gymnasium = seed, episode_len, lives, gymnasium
        self._lives, self._t, self.resets = 0, 0, 0
        self.action_space = self._Space(n_actions, self._rng)
        self.ale = self._ALE(self)
        self.frames = []                                         # every frame handed out, in order

    def _frame(self):
        f = self._rng.integers(0, 256, (210, 160, 3), dtype=self._np.uint8)
        self.frames.append(f)
        return f

    def reset(self):
        self._lives, self._t = self._max_lives, 0
        self.resets += 1
        f = self._frame()
        return (f, {}) if self._gymnasium else f

    def step(self, a):
        self._t += 1
        done = self._t % self._len == 0
        if done:
            self._lives -= 1
        f = self._frame()
        if self._gymnasium:
            return f, float(a), done, False, {}
        return f, float(a), done, {}

    def render(self):
        pass
