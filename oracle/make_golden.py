"""Generate tests/golden/preprocess_golden.npz by EXECUTING the reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference,
which does not exist on the GPU box):

    python oracle/make_golden.py

What it executes, unmodified, from /root/reference:
  * src/environment.py  ``Environment.screen`` (:49-53), ``new_random_game``
    (:35-40) and ``GymEnvironment.act`` (:78-96), behind a stub ``gym`` module
    (gym/ALE are absent here) and ``builtins.xrange = range`` (the file is py2).
    The import takes the ``cv2.resize`` branch of environment.py:5-12.
  * src/history.py      ``History`` (add/get/copy), as-is.
  * config.py           ``M1`` with cnn_format='NHWC' (what main.py:45 forces).

Inputs are deterministic (numpy PCG64 seeds + closed-form frames) and stored in
the fixture together with the outputs so the GPU box needs neither numpy-RNG
stability nor the reference.
"""
import builtins
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden",
                   "preprocess_golden.npz")
H, W = 210, 160


class _StubALE:
    def __init__(self, env):
        self._env = env

    def lives(self):
        return self._env._lives


class _StubSpace:
    n = 6

    def sample(self):
        return 0


class StubGymEnv:
    """Feeds a fixed list of frames/rewards/terminals/lives through gym's 4-tuple API."""

    def __init__(self):
        self.frames = None
        self.rewards = None
        self.terminals = None
        self.lives_seq = None
        self._i = 0
        self._lives = 0
        self.ale = _StubALE(self)
        self.action_space = _StubSpace()

    def load(self, frames, rewards=None, terminals=None, lives=None):
        n = len(frames)
        self.frames = frames
        self.rewards = rewards if rewards is not None else [0.0] * n
        self.terminals = terminals if terminals is not None else [False] * n
        self.lives_seq = lives if lives is not None else [3] * n
        self._i = 0

    def reset(self):
        self._lives = self.lives_seq[0]
        return self.frames[0]

    def step(self, action):
        i = self._i % len(self.frames)
        self._i += 1
        self._lives = self.lives_seq[i]
        return self.frames[i], self.rewards[i], self.terminals[i], {}

    def render(self):
        pass


def load_reference():
    """Import the reference's config/history/environment with the stub gym."""
    sys.dont_write_bytecode = True
    builtins.xrange = range
    gym = types.ModuleType("gym")
    gym.make = lambda name: StubGymEnv()
    sys.modules["gym"] = gym
    if "scipy.misc" in sys.modules and hasattr(sys.modules["scipy.misc"], "imresize"):
        raise RuntimeError("scipy.misc.imresize exists: reference would not take the cv2 branch")
    sys.path.insert(0, REF)
    import config as ref_config                      # noqa
    from src import environment as ref_env           # noqa
    from src import history as ref_hist              # noqa
    assert ref_env.imresize.__name__ == "resize", ref_env.imresize

    class Cfg(ref_config.M1):
        cnn_format = "NHWC"
        display = False
    return Cfg, ref_env, ref_hist


def luma_exact_integer_triples():
    """All (R,G,B) with (2126R+7152G+722B) % 10000 == 0 (3384 of 2^24)."""
    r, g = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
    out = []
    for b in range(256):
        s = 2126 * r + 7152 * g + 722 * b
        m = (s % 10000) == 0
        out.append(np.stack([r[m], g[m], np.full(m.sum(), b)], axis=1))
    return np.concatenate(out).astype(np.uint8)


def make_frames():
    rng = np.random.default_rng(123)
    frames = {}
    frames["random0"] = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    frames["random1"] = np.roll(frames["random0"][::-1, :, ::-1], (3, 5), axis=(0, 1)).copy()
    # Atari-like: few colours, flat regions, some sprites
    pal = rng.integers(0, 256, (8, 3), dtype=np.uint8)
    pal[0] = 0
    f = np.zeros((H, W, 3), np.uint8)
    f[:] = pal[0]
    f[17:26, :, :] = pal[1]
    f[57:93:6, 8:152, :] = pal[2]
    f[60:96:6, 8:152, :] = pal[3]
    f[189:193, 70:86, :] = pal[4]
    f[100:104, 50:52, :] = pal[5]
    f[:, :8, :] = pal[6]
    f[:, 152:, :] = pal[6]
    frames["atari_like"] = f
    frames["white"] = np.full((H, W, 3), 255, np.uint8)
    frames["black"] = np.zeros((H, W, 3), np.uint8)
    g = np.zeros((H, W, 3), np.uint8)
    g[...] = (np.arange(H * W).reshape(H, W) % 256)[..., None]
    frames["greys"] = g
    # every exact-integer luma triple (incl. the 774 that truncate one low), tiled
    tri = luma_exact_integer_triples()
    t = np.zeros((H * W, 3), np.uint8)
    reps = (H * W + len(tri) - 1) // len(tri)
    t[:] = np.tile(tri, (reps, 1))[:H * W]
    frames["luma_edge"] = t.reshape(H, W, 3)
    # vertical / horizontal ramps hit every tap pair with distinct values
    ramp = np.zeros((H, W, 3), np.uint8)
    ramp[..., 0] = (np.arange(W) * 255 // (W - 1))[None, :]
    ramp[..., 1] = (np.arange(H) * 255 // (H - 1))[:, None]
    ramp[..., 2] = ((np.arange(H)[:, None] * 7 + np.arange(W)[None, :] * 13) % 256)
    frames["ramps"] = ramp
    return frames


def main():
    Cfg, ref_env, ref_hist = load_reference()
    cfg = Cfg()
    frames = make_frames()
    names = sorted(frames)
    out = {"names": np.array(names)}

    # 1. Environment.screen on single frames (environment.py:49-53, executed)
    env = ref_env.GymEnvironment(cfg)
    for n in names:
        env._screen = frames[n]
        scr = env.screen
        assert scr.dtype == np.uint8 and scr.shape == (84, 84)
        out["frame_" + n] = frames[n]
        out["screen_" + n] = scr

    # 2. the full set of exact-integer luma triples and the reference's luma on them
    tri = luma_exact_integer_triples()
    env._screen = tri.reshape(1, -1, 3)
    y = 0.2126 * env._screen[:, :, 0] + 0.7152 * env._screen[:, :, 1] + 0.0722 * env._screen[:, :, 2]
    out["luma_triples"] = tri
    out["luma_triples_y"] = y.astype(np.uint8).reshape(-1)

    # 3. act()/new_random_game() sequence + History (history.py executed as-is)
    import random
    random.seed(123)                                   # main.py:41
    # the 10 sequence frames are rolls of frame "random0" (the fixture stores only the shifts)
    shifts = [(7 * k + 1, 13 * k + 2) for k in range(10)]
    seq = [np.roll(frames["random0"], s, axis=(0, 1)) for s in shifts]
    rewards = [0.0, 1.0, 0.0, 5.0, -3.0, 0.0, 0.0, 1.0, 0.0, 0.0]
    terminals = [False] * 10
    lives = [3, 3, 3, 3, 2, 2, 2, 2, 2, 2]
    env = ref_env.GymEnvironment(cfg)
    env.env.load(seq, rewards, terminals, lives)
    hist = ref_hist.History(cfg)
    screen0, _, _, _ = env.new_game()                  # environment.py:28-33 (steps once)
    for _ in range(cfg.history_length):                # agent.py:37-38
        hist.add(screen0)
    stacks = [hist.copy()]
    acts, rews, terms, screens = [], [], [], [screen0]
    for t in range(6):
        s, r, term = env.act(t % 6, is_training=True)  # environment.py:78-96
        hist.add(s)                                    # agent.py:156
        stacks.append(hist.copy())                     # agent.py:157
        screens.append(s)
        rews.append(r)
        terms.append(term)
        acts.append(t % 6)
    out["seq_shifts"] = np.array(shifts)
    out["seq_screens"] = np.stack(screens)
    st = np.stack(stacks)                              # [7, 84, 84, 4] float32 NHWC
    assert st.dtype == np.float32 and np.array_equal(st, st.astype(np.uint8))
    out["seq_stacks"] = st.astype(np.uint8)            # exact: u8 values held in f32
    out["seq_rewards"] = np.array(rews, np.float64)
    out["seq_terminals"] = np.array(terms)
    out["seq_lives"] = np.array(lives)

    # 4. History ordering probe: 6 adds of constant planes -> channels [2,3,4,5]
    h2 = ref_hist.History(cfg)
    for k in range(6):
        h2.add(np.full((84, 84), k, np.uint8))
    out["hist_order"] = h2.get()[0, 0, :].copy()

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **out)
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes;",
          "hist_order", out["hist_order"], "white ->", int(out["screen_white"][0, 0]))


if __name__ == "__main__":
    main()
