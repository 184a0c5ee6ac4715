"""Environment wrappers with the reference's frame API (src/environment.py:14-106), batched
over ``num_envs`` environments per process.

The emulator itself (gym/ALE) is outside the hot path; ``SyntheticAtari`` stands in for it
with Atari-shaped u8 frames (BASELINE.json configs).  Any backend with the same batched
gym-like surface can be passed as ``env=``:

    reset(mask=None) -> frames u8 [B,210,160,3] (cuda)
    step(actions i32[B]) -> (frames, rewards f32[B], terminals bool[B], info)
    ale.lives() -> i32 [B];  action_space.n;  action_space.sample() -> i32 [B]
"""
import random

import torch

from .. import _cabi
from .history import SCREEN, FRAME_SHAPE, from_blocked


class _ALE(object):
    def __init__(self, env):
        self._env = env

    def lives(self):
        return self._env._lives


class _ActionSpace(object):
    def __init__(self, env, n):
        self._env, self.n = env, n

    def sample(self):
        return torch.randint(0, self.n, (self._env.num_envs,), device=self._env.device,
                             dtype=torch.int32, generator=self._env._gen)


class SyntheticAtari(object):
    """Deterministic stand-in for ``gym.make(name)`` x num_envs.

    Frames come from a rotating pool of ``pool`` pre-generated steps so that a benchmark
    reads HBM (pool * B * 100 800 B  >>  L2).  rewards in {-1,0,+1} w.p. {.05,.9,.05} times
    ``reward_scale``; terminals ~ Bernoulli(p_terminal).  ``host=True`` keeps the pool in
    pinned host memory and copies each step's frames host->device inside ``step`` (the
    end-to-end path a real emulator would take).
    """

    def __init__(self, num_envs, action_size=6, seed=123, pool=8, device='cuda', host=False,
                 p_terminal=0.01, reward_scale=1.0, structured=False, lives=3, shard=None):
        self.num_envs, self.device, self.host = int(num_envs), torch.device(device), bool(host)
        self.pool = int(pool)
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(int(seed))
        # shard=(rank, world): draw the pool for the world*num_envs GLOBAL envs from the one seed
        # and keep this rank's slice, so an N-rank run sees exactly the single-process data
        # (multi-GPU parity tests; benchmarks use per-rank seeds instead -- no extra memory)
        rank, world = shard if shard is not None else (0, 1)
        B, P = self.num_envs * int(world), self.pool
        shape = (P, B) + FRAME_SHAPE
        if structured:   # few colours, large flat regions (Atari-like)
            pal = torch.randint(0, 256, (16, 3), device=self.device, dtype=torch.uint8,
                                generator=self._gen)
            idx = torch.randint(0, 16, (P, B, 21, 16), device=self.device, generator=self._gen)
            idx = idx.repeat_interleave(10, dim=2).repeat_interleave(10, dim=3)
            frames = pal[idx]
        else:
            frames = torch.randint(0, 256, shape, device=self.device, dtype=torch.uint8,
                                   generator=self._gen)
        u = torch.rand(P, B, device=self.device, generator=self._gen)
        self._rewards = (torch.where(u < 0.05, -1.0, torch.where(u > 0.95, 1.0, 0.0))
                         * reward_scale).float()
        self._terminals = torch.rand(P, B, device=self.device, generator=self._gen) < p_terminal
        if shard is not None:
            lo, hi = rank * self.num_envs, (rank + 1) * self.num_envs
            frames = frames[:, lo:hi].contiguous()
            self._rewards = self._rewards[:, lo:hi].contiguous()
            self._terminals = self._terminals[:, lo:hi].contiguous()
            B = self.num_envs
            shape = (P, B) + FRAME_SHAPE
        if self.host:
            self._frames = torch.empty(shape, dtype=torch.uint8, pin_memory=True)
            self._frames.copy_(frames)
            # double-buffered staging: the upload of step i+1 overlaps the compute of step i
            self._stage = [torch.zeros((B,) + FRAME_SHAPE, dtype=torch.uint8, device=self.device)
                           for _ in range(2)]
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._ready = [None, None]
            self._staged_for = [-1, -1]
            del frames
        else:
            self._frames = frames
        self._lives = torch.full((B,), int(lives), dtype=torch.int32, device=self.device)
        self._i = 0
        self.ale = _ALE(self)
        self.action_space = _ActionSpace(self, int(action_size))
        self.set_resize('cv2')

    def set_resize(self, resize):
        """Called by Environment with its resize branch (environment.py:5-12).  'cv2' reads 168 of
        the 210 source rows, so only those cross PCIe; 'pil' (scipy.misc.imresize) reads every
        row, so the whole frame is uploaded."""
        self.resize = resize
        rows = 168 if resize == 'cv2' else FRAME_SHAPE[0]
        self._upload_fn = "arl_upload_frames" if resize == 'cv2' else "arl_upload_frames_full"
        self.h2d_bytes_per_step = self.num_envs * rows * FRAME_SHAPE[1] * FRAME_SHAPE[2] if self.host else 0
        if self.host:
            self._staged_for = [-1, -1]                          # staged frames may lack rows

    def _upload(self, i):
        """Queue the pinned-host -> device copy of step i's frames on the copy stream.  The copy
        first waits for everything already queued on the compute stream, which includes every
        reader of the buffer's previous content (step i-2)."""
        k = i & 1
        if self._staged_for[k] == i:
            return
        self._copy_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._copy_stream):
            # cv2 branch: only the 168 of 210 rows K1 reads cross PCIe (80 640 B per frame)
            _cabi.call(self._upload_fn, self._frames[i % self.pool].data_ptr(),
                       _cabi.ptr(self._stage[k]), self.num_envs, self._copy_stream.cuda_stream)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._ready[k], self._staged_for[k] = ev, i

    def _frame(self, i):
        if not self.host:
            return self._frames[i % self.pool]
        self._upload(i)
        k = i & 1
        torch.cuda.current_stream().wait_event(self._ready[k])
        self._upload(i + 1)                                      # prefetch the next step
        return self._stage[k]

    def reset(self, mask=None):
        return self._frame(self._i)

    def step(self, actions):
        i = self._i
        self._i += 1
        k = i % self.pool
        return self._frame(i), self._rewards[k], self._terminals[k], {}

    def render(self):
        pass


class GymVectorAdapter(object):
    """Real emulators behind the batched surface above: ``envs`` is a list of gym-style
    environments, one per env slot, exactly what the reference's ``gym.make(config.env_name)``
    returns (environment.py:16): ``reset() -> frame`` (or ``(frame, info)``), ``step(a) ->
    (frame, reward, terminal, info)`` (or gymnasium's 5-tuple), ``ale.lives()``,
    ``action_space.n`` / ``.sample()``.  The emulators run on the host; each step's frames are
    gathered into one pinned buffer u8 [B,210,160,3] and cross PCIe once -- on a CUDA device only
    the 168 of 210 rows K1 reads (arl_upload_frames), double-buffered so that the copy of step
    i+1 can overlap the kernels of step i.  ``restart(mask, noop_steps)`` is the per-env
    new_random_game of the reference's training loop (agent.py:66-67, environment.py:28-40)."""

    def __init__(self, envs, device='cuda', auto_reset=True):
        self.envs = list(envs)
        self.num_envs = len(self.envs)
        self.device = torch.device(device)
        self.auto_reset = bool(auto_reset)
        B = self.num_envs
        pin = self.device.type == 'cuda'
        self._host = [torch.zeros((B,) + FRAME_SHAPE, dtype=torch.uint8, pin_memory=pin) for _ in range(2)]
        self._dev = [torch.zeros((B,) + FRAME_SHAPE, dtype=torch.uint8, device=self.device) for _ in range(2)]
        self._k = 0
        self._rew = torch.zeros(B, dtype=torch.float32)
        self._term = torch.zeros(B, dtype=torch.bool)
        self._lives_host = torch.zeros(B, dtype=torch.int32)
        self._over = [True] * B                                  # game over, waiting for a reset
        self._uploaded = [None, None]                            # event after each buffer's upload
        self.ale = self                                          # ale.lives()
        self.action_space = self                                 # action_space.n / .sample()
        self.n = int(self.envs[0].action_space.n)
        self.set_resize('cv2')

    def set_resize(self, resize):
        """See SyntheticAtari.set_resize: 'pil' needs every source row on the device."""
        self.resize = resize
        rows = 168 if resize == 'cv2' else FRAME_SHAPE[0]
        self._upload_fn = "arl_upload_frames" if resize == 'cv2' else "arl_upload_frames_full"
        pin = self.device.type == 'cuda'
        self.h2d_bytes_per_step = self.num_envs * rows * FRAME_SHAPE[1] * FRAME_SHAPE[2] if pin else 0

    # -- gym surface of one emulator ---------------------------------------------------------
    @staticmethod
    def _reset_one(env):
        out = env.reset()
        return out[0] if isinstance(out, tuple) else out

    @staticmethod
    def _step_one(env, a):
        out = env.step(int(a))
        if len(out) == 5:                                        # gymnasium: terminated, truncated
            frame, reward, terminated, truncated, _ = out
            return frame, float(reward), bool(terminated or truncated)
        frame, reward, terminal, _ = out
        return frame, float(reward), bool(terminal)

    @staticmethod
    def _lives_one(env):
        """ale.lives(); 0 for emulators without an ALE (no lives: every terminal is a game over,
        so new_game's ``lives == 0`` reset rule of environment.py:29-31 resets them)."""
        ale = getattr(env, 'ale', None) or getattr(getattr(env, 'unwrapped', env), 'ale', None)
        return int(ale.lives()) if ale is not None else 0

    # -- batched surface ---------------------------------------------------------------------
    def lives(self):
        for b, e in enumerate(self.envs):
            self._lives_host[b] = self._lives_one(e)
        return self._lives_host.to(self.device)

    def sample(self):
        return torch.tensor([int(e.action_space.sample()) for e in self.envs], dtype=torch.int32,
                            device=self.device)

    def _upload(self):
        k = self._k
        self._k ^= 1
        host, dev = self._host[k], self._dev[k]
        if self.device.type == 'cuda':
            _cabi.call(self._upload_fn, host.data_ptr(), _cabi.ptr(dev), self.num_envs,
                       _cabi.stream_ptr())
            self._uploaded[k] = torch.cuda.Event()
            self._uploaded[k].record()
        else:
            dev.copy_(host)
        return dev

    def _host_buffer(self):
        """The pinned buffer the next upload will read; its previous upload (two steps ago) must
        have completed before the emulators overwrite it (one event per buffer, not a stream
        synchronize: the kernels of the previous step keep running while the emulators step)."""
        ev = self._uploaded[self._k]
        if ev is not None:
            ev.synchronize()
        return self._host[self._k]

    def reset(self, mask=None):
        buf = self._host_buffer()
        if mask is not None:
            mask = mask.cpu().tolist()
        prev = self._host[self._k ^ 1]
        for b, e in enumerate(self.envs):
            if mask is None or mask[b]:
                buf[b].copy_(torch.as_tensor(self._reset_one(e)))
                self._over[b] = False
            else:
                buf[b].copy_(prev[b])
        return self._upload()

    def step(self, actions):
        """One emulator step per env.  The frame of a terminal step IS the terminal frame (the
        reference pushes it into the history, agent.py:62-64, and only then restarts the env).  An
        emulator whose game is over and that has not been restarted is not stepped again: with
        ``auto_reset`` it is reset on its next step (reward 0, not terminal), otherwise it idles on
        its last frame (reward 0, terminal)."""
        acts = actions.cpu().tolist()
        buf = self._host_buffer()
        prev = self._host[self._k ^ 1]
        for b, e in enumerate(self.envs):
            if self._over[b]:
                if self.auto_reset:
                    buf[b].copy_(torch.as_tensor(self._reset_one(e)))
                    self._rew[b], self._term[b], self._over[b] = 0.0, False, False
                else:
                    buf[b].copy_(prev[b])
                    self._rew[b], self._term[b] = 0.0, True
                continue
            frame, self._rew[b], self._term[b] = self._step_one(e, acts[b])
            self._over[b] = bool(self._term[b]) and self._lives_one(e) == 0
            buf[b].copy_(torch.as_tensor(frame))
        return (self._upload(), self._rew.to(self.device), self._term.to(self.device), {})

    def restart(self, mask, noop_steps):
        """new_random_game for the envs in ``mask`` only (agent.py:66-67 restarts the env that
        died; environment.py:28-40): reset if the game is over (lives == 0), one no-op step, then
        ``noop_steps[b]`` further no-op steps -- each env its own count.  The other envs keep their
        last frame.  Returns (frames, rewards, terminals) like ``step``."""
        buf = self._host_buffer()
        prev = self._host[self._k ^ 1]
        for b, e in enumerate(self.envs):
            if not mask[b]:
                buf[b].copy_(prev[b])
                continue
            if self._over[b] or self._lives_one(e) == 0:          # environment.py:29-30
                frame = self._reset_one(e)
                self._over[b] = False
            for _ in range(1 + int(noop_steps[b])):               # environment.py:31, 37-38
                frame, self._rew[b], self._term[b] = self._step_one(e, 0)
                if self._term[b] and self._lives_one(e) == 0:      # died during its random start
                    self._over[b] = True
                    break
            buf[b].copy_(torch.as_tensor(frame))
        return self._upload(), self._rew.to(self.device), self._term.to(self.device)

    def render(self):
        for e in self.envs:
            e.render()


class Environment(object):
    def __init__(self, config, env=None, device=None):
        self.device = torch.device(device if device is not None else 'cuda')
        _cabi.init(self.device)
        self.num_envs = int(getattr(config, 'num_envs', 1))
        # environment.py:5-12: scipy.misc.imresize (PIL antialiased bilinear) when SciPy still has
        # it, else cv2.resize (the branch the reference takes with any current SciPy: the default)
        self.resize = getattr(config, 'resize', 'cv2')
        if self.resize not in ('cv2', 'pil'):
            raise NotImplementedError("resize=%r: environment.py:5-12 has two branches, 'cv2' "
                                      "(cv2.resize) and 'pil' (scipy.misc.imresize)" % (self.resize,))
        self._push = "arl_preprocess_push" if self.resize == 'cv2' else "arl_preprocess_push_pil"
        self.env = env if env is not None else SyntheticAtari(
            self.num_envs, seed=getattr(config, 'seed', 123), device=self.device)
        if hasattr(self.env, 'set_resize'):                      # host-fed backends: which rows to upload
            self.env.set_resize(self.resize)
        screen_width, screen_height, self.action_repeat, self.random_start = \
            config.screen_width, config.screen_height, config.action_repeat, config.random_start
        self.display = config.display
        self.dims = (screen_width, screen_height)
        if self.dims != (SCREEN, SCREEN):
            raise ValueError("the preprocessing kernel produces 84x84 screens only")
        self._screen = None                            # raw frames u8 [B,210,160,3]
        self.reward = torch.zeros(self.num_envs, device=self.device)
        self.terminal = torch.ones(self.num_envs, dtype=torch.bool, device=self.device)
        self._scratch = torch.empty(self.num_envs, 4, SCREEN, SCREEN, dtype=torch.uint8,
                                    device=self.device)

    @property
    def per_env_restart(self):
        """True when the backend can restart single envs (real emulators): the training loop then
        calls new_random_game(mask=terminal) after every step, as agent.py:66-67 does."""
        return hasattr(self.env, 'restart')

    def new_game(self, from_random_game=False):
        """environment.py:28-33: reset the envs that are out of lives, then one no-op step."""
        dead = self.lives == 0
        if self._screen is None or bool(dead.any()):
            self._screen = self.env.reset(dead)
        self._step(self._noop())
        self.render()
        return self.screen, 0, 0, self.terminal

    def new_random_game(self, mask=None):
        """environment.py:35-40.  With a backend that can restart single envs every env in
        ``mask`` (default: all) draws its OWN ``random.randint(0, random_start - 1)`` no-op count
        (environment.py:37, in env order) -- one reference worker per env.  Backends that only step
        in lockstep (SyntheticAtari: frames come from a fixed pool, there is no emulator state to
        restart) share one draw, and a masked restart is a no-op for them."""
        if self.per_env_restart:
            m = [True] * self.num_envs if mask is None else [bool(v) for v in mask.cpu().tolist()]
            if self._screen is None:
                m = [True] * self.num_envs
            steps = [random.randint(0, self.random_start - 1) if v else 0 for v in m]
            if any(m):
                self._screen, self.reward, self.terminal = self.env.restart(m, steps)
            self.render()
            return self.screen, 0, 0, self.terminal
        if mask is not None:
            return self._screen, 0, 0, self.terminal
        self.new_game(True)
        for _ in range(random.randint(0, self.random_start - 1)):
            self._step(self._noop())
        self.render()
        return self.screen, 0, 0, self.terminal

    def _noop(self):
        return torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)

    def _step(self, action):
        self._screen, self.reward, self.terminal, _ = self.env.step(action)

    def _random_step(self):
        self._step(self.env.action_space.sample())

    @property
    def frames(self):
        """Raw u8 [B,210,160,3] frames of the last step: feed to History.add for the fused path."""
        return self._screen

    @property
    def screen(self):
        """environment.py:49-53 on the device (K1, bit-exact): u8 [B,84,84]."""
        _cabi.call(self._push, _cabi.ptr(self._screen), _cabi.ptr(self._scratch),
                   self.num_envs, 4, 0, 1, _cabi.stream_ptr())
        return from_blocked(self._scratch[:, 0]).contiguous()        # the ring layout is 4x4-blocked

    @property
    def action_size(self):
        return self.env.action_space.n

    @property
    def lives(self):
        return self.env.ale.lives()

    @property
    def state(self):
        return self.screen, self.reward, self.terminal

    def render(self):
        if self.display:
            self.env.render()

    def after_act(self, action):
        self.render()


class GymEnvironment(Environment):
    def act(self, action, is_training=True, fused=False):
        """environment.py:78-96, per env.  ``fused=True`` returns the raw frames instead of the
        84x84 screen so that History.add can run preprocess+push as one kernel."""
        start_lives = self.lives.clone()
        if self.action_repeat == 1:
            # config.py:50 default: the loop body once; the life-loss rule of environment.py:86-88
            # is one kernel (arl_act_update) writing into two alternating result buffers
            self._step(action)
            k = self._act_k = getattr(self, '_act_k', 0) ^ 1
            if not hasattr(self, '_act_out'):
                self._act_out = [(torch.empty(self.num_envs, device=self.device),
                                  torch.empty(self.num_envs, dtype=torch.bool, device=self.device))
                                 for _ in range(2)]
            rew, term = self._act_out[k]
            # converted copies are bound to names that outlive the launch (a temporary's memory
            # could be handed to the next allocation before the kernel has read it)
            rew_in, term_in, lives_now = self.reward.float(), self.terminal.bool(), self.lives.int()
            _cabi.call("arl_act_update", _cabi.ptr(rew_in), _cabi.ptr(term_in),
                       _cabi.ptr(start_lives), _cabi.ptr(lives_now), 1 if is_training else 0,
                       _cabi.ptr(rew), _cabi.ptr(term), self.num_envs, _cabi.stream_ptr())
            self._act_in = (rew_in, term_in, lives_now, start_lives)
            self.reward, self.terminal = rew, term
            self.after_act(action)
            if fused:
                return self.frames, self.reward, self.terminal
            return self.state
        cumulated = torch.zeros(self.num_envs, device=self.device)
        done = torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)
        for k in range(self.action_repeat):
            self._step(action)
            live = ~done
            cumulated = cumulated + torch.where(live, self.reward, torch.zeros_like(self.reward))
            term = self.terminal
            if is_training:
                lost = (start_lives > self.lives) & live
                cumulated = cumulated - lost.float()              # environment.py:86-88
                term = term | lost
            done = done | term
            if self.action_repeat > 1 and bool(done.all()):
                break
        self.reward, self.terminal = cumulated, done
        self.after_act(action)
        if fused:
            return self.frames, self.reward, self.terminal
        return self.state


class SimpleGymEnvironment(Environment):
    def act(self, action, is_training=True, fused=False):
        """environment.py:102-106."""
        self._step(action)
        self.after_act(action)
        if fused:
            return self.frames, self.reward, self.terminal
        return self.state
