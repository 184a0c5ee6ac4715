"""History: the 4-frame stack (reference src/history.py:3-27), batched over envs and kept
on the device as a ring of u8 planes so the T+1 overlapping stacks of a rollout are windows
into one buffer instead of T+1 float32 copies (reference agent.py:157 copies 113 KB/step).

    ring u8 [num_envs, ring_slots, 7056];  ``head`` = slot of the newest frame.
    stack at offset t (t = 0 newest state) = slots head-3-t .. head-t, oldest first --
    the channel order of History.get() (history.py:20-24).

Each 84x84 plane is stored in 4x4 blocks (space-to-depth): byte (y, x) sits at
((y//4)*21 + x//4)*16 + (y%4)*4 + x%4.  An 8x8 stride-4 conv window is then four 16-byte
vectors, which is how the conv1 kernels read it; ``get``/``copy``/``planes`` return the
row-major layout of the reference.
"""
import torch

from .. import _cabi

SCREEN = 84
FRAME_SHAPE = (210, 160, 3)


def to_blocked(screens):
    """[..., 84, 84] row-major -> the ring's 4x4-block order (same shape)."""
    lead = screens.shape[:-2]
    x = screens.reshape(lead + (21, 4, 21, 4))
    n = len(lead)
    return x.permute(*range(n), n, n + 2, n + 1, n + 3).reshape(lead + (SCREEN, SCREEN))


def from_blocked(planes):
    """Inverse of to_blocked."""
    lead = planes.shape[:-2]
    x = planes.reshape(lead + (21, 21, 4, 4))
    n = len(lead)
    return x.permute(*range(n), n, n + 2, n + 1, n + 3).reshape(lead + (SCREEN, SCREEN))


def default_ring_slots(t_max):
    """Slots of the ring: at least t_max + 4 (the T+1 overlapping stacks of a rollout need T+4
    planes).  The head advances by t_max per cycle, so the slot pattern of a cycle repeats every
    R / gcd(t_max, R) cycles; a slightly larger R with a short period (<= 3 cycles) keeps the
    number of distinct CUDA graphs of the loop small (t_max 5: 10 slots, period 2, instead of 9
    slots with period 9; t_max 20: 30 slots, period 3, instead of 24 with period 6)."""
    import math
    for r in range(t_max + 4, 2 * t_max + 9):
        if r // math.gcd(t_max, r) <= 3:
            return r
    return t_max + 4


class History(object):
    def __init__(self, config, num_envs=None, ring_slots=None, device=None):
        self.cnn_format = getattr(config, 'cnn_format', 'NHWC')
        self.history_length = config.history_length
        assert self.history_length == 4, "the kernels are built for history_length 4"
        assert (config.screen_height, config.screen_width) == (SCREEN, SCREEN)
        self.num_envs = int(num_envs if num_envs is not None else getattr(config, 'num_envs', 1))
        t_max = int(getattr(config, 't_max', 5))
        self.ring_slots = int(ring_slots if ring_slots is not None else default_ring_slots(t_max))
        if self.ring_slots < 4:
            raise ValueError("ring_slots must be >= 4")
        self.device = torch.device(device if device is not None else 'cuda')
        _cabi.init(self.device)
        self.ring = torch.zeros(self.num_envs, self.ring_slots, SCREEN, SCREEN,
                                dtype=torch.uint8, device=self.device)   # history.py:10-11
        self.head = self.ring_slots - 1
        self.timer = None                     # bench.py: callable that brackets K1 with an event pair
        self.timer_in_graph = False           # ... as event-record nodes INSIDE the captured graph
        resize = getattr(config, 'resize', 'cv2')                # environment.py:5-12 branch
        if resize not in ('cv2', 'pil'):
            raise NotImplementedError("resize=%r" % (resize,))
        self._push = "arl_preprocess_push" if resize == 'cv2' else "arl_preprocess_push_pil"

    # -- reference API ------------------------------------------------------------------
    def add(self, screen, replicate=1):
        """history.py:13-15.  ``screen`` is either raw frames u8 [B,210,160,3] (fused
        Environment.screen + add: one kernel, K1) or ready 84x84 screens u8 [B,84,84]."""
        new_head = (self.head + 1) % self.ring_slots
        self.push_into(screen, new_head, replicate)
        self.head = (new_head + replicate - 1) % self.ring_slots

    def push_into(self, screen, slot, replicate=1):
        """The device work of ``add`` for an explicit ring slot; no host state changes (the form a
        captured CUDA graph can hold)."""
        if tuple(screen.shape[1:]) == FRAME_SHAPE:
            if screen.dtype != torch.uint8:
                raise TypeError("frames must be uint8")
            args = (self._push, _cabi.ptr(screen), _cabi.ptr(self.ring),
                    self.num_envs, self.ring_slots, slot, int(replicate), _cabi.stream_ptr())
            if self.timer is not None:
                self.timer(*args)                    # bench.py: event pair around K1
            else:
                _cabi.call(*args)
        elif tuple(screen.shape[1:]) == (SCREEN, SCREEN):
            for r in range(replicate):
                self.ring[:, (slot + r) % self.ring_slots].copy_(to_blocked(screen))
        else:
            raise ValueError("expected [B,210,160,3] frames or [B,84,84] screens, got %s"
                             % (tuple(screen.shape),))

    def reset(self):
        """history.py:17-18."""
        _cabi.call("arl_history_reset", _cabi.ptr(self.ring), self.num_envs, self.ring_slots,
                   _cabi.stream_ptr())

    def get(self, back=0, dtype=torch.float32):
        """history.py:20-24: [B,84,84,4] (NHWC) or [B,4,84,84]; float32 like the reference."""
        first = self.first_slot(back)
        out = torch.empty(self.num_envs, SCREEN, SCREEN, 4, dtype=dtype, device=self.device)
        _cabi.call("arl_history_get", _cabi.ptr(self.ring), _cabi.ptr(out),
                   1 if dtype == torch.uint8 else 0, self.num_envs, self.ring_slots, first,
                   _cabi.stream_ptr())
        if self.cnn_format == 'NHWC':
            return out
        return out.permute(0, 3, 1, 2)

    def copy(self):
        """history.py:26-27."""
        return self.get().contiguous()

    def planes(self, slot=None):
        """Row-major u8 view-copy of the ring planes: [B, ring_slots, 84, 84], or [B, 84, 84] of
        one slot (the ring itself is 4x4-blocked)."""
        src = self.ring if slot is None else self.ring[:, slot]
        return from_blocked(src).contiguous()

    # -- ring bookkeeping used by Network/Agent --------------------------------------------
    def first_slot(self, back=0):
        """Slot of the OLDEST plane of the stack that ended ``back`` pushes ago."""
        return (self.head - 3 - back) % self.ring_slots
