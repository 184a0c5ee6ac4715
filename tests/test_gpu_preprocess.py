"""GPU: K1 (Environment.screen + History.add) through the C-ABI, bit-exact against the golden
vectors made by executing the reference and against the oracle on seeded inputs."""
import numpy as np
import pytest
import torch

from oracle import preprocess as P
from util import golden_seq_frames, unblock

pytestmark = pytest.mark.gpu


def push(pkg, frames_np, ring_slots=4, slot=0, replicate=1, ring=None):
    dev = torch.device("cuda:0")
    frames = torch.as_tensor(np.ascontiguousarray(frames_np), device=dev)
    B = frames.shape[0]
    if ring is None:
        ring = torch.full((B, ring_slots, 84, 84), 7, dtype=torch.uint8, device=dev)
    pkg._cabi.call("arl_preprocess_push", pkg._cabi.ptr(frames), pkg._cabi.ptr(ring), B,
                   ring.shape[1], slot, replicate, pkg._cabi.stream_ptr())
    torch.cuda.synchronize()
    return unblock(ring)                                         # row-major screens


def test_golden_frames_bit_exact(pkg, cuda, golden):
    names = [str(n) for n in golden["names"]]
    frames = np.stack([golden["frame_" + n] for n in names])
    ring = push(pkg, frames).cpu().numpy()
    for i, n in enumerate(names):
        assert np.array_equal(ring[i, 0], golden["screen_" + n]), n
    assert (ring[:, 1:] == 7).all()                             # other slots untouched


def test_luma_edge_triples_all_3384(pkg, cuda, golden):
    """Every exact-integer luma triple (774 of them truncate one low in float64) in flat frames:
    a constant frame resizes to the constant, so the screen value IS the luma."""
    tri, y = golden["luma_triples"], golden["luma_triples_y"]
    for lo in range(0, len(tri), 512):
        t = tri[lo:lo + 512]
        frames = np.broadcast_to(t[:, None, None, :], (len(t), 210, 160, 3))
        ring = push(pkg, frames).cpu().numpy()
        assert np.array_equal(ring[:, 0, 0, 0], y[lo:lo + 512])
        assert (ring[:, 0] == ring[:, 0, :1, :1]).all()


def test_all_256_greys(pkg, cuda):
    g = np.arange(256, dtype=np.uint8)
    frames = np.broadcast_to(g[:, None, None, None], (256, 210, 160, 3))
    ring = push(pkg, frames).cpu().numpy()
    assert np.array_equal(ring[:, 0, 5, 5], P.luma_truncate(np.repeat(g[:, None], 3, 1)))
    assert ring[255, 0, 0, 0] == 254


@pytest.mark.parametrize("B", [1, 2, 147, 149, 300])
def test_random_batches_vs_oracle_ragged_sizes(pkg, cuda, B):
    rng = np.random.default_rng(100 + B)
    frames = rng.integers(0, 256, (B, 210, 160, 3), dtype=np.uint8)
    ring = push(pkg, frames, ring_slots=5, slot=3).cpu().numpy()
    assert np.array_equal(ring[:, 3], P.screen(frames))


def test_structured_frames_vs_oracle(pkg, cuda):
    rng = np.random.default_rng(5)
    pal = rng.integers(0, 256, (16, 3), dtype=np.uint8)
    pal[0] = 0
    idx = rng.integers(0, 16, (64, 21, 16))
    frames = pal[np.repeat(np.repeat(idx, 10, 1), 10, 2)]
    ring = push(pkg, frames).cpu().numpy()
    assert np.array_equal(ring[:, 0], P.screen(frames))


def test_replicate_and_ring_wrap(pkg, cuda):
    rng = np.random.default_rng(6)
    frames = rng.integers(0, 256, (3, 210, 160, 3), dtype=np.uint8)
    ring = push(pkg, frames, ring_slots=6, slot=4, replicate=4).cpu().numpy()   # slots 4,5,0,1
    ref = P.screen(frames)
    for s in (4, 5, 0, 1):
        assert np.array_equal(ring[:, s], ref)
    assert (ring[:, 2:4] == 7).all()


def test_empty_batch_and_bad_arguments(pkg, cuda):
    lib = pkg._cabi.load()
    ring = torch.zeros(1, 4, 84, 84, dtype=torch.uint8, device=cuda)
    fr = torch.zeros(1, 210, 160, 3, dtype=torch.uint8, device=cuda)
    st = pkg._cabi.stream_ptr()
    assert lib.arl_preprocess_push(fr.data_ptr(), ring.data_ptr(), 0, 4, 0, 1, st) == 0
    assert lib.arl_preprocess_push(fr.data_ptr(), ring.data_ptr(), 1, 4, 4, 1, st) == -1
    assert b"slot" in lib.arl_last_error()
    assert lib.arl_preprocess_push(fr.data_ptr(), ring.data_ptr(), 1, 3, 0, 1, st) == -1
    assert lib.arl_preprocess_push(None, ring.data_ptr(), 1, 4, 0, 1, st) == -1
    assert lib.arl_preprocess_push(fr.data_ptr() + 1, ring.data_ptr(), 1, 4, 0, 1, st) == -1


def test_history_sequence_matches_executed_reference(pkg, cuda, golden):
    """agent.py:37-38 + 156-157 through History/GymEnvironment-shaped calls: 4 copies of the first
    screen, then 6 pushes; every stack equals the reference's history.copy()."""
    cfg = pkg.config.get_config({"model": "m1", "num_envs": 2, "t_max": 5})
    hist = pkg.History(cfg, num_envs=2, device=cuda)
    frames = golden_seq_frames(golden)
    f = lambda i: torch.as_tensor(np.stack([frames[i], frames[(i + 3) % 10]]), device=cuda)
    hist.add(f(0), replicate=4)
    stacks = [hist.get().cpu().numpy()]
    for t in range(6):
        hist.add(f(1 + t))
        stacks.append(hist.get().cpu().numpy())
    got = np.stack(stacks)                                      # [7, 2, 84, 84, 4] float32
    assert got.dtype == np.float32
    assert np.array_equal(got[:, 0].astype(np.uint8), golden["seq_stacks"])
    assert np.array_equal(hist.get(dtype=torch.uint8).cpu().numpy()[0], golden["seq_stacks"][6])
    # ring indexing: stack ``back`` pushes ago
    assert np.array_equal(hist.get(back=2).cpu().numpy()[0].astype(np.uint8), golden["seq_stacks"][4])
    hist.reset()
    assert float(hist.get().abs().max()) == 0.0


def test_environment_screen_property(pkg, cuda, golden):
    cfg = pkg.config.get_config({"model": "m1", "num_envs": 4})
    env = pkg.GymEnvironment(cfg, device=cuda)
    scr, r, a, term = env.new_random_game()
    assert scr.shape == (4, 84, 84) and scr.dtype == torch.uint8
    assert np.array_equal(scr.cpu().numpy(), P.screen(env.frames.cpu().numpy()))
    s2, rew, term = env.act(torch.zeros(4, dtype=torch.int32, device=cuda))
    assert np.array_equal(s2.cpu().numpy(), P.screen(env.frames.cpu().numpy()))
    assert rew.shape == (4,) and term.dtype == torch.bool


def test_full_size_properties_4096_envs(pkg, cuda):
    """BASELINE config size.  Size-independent properties: (1) identical frames -> identical
    screens wherever they sit in the batch, (2) a random subset equals the oracle, (3) the
    screen of a constant frame is its luma."""
    B = 4096
    g = torch.Generator(device=cuda).manual_seed(11)
    base = torch.randint(0, 256, (64, 210, 160, 3), dtype=torch.uint8, device=cuda, generator=g)
    perm = torch.randint(0, 64, (B,), device=cuda, generator=g)
    frames = base[perm].contiguous()
    ring = torch.zeros(B, 9, 84, 84, dtype=torch.uint8, device=cuda)
    pkg._cabi.call("arl_preprocess_push", pkg._cabi.ptr(frames), pkg._cabi.ptr(ring), B, 9, 8, 1,
                   pkg._cabi.stream_ptr())
    torch.cuda.synchronize()
    out = unblock(ring[:, 8])
    first = torch.stack([out[(perm == k).nonzero()[0, 0]] for k in range(64)])
    assert bool((out == first[perm]).all())
    assert np.array_equal(first.cpu().numpy(), P.screen(base.cpu().numpy()))
    assert int(ring[:, :8].max()) == 0


def test_upload_frames_skips_only_unread_rows(pkg, cuda):
    """arl_upload_frames moves rows != 2 (mod 5) from pinned host memory; K1 on the uploaded
    buffer equals the oracle on the full host frames, and the skipped rows are left untouched."""
    B = 37
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, (B, 210, 160, 3), dtype=np.uint8)
    host = torch.empty((B, 210, 160, 3), dtype=torch.uint8, pin_memory=True)
    host.copy_(torch.as_tensor(frames))
    dev = torch.full((B, 210, 160, 3), 9, dtype=torch.uint8, device=cuda)
    pkg._cabi.call("arl_upload_frames", host.data_ptr(), pkg._cabi.ptr(dev), B,
                   pkg._cabi.stream_ptr())
    ring = torch.zeros(B, 4, 84, 84, dtype=torch.uint8, device=cuda)
    pkg._cabi.call("arl_preprocess_push", pkg._cabi.ptr(dev), pkg._cabi.ptr(ring), B, 4, 2, 1,
                   pkg._cabi.stream_ptr())
    torch.cuda.synchronize()
    got = dev.cpu().numpy()
    keep = np.arange(210) % 5 != 2
    assert np.array_equal(got[:, keep], frames[:, keep])
    assert (got[:, ~keep] == 9).all()
    assert np.array_equal(unblock(ring[:, 2]).cpu().numpy(), P.screen(frames))


def test_pil_branch_bit_exact(pkg, cuda):
    """environment.py:5-7: the scipy.misc.imresize branch = PIL BILINEAR.  arl_preprocess_push_pil
    against the oracle restatement and, when Pillow is importable, against Pillow itself."""
    rng = np.random.default_rng(21)
    frames = rng.integers(0, 256, (9, 210, 160, 3), dtype=np.uint8)
    frames[0] = 255; frames[1] = 0
    frames[2, :, ::2] = 0; frames[2, :, 1::2] = 255                 # vertical stripes: worst case for the filter
    pal = rng.integers(0, 256, (16, 3), dtype=np.uint8)
    frames[3] = pal[rng.integers(0, 16, (21, 16))].repeat(10, 0).repeat(10, 1)
    dev_frames = torch.as_tensor(frames, device=cuda)
    ring = torch.full((9, 5, 84, 84), 7, dtype=torch.uint8, device=cuda)
    pkg._cabi.call("arl_preprocess_push_pil", pkg._cabi.ptr(dev_frames), pkg._cabi.ptr(ring), 9, 5, 4, 2,
                   pkg._cabi.stream_ptr())                           # slots 4 and 0 (wrap)
    torch.cuda.synchronize()
    got = unblock(ring).cpu().numpy()
    ref = P.screen(frames, resize="pil")
    assert np.array_equal(got[:, 4], ref) and np.array_equal(got[:, 0], ref)
    assert (got[:, 1:4] == 7).all()
    try:
        from PIL import Image
    except ImportError:
        return
    y = P.luma_truncate(frames)
    for i in range(len(frames)):
        pil = np.asarray(Image.fromarray(y[i]).resize((84, 84), Image.BILINEAR))
        assert np.array_equal(got[i, 4], pil), i


def test_history_uses_the_configured_resize_branch(pkg, cuda):
    rng = np.random.default_rng(4)
    frames = rng.integers(0, 256, (5, 210, 160, 3), dtype=np.uint8)
    for mode in ("cv2", "pil"):
        cfg = pkg.config.get_config({"model": "m1", "num_envs": 5, "resize": mode})
        hist = pkg.History(cfg, num_envs=5, device=cuda)
        hist.add(torch.as_tensor(frames, device=cuda))
        torch.cuda.synchronize()
        assert np.array_equal(hist.planes(hist.head).cpu().numpy(), P.screen(frames, resize=mode))
