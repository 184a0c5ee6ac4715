"""CPU: the oracle's preprocessing/history restatement against (a) the golden vectors made by
EXECUTING the reference (oracle/make_golden.py) and (b) cv2 / Pillow directly."""
import numpy as np
import pytest

from oracle import preprocess as P
from util import golden_seq_frames


def test_screen_matches_executed_reference(golden):
    for name in golden["names"]:
        got = P.screen(golden["frame_" + str(name)])
        assert got.dtype == np.uint8 and got.shape == (84, 84)
        assert np.array_equal(got, golden["screen_" + str(name)]), name


def test_white_is_254_and_luma_edge_set(golden):
    assert int(golden["screen_white"].min()) == 254 and int(golden["screen_white"].max()) == 254
    tri, y = golden["luma_triples"], golden["luma_triples_y"]
    assert len(tri) == 3384
    assert np.array_equal(P.luma_truncate(tri), y)
    assert int((P.luma_int_floor(tri) != y).sum()) == 774
    assert np.all(P.luma_int_floor(tri).astype(int) - y.astype(int) >= 0)


def test_all_greys_and_random_triples_int_floor_agrees_off_the_edge_set():
    rng = np.random.default_rng(0)
    t = rng.integers(0, 256, (200000, 3), dtype=np.uint8)
    s = 2126 * t[:, 0].astype(int) + 7152 * t[:, 1].astype(int) + 722 * t[:, 2].astype(int)
    off = s % 10000 != 0
    assert np.array_equal(P.luma_truncate(t)[off], P.luma_int_floor(t)[off])
    g = np.repeat(np.arange(256, dtype=np.uint8)[:, None], 3, axis=1)
    assert P.luma_truncate(g)[255] == 254


def test_tap_tables_match_survey():
    xi, c0, c1 = P.cv2_linear_taps(160, 84)
    assert list(xi[:12]) == [0, 2, 4, 6, 8, 9, 11, 13, 15, 17, 19, 21]
    assert (int(c0[0]), int(c1[0])) == (1122, 926) and (int(c0[5]), int(c1[5])) == (49, 1999)
    assert len(set(zip(c0.tolist(), c1.tolist()))) == 21
    yi, b0, b1 = P.cv2_linear_taps(210, 84)
    assert list(yi[:6]) == [0, 3, 5, 8, 10, 13] and int(yi[-1]) == 208
    assert set(zip(b0.tolist(), b1.tolist())) == {(512, 1536), (1536, 512)}
    used = set(yi.tolist()) | set((yi + 1).tolist())
    assert len(used) == 168 and all(r % 5 != 2 for r in used)


def test_resize_matches_cv2_live():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    for _ in range(4):
        y = rng.integers(0, 256, (210, 160), dtype=np.uint8)
        assert np.array_equal(P.cv2_resize_linear_u8(y, (84, 84)), cv2.resize(y, (84, 84)))
    y = rng.integers(0, 256, (3, 210, 160), dtype=np.uint8)        # batched restatement
    ref = np.stack([cv2.resize(v, (84, 84)) for v in y])
    assert np.array_equal(P.cv2_resize_linear_u8(y, (84, 84)), ref)


def test_pil_mode_matches_pillow_live():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(2)
    y = rng.integers(0, 256, (210, 160), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(y).resize((84, 84), Image.BILINEAR))
    assert np.array_equal(P.pil_resize_bilinear_u8(y, (84, 84)), ref)


def test_history_matches_executed_reference(golden):
    assert list(golden["hist_order"]) == [2.0, 3.0, 4.0, 5.0]       # channel 0 oldest .. 3 newest
    h = P.History()
    for k in range(6):
        h.add(np.full((84, 84), k, np.uint8))
    assert list(h.get()[0, 0, :]) == [2.0, 3.0, 4.0, 5.0]
    # the act()/add() sequence of the fixture
    frames = golden_seq_frames(golden)
    screens = [P.screen(frames[0])] + [P.screen(frames[1 + t]) for t in range(6)]
    assert np.array_equal(np.stack(screens), golden["seq_screens"])
    hist = P.History()
    for _ in range(4):
        hist.add(screens[0])
    stacks = [hist.copy()]
    for t in range(6):
        hist.add(screens[1 + t])
        stacks.append(hist.copy())
    assert np.array_equal(np.stack(stacks).astype(np.uint8), golden["seq_stacks"])
    # reference act(): life loss at step index 4 -> reward-1 and terminal (environment.py:86-88)
    assert list(golden["seq_rewards"]) == [1.0, 0.0, 5.0, -4.0, 0.0, 0.0]
    assert list(golden["seq_terminals"]) == [False, False, False, True, False, False]
