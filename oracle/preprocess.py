"""CPU restatement of the reference's frame preprocessing and 4-frame stack.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows, line by line:
  * /root/reference/src/environment.py:49-53  ``Environment.screen``:
        y = 0.2126*R + 0.7152*G + 0.0722*B      (numpy: u8 * python float -> float64,
                                                  evaluated left to right)
        y = y.astype(np.uint8)                   (C truncation)
        return imresize(y, self.dims)
  * /root/reference/src/environment.py:5-12   ``imresize`` selection: the
    ``scipy.misc.imresize`` branch no longer exists in any current SciPy, so the
    executed branch is ``cv2.resize(y, (84, 84))`` = INTER_LINEAR.  The 2016
    first-choice branch (PIL BILINEAR through scipy.misc) is restated too as
    the optional ``resize='pil'`` mode.
  * /root/reference/src/history.py:3-27        ``History``.

The arithmetic of ``cv2.resize`` / PIL lives in third-party libraries that are
not vendored under /root/reference (README.md:10-14 names them without
versions).  The restatements below are of their published 8-bit algorithms and
are pinned in tests/ against (a) the executed reference expression
(tests/golden/preprocess_golden.npz, made by oracle/make_golden.py) and
(b) cv2 4.13 / Pillow 12.2 directly when those libraries are importable.
"""
import numpy as np

LUMA_R, LUMA_G, LUMA_B = 0.2126, 0.7152, 0.0722       # environment.py:51
INTER_RESIZE_COEF_BITS = 11
INTER_RESIZE_COEF_SCALE = 1 << INTER_RESIZE_COEF_BITS   # 2048, cv2 imgproc


def luma_truncate(frame):
    """environment.py:51-52.  frame u8 [..., H, W, 3] -> u8 [..., H, W].

    float64, separate roundings, left-to-right, then truncation.  White -> 254.
    """
    f = np.asarray(frame)
    assert f.dtype == np.uint8 and f.shape[-1] == 3
    y = LUMA_R * f[..., 0] + LUMA_G * f[..., 1] + LUMA_B * f[..., 2]
    return y.astype(np.uint8)


def luma_int_floor(frame):
    """Exact floor((2126R+7152G+722B)/10000) -- NOT the reference value; it is
    what the CUDA kernel computes before applying its correction bitmap.  Used
    by tests to enumerate the triples where the two differ (774 of 2^24)."""
    f = np.asarray(frame).astype(np.int64)
    s = 2126 * f[..., 0] + 7152 * f[..., 1] + 722 * f[..., 2]
    return (s // 10000).astype(np.uint8)


def cv2_linear_taps(src, dst):
    """Tap table of cv2's 8-bit INTER_LINEAR along one axis.

    Returns (idx int32[dst], c0 int16[dst], c1 int16[dst]) such that the
    horizontal pass is  S[idx]*c0 + S[min(idx+1, src-1)]*c1.
    cv2: fx = (float)((dx+0.5)*scale - 0.5); sx = floor(fx); fx -= sx;
         clamp; coeffs = saturate_cast<short>((1-fx, fx) * 2048)  (round half even).
    """
    scale = 1.0 / (float(dst) / float(src))                 # double, as cv2 does
    idx = np.zeros(dst, np.int32)
    c0 = np.zeros(dst, np.int16)
    c1 = np.zeros(dst, np.int16)
    for d in range(dst):
        fx = np.float32((d + 0.5) * scale - 0.5)
        sx = int(np.floor(fx))
        fx = np.float32(fx - np.float32(sx))
        if sx < 0:
            fx, sx = np.float32(0), 0
        if sx >= src - 1:
            fx, sx = np.float32(0), src - 1
        a1 = np.float32(fx) * np.float32(INTER_RESIZE_COEF_SCALE)
        a0 = (np.float32(1.0) - np.float32(fx)) * np.float32(INTER_RESIZE_COEF_SCALE)
        idx[d] = sx
        c0[d] = int(np.rint(a0))                           # rint = half-to-even = cvRound
        c1[d] = int(np.rint(a1))
    return idx, c0, c1


def cv2_resize_linear_u8(y, dsize):
    """cv2.resize(y, dsize) for 2-D u8, INTER_LINEAR, restated.

    dsize = (width, height) like cv2.  Horizontal pass in int32, vertical pass
    ((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2  (cv2 VResizeLinear<uchar>).
    """
    y = np.asarray(y)
    assert y.dtype == np.uint8 and y.ndim >= 2
    sh, sw = y.shape[-2:]
    dw, dh = dsize
    xi, xc0, xc1 = cv2_linear_taps(sw, dw)
    yi, yc0, yc1 = cv2_linear_taps(sh, dh)
    xi1 = np.minimum(xi + 1, sw - 1)
    yi1 = np.minimum(yi + 1, sh - 1)
    s = y.astype(np.int32)
    rows = s[..., :, xi] * xc0.astype(np.int32) + s[..., :, xi1] * xc1.astype(np.int32)
    r0 = rows[..., yi, :]
    r1 = rows[..., yi1, :]
    b0 = yc0.astype(np.int32)[:, None]
    b1 = yc1.astype(np.int32)[:, None]
    out = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def _pil_bilinear_coeffs(src, dst):
    """Pillow precompute_coeffs for BILINEAR (support 1.0), 8-bit normalisation."""
    scale = float(src) / float(dst)
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    taps = []
    for i in range(dst):
        center = (i + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > src:
            xmax = src
        n = xmax - xmin
        w = np.zeros(n, np.float64)
        for j in range(n):
            x = (j + xmin - center + 0.5) / filterscale
            x = -x if x < 0 else x
            w[j] = 1.0 - x if x < 1.0 else 0.0
        tot = w.sum()
        if tot != 0.0:
            w = w / tot
        k = np.array([int(0.5 + v * (1 << 22)) if v >= 0 else int(-0.5 + v * (1 << 22))
                      for v in w], np.int64)
        taps.append((xmin, k))
    return taps


def pil_resize_bilinear_u8(y, dsize):
    """PIL Image.resize(dsize, BILINEAR) on mode 'L' (scipy.misc.imresize branch,
    environment.py:6-7), restated: horizontal pass to u8, then vertical pass."""
    y = np.asarray(y)
    assert y.dtype == np.uint8 and y.ndim == 2
    sh, sw = y.shape
    dw, dh = dsize

    def clip8(v):
        return np.clip(v >> 22, 0, 255).astype(np.uint8)

    tmp = np.zeros((sh, dw), np.uint8)
    src = y.astype(np.int64)
    for i, (xmin, k) in enumerate(_pil_bilinear_coeffs(sw, dw)):
        acc = (1 << 21) + (src[:, xmin:xmin + len(k)] * k[None, :]).sum(axis=1)
        tmp[:, i] = clip8(acc)
    out = np.zeros((dh, dw), np.uint8)
    src = tmp.astype(np.int64)
    for i, (ymin, k) in enumerate(_pil_bilinear_coeffs(sh, dh)):
        acc = (1 << 21) + (src[ymin:ymin + len(k), :] * k[:, None]).sum(axis=0)
        out[i, :] = clip8(acc)
    return out


def screen(frame, dims=(84, 84), resize="cv2"):
    """``Environment.screen`` (environment.py:49-53) for one frame or a batch.

    frame u8 [..., 210, 160, 3] -> u8 [..., dims[1], dims[0]].
    """
    y = luma_truncate(frame)
    if resize == "cv2":
        return cv2_resize_linear_u8(y, dims)
    if resize == "pil":
        if y.ndim == 2:
            return pil_resize_bilinear_u8(y, dims)
        flat = y.reshape((-1,) + y.shape[-2:])
        out = np.stack([pil_resize_bilinear_u8(f, dims) for f in flat])
        return out.reshape(y.shape[:-2] + out.shape[-2:])
    raise ValueError("unknown resize mode: %s" % resize)


class History:
    """history.py:3-27 restated (NHWC/NCHW ``get``; ``add`` shifts left)."""

    def __init__(self, history_length=4, screen_height=84, screen_width=84, cnn_format="NHWC"):
        self.cnn_format = cnn_format
        self.history = np.zeros([history_length, screen_height, screen_width], np.float32)

    def add(self, screen_):
        self.history[:-1] = self.history[1:]          # history.py:14
        self.history[-1] = screen_                      # history.py:15

    def reset(self):
        self.history *= 0                              # history.py:18

    def get(self):
        if self.cnn_format == "NHWC":                  # history.py:21-22
            return np.transpose(self.history, (1, 2, 0))
        return self.history

    def copy(self):
        return self.get().copy()                       # history.py:27
