"""ORACLE (test infrastructure, not the product): CPU restatement of the reference's frame ingest, SURVEY.md section 8(f)
row 1 - the step in front of the Swin encoder that a future uint8-ingest kernel has to reproduce.

Reference: `utils/dataset.py:47-69` `from_image_to_embedding_no_IncepRes`:
    im = cv2.imread(path)                                    # uint8 H x W x 3, B,G,R order
    if H > 224: im = cv2.resize(im, (224,224), INTER_AREA)   # decided on the HEIGHT only; always to a square
    if H < 224: im = cv2.resize(im, (224,224), INTER_CUBIC)
    im = Image.fromarray(im, mode='RGB')                     # the BGR bytes are *labelled* RGB: no channel swap happens
    x = Normalize(.5,.5)(ToTensor()(im))                     # CHW float32, (v/255 - 0.5)/0.5

Third-party arithmetic: OpenCV (`opencv-python`, UNPINNED in requirements.txt:2; checked here against 4.13.0). The resize is
restated from OpenCV's own (non-IPP) code path, modules/imgproc/src/resize.cpp:
  * INTER_CUBIC on 8U: Keys cubic with A = -0.75 at fx = (dx + 0.5) * scale - 0.5 (scale = 1 / (dst / src), float32), taps clamped
    to the image, coefficients rounded to int16 at 2^11, exact int32 horizontal pass, vertical pass in float32 FMA order
    S0*b0 + (S1*b1 + (S2*b2 + S3*b3)) with b = coef / 2^22 and round-half-even (VResizeCubicVec_32s8u);
  * INTER_AREA: integer ratios = block mean with round-half-up ((sum + area/2) / area); other ratios = separable float32
    weights of the covered source cells (computeResizeAreaTab), round-half-even.
Pinned (tests/test_frame_ingest.py): bit-exact against cv2 for INTER_AREA and for integer-ratio INTER_CUBIC (the 112 -> 224 case
of BASELINE.json) on any build; for other INTER_CUBIC ratios bit-exact against cv2 with IPP switched off
(`cv2.ipp.setUseIPP(False)`), while IPP-enabled wheels differ from OpenCV's own code by 1 LSB on ~4-5 % of the pixels.
"""
from __future__ import annotations

import numpy as np

SWIN_IMG_SIZE = 224   # utils/dataset.py:20


def _cubic_taps(ssize: int, dsize: int):
    """Source indices (dsize, 4) and int coefficients at 2^11 (dsize, 4) of OpenCV's 8U bicubic resize."""
    f32 = np.float32
    scale = 1.0 / (dsize / ssize)
    d = np.arange(dsize, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    x = (f - s.astype(np.float32)).astype(np.float32)
    A, one = f32(-0.75), f32(1)
    c0 = ((A * (x + one) - f32(5) * A) * (x + one) + f32(8) * A) * (x + one) - f32(4) * A
    c1 = ((A + f32(2)) * x - (A + f32(3))) * x * x + one
    c2 = ((A + f32(2)) * (one - x) - (A + f32(3))) * (one - x) * (one - x) + one
    c3 = one - c0 - c1 - c2
    coef = np.rint(np.stack([c0, c1, c2, c3], 1).astype(np.float32) * f32(2048)).astype(np.int64)
    idx = np.clip(s[:, None] + np.arange(-1, 3)[None, :], 0, ssize - 1)
    return idx, coef


def resize_cubic_u8(img: np.ndarray, dsize: int = SWIN_IMG_SIZE) -> np.ndarray:
    """cv2.resize(img, (dsize, dsize), interpolation=cv2.INTER_CUBIC) for uint8 H x W x C."""
    assert img.dtype == np.uint8 and img.ndim == 3
    H, W, _ = img.shape
    xi, xc = _cubic_taps(W, dsize)
    yi, yc = _cubic_taps(H, dsize)
    src = img.astype(np.int64)
    hor = np.zeros((H, dsize, img.shape[2]), dtype=np.int64)
    for k in range(4):
        hor += src[:, xi[:, k], :] * xc[:, k][None, :, None]
    b = (yc.astype(np.float32) * np.float32(1.0 / (2048.0 * 2048.0))).astype(np.float32)
    S = [hor[yi[:, k]].astype(np.float32) for k in range(4)]
    bk = [b[:, k][:, None, None].astype(np.float64) for k in range(4)]

    def fma(a, w, c):   # float32 fused multiply-add: the product of two float32 is exact in float64
        return (a.astype(np.float64) * w + c.astype(np.float64)).astype(np.float32)

    r = (S[3] * b[:, 3][:, None, None]).astype(np.float32)
    r = fma(S[2], bk[2], r)
    r = fma(S[1], bk[1], r)
    r = fma(S[0], bk[0], r)
    return np.clip(np.rint(r), 0, 255).astype(np.uint8)


def _area_tab(ssize: int, dsize: int):
    scale = ssize / dsize
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1 = int(np.ceil(fsx1))
        sx2 = min(int(np.floor(fsx2)), ssize - 1)
        sx1 = min(sx1, sx2)
        ent = []
        if sx1 - fsx1 > 1e-3:
            ent.append((sx1 - 1, (sx1 - fsx1) / cell))
        for sx in range(sx1, sx2):
            ent.append((sx, 1.0 / cell))
        if fsx2 - sx2 > 1e-3:
            ent.append((sx2, min(min(fsx2 - sx2, 1.0), cell) / cell))
        tab.append(ent)
    return tab


def resize_area_u8(img: np.ndarray, dsize: int = SWIN_IMG_SIZE) -> np.ndarray:
    """cv2.resize(img, (dsize, dsize), interpolation=cv2.INTER_AREA) for uint8 H x W x C, shrinking."""
    assert img.dtype == np.uint8 and img.ndim == 3
    H, W, C = img.shape
    if H % dsize == 0 and W % dsize == 0:
        ky, kx = H // dsize, W // dsize
        s = img.reshape(dsize, ky, dsize, kx, C).astype(np.int64).sum(axis=(1, 3))
        area = ky * kx
        return ((s + area // 2) // area).astype(np.uint8)
    tx, ty = _area_tab(W, dsize), _area_tab(H, dsize)
    src = img.astype(np.float32)
    hor = np.zeros((H, dsize, C), dtype=np.float32)
    for dx, ent in enumerate(tx):
        for sx, w in ent:
            hor[:, dx, :] += src[:, sx, :] * np.float32(w)
    out = np.zeros((dsize, dsize, C), dtype=np.float32)
    for dy, ent in enumerate(ty):
        for sy, w in ent:
            out[dy] += hor[sy] * np.float32(w)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def ingest_frame(bgr_u8: np.ndarray) -> np.ndarray:
    """One decoded crop (uint8 H x W x 3 as cv2.imread returns it) -> float32 (3, 224, 224) as the Swin encoder receives it
    (utils/dataset.py:52-65): channel c of the output is channel c of the input (B, G, R), values in [-1, 1]."""
    im = bgr_u8
    if im.shape[0] > SWIN_IMG_SIZE:
        im = resize_area_u8(im)
    if im.shape[0] < SWIN_IMG_SIZE:
        im = resize_cubic_u8(im)
    if im.shape[:2] != (SWIN_IMG_SIZE, SWIN_IMG_SIZE):
        raise ValueError(f"a crop of height 224 must also be 224 wide (got {im.shape}); the reference would fail in X[i,:] = x")
    x = im.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)       # ToTensor
    return ((x - np.float32(0.5)) / np.float32(0.5)).astype(np.float32)    # Normalize(0.5, 0.5)


def ingest_frames(crops) -> np.ndarray:
    return np.stack([ingest_frame(c) for c in crops], 0)
