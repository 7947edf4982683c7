"""Deterministic synthetic stereo pairs.

Left image: per channel, a sigma=2 gaussian-blurred N(0,1) field rescaled to unit std, mixed with
white noise and quantised to 8 bits; ground-truth disparity is constant on 40-row bands,
d_gt(y) = 1 + ((y // 40) * 17) mod (D - 2); the right image is the left one shifted by d_gt(y)
(cyclically), so the left map has a known interior answer.
"""
import numpy as np

SEEDS = {(1280, 720, 64): 1280720064, (1920, 1080, 128): 19201080128, (1920, 1080, 256): 19201080256}


def _blur_rows_cols(a, sigma=2.0):
    r = int(4 * sigma + 0.5)
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2)
    k /= k.sum()
    a = np.apply_along_axis(lambda v: np.convolve(np.pad(v, r, mode="reflect"), k, mode="valid"), 1, a)
    a = np.apply_along_axis(lambda v: np.convolve(np.pad(v, r, mode="reflect"), k, mode="valid"), 0, a)
    return a


def stereo_pair_u8(W, H, D, seed=None):
    """-> (left_u8[H,W,3], right_u8[H,W,3], d_gt[H]) ; BGR order is irrelevant for synthetic data."""
    if seed is None:
        seed = SEEDS.get((W, H, D), W * 1000003 + H * 1009 + D)
    rng = np.random.default_rng(seed)
    left = np.empty((H, W, 3), np.uint8)
    for c in range(3):
        tex = _blur_rows_cols(rng.standard_normal((H, W)))
        tex /= tex.std()
        v = 128.0 + 48.0 * tex + 16.0 * rng.standard_normal((H, W))
        left[:, :, c] = np.clip(np.rint(v), 0, 255).astype(np.uint8)
    d_gt = 1 + ((np.arange(H) // 40) * 17) % (D - 2)
    cols = (np.arange(W)[None, :] + d_gt[:, None]) % W
    right = np.take_along_axis(left, cols[:, :, None].repeat(3, axis=2), axis=1)
    return left, right, d_gt


def to_f32(img_u8):
    """StereoMatch.cpp:193-197: convertTo(CV_32F, 1/255.0f)."""
    return img_u8.astype(np.float32) * np.float32(1 / np.float32(255.0))


def stereo_pair_f32(W, H, D, seed=None):
    l, r, d = stereo_pair_u8(W, H, D, seed)
    return to_f32(l), to_f32(r), d
