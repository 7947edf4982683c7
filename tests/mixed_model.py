"""CPU restatement (numpy, fp32) of PSM_CVF_MIXED's second stage -- TEST INFRASTRUCTURE.

PSM_CVF_MIXED keeps the four stage-1 box sums of GuidedFilter_cv (reference src/CVF.cpp:81-89) in
fp64, so a and b (CVF.cpp:131-155) stay bit-exact, and evaluates the four stage-2 box filters
(CVF.cpp:157-160: box(b), box(a_c)) in fp32 with a fixed summation tree:

  vertical    V(y,x)  = ((X[r0]+X[r1]) + (X[r2]+X[r3])) + ((X[r4]+X[r5]) + (X[r6]+X[r7])),
              r_k = reflect101(y-4+k)                       (window rows y-4 .. y+3, anchor 4)
  horizontal  over ALIGNED 4-column groups g (columns 4g .. 4g+3, reflected outside the image),
              c_i = V(y, 4g+i):  P2=c0+c1, P3=P2+c2, T=P3+c3, S2=c2+c3, S3=c1+S2
              output column 4g+j uses group g-1 (suffix), g (total), g+1 (prefix):
                j=0: T[g-1] + T[g]
                j=1: (S3[g-1] + T[g]) + c0[g+1]
                j=2: (S2[g-1] + T[g]) + P2[g+1]
                j=3: (c3[g-1] + T[g]) + P3[g+1]
  scale       * (1/64)  (exact)
  q = box(b) + box(a0)*I0 + box(a1)*I1 + box(a2)*I2, separate fp32 roundings in the reference's
      order (CVF.cpp:157-163)

The GPU kernel (primestereomatch_b200/csrc/psm_cvf_stream.cuh, S2M = kS2Mixed) must reproduce this
model bit for bit; the model itself is compared with the exact oracle for the north-star tolerance
(1e-4 on a/b -- here a,b are bit-exact -- and +-1 disparity level)."""
import numpy as np

f32 = np.float32


def _refl(i, n):
    i = np.abs(np.asarray(i))
    if n == 1:
        return np.zeros_like(i)
    return np.where(i >= n, 2 * (n - 1) - i, i)


def box8_mixed(X):
    """fp32 8x8 box mean of one plane with the MIXED summation tree."""
    X = np.ascontiguousarray(X, dtype=f32)
    H, W = X.shape
    G = (W + 3) // 4 + 2                       # groups -1 .. ceil(W/4)
    cols = _refl(np.arange(-4, 4 * (G - 1)), W)
    Xp = X[:, cols]
    r = [Xp[_refl(np.arange(H) - 4 + k, H)] for k in range(8)]
    V = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    assert V.dtype == f32
    c = V.reshape(H, G, 4)
    c0, c1, c2, c3 = c[..., 0], c[..., 1], c[..., 2], c[..., 3]
    P2 = c0 + c1
    P3 = P2 + c2
    T = P3 + c3
    S2 = c2 + c3
    S3 = c1 + S2
    n = G - 2
    a, b, d = slice(0, n), slice(1, n + 1), slice(2, n + 2)
    h = np.empty((H, n, 4), f32)
    h[..., 0] = T[:, a] + T[:, b]
    h[..., 1] = (S3[:, a] + T[:, b]) + c0[:, d]
    h[..., 2] = (S2[:, a] + T[:, b]) + P2[:, d]
    h[..., 3] = (c3[:, a] + T[:, b]) + P3[:, d]
    return (h.reshape(H, 4 * n)[:, :W] * f32(1 / 64)).astype(f32)


def q_mixed(rgb, a, b):
    """q of one slice from the exact coefficient planes a[3], b and the guide channels rgb[3]."""
    q = box8_mixed(b)
    for c in range(3):
        q = q + box8_mixed(a[c]) * rgb[c]
    return q.astype(f32)


def cost_filter_mixed(oracle, img, vol):
    """MIXED-filtered volume of one view: exact a,b from the oracle, stage 2 by the model."""
    rgb, mean, var = oracle.cvf_preprocess(img)
    out = np.empty_like(vol)
    for d in range(vol.shape[0]):
        _, a, b = oracle.guided_filter(rgb, mean, var, vol[d], want_ab=True)
        out[d] = q_mixed(rgb, a, b)
    return out
