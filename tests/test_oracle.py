"""CPU-only: the plain-C oracle is pinned against the cv2-driven golden run."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN, read_png


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_primitives_match_cv2_fixtures(oracle, golden):
    P = np.load(os.path.join(GOLDEN, "primitives.npz"))
    assert np.array_equal(oracle.box8(P["plane"]), P["box"])
    assert np.array_equal(oracle.rgb2gray(P["img"]), P["gray"])
    assert np.array_equal(oracle.cvc_preprocess(P["img"]), P["sobel"])
    assert sha(oracle.box8(P["plane"])) == golden["primitives"]["box8_61x83"]


def test_primitives_match_live_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    for (h, w) in [(9, 11), (16, 16), (37, 130), (375, 450)]:
        p = (rng.standard_normal((h, w)) * np.exp(rng.uniform(-10, 4, (h, w)))).astype(np.float32)
        assert np.array_equal(oracle.box8(p), cv2.boxFilter(p, -1, (8, 8))), (h, w)
        img = rng.random((h, w, 3), dtype=np.float32)
        g = cv2.cvtColor(img, cv2.COLOR_RGB2GRAY)
        assert np.array_equal(oracle.rgb2gray(img), g)
        assert np.array_equal(oracle.sobel_x(g), cv2.Sobel(g, cv2.CV_32F, 1, 0, ksize=1))


def test_u8_scaling(oracle):
    u = np.arange(256, dtype=np.uint8)
    assert np.array_equal(oracle.u8_to_f32(u), u.astype(np.float32) * np.float32(1 / np.float32(255.0)))


@pytest.mark.parametrize("scene", ["Cones", "Teddy"])
def test_oracle_reproduces_golden(scene, golden, oracle_scene_results):
    g, r = golden["scenes"][scene], oracle_scene_results[scene]
    assert sha(r["lg"]) == g["lGrd"] and sha(r["rg"]) == g["rGrd"]
    assert sha(r["lraw"]) == g["lRaw"] and sha(r["rraw"]) == g["rRaw"]
    assert sha(r["lf"]) == g["lFilt"] and sha(r["rf"]) == g["rFilt"]
    assert sha(r["ld"]) == g["lDis"] and sha(r["rd"]) == g["rDis"]
    s = scene.lower()
    assert np.array_equal(r["ld"], read_png(os.path.join(GOLDEN, f"{s}_lDis.png")))
    assert np.array_equal(r["rd"], read_png(os.path.join(GOLDEN, f"{s}_rDis.png")))
    assert r["ld"].min() >= 1 and r["ld"].max() <= 63


@pytest.mark.parametrize("scene", ["Cones", "Teddy"])
def test_oracle_guide_and_coefficients(scene, golden, oracle, scenes, oracle_scene_results):
    g = golden["scenes"][scene]
    _, _, l, r = scenes[scene]
    rgb, mean, var = oracle.cvf_preprocess(l)
    assert sha(mean) == g["lMean"] and sha(var) == g["lVar"]
    crops = np.load(os.path.join(GOLDEN, f"{scene.lower()}_crops.npz"))
    ys, xs = slice(96, 128), slice(180, 244)
    for d in (1, 20, 63):
        q, a, b = oracle.guided_filter(rgb, mean, var, oracle_scene_results[scene]["lraw"][d], want_ab=True)
        assert sha(a) == g[f"l_a_d{d}"] and sha(b) == g[f"l_b_d{d}"]
        assert np.array_equal(a[:, ys, xs], crops[f"l_a_d{d}"])
        assert np.array_equal(q[:12, :24], crops[f"l_q_top_d{d}"])
        assert np.array_equal(q, oracle_scene_results[scene]["lf"][d])


def test_wta_semantics(oracle):
    """d starts at 1, strict <, ties -> lowest d, NaN never selected (DispSel.cpp:93-102)."""
    vol = np.full((5, 2, 4), 3.0, np.float32)
    vol[0] = -100.0                     # d=0 is never a candidate
    vol[2, 0, 0] = 1.0; vol[3, 0, 0] = 1.0  # tie -> 2
    vol[4, 0, 1] = -0.0; vol[1, 0, 1] = 0.0  # -0 == +0 -> lowest d (1)
    vol[:, 0, 2] = np.nan               # nothing compares below +inf -> 0
    vol[1:, 0, 3] = np.inf              # +inf < +inf is false -> 0
    d = oracle.wta(vol)
    assert d[0, 0] == 2 and d[0, 1] == 1 and d[0, 2] == 0 and d[0, 3] == 0 and d[1, 0] == 1


def test_thread_count_does_not_change_results(oracle, scenes):
    _, _, l, r = scenes["Teddy"]
    l, r = l[:64, :96].copy(), r[:64, :96].copy()
    a = oracle.pipeline(l, r, 16, threads=1, keep_volumes=True)
    b = oracle.pipeline(l, r, 16, threads=7, keep_volumes=True)
    assert np.array_equal(a["lVol"], b["lVol"]) and np.array_equal(a["rDis"], b["rDis"])


def test_bad_pixel_metric_matches_golden(golden, oracle_scene_results):
    """StereoMatch.cpp:275-311 (non-occluded mask): the oracle's Cones map scores the golden %BP."""
    cv2 = pytest.importorskip("cv2")
    for scene in ("Cones", "Teddy"):
        s = scene.lower()
        gt = read_png(os.path.join(GOLDEN, f"{s}_disp2.png"), gray=True)
        occl = read_png(os.path.join(GOLDEN, f"{s}_occl.png"), gray=True)
        disp = cv2.convertScaleAbs(oracle_scene_results[scene]["ld"], alpha=4)
        e = cv2.absdiff(disp, gt)
        e[:, :65] = 0
        _, e = cv2.threshold(e, 4 * (127 // 64), 255, cv2.THRESH_TOZERO)
        e = cv2.multiply(e, occl, scale=1 / 255.0)
        bp = float(np.count_nonzero(e)) * 100.0 / gt.size
        assert abs(bp - golden["scenes"][scene]["bp_nonocc_left"]) < 1e-9


def test_box_filter_matches_cv2_on_random_shapes_and_ranges(oracle):
    """Property test of the one primitive everything hangs on: cv::boxFilter(f32, 8x8) == the oracle's
    RowSum/ColumnSum restatement, on many shapes (incl. smaller than the window) and value ranges
    (wide dynamic range, exact zeros, negatives, denormals)."""
    cv2 = pytest.importorskip("cv2")
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 70), st.integers(1, 70), st.integers(0, 2**31 - 1), st.sampled_from(["unit", "wide", "sparse", "tiny"]))
    def check(h, w, seed, kind):
        rng = np.random.default_rng(seed)
        if kind == "unit":
            p = rng.random((h, w), dtype=np.float32)
        elif kind == "wide":
            p = (rng.standard_normal((h, w)) * np.exp(rng.uniform(-20, 10, (h, w)))).astype(np.float32)
        elif kind == "sparse":
            p = (rng.random((h, w)) * (rng.random((h, w)) < 0.2)).astype(np.float32)
        else:
            p = (rng.standard_normal((h, w)) * 1e-41).astype(np.float32)  # denormals
        assert np.array_equal(oracle.box8(p), cv2.boxFilter(p, -1, (8, 8)))

    check()
