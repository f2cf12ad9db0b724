"""GPU parity of the two morphological closes of the segmentation:
  * binary edge-map close  (dilate_Kernel / erode_Kernel, segmentation.cu:217-255, host loop :334-354) at the header default
    morphEdgeIterations = 3 (MfSegmentation.h:51) and other radii;
  * mask-id close          (cv::morphologyEx(MORPH_CLOSE, MORPH_ELLIPSE), MfSegmentation.cpp:424-426) at morphMaskIterations 1..3;
both bit-exact against the oracle (whose elliptic close is itself pinned against cv2 in tests/test_cpu.py), and a free-running
multi-model replay with both closes switched on."""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

from tests import oracle_lib as ol

pytestmark = pytest.mark.gpu
W, H = 640, 480


def _mf():
    import maskfusion_b200 as mfb
    return mfb.MaskFusion(mfb.default_config(W, H, capacityGlobal=300000))


def test_edge_close_matches_oracle():
    rng = np.random.default_rng(0)
    mf = _mf()
    L = ol.lib()
    yy, xx = np.mgrid[0:H, 0:W]
    for img in ((rng.random((H, W)) < 0.08).astype(np.uint8) * 255,
                ((np.abs(np.sin(xx / 11.0) * np.cos(yy / 13.0)) < 0.06) * 255).astype(np.uint8)):
        for r in (1, 2, 3):
            for it in (0, 1, 3):
                a = img.copy(); buf = np.zeros_like(a)
                L.orc_morph_close(ol.ptr(a), ol.ptr(buf), W, H, r, it)
                inv_o = np.zeros_like(a); L.orc_invert(ol.ptr(a), W * H, ol.ptr(inv_o))
                g, inv = mf.morphClose(img, r, it, ellipse=False)
                assert np.array_equal(g, a), (r, it, int((g != a).sum()))
                assert np.array_equal(inv, inv_o)
    mf.close()


def test_mask_close_matches_oracle():
    rng = np.random.default_rng(1)
    mf = _mf()
    L = ol.lib()
    L.orc_morph_close_ellipse.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    seg = np.zeros((H, W), np.uint8)
    yy, xx = np.mgrid[0:H, 0:W]
    for k in range(1, 9):
        cx, cy, r = rng.integers(20, W - 20), rng.integers(20, H - 20), rng.integers(10, 90)
        seg[(xx - cx) ** 2 + (yy - cy) ** 2 < r * r] = k
    seg[rng.random((H, W)) < 0.1] = 0
    seg[rng.random((H, W)) < 0.01] = 255
    seg[:3, :] = 7; seg[:, -2:] = 3                      # labels touching the border: taps outside the image are ignored
    for r in (0, 1, 2, 3, 5, 8):
        for it in (1, 2, 3):
            a = seg.copy()
            L.orc_morph_close_ellipse(ol.ptr(a), W, H, r, it)
            g = mf.morphClose(seg, r, it, ellipse=True)
            assert np.array_equal(g, a), (r, it, int((g != a).sum()))
    mf.close()


def test_multi_model_with_both_closes():
    """the multi-model schedule with morphEdgeIterations = 3 and morphMaskIterations = 2 (the CUDA path used to throw on the
    latter): lifecycle, segmentation and ID images against the free-running oracle"""
    from tests.test_gpu_multi import run, check
    from tests.test_gpu_multi import check_exact
    log = run(20, track_all=False, segMorphEdgeIterations=3, segMorphEdgeRadius=1, segMorphMaskIterations=2, segMorphMaskRadius=2, modelSpawnOffset=5,
              tag="closes")
    check(log, 2)
    check_exact(log, 2)
