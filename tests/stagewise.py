"""Drives the CUDA library and the CPU oracle through MaskFusion::processFrame's schedule
(Core/MaskFusion.cpp:200-607) ONE STAGE AT A TIME, so that every stage of the CUDA path can
be compared with the oracle on identical inputs.  After tracking, the CUDA model is given
the oracle's pose ("teacher forcing"): the Gauss-Newton reductions differ in summation
order (tolerance-checked separately), everything downstream is then bit-comparable.
"""
from __future__ import annotations

import ctypes as C
import numpy as np

from tests import oracle_lib as ol


class OracleStages:
    def __init__(self, cfg: ol.Config):
        self.p = ol.OraclePipeline(cfg)
        self.L = self.p.L
        self.h = self.p.h
        self.cfg = cfg
        for name in ("orc_mf_set_frame", "orc_generate_frame_maps", "orc_model_track", "orc_model_predict_indices",
                     "orc_model_fuse", "orc_model_clean", "orc_model_combined_predict", "orc_model_fill_in"):
            getattr(self.L, name).restype = None
        self.L.orc_mf_set_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self.L.orc_generate_frame_maps.argtypes = [C.c_void_p]
        self.L.orc_model_track.argtypes = [C.c_void_p, C.POINTER(ol.Model), C.c_void_p]
        self.L.orc_model_predict_indices.argtypes = [C.c_void_p, C.POINTER(ol.Model), C.c_int]
        self.L.orc_model_fuse.argtypes = [C.c_void_p, C.POINTER(ol.Model), C.c_int, C.c_float, C.c_float]
        self.L.orc_model_clean.argtypes = [C.c_void_p, C.POINTER(ol.Model), C.c_int]
        self.L.orc_model_combined_predict.argtypes = [C.c_void_p, C.POINTER(ol.Model), C.c_int, C.c_int]
        self.L.orc_model_fill_in.argtypes = [C.c_void_p, C.POINTER(ol.Model)]
        self.L.orc_init_model.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, ol.Cam, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int]
        self.L.orc_odom_init_first_rgb.argtypes = [C.c_void_p, C.c_void_p]
        self.tick = 1

    def mptr(self, i=0):
        return self.L.orc_mf_model(self.h, i)

    def mf(self):
        class MF(C.Structure):
            pass
        return None

    def set_frame(self, rgb, depth):
        self.rgb = np.ascontiguousarray(rgb); self.depth = np.ascontiguousarray(depth)
        self.L.orc_mf_set_frame(self.h, ol.ptr(self.rgb), ol.ptr(self.depth), None)

    def frame_arrays(self):
        """views of the oracle's frame state: depthFilt, vmap[3], nmap[3] (planar)"""
        W, H = self.cfg.width, self.cfg.height

        class MFS(C.Structure):
            _fields_ = [("cfg", ol.Config), ("cam", ol.Cam), ("tick", C.c_int), ("rgb", ol.u8p), ("depthRaw", ol.f32p),
                        ("depthFilt", ol.f32p), ("mask", ol.u8p), ("depthPyr", ol.f32p * 3), ("maskPyr", ol.u8p * 3),
                        ("vmap", ol.f32p * 3), ("nmap", ol.f32p * 3)]
        s = C.cast(self.h, C.POINTER(MFS)).contents
        out = {"depthFilt": ol.arr(s.depthFilt, (H, W), np.float32)}
        for l in range(3):
            out[f"depth{l}"] = ol.arr(s.depthPyr[l], (H >> l, W >> l), np.float32)
            out[f"vmap{l}"] = ol.arr(s.vmap[l], (3, H >> l, W >> l), np.float32)
            out[f"nmap{l}"] = ol.arr(s.nmap[l], (3, H >> l, W >> l), np.float32)
        return out

    def odom(self, i=0):
        class Odom(C.Structure):
            _fields_ = [("W", C.c_int), ("H", C.c_int), ("cam", ol.Cam), ("vtex_tmp", ol.f32p), ("vmap_g", ol.f32p * 3), ("nmap_g", ol.f32p * 3),
                        ("lastDepth", ol.f32p * 3), ("nextDepth", ol.f32p * 3), ("lastImage", ol.u8p * 3), ("nextImage", ol.u8p * 3),
                        ("lastNextImage", ol.u8p * 3), ("dIdx", C.POINTER(C.c_int16) * 3), ("dIdy", C.POINTER(C.c_int16) * 3),
                        ("cloud", ol.f32p * 3), ("corres", C.c_void_p * 3),
                        ("lastICPError", C.c_float), ("lastICPCount", C.c_float), ("lastRGBError", C.c_float), ("lastRGBCount", C.c_float),
                        ("lastSO3Error", C.c_float), ("lastSO3Count", C.c_float), ("lastA", C.c_double * 36), ("lastb", C.c_double * 6)]
        return C.cast(self.p.model(i).odom, C.POINTER(Odom)).contents

    def init_first(self):
        m = self.p.model(0)
        cfg = self.cfg
        cam = ol.cam(cfg.fx, cfg.fy, cfg.cx, cfg.cy)
        fa = self.frame_arrays()
        n = self.L.orc_init_model(ol.ptr(self.rgb), ol.ptr(self.depth), ol.ptr(fa["depthFilt"]), cam, cfg.width, cfg.height, self.tick,
                                  C.c_float(cfg.maxDepthProcessed), m.surf[m.target], m.capacity)
        self.mptr(0).contents.count = n
        self.L.orc_odom_init_first_rgb(m.odom, ol.ptr(self.rgb))

    def generate_maps(self):
        self.L.orc_generate_frame_maps(self.h)

    def track(self, i=0):
        T = np.zeros(16, np.float32)
        self.L.orc_model_track(self.h, self.mptr(i), ol.ptr(T))
        return T.reshape(4, 4)

    def predict_indices(self, i=0):
        self.L.orc_model_predict_indices(self.h, self.mptr(i), self.tick)

    def fuse(self, i=0, weight=1.0):
        self.L.orc_model_fuse(self.h, self.mptr(i), self.tick, C.c_float(self.cfg.depthCutoff), C.c_float(weight))

    def clean(self, i=0):
        self.L.orc_model_clean(self.h, self.mptr(i), self.tick)

    def predict(self, i=0):
        self.L.orc_model_combined_predict(self.h, self.mptr(i), self.tick, self.tick)
        self.L.orc_model_fill_in(self.h, self.mptr(i))

    def pose(self, i=0):
        return np.array(self.p.model(i).pose, np.float32).reshape(4, 4)

    def last_pose(self, i=0):
        return np.array(self.p.model(i).lastPose, np.float32).reshape(4, 4)


def nan_equal(a, b):
    """bit-for-bit equality treating any NaN as equal to any NaN"""
    a = np.asarray(a); b = np.asarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype.kind == "f":
        both_nan = np.isnan(a) & np.isnan(b)
        return bool(np.all((a == b) | both_nan))
    return bool(np.array_equal(a, b))


def mismatch(a, b):
    a = np.asarray(a); b = np.asarray(b)
    if a.dtype.kind == "f":
        bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
    else:
        bad = a != b
    return int(bad.sum()), bad


def planar_valid_equal(a, b):
    """planar 3xHxW maps: x planes must agree in NaN-ness; y/z compared only where x is valid"""
    nan_a, nan_b = np.isnan(a[0]), np.isnan(b[0])
    if not np.array_equal(nan_a, nan_b):
        return False, int((nan_a != nan_b).sum())
    ok = ~nan_a
    bad = 0
    for p in range(3):
        bad += int((a[p][ok] != b[p][ok]).sum())
    return bad == 0, bad
