"""ctypes binding of the CPU oracle (oracle/libmf_oracle.so).  TEST INFRASTRUCTURE:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package never does."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None

f32p = C.POINTER(C.c_float)
u8p = C.POINTER(C.c_uint8)
u16p = C.POINTER(C.c_uint16)
u32p = C.POINTER(C.c_uint32)
f64p = C.POINTER(C.c_double)


class Cam(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float)]


class Config(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32),
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
        ("depthCutoff", C.c_float), ("maxDepthProcessed", C.c_float), ("icpWeight", C.c_float),
        ("rgbOnly", C.c_int32), ("pyramid", C.c_int32), ("fastOdom", C.c_int32), ("so3", C.c_int32),
        ("frameToFrameRGB", C.c_int32),
        ("confGlobal", C.c_float), ("confObject", C.c_float),
        ("timeDelta", C.c_int32), ("outlierCoeff", C.c_float),
        ("capacityGlobal", C.c_int32), ("capacityObject", C.c_int32),
        ("enableMultipleModels", C.c_int32), ("trackAllModels", C.c_int32), ("modelSpawnOffset", C.c_int32),
        ("minRelSizeNew", C.c_float), ("maxRelSizeNew", C.c_float),
        ("segThreshold", C.c_float), ("segWeightDistance", C.c_float), ("segWeightConvexity", C.c_float),
        ("segMorphEdgeIterations", C.c_int32), ("segMorphEdgeRadius", C.c_int32),
        ("segMorphMaskIterations", C.c_int32), ("segMorphMaskRadius", C.c_int32),
    ]


class Model(C.Structure):
    _fields_ = [
        ("id", C.c_int), ("classID", C.c_int),
        ("pose", C.c_float * 16), ("lastPose", C.c_float * 16), ("initialC2Winv", C.c_float * 16),
        ("isStatic", C.c_int), ("age", C.c_int), ("nonstatic", C.c_int),
        ("confThreshold", C.c_float), ("maxDepth", C.c_float),
        ("capacity", C.c_int), ("count", C.c_int),
        ("surf", f32p * 2), ("target", C.c_int), ("allowFillIn", C.c_int),
        ("idx", u32p), ("vertConf", f32p), ("colorTime", f32p), ("normRad", f32p),
        ("splatImage", u8p), ("splatVertex", f32p), ("splatNormal", f32p), ("splatTime", u16p),
        ("fillVertex", f32p), ("fillNormal", f32p), ("fillImage", u8p),
        ("updateId", u8p), ("best", u32p), ("meas", f32p),
        ("odom", C.c_void_p),
        ("nlog", C.c_int), ("caplog", C.c_int), ("log", f64p),
    ]


def build():
    subprocess.run(["make", "-C", ORACLE_DIR, "libmf_oracle.so"], check=True, capture_output=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ORACLE_DIR, "libmf_oracle.so")
        srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".c", ".h"))]
        if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
            build()
        L = C.CDLL(path)
        L.orc_expf.restype = C.c_float; L.orc_expf.argtypes = [C.c_float]
        L.orc_acosf.restype = C.c_float; L.orc_acosf.argtypes = [C.c_float]
        L.orc_mf_create.restype = C.c_void_p; L.orc_mf_create.argtypes = [C.POINTER(Config)]
        L.orc_mf_destroy.argtypes = [C.c_void_p]
        L.orc_mf_process_frame.restype = C.c_int
        L.orc_mf_process_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_mf_model.restype = C.POINTER(Model); L.orc_mf_model.argtypes = [C.c_void_p, C.c_int]
        L.orc_config_defaults.argtypes = [C.POINTER(Config), C.c_int, C.c_int]
        L.orc_model_fusion_weight.restype = C.c_float
        L.orc_requires_fill_in.restype = C.c_int
        L.orc_init_model.restype = C.c_int
        L.orc_clean.restype = C.c_int
        _LIB = L
    return _LIB


def ptr(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def cam(fx, fy, cx, cy):
    return Cam(fx, fy, cx, cy)


def default_config(width=640, height=480, **kw) -> Config:
    c = Config()
    lib().orc_config_defaults(C.byref(c), width, height)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def arr(p, shape, dtype):
    """numpy view (no copy) of a ctypes pointer"""
    n = int(np.prod(shape))
    return np.ctypeslib.as_array(p, shape=(n,)).view(dtype).reshape(shape) if n else np.zeros(shape, dtype)


class OraclePipeline:
    """Stateful oracle == MaskFusion::processFrame restatement (oracle/orc_pipeline.c)."""

    def __init__(self, cfg: Config):
        self.L = lib()
        self.cfg = cfg
        self.h = self.L.orc_mf_create(C.byref(cfg))
        self.W, self.H = cfg.width, cfg.height

    def close(self):
        if self.h:
            self.L.orc_mf_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def process_frame(self, rgb, depth, ts=0):
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        return self.L.orc_mf_process_frame(self.h, ptr(rgb), ptr(depth), ts)

    def model(self, i=0) -> Model:
        return self.L.orc_mf_model(self.h, i).contents

    def pose(self, i=0):
        return np.array(self.model(i).pose, dtype=np.float32).reshape(4, 4)

    def count(self, i=0):
        return self.model(i).count

    def surfels(self, i=0):
        m = self.model(i)
        return arr(m.surf[m.target], (m.count, 12), np.float32)

    def tex(self, i, name):
        m = self.model(i)
        P = (self.H, self.W)
        spec = {"idx": (P, np.uint32), "vertConf": (P + (4,), np.float32), "colorTime": (P + (4,), np.float32),
                "normRad": (P + (4,), np.float32), "splatImage": (P + (4,), np.uint8),
                "splatVertex": (P + (4,), np.float32), "splatNormal": (P + (4,), np.float32),
                "splatTime": (P, np.uint16), "fillVertex": (P + (4,), np.float32),
                "fillNormal": (P + (4,), np.float32), "fillImage": (P + (4,), np.uint8),
                "updateId": ((self.W, self.H), np.uint8), "best": ((self.W, self.H), np.uint32),
                "meas": ((self.W, self.H, 12), np.float32)}[name]
        return arr(getattr(m, name), spec[0], spec[1])


# ---- MfSegmentation CPU tail (oracle/orc_mfseg.h): used by the thread-invariance test and by bench.py's configs[0] leg ----
class MfsegState(C.Structure):
    _fields_ = [("semanticIgnoreMap", u8p), ("maskToID", C.c_uint8 * 256), ("minMaskModelOverlap", C.c_float),
                ("minMappedComponentSize", C.c_int32), ("personClassID", C.c_int32), ("removeEdges", C.c_int32)]


class MfsegIn(C.Structure):
    _fields_ = [("W", C.c_int32), ("H", C.c_int32), ("edgesInv", u8p), ("depth", f32p), ("mask", u8p), ("nMasks", C.c_int32),
                ("classIDs", C.POINTER(C.c_int32)), ("projectedIDs", u8p), ("nModels", C.c_int32), ("modelIDs", u8p),
                ("modelClassIDs", C.POINTER(C.c_int32)), ("nextModelID", C.c_uint8), ("allowNew", C.c_int32),
                ("minRelSizeNew", C.c_float), ("maxRelSizeNew", C.c_float), ("morphMaskRadius", C.c_int32), ("morphMaskIterations", C.c_int32)]


class MfsegOut(C.Structure):
    _fields_ = [("fullSegmentation", u8p), ("hasNewLabel", C.c_int32), ("newClassID", C.c_int32), ("isEmpty", C.c_int32 * 256),
                ("pixelCount", C.c_int32 * 256), ("labels", C.POINTER(C.c_int32)), ("nComponents", C.c_int32)]


def segmentation_frame(W=640, H=480, n_objects=3, seed=0, t=8):
    """inputs of MfSegmentation::performSegmentation's CPU part for one synthetic frame: inverted edge map (edge-ness on the oracle's
    frame maps, threshold 0.3, no close), raw depth, the frame's instance mask + class ids, and a projected-ID image as the global
    projection of already-spawned models would give it (the previous frame's instances)"""
    from maskfusion_b200.synth import SynthScene
    from tests.stagewise import OracleStages
    sc = SynthScene(W, H, n_objects=n_objects, seed=seed)
    rgb, depth, mask, *_ = sc.render(t)
    _, _, prev, *_ = sc.render(t - 1)
    st = OracleStages(default_config(W, H, capacityGlobal=1000))
    st.set_frame(rgb, depth); st.generate_maps()
    fa = st.frame_arrays()
    L = lib()
    e = np.zeros((H, W), np.float32); b = np.zeros((H, W), np.uint8); inv = np.zeros((H, W), np.uint8)
    L.orc_geometric_edges(ptr(fa["vmap0"]), ptr(fa["nmap0"]), W, H, C.c_float(150.0), C.c_float(2.8), ptr(e))
    L.orc_threshold(ptr(e), W * H, C.c_float(0.3), ptr(b)); L.orc_invert(ptr(b), W * H, ptr(inv))
    cls = np.array([0] + [o.class_id for o in sc.objects], np.int32)
    return {"W": W, "H": H, "edgesInv": inv, "depth": np.ascontiguousarray(depth), "mask": np.ascontiguousarray(mask), "classIDs": cls,
            "projectedIDs": np.ascontiguousarray(prev), "modelIDs": np.arange(1 + n_objects, dtype=np.uint8),
            "modelClassIDs": cls.copy()}


def run_mfseg_cpu(fr, threads=1, repeats=1, morphMaskIterations=0):
    """orc_mfseg_cpu on a segmentation_frame(); returns (fullSegmentation, nComponents, hasNewLabel, seconds per call)"""
    import time
    L = lib()
    L.orc_mfseg_set_threads(int(threads))
    W, H = fr["W"], fr["H"]
    st = MfsegState(); L.orc_mfseg_state_init(C.byref(st), W, H)
    seg = np.zeros((H, W), np.uint8)
    dt = []
    for _ in range(repeats):
        edges = fr["edgesInv"].copy()                    # the call overwrites its edge buffer
        i = MfsegIn(W, H, edges.ctypes.data_as(u8p), fr["depth"].ctypes.data_as(f32p), fr["mask"].ctypes.data_as(u8p), len(fr["classIDs"]),
                    fr["classIDs"].ctypes.data_as(C.POINTER(C.c_int32)), fr["projectedIDs"].ctypes.data_as(u8p), len(fr["modelIDs"]),
                    fr["modelIDs"].ctypes.data_as(u8p), fr["modelClassIDs"].ctypes.data_as(C.POINTER(C.c_int32)), 200, 1, 0.015, 0.4, 2,
                    morphMaskIterations)
        o = MfsegOut(); o.fullSegmentation = seg.ctypes.data_as(u8p); o.labels = None
        t0 = time.perf_counter()
        L.orc_mfseg_cpu(C.byref(st), C.byref(i), C.byref(o))
        dt.append(time.perf_counter() - t0)
    L.orc_mfseg_state_free(C.byref(st))
    L.orc_mfseg_set_threads(1)
    return seg, int(o.nComponents), int(o.hasNewLabel), float(np.median(dt))
