"""Python host-side mirror of the reference interface over the C ABI
(include/maskfusion_b200.h).  Names follow the reference: MaskFusion.processFrame,
Model.performTracking / predictIndices / fuse / clean / combinedPredict
(Core/MaskFusion.h:69-70, Core/Model/Model.h:128-164).

There is NO CPU fallback: if libmaskfusion_b200.so is missing or no CUDA device is
present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmaskfusion_b200.so")
_LIB = None


class MFError(RuntimeError):
    pass


class Config(C.Structure):
    """mf_config (include/maskfusion_b200.h)"""
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


EXPORTS = [
    "mf_last_error", "mf_abi_version", "mf_config_defaults", "mf_create", "mf_destroy", "mf_process_frame",
    "mf_process_frame_device", "mf_set_input_event", "mf_sync", "mf_tick", "mf_kernel_launches", "mf_model_count", "mf_model_id", "mf_get_pose",
    "mf_set_pose", "mf_model_surfel_count", "mf_model_set_conf_threshold", "mf_download_surfels", "mf_upload_surfels",
    "mf_pose_log_size", "mf_get_pose_log", "mf_set_frame", "mf_model_perform_tracking", "mf_model_predict_indices",
    "mf_model_fuse", "mf_model_clean", "mf_model_combined_predict", "mf_model_init_from_frame",
    "mf_download_filtered_depth", "mf_download_frame_maps", "mf_download_model_maps", "mf_download_index_map",
    "mf_download_prediction", "mf_download_fill_in", "mf_download_association", "mf_download_track_stats",
    "mf_download_edge_map", "mf_morph_close", "mf_debug_track_timing", "mf_attach_backbone", "mf_backbone_stream", "mf_icp_step", "mf_debug_set_poses", "mf_set_profiling", "mf_get_stage_times", "mf_set_frame_classes", "mf_download_segmentation", "mf_model_class_id", "mf_klg_open", "mf_klg_num_frames", "mf_klg_has_more", "mf_klg_get_next",
    "mf_klg_close", "mf_klg_write", "mf_dir_open", "mf_dir_num_frames", "mf_dir_has_more", "mf_dir_has_masks", "mf_dir_set_max_masks", "mf_dir_size",
    "mf_dir_get_next", "mf_dir_close", "mf_decode_jpeg", "mf_decode_exr_depth", "mf_export_poses", "mf_generate_id_image", "mf_pre_segmentation", "mf_write_ply", "mf_cnn_last_error", "mf_gemm_bf16", "mf_conv3x3_bf16", "mf_backbone_create", "mf_backbone_destroy", "mf_backbone_num_layers",
    "mf_backbone_layer", "mf_backbone_get_weights", "mf_backbone_mold", "mf_backbone_input_buffer", "mf_backbone_forward", "mf_backbone_output",
    "mf_backbone_flops", "mf_backbone_num_gemms", "mf_backbone_download",
    "mf_shard_configure", "mf_shard_unique_id", "mf_shard_comm_init", "mf_shard_process_frame", "mf_shard_stats", "mf_shard_frame_begin", "mf_shard_get_poses", "mf_shard_set_poses", "mf_shard_project",
    "mf_shard_projection_keys", "mf_shard_frame_end", "mf_model_owner", "mf_shard_pick_owner", "mf_track_shares",
]


def load_library():
    """dlopen the in-tree CUDA library; loud failure if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = LIB_PATH
    tag = os.environ.get("MFB200_TAG")          # A/B builds of the same library (maskfusion_b200/build.py), e.g. another CTA shape
    if tag:
        path = LIB_PATH.replace(".so", f"_{tag}.so")
    if not os.path.exists(path):
        raise MFError(f"{path} not found: run `python -m maskfusion_b200.build` (there is no CPU fallback)")
    L = C.CDLL(path)
    L.mf_last_error.restype = C.c_char_p
    L.mf_create.restype = C.c_void_p
    L.mf_create.argtypes = [C.POINTER(Config), C.c_int, C.c_void_p]
    L.mf_destroy.argtypes = [C.c_void_p]
    L.mf_config_defaults.argtypes = [C.POINTER(Config), C.c_int, C.c_int]
    L.mf_process_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_float, C.c_int]
    L.mf_process_frame_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_float, C.c_int]
    L.mf_set_input_event.argtypes = [C.c_void_p, C.c_void_p]
    L.mf_kernel_launches.restype = C.c_int64
    for name in ("mf_sync", "mf_tick", "mf_kernel_launches", "mf_model_count"):
        getattr(L, name).argtypes = [C.c_void_p]
    for name in ("mf_model_id", "mf_model_surfel_count", "mf_pose_log_size"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_int]
    L.mf_get_pose.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.mf_set_pose.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.mf_model_set_conf_threshold.argtypes = [C.c_void_p, C.c_int, C.c_float]
    L.mf_download_surfels.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.mf_upload_surfels.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.mf_get_pose_log.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.mf_set_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mf_model_perform_tracking.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.mf_model_predict_indices.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.mf_model_fuse.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float]
    L.mf_model_clean.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.mf_model_combined_predict.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.mf_model_init_from_frame.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.mf_download_filtered_depth.argtypes = [C.c_void_p, C.c_void_p]
    L.mf_download_frame_maps.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mf_download_model_maps.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.mf_download_index_map.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
    L.mf_download_prediction.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
    L.mf_download_fill_in.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 3
    L.mf_download_association.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 3
    L.mf_download_track_stats.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 3
    L.mf_download_edge_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.mf_debug_track_timing.argtypes = [C.c_void_p, C.c_int]
    L.mf_attach_backbone.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.mf_backbone_stream.restype = C.c_void_p; L.mf_backbone_stream.argtypes = [C.c_void_p]
    L.mf_morph_close.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.mf_set_frame_classes.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.mf_download_segmentation.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.mf_model_class_id.argtypes = [C.c_void_p, C.c_int]
    L.mf_set_profiling.argtypes = [C.c_void_p, C.c_int]
    L.mf_get_stage_times.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    L.mf_debug_set_poses.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.mf_icp_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mf_klg_open.restype = C.c_void_p
    L.mf_klg_open.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
    for name in ("mf_klg_num_frames", "mf_klg_has_more", "mf_klg_close"):
        getattr(L, name).argtypes = [C.c_void_p]
    L.mf_klg_close.restype = None
    L.mf_klg_get_next.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mf_klg_write.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mf_dir_open.restype = C.c_void_p
    L.mf_dir_open.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p]
    for name in ("mf_dir_num_frames", "mf_dir_has_more", "mf_dir_has_masks"):
        getattr(L, name).argtypes = [C.c_void_p]
    L.mf_dir_set_max_masks.argtypes = [C.c_void_p, C.c_int]
    L.mf_dir_size.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.mf_dir_get_next.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int64)]
    L.mf_dir_close.argtypes = [C.c_void_p]
    L.mf_cnn_last_error.restype = C.c_char_p
    L.mf_gemm_bf16.argtypes = [C.c_void_p] * 5 + [C.c_int] * 4 + [C.c_void_p]
    L.mf_conv3x3_bf16.argtypes = [C.c_void_p] * 5 + [C.c_int] * 5 + [C.c_void_p]
    L.mf_backbone_create.restype = C.c_void_p
    L.mf_backbone_create.argtypes = [C.c_int, C.c_uint, C.c_void_p]
    L.mf_backbone_destroy.argtypes = [C.c_void_p]; L.mf_backbone_destroy.restype = None
    L.mf_backbone_num_layers.argtypes = [C.c_void_p]
    L.mf_backbone_layer.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.mf_backbone_get_weights.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.mf_backbone_mold.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.mf_backbone_input_buffer.restype = C.c_void_p; L.mf_backbone_input_buffer.argtypes = [C.c_void_p]
    L.mf_backbone_forward.argtypes = [C.c_void_p, C.c_void_p]
    L.mf_backbone_output.restype = C.c_void_p; L.mf_backbone_output.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.mf_backbone_flops.restype = C.c_double; L.mf_backbone_flops.argtypes = [C.c_void_p]
    L.mf_backbone_num_gemms.argtypes = [C.c_void_p]
    L.mf_backbone_download.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.mf_shard_configure.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.mf_shard_frame_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
    L.mf_shard_unique_id.argtypes = [C.c_void_p]
    L.mf_shard_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.mf_shard_process_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int]
    L.mf_shard_stats.argtypes = [C.c_void_p, C.c_void_p]
    L.mf_shard_get_poses.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.mf_shard_set_poses.argtypes = [C.c_void_p, C.c_void_p]
    L.mf_shard_project.argtypes = [C.c_void_p]
    L.mf_shard_projection_keys.restype = C.c_void_p; L.mf_shard_projection_keys.argtypes = [C.c_void_p]
    L.mf_shard_frame_end.argtypes = [C.c_void_p, C.c_float]
    L.mf_model_owner.argtypes = [C.c_void_p, C.c_int]
    L.mf_shard_pick_owner.argtypes = [C.c_void_p, C.c_int]
    _LIB = L
    return L


def default_config(width=640, height=480, **kw) -> Config:
    c = Config()
    load_library().mf_config_defaults(C.byref(c), width, height)
    for k, v in kw.items():
        if not hasattr(c, k):
            raise AttributeError(k)
        setattr(c, k, v)
    return c


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class Model:
    """Handle on one surfel model (reference: class Model, Core/Model/Model.h)."""

    def __init__(self, mf: "MaskFusion", index: int):
        self.mf, self.i = mf, index

    def _ck(self, r):
        return self.mf._ck(r)

    def getPose(self) -> np.ndarray:
        out = np.zeros(16, np.float32)
        self._ck(self.mf.L.mf_get_pose(self.mf.h, self.i, _p(out)))
        return out.reshape(4, 4).T.copy()          # column-major ABI -> numpy row-major

    def overridePose(self, T: np.ndarray):
        a = np.ascontiguousarray(np.asarray(T, np.float32).T).ravel()
        self._ck(self.mf.L.mf_set_pose(self.mf.h, self.i, _p(a)))

    def debugSetPoses(self, pose: np.ndarray, lastPose: np.ndarray):
        a = np.ascontiguousarray(np.asarray(pose, np.float32).T).ravel()
        b = np.ascontiguousarray(np.asarray(lastPose, np.float32).T).ravel()
        self._ck(self.mf.L.mf_debug_set_poses(self.mf.h, self.i, _p(a), _p(b)))

    def getClassID(self) -> int:
        return self.mf.L.mf_model_class_id(self.mf.h, self.i)

    def getID(self) -> int:
        return self.mf.L.mf_model_id(self.mf.h, self.i)

    def lastCount(self) -> int:
        return self._ck(self.mf.L.mf_model_surfel_count(self.mf.h, self.i))

    def setConfidenceThreshold(self, t: float):
        self._ck(self.mf.L.mf_model_set_conf_threshold(self.mf.h, self.i, float(t)))

    def downloadMap(self) -> np.ndarray:
        n = self.lastCount()
        out = np.zeros((max(n, 1), 12), np.float32)
        got = self._ck(self.mf.L.mf_download_surfels(self.mf.h, self.i, _p(out), n))
        return out[:got]

    def uploadMap(self, surfels: np.ndarray):
        s = np.ascontiguousarray(surfels, np.float32)
        self._ck(self.mf.L.mf_upload_surfels(self.mf.h, self.i, _p(s), s.shape[0]))

    def initialise(self, time: int):
        self._ck(self.mf.L.mf_model_init_from_frame(self.mf.h, self.i, time))

    def performTracking(self) -> np.ndarray:
        out = np.zeros(16, np.float32)
        self._ck(self.mf.L.mf_model_perform_tracking(self.mf.h, self.i, _p(out)))
        return out.reshape(4, 4).T.copy()

    def predictIndices(self, time: int):
        self._ck(self.mf.L.mf_model_predict_indices(self.mf.h, self.i, time))

    def fuse(self, time: int, depthCutoff: float, weightMultiplier: float = 1.0):
        self._ck(self.mf.L.mf_model_fuse(self.mf.h, self.i, time, depthCutoff, weightMultiplier))

    def clean(self, time: int):
        self._ck(self.mf.L.mf_model_clean(self.mf.h, self.i, time))

    def combinedPredict(self, time: int, maxTime: int):
        self._ck(self.mf.L.mf_model_combined_predict(self.mf.h, self.i, time, maxTime))

    # ---- read-back (reference layouts) ----
    def indexMap(self):
        H, W = self.mf.H, self.mf.W
        idx = np.zeros((H, W), np.uint32); vc = np.zeros((H, W, 4), np.float32)
        ct = np.zeros((H, W, 4), np.float32); nr = np.zeros((H, W, 4), np.float32)
        self._ck(self.mf.L.mf_download_index_map(self.mf.h, self.i, _p(idx), _p(vc), _p(ct), _p(nr)))
        return idx, vc, ct, nr

    def prediction(self):
        H, W = self.mf.H, self.mf.W
        im = np.zeros((H, W, 4), np.uint8); vc = np.zeros((H, W, 4), np.float32)
        nr = np.zeros((H, W, 4), np.float32); tt = np.zeros((H, W), np.uint16)
        self._ck(self.mf.L.mf_download_prediction(self.mf.h, self.i, _p(im), _p(vc), _p(nr), _p(tt)))
        return im, vc, nr, tt

    def fillIn(self):
        H, W = self.mf.H, self.mf.W
        im = np.zeros((H, W, 4), np.uint8); v = np.zeros((H, W, 4), np.float32); n = np.zeros((H, W, 4), np.float32)
        self._ck(self.mf.L.mf_download_fill_in(self.mf.h, self.i, _p(im), _p(v), _p(n)))
        return im, v, n

    def association(self):
        H, W = self.mf.H, self.mf.W
        flag = np.zeros((W, H), np.uint8); best = np.zeros((W, H), np.uint32); meas = np.zeros((W, H, 12), np.float32)
        self._ck(self.mf.L.mf_download_association(self.mf.h, self.i, _p(flag), _p(best), _p(meas)))
        return flag, best, meas

    def modelMaps(self, level: int):
        H, W = self.mf.H >> level, self.mf.W >> level
        v = np.zeros((3, H, W), np.float32); n = np.zeros((3, H, W), np.float32)
        self._ck(self.mf.L.mf_download_model_maps(self.mf.h, self.i, level, _p(v), _p(n)))
        return v, n

    def trackStats(self):
        A = np.zeros((6, 6)); b = np.zeros(6); e = np.zeros(6, np.float32)
        self._ck(self.mf.L.mf_download_track_stats(self.mf.h, self.i, _p(A), _p(b), _p(e)))
        return A, b, e

    def icpStep(self, level: int, Rcurr: np.ndarray, tcurr: np.ndarray) -> np.ndarray:
        out = np.zeros(29, np.float32)
        R = np.ascontiguousarray(Rcurr, np.float32).ravel(); t = np.ascontiguousarray(tcurr, np.float32).ravel()
        self._ck(self.mf.L.mf_icp_step(self.mf.h, self.i, level, _p(R), _p(t), _p(out)))
        return out

    def poseLog(self) -> np.ndarray:
        n = self._ck(self.mf.L.mf_pose_log_size(self.mf.h, self.i))
        out = np.zeros((max(n, 1), 8))
        got = self._ck(self.mf.L.mf_get_pose_log(self.mf.h, self.i, _p(out), n))
        return out[:got]


class MaskFusion:
    """reference: class MaskFusion (Core/MaskFusion.h).  `stream` is a raw cudaStream_t
    (int); pass torch.cuda.current_stream().cuda_stream to time with torch events."""

    def __init__(self, cfg: Config | None = None, device: int = 0, stream: int | None = None, **kw):
        self.L = load_library()
        self.cfg = cfg if cfg is not None else default_config(**kw)
        self.W, self.H = self.cfg.width, self.cfg.height
        self.h = self.L.mf_create(C.byref(self.cfg), device, C.c_void_p(stream) if stream else None)
        if not self.h:
            raise MFError(self.L.mf_last_error().decode())

    def _ck(self, r):
        if r < 0:
            raise MFError(self.L.mf_last_error().decode())
        return r

    def close(self):
        if getattr(self, "h", None):
            self.L.mf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def exportPoses(self, exportDir: str) -> int:
        """MaskFusion::exportPoses: <exportDir>poses-<id>.txt per model"""
        self.L.mf_export_poses.argtypes = [C.c_void_p, C.c_char_p]
        return self._ck(self.L.mf_export_poses(self.h, exportDir.encode()))

    def processFrame(self, rgb: np.ndarray, depth: np.ndarray, timestamp: int = 0, mask=None, inPose=None,
                     weightMultiplier: float = 1.0, bootstrap: bool = False, classIDs=None):
        """bool MaskFusion::processFrame(FrameDataPointer, const Eigen::Matrix4f*, float, bool); mask/classIDs are
        FrameData::mask / FrameData::classIDs (external instance masks, Core/FrameData.h:36-40)"""
        if classIDs is not None:
            c = np.ascontiguousarray(classIDs, np.int32)
            self._ck(self.L.mf_set_frame_classes(self.h, _p(c), int(c.shape[0])))
        if mask is not None:
            if mask.dtype != np.uint8 or mask.shape != (self.H, self.W):
                raise MFError("mask must be HxW uint8 (CV_8UC1)")
            mask = np.ascontiguousarray(mask)
        if rgb.dtype != np.uint8 or rgb.shape != (self.H, self.W, 3):
            raise MFError("rgb must be HxWx3 uint8 (CV_8UC3, MaskFusion.cpp:202)")
        if depth.dtype != np.float32 or depth.shape != (self.H, self.W):
            raise MFError("depth must be HxW float32 metres (CV_32FC1, MaskFusion.cpp:201)")
        ip = None if inPose is None else np.ascontiguousarray(np.asarray(inPose, np.float32).T).ravel()
        self._ck(self.L.mf_process_frame(self.h, _p(np.ascontiguousarray(rgb)), _p(np.ascontiguousarray(depth)), int(timestamp),
                                         _p(mask), _p(ip), float(weightMultiplier), int(bootstrap)))
        return False

    def processFramePtr(self, rgb_ptr: int, depth_ptr: int, timestamp: int = 0, on_device: bool = False, mask_ptr: int = 0):
        """raw-pointer variant (pinned host or device memory), used by bench.py; class ids through setFrameClasses"""
        fn = self.L.mf_process_frame_device if on_device else self.L.mf_process_frame
        self._ck(fn(self.h, C.c_void_p(rgb_ptr), C.c_void_p(depth_ptr), int(timestamp), C.c_void_p(mask_ptr) if mask_ptr else None, None, 1.0, 0))

    def attachBackbone(self, backbone, every_k: int = 5):
        """Mask R-CNN backbone on the frame path: every k-th processFrame enqueues mold + forward on the backbone's stream"""
        self._ck(self.L.mf_attach_backbone(self.h, C.c_void_p(backbone.h) if backbone is not None else None, int(every_k)))

    def setFrameClasses(self, classIDs):
        c = np.ascontiguousarray(classIDs, np.int32)
        self._ck(self.L.mf_set_frame_classes(self.h, c.ctypes.data_as(C.c_void_p), int(c.shape[0])))

    def setFrame(self, rgb, depth, mask=None):
        self._ck(self.L.mf_set_frame(self.h, _p(np.ascontiguousarray(rgb)), _p(np.ascontiguousarray(depth)), _p(mask)))

    def segmentation(self):
        m = np.zeros((self.H, self.W), np.uint8); p = np.zeros((self.H, self.W), np.uint8)
        multi = bool(self.cfg.enableMultipleModels)
        self._ck(self.L.mf_download_segmentation(self.h, _p(m), _p(p) if multi else None))
        return m, p

    def sync(self):
        self._ck(self.L.mf_sync(self.h))

    def setProfiling(self, on: bool):
        self._ck(self.L.mf_set_profiling(self.h, int(on)))

    def stageTimes(self) -> dict:
        """{kernel name: (launch count, total device ms)} measured with CUDA events on the pipeline's stream"""
        buf = C.create_string_buffer(1 << 16)
        self._ck(self.L.mf_get_stage_times(self.h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            n, c, ms = line.split()
            out[n] = (int(c), float(ms))
        return out

    def getTick(self) -> int:
        return self.L.mf_tick(self.h)

    def kernelLaunches(self) -> int:
        return int(self.L.mf_kernel_launches(self.h))

    def getModels(self):
        return [Model(self, i) for i in range(self.L.mf_model_count(self.h))]

    def getBackgroundModel(self) -> Model:
        return Model(self, 0)

    def filteredDepth(self):
        out = np.zeros((self.H, self.W), np.float32)
        self._ck(self.L.mf_download_filtered_depth(self.h, _p(out)))
        return out

    def frameMaps(self, level: int):
        H, W = self.H >> level, self.W >> level
        d = np.zeros((H, W), np.float32); v = np.zeros((3, H, W), np.float32); n = np.zeros((3, H, W), np.float32)
        self._ck(self.L.mf_download_frame_maps(self.h, level, _p(d), _p(v), _p(n)))
        return d, v, n

    def edgeMap(self):
        e = np.zeros((self.H, self.W), np.float32); b = np.zeros((self.H, self.W), np.uint8)
        self._ck(self.L.mf_download_edge_map(self.h, _p(e), _p(b)))
        return e, b

    def morphClose(self, image, radius: int, iterations: int, ellipse: bool = True):
        """test hook: GPU close of a host image; ellipse=True: the mask-id close (MfSegmentation.cpp:424-426), False: the binary
        edge-map close (segmentation.cu:217-255) -> (closed, inverted)"""
        a = np.ascontiguousarray(image, np.uint8).copy()
        assert a.shape == (self.H, self.W)
        inv = np.zeros_like(a)
        self._ck(self.L.mf_morph_close(self.h, _p(a), int(radius), int(iterations), int(ellipse), None if ellipse else _p(inv)))
        return a if ellipse else (a, inv)


class KlgLogReader:
    """reference: class KlgLogReader (GUI/Tools/KlgLogReader.{h,cpp})"""

    def __init__(self, path: str, width: int, height: int, flipColors: bool = False):
        self.L = load_library()
        self.W, self.H = width, height
        self.k = self.L.mf_klg_open(path.encode(), width, height, int(flipColors))
        if not self.k:
            raise MFError(self.L.mf_last_error().decode())

    def getNumFrames(self):
        return self.L.mf_klg_num_frames(self.k)

    def hasMore(self):
        return bool(self.L.mf_klg_has_more(self.k))

    def getNext(self):
        rgb = np.zeros((self.H, self.W, 3), np.uint8); depth = np.zeros((self.H, self.W), np.float32)
        ts = C.c_int64(0)
        if self.L.mf_klg_get_next(self.k, _p(rgb), _p(depth), C.byref(ts)) != 0:
            raise MFError(self.L.mf_last_error().decode())
        return rgb, depth, ts.value

    def close(self):
        if self.k:
            self.L.mf_klg_close(self.k)
            self.k = None


def write_ply(path: str, surfels: np.ndarray, conf_threshold: float) -> int:
    """one model's cloud as MaskFusion::savePly writes it (MaskFusion.cpp:733-848); surfels = Model.downloadMap()"""
    L = load_library()
    a = np.ascontiguousarray(surfels, np.float32).reshape(-1, 12)
    L.mf_write_ply.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_float]
    n = L.mf_write_ply(path.encode(), _p(a) if a.size else None, int(a.shape[0]), float(conf_threshold))
    if n < 0:
        raise MFError(L.mf_last_error().decode())
    return n


def generate_id_image(result: dict, min_score: float, class_filter=(), special_assignments=()):
    """reference: generate_id_image(result, min_score, class_filter, special_assignments), MaskRCNN/helpers.py:70-98.
    result = {'masks': HxWxN uint8, 'scores': N, 'class_ids': N, 'rois': Nx4} -> (id_image HxW uint8, class ids, rois)"""
    L = load_library()
    masks = np.ascontiguousarray(result["masks"], np.uint8)
    H, W, N = masks.shape
    scores = np.ascontiguousarray(result["scores"], np.float32); cls = np.ascontiguousarray(result["class_ids"], np.int32)
    rois = np.ascontiguousarray(result["rois"], np.int32).reshape(N, 4)
    cf = np.ascontiguousarray(list(class_filter), np.int32); sa = np.ascontiguousarray(list(special_assignments), np.int32)
    img = np.zeros((H, W), np.uint8); ec = np.zeros(max(N, 1), np.int32); er = np.zeros((max(N, 1), 4), np.int32)
    L.mf_generate_id_image.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    n = L.mf_generate_id_image(_p(masks), H, W, N, _p(scores), _p(cls), _p(rois), float(min_score), _p(cf) if cf.size else None, int(cf.size),
                               _p(sa) if sa.size else None, int(sa.size), _p(img), _p(ec), _p(er))
    if n < 0:
        raise MFError(L.mf_last_error().decode())
    return img, ec[:n].tolist(), er[:n].tolist()


def pre_segmentation(mask: np.ndarray, depth: np.ndarray, model_ids, next_model_id: int, allow_new: bool, mapping: np.ndarray):
    """PreSegmentation::performSegmentation (PreSegmentation.cpp:28-90).  mapping: uint8[256], updated in place (state across frames).
    -> (fullSegmentation HxW u8, hasNewLabel, superPixelCount, depthMean, depthStd) with one entry per model (+1 with a new label)"""
    L = load_library()
    m = np.ascontiguousarray(mask, np.uint8); d = np.ascontiguousarray(depth, np.float32)
    H, W = m.shape
    ids = np.ascontiguousarray(model_ids, np.uint8)
    assert mapping.dtype == np.uint8 and mapping.size == 256 and mapping.flags["C_CONTIGUOUS"]
    seg = np.zeros((H, W), np.uint8); has_new = C.c_int(0)
    spc = np.zeros(len(ids) + 1, np.uint32); mean = np.zeros(len(ids) + 1, np.float32); std = np.zeros(len(ids) + 1, np.float32)
    L.mf_pre_segmentation.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p]
    n = L.mf_pre_segmentation(_p(m), _p(d), W, H, _p(ids), len(ids), int(next_model_id), int(bool(allow_new)), _p(mapping), _p(seg), C.byref(has_new),
                              _p(spc), _p(mean), _p(std))
    if n < 0:
        raise MFError(L.mf_last_error().decode())
    return seg, bool(has_new.value), spc[:n].copy(), mean[:n].copy(), std[:n].copy()


def decode_exr_depth(buf: bytes) -> np.ndarray:
    """OpenEXR scan-line file -> HxW float32 depth as the -dir reader delivers it (csrc/mf_loader.cu: decodeEXRDepth)"""
    L = load_library()
    w, h = C.c_int(0), C.c_int(0)
    a = np.frombuffer(buf, np.uint8)
    if L.mf_decode_exr_depth(_p(a), int(a.shape[0]), None, 0, C.byref(w), C.byref(h)) != 0:
        raise MFError(L.mf_last_error().decode())
    out = np.zeros((h.value, w.value), np.float32)
    if L.mf_decode_exr_depth(_p(a), int(a.shape[0]), _p(out), out.size, C.byref(w), C.byref(h)) != 0:
        raise MFError(L.mf_last_error().decode())
    return out


def decode_jpeg(buf: bytes) -> np.ndarray:
    """baseline JPEG -> HxWx3 uint8 RGB, bit-identical to libjpeg's default decode (csrc/mf_jpeg.cu)"""
    L = load_library()
    w, h = C.c_int(0), C.c_int(0)
    a = np.frombuffer(buf, np.uint8)
    if L.mf_decode_jpeg(_p(a), int(a.shape[0]), None, 0, C.byref(w), C.byref(h)) != 0:
        raise MFError(L.mf_last_error().decode())
    out = np.zeros((h.value, w.value, 3), np.uint8)
    if L.mf_decode_jpeg(_p(a), int(a.shape[0]), _p(out), out.size, C.byref(w), C.byref(h)) != 0:
        raise MFError(L.mf_last_error().decode())
    return out


class ImageLogReader:
    """reference: class ImageLogReader (GUI/Tools/ImageLogReader.{h,cpp}), the "-dir" loader: colour / depth / optional mask images
    named <prefix><zero-padded index><ext> (+ "<mask>.txt" with class ids and boxes)"""

    def __init__(self, colorDirectory: str, depthDirectory: str | None = None, maskDirectory: str | None = None, indexWidth: int = 4,
                 colorPrefix: str = "", depthPrefix: str = "", maskPrefix: str = ""):
        self.L = load_library()
        enc = lambda v: v.encode() if v else None          # noqa: E731
        self.r = self.L.mf_dir_open(colorDirectory.encode(), enc(depthDirectory), enc(maskDirectory), indexWidth, colorPrefix.encode(),
                                    depthPrefix.encode(), maskPrefix.encode())
        if not self.r:
            raise MFError(self.L.mf_last_error().decode())
        w, h = C.c_int(0), C.c_int(0)
        self.L.mf_dir_size(self.r, C.byref(w), C.byref(h))
        self.W, self.H = w.value, h.value

    def getNumFrames(self):
        return self.L.mf_dir_num_frames(self.r)

    def hasMore(self):
        return bool(self.L.mf_dir_has_more(self.r))

    def hasMasks(self):
        return bool(self.L.mf_dir_has_masks(self.r))

    def setMaxMasks(self, n: int):
        self.L.mf_dir_set_max_masks(self.r, n)

    def getNext(self):
        """-> rgb, depth, timestamp, mask or None, classIDs or None, rois (n,4 as x,y,w,h) or None   (FrameData fields)"""
        rgb = np.zeros((self.H, self.W, 3), np.uint8); depth = np.zeros((self.H, self.W), np.float32); mask = np.zeros((self.H, self.W), np.uint8)
        ids = np.zeros(256, np.int32); boxes = np.zeros((255, 4), np.int32); n = C.c_int(256); ts = C.c_int64(0)
        rc = self.L.mf_dir_get_next(self.r, _p(rgb), _p(depth), _p(mask), _p(ids), _p(boxes), C.byref(n), C.byref(ts))
        if rc < 0:
            raise MFError(self.L.mf_last_error().decode())
        cls = ids[:n.value].copy() if n.value else None
        return rgb, depth, ts.value, (mask if rc == 1 else None), cls, (boxes[:n.value - 1].copy() if n.value > 1 else None)

    def close(self):
        if self.r:
            self.L.mf_dir_close(self.r)
            self.r = None


def write_klg(path: str, timestamps, depth_mm: np.ndarray, rgb: np.ndarray):
    """raw .klg (layout: KlgLogReader.cpp:29,53-89)"""
    L = load_library()
    n, H, W = depth_mm.shape
    ts = np.ascontiguousarray(timestamps, np.int64)
    d = np.ascontiguousarray(depth_mm, np.uint16); c = np.ascontiguousarray(rgb, np.uint8)
    if L.mf_klg_write(path.encode(), W, H, n, _p(ts), _p(d), _p(c)) != 0:
        raise MFError(L.mf_last_error().decode())


class Backbone:
    """Mask R-CNN ResNet-101-FPN backbone on tcgen05 GEMMs (csrc/mf_cnn.cu).  Weights are synthetic (seeded)."""

    def __init__(self, input_size=1024, seed=1, stream: int | None = None):
        self.L = load_library()
        self.S = input_size
        self.h = self.L.mf_backbone_create(input_size, seed, C.c_void_p(stream) if stream else None)
        if not self.h:
            raise MFError(self.L.mf_cnn_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.mf_backbone_destroy(self.h); self.h = None

    def layers(self):
        out = []
        for i in range(self.L.mf_backbone_num_layers(self.h)):
            d = np.zeros(6, np.int32)
            self.L.mf_backbone_layer(self.h, i, _p(d))
            out.append(tuple(int(v) for v in d))
        return out

    def weights(self, i):
        cin, cout, k, stride, pad, kpad = self.layers()[i]
        w = np.zeros((cout, kpad), np.float32); b = np.zeros(cout, np.float32)
        self.L.mf_backbone_get_weights(self.h, i, _p(w), _p(b))
        return w[:, :k * k * cin].reshape(cout, k, k, cin), b

    def forward(self, input_ptr: int):
        if self.L.mf_backbone_forward(self.h, C.c_void_p(input_ptr)) != 0:
            raise MFError(self.L.mf_cnn_last_error().decode())

    def output(self, level):
        d = np.zeros(3, np.int32)
        ptr = self.L.mf_backbone_output(self.h, level, _p(d))
        return ptr, tuple(int(v) for v in d)

    def download(self, level) -> np.ndarray:
        """bf16 output as float32 numpy (H, W, C)"""
        _, (h, w, c) = self.output(level)
        raw = np.zeros((h, w, c), np.uint16)
        if self.L.mf_backbone_download(self.h, level, _p(raw)) != 0:
            raise MFError("backbone download failed")
        return (raw.astype(np.uint32) << 16).view(np.float32)

    def flops(self):
        return float(self.L.mf_backbone_flops(self.h))

    def numGemms(self):
        return int(self.L.mf_backbone_num_gemms(self.h))
