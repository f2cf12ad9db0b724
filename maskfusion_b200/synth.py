"""Seeded synthetic RGB-D sequence generator (SURVEY.md section 8(d) "Synthetic inputs").

Analytic ray-casting of a room box (background) plus N rigid objects (spheres and
boxes) with a smooth seeded camera trajectory and per-object rigid motions.
Produces exactly what the reference's loaders deliver to MaskFusion::processFrame
(Core/FrameData.h:25-46): rgb HxWx3 u8, depth HxW f32 metres quantised to u16 mm
like a .klg log (GUI/Tools/KlgLogReader.cpp:68-70), optional instance masks and
ground-truth poses.  Pure numpy; used by tests and bench.py (data: "synthetic").
"""
from __future__ import annotations

import dataclasses
import numpy as np


def default_intrinsics(width: int, height: int):
    """VGA 528/528/320/240 (GUI/MainController.cpp:124-125); 720p 792/792/640/360."""
    if (width, height) == (640, 480):
        return 528.0, 528.0, 320.0, 240.0
    if (width, height) == (1280, 720):
        return 792.0, 792.0, 640.0, 360.0
    f = 528.0 * width / 640.0
    return f, f, width / 2.0, height / 2.0


def _rot(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


@dataclasses.dataclass
class SynthObject:
    kind: str            # "sphere" | "box"
    size: np.ndarray     # radius (1,) or half extents (3,)
    pose0: np.ndarray    # 4x4 object->world at t=0
    vel: np.ndarray      # world translation per frame once moving (3,)
    start: int           # first moving frame
    class_id: int = 1


class SynthScene:
    """Room 4 x 3 x 4 m around the origin, camera near the origin looking down +z."""

    def __init__(self, width=640, height=480, n_objects=0, seed=0, noise=False, holes=0.0,
                 max_step_mm=4.0, move_after=30):
        self.W, self.H = width, height
        self.fx, self.fy, self.cx, self.cy = default_intrinsics(width, height)
        self.rng = np.random.default_rng(seed)
        self.seed = seed
        self.noise, self.holes = noise, holes
        self.room_min = np.array([-1.1, -0.8, -1.0])
        self.room_max = np.array([1.1, 0.8, 2.6])
        self.objects: list[SynthObject] = []
        # static furniture: part of the background (mask 0), gives ICP all 6 DoF
        self.furniture: list[SynthObject] = []
        for c, hs, r in (((-0.7, 0.55, 2.0), (0.25, 0.25, 0.3), (0.0, 0.5, 0.0)),
                         ((0.75, 0.5, 2.2), (0.2, 0.3, 0.25), (0.0, -0.4, 0.1)),
                         ((0.1, -0.55, 2.3), (0.45, 0.12, 0.2), (0.2, 0.3, 0.0))):
            T = np.eye(4); T[:3, :3] = _rot(*r); T[:3, 3] = c
            self.furniture.append(SynthObject("box", np.array(hs), T, np.zeros(3), 1 << 30, 0))
        for k in range(n_objects):
            ang = (k + 0.5) / max(n_objects, 1) * 2.0 - 1.0
            centre = np.array([ang * 0.7, 0.3 - 0.22 * (k % 3), 1.1 + 0.2 * (k % 4)])
            T = np.eye(4)
            T[:3, :3] = _rot(*(self.rng.uniform(-0.4, 0.4, 3)))
            T[:3, 3] = centre
            if k % 2 == 0:
                size = np.array([self.rng.uniform(0.12, 0.2)])
                kind = "sphere"
            else:
                size = self.rng.uniform(0.08, 0.18, 3)
                kind = "box"
            vel = self.rng.uniform(-1.0, 1.0, 3) * np.array([1.0, 0.2, 0.5]) * max_step_mm * 1e-3
            self.objects.append(SynthObject(kind, size, T, vel, move_after, class_id=1 + k % 5))
        u, v = np.meshgrid(np.arange(width, dtype=np.float64), np.arange(height, dtype=np.float64))
        self.dirs = np.stack([(u - self.cx) / self.fx, (v - self.cy) / self.fy, np.ones_like(u)], -1)
        self.max_step = max_step_mm * 1e-3

    # ---- trajectories -------------------------------------------------
    def camera_pose(self, t: int) -> np.ndarray:
        """camera->world, smooth Lissajous: <= ~max_step per frame and <= 0.3 deg per frame."""
        a = self.max_step * 18.0
        T = np.eye(4)
        T[:3, 3] = [a * np.sin(t * 0.05), 0.6 * a * np.sin(t * 0.07 + 0.5), 0.5 * a * (1 - np.cos(t * 0.04))]
        T[:3, :3] = _rot(0.03 * np.sin(t * 0.045), 0.05 * np.sin(t * 0.06 + 1.0), 0.02 * np.sin(t * 0.03))
        return T

    def object_pose(self, k: int, t: int) -> np.ndarray:
        o = self.objects[k]
        T = o.pose0.copy()
        dt = max(0, t - o.start)
        T[:3, 3] = T[:3, 3] + o.vel * dt
        return T

    # ---- rendering ----------------------------------------------------
    def _texture(self, p, surf_id):
        """procedural colour from world/object point p (N,3): strong gradients, never 0."""
        k = 23.0 + 3.0 * (surf_id % 5)
        s = np.sin(p[:, 0] * k) * np.sin(p[:, 1] * (k + 4.0) + surf_id) * np.sin(p[:, 2] * (k - 5.0) + 0.7)
        chk = ((np.floor(p[:, 0] * 6) + np.floor(p[:, 1] * 6) + np.floor(p[:, 2] * 6)) % 2) * 50.0
        base = np.array([90.0, 120.0, 150.0]) + 25.0 * np.array([surf_id % 3, (surf_id + 1) % 3, (surf_id + 2) % 3])
        c = base[None, :] + 60.0 * s[:, None] * np.array([1.0, 0.8, 0.9])[None, :] + chk[:, None] * np.array([0.5, 1.0, 0.3])
        return np.clip(c, 8, 250)

    def render(self, t: int):
        """returns rgb u8 (H,W,3), depth f32 metres (H,W) (u16-mm quantised), mask u8 (H,W), cam pose 4x4"""
        W, H = self.W, self.H
        Tc = self.camera_pose(t)
        R, o = Tc[:3, :3], Tc[:3, 3]
        d = (self.dirs.reshape(-1, 3) @ R.T)
        N = d.shape[0]
        tbest = np.full(N, np.inf)
        sid = np.zeros(N, dtype=np.int32)
        # room planes (camera is inside)
        with np.errstate(divide="ignore", invalid="ignore"):
            for ax in range(3):
                for side, bound in ((0, self.room_min[ax]), (1, self.room_max[ax])):
                    tt = (bound - o[ax]) / d[:, ax]
                    ok = (tt > 1e-6) & (tt < tbest)
                    p = o[None, :] + tt[:, None] * d
                    inside = np.all((p >= self.room_min - 1e-6) & (p <= self.room_max + 1e-6), axis=1)
                    ok &= inside
                    tbest = np.where(ok, tt, tbest)
                    sid = np.where(ok, ax * 2 + side, sid)
        mask = np.zeros(N, dtype=np.uint8)
        hit_local = o[None, :] + np.where(np.isfinite(tbest), tbest, 0.0)[:, None] * d
        nf = len(self.furniture)
        for k, ob in enumerate(self.furniture + self.objects):
            To = ob.pose0 if k < nf else self.object_pose(k - nf, t)
            Ro, to = To[:3, :3], To[:3, 3]
            ol = (o - to) @ Ro
            dl = d @ Ro
            if ob.kind == "sphere":
                r = ob.size[0]
                b = dl @ ol
                a = np.einsum("ij,ij->i", dl, dl)
                c = ol @ ol - r * r
                disc = b * b - a * c
                with np.errstate(invalid="ignore"):
                    tt = (-b - np.sqrt(disc)) / a
                ok = (disc > 0) & (tt > 1e-6) & (tt < tbest)
            else:
                with np.errstate(divide="ignore", invalid="ignore"):
                    t1 = (-ob.size[None, :] - ol[None, :]) / dl
                    t2 = (ob.size[None, :] - ol[None, :]) / dl
                with np.errstate(invalid="ignore"):
                    tn = np.max(np.minimum(t1, t2), axis=1)
                tf = np.min(np.maximum(t1, t2), axis=1)
                tt = tn
                ok = (tn < tf) & (tn > 1e-6) & (tn < tbest)
            tbest = np.where(ok, tt, tbest)
            sid = np.where(ok, 10 + k, sid)
            mask = np.where(ok, max(0, k - nf + 1), mask).astype(np.uint8)
            with np.errstate(invalid="ignore"):
                pl = ol[None, :] + tt[:, None] * dl
            hit_local = np.where(ok[:, None], pl, hit_local)
        depth = np.where(np.isfinite(tbest), tbest, 0.0)          # dirs.z == 1 -> camera-frame z == t
        if self.noise:
            rng = np.random.default_rng(self.seed * 100003 + t)
            depth = depth + rng.normal(0.0, 1.0, N) * 1e-3 * depth * depth
        self.last_depth_exact = depth.reshape(H, W).astype(np.float32)
        d16 = np.clip(np.round(depth * 1000.0), 0, 65535).astype(np.uint16)
        if self.holes > 0:
            rng = np.random.default_rng(self.seed * 7919 + t)
            d16 = np.where(rng.random(N) < self.holes, 0, d16).astype(np.uint16)
        rgb = np.zeros((N, 3))
        for s in np.unique(sid):
            sel = sid == s
            rgb[sel] = self._texture(hit_local[sel], int(s))
        rgb = rgb.astype(np.uint8).reshape(H, W, 3)
        return rgb, depth_from_u16(d16).reshape(H, W), mask.reshape(H, W), Tc, d16.reshape(H, W)


def depth_from_u16(d16: np.ndarray) -> np.ndarray:
    """u16 millimetres -> float metres exactly as cv::Mat::convertTo(CV_32FC1, 0.001)
    (GUI/Tools/KlgLogReader.cpp:68-70): double multiply, then round to float."""
    return (d16.astype(np.float64) * 0.001).astype(np.float32)


def dense_room_surfels(scene: SynthScene, n_target: int, time: int = 1, conf: float = 20.0) -> np.ndarray:
    """A dense surfel cloud of the room walls (world frame), used to pre-populate the
    background store up to the 5M / 50M capacities BASELINE.json's configs name
    (fusion alone adds <= P/4 surfels per frame).  Returns (n,12) float32 in the
    reference's surfel layout (Model.h:190-192)."""
    rng = np.random.default_rng(scene.seed + 12345)
    lo, hi = scene.room_min, scene.room_max
    ext = hi - lo
    areas = np.array([ext[1] * ext[2], ext[1] * ext[2], ext[0] * ext[2], ext[0] * ext[2], ext[0] * ext[1], ext[0] * ext[1]])
    counts = np.floor(areas / areas.sum() * n_target).astype(np.int64)
    counts[-1] += n_target - counts.sum()
    out = np.zeros((n_target, 12), dtype=np.float32)
    k = 0
    for face in range(6):
        ax, side = face // 2, face % 2
        n = int(counts[face])
        p = rng.random((n, 3)) * ext[None, :] + lo[None, :]
        p[:, ax] = hi[ax] if side else lo[ax]
        nrm = np.zeros((n, 3)); nrm[:, ax] = -1.0 if side else 1.0
        spacing = np.sqrt(areas[face] / max(n, 1))
        col = scene._texture(p, face)
        enc = (col[:, 0].astype(np.int64) << 16) + (col[:, 1].astype(np.int64) << 8) + col[:, 2].astype(np.int64)
        out[k:k + n, 0:3] = p
        out[k:k + n, 3] = conf
        out[k:k + n, 4] = enc.astype(np.float32)
        out[k:k + n, 6] = time
        out[k:k + n, 7] = time
        out[k:k + n, 8:11] = nrm
        out[k:k + n, 11] = spacing * 1.5
        k += n
    return out
