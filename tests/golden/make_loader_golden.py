"""Generates tests/golden/loader_golden.npz: JPEG / PNG byte streams encoded by OpenCV together with OpenCV's own decode of
them (cv2.imdecode == libjpeg-turbo / libpng, the decoders behind the reference's cv::imread and -- for JPEG -- bit-compatible
with the libjpeg the reference's JPEGLoader uses).  tests/test_cpu_loader.py checks the in-tree decoders against these vectors
without needing cv2.  Run:  python tests/golden/make_loader_golden.py"""
import os
import numpy as np
import cv2

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rng = np.random.default_rng(11)
out = {}
k = 0
for (W, H) in [(37, 21), (16, 16), (9, 30)]:
    yy, xx = np.mgrid[0:H, 0:W]
    img = np.stack([128 + 100 * np.sin(xx / 5.0) * np.cos(yy / 4.0), (xx * 7 + yy * 3) % 256, rng.integers(0, 256, (H, W))], -1).clip(0, 255).astype(np.uint8)
    for sf in (cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420):
        for q, rst in ((85, 0), (40, 2)):
            ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf, cv2.IMWRITE_JPEG_RST_INTERVAL, rst])
            out[f"jpg{k}"] = np.frombuffer(buf.tobytes(), np.uint8)
            out[f"jpg{k}_rgb"] = cv2.imdecode(buf, cv2.IMREAD_COLOR)[:, :, ::-1].copy()
            k += 1
ok, buf = cv2.imencode(".jpg", rng.integers(0, 256, (19, 23), dtype=np.uint8), [cv2.IMWRITE_JPEG_QUALITY, 70])
out[f"jpg{k}"] = np.frombuffer(buf.tobytes(), np.uint8); out[f"jpg{k}_rgb"] = cv2.imdecode(buf, cv2.IMREAD_COLOR)[:, :, ::-1].copy(); k += 1
out["njpg"] = k
# PNG: colour, BGRA (alpha dropped by imread), 16-bit gray (depth), 8-bit gray (mask)
H, W = 20, 27
yy, xx = np.mgrid[0:H, 0:W]
pngs = {"png_rgb": np.stack([(xx * 9) % 256, (yy * 11) % 256, (xx + yy) % 256], -1).astype(np.uint8),
        "png_rgba": rng.integers(0, 256, (H, W, 4), dtype=np.uint8),
        "png_d16": ((xx * 977 + yy * 1313) % 65536).astype(np.uint16),
        "png_m8": ((xx // 5 + yy // 4) % 5).astype(np.uint8)}
for name, a in pngs.items():
    ok, buf = cv2.imencode(".png", a)
    out[name] = np.frombuffer(buf.tobytes(), np.uint8)
    flag = cv2.IMREAD_UNCHANGED if a.dtype == np.uint16 else (cv2.IMREAD_GRAYSCALE if a.ndim == 2 else cv2.IMREAD_COLOR)
    dec = cv2.imdecode(buf, flag)
    out[name + "_dec"] = dec[:, :, ::-1].copy() if dec.ndim == 3 else dec
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "loader_golden.npz"), **out)
print("wrote", k, "jpeg streams and", len(pngs), "png files:", os.path.getsize(os.path.join(ROOT, "tests", "golden", "loader_golden.npz")), "bytes")
