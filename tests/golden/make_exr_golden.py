"""Generates tests/golden/exr_golden.npz: OpenEXR byte streams written by OpenCV (its bundled OpenEXR: the encoder behind the reference's
datasets' depth files is the same library) together with what the reference keeps of OpenCV's own decode of them
(GUI/Tools/ImageLogReader.cpp:251-258: cv::imread(IMREAD_UNCHANGED); CV_32FC1 as is, element 0 of a CV_32FC3 pixel).
tests/test_cpu_loader.py checks the in-tree decoder against these vectors without needing cv2.
Run:  OPENCV_IO_ENABLE_OPENEXR=1 python tests/golden/make_exr_golden.py"""
import os
os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import numpy as np
import cv2

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rng = np.random.default_rng(5)
out = {}
k = 0
comp = {"none": cv2.IMWRITE_EXR_COMPRESSION_NO, "rle": cv2.IMWRITE_EXR_COMPRESSION_RLE, "zips": cv2.IMWRITE_EXR_COMPRESSION_ZIPS, "zip": cv2.IMWRITE_EXR_COMPRESSION_ZIP}
for (W, H) in [(37, 21), (20, 17), (5, 35)]:
    yy, xx = np.mgrid[0:H, 0:W]
    d1 = (0.4 + 0.01 * xx + 0.02 * yy + 0.05 * rng.random((H, W))).astype(np.float32)
    d1[yy % 7 == 0] = 0.0                                           # holes and runs (the RLE path needs runs)
    d3 = np.stack([d1, d1 * 2, d1 * 3], -1)                         # B, G, R in OpenCV's order: element 0 is what the reference keeps
    for name, c in comp.items():
        for typ in (cv2.IMWRITE_EXR_TYPE_FLOAT, cv2.IMWRITE_EXR_TYPE_HALF):
            for img in (d1, d3):
                ok, buf = cv2.imencode(".exr", img, [cv2.IMWRITE_EXR_TYPE, typ, cv2.IMWRITE_EXR_COMPRESSION, c])
                assert ok
                dec = cv2.imdecode(buf, cv2.IMREAD_UNCHANGED)
                assert dec.dtype == np.float32
                ref = dec if dec.ndim == 2 else dec[:, :, 0]
                out[f"exr{k}"] = np.frombuffer(buf.tobytes(), np.uint8)
                out[f"exr{k}_depth"] = np.ascontiguousarray(ref)
                k += 1
out["nexr"] = k
ok, buf = cv2.imencode(".exr", d1, [cv2.IMWRITE_EXR_COMPRESSION, cv2.IMWRITE_EXR_COMPRESSION_PIZ])
out["exr_piz"] = np.frombuffer(buf.tobytes(), np.uint8)           # refused with a message
p = os.path.join(ROOT, "tests", "golden", "exr_golden.npz")
np.savez_compressed(p, **out)
print("wrote", k, "exr streams:", os.path.getsize(p), "bytes")
