#!/usr/bin/env python
"""Stage clock of k_track_persistent (profiling build: MFB200_TAG=timing MFB200_DEFINES=-DMF_TRACK_TIMING python -m maskfusion_b200.build):
per pyramid level, the average SM-clock time CTA 0 spends in each stage of a reduction step.  Usage: MFB200_TAG=timing python scripts/track_timing.py"""
import collections
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MFB200_TAG", "timing")


def main():
    import maskfusion_b200 as mfb
    from maskfusion_b200.synth import SynthScene
    W, H = 640, 480
    sc = SynthScene(W, H, n_objects=0, seed=0)
    mf = mfb.MaskFusion(mfb.default_config(W, H, capacityGlobal=700000))
    for t in range(6):
        rgb, depth, *_ = sc.render(t)
        mf.processFrame(rgb, depth, t * 33333)
    mf.sync()
    buf = np.zeros(8192, np.int64)
    n = mf.L.mf_debug_track_timing(buf.ctypes.data, 8192)
    ev = buf[:n].reshape(-1, 2)
    ghz = 1.965
    names = {2: "A pixels", 5: "A reduce+exchange+sum", 6: "B pixels (+sigma)", 9: "B reduce+exchange+sum", 20: "solve: assemble A, b", 21: "solve: pivoted LDLT", 22: "solve: rodrigues", 23: "solve: pose composition", 10: "solve: warp constants + barrier",
             12: "so3 pixels", 15: "so3 reduce+exchange+sum", 16: "so3 solve"}
    acc = collections.defaultdict(lambda: [0, 0.0])
    level = "so3"
    prev = None
    for tag, clk in ev:
        tag = int(tag)
        if tag in (100, 101, 102):
            level = f"L{tag - 100}"
        if tag == 11:
            level = "so3"
        if prev is not None and tag in names:
            a = acc[(level, names[tag])]; a[0] += 1; a[1] += (clk - prev) / ghz / 1e3
        prev = clk
    out = {"total_us": round(float(ev[-1, 1] - ev[0, 1]) / ghz / 1e3, 1), "stages_us": {}}
    for (lv, nm), (c, us) in sorted(acc.items()):
        out["stages_us"].setdefault(lv, {})[nm] = {"n": c, "avg_us": round(us / c, 2), "total_us": round(us, 1)}
    mf.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
