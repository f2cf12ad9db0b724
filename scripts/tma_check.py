#!/usr/bin/env python
"""bilateral filter with the halo tile staged by TMA (MFB200_BILATERAL_TMA=1) against the hand-staged kernel: same bits?"""
import os, sys, subprocess
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)


def one():
    import maskfusion_b200 as mfb
    from maskfusion_b200.synth import SynthScene
    W, H = 640, 480
    sc = SynthScene(W, H, n_objects=0, seed=0, holes=0.02)
    mf = mfb.MaskFusion(mfb.default_config(W, H, capacityGlobal=100000))
    rgb, depth, *_ = sc.render(3)
    mf.setFrame(rgb, depth)
    out = mf.filteredDepth()
    mf.close()
    np.save(sys.argv[2], out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        res = {}
        for tag, env in (("plain", "0"), ("tma", "1")):
            e = dict(os.environ, MFB200_BILATERAL_TMA=env)
            r = subprocess.run([sys.executable, __file__, "one", f"/tmp/bil_{tag}.npy"], env=e, capture_output=True, text=True, timeout=120)
            res[tag] = r.returncode
            if r.returncode != 0:
                print(tag, "failed:", (r.stderr or r.stdout)[-600:])
        if res["plain"] == 0 and res["tma"] == 0:
            a, b = np.load("/tmp/bil_plain.npy"), np.load("/tmp/bil_tma.npy")
            print("bilateral TMA == hand-staged:", bool(np.array_equal(a.view(np.uint32), b.view(np.uint32))), "mismatches", int((a.view(np.uint32) != b.view(np.uint32)).sum()))
