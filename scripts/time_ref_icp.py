#!/usr/bin/env python
"""Times the REFERENCE's own icpStep (reduce.cu, compiled unmodified into oracle/_ref/libmf_ref.so) on this GPU in the reference's
calling pattern (one call per Gauss-Newton iteration, sync + D2H inside) on the maps of a synthetic VGA frame.  Test infrastructure."""
import ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tests import oracle_lib as ol
from tests.stagewise import OracleStages
from maskfusion_b200.synth import SynthScene
W, H = 640, 480
ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libmf_ref.so"))
ref.ref_icp_step_time_ms.restype = C.c_float
sc = SynthScene(W, H, n_objects=0, seed=0)
orc = OracleStages(ol.default_config(W, H, capacityGlobal=600000, icpWeight=100.0, so3=0))
for t in range(2):
    rgb, depth, *_ = sc.render(t); orc.p.process_frame(rgb, depth, t)
rgb, depth, *_ = sc.render(2)
P = orc.pose(0).copy()
orc.set_frame(rgb, depth); orc.generate_maps(); orc.track()
fa = orc.frame_arrays(); od = orc.odom(0)
Rpi = np.ascontiguousarray(np.linalg.inv(P[:3, :3].astype(np.float64)).astype(np.float32)); Rc = np.ascontiguousarray(P[:3, :3]); tc = np.ascontiguousarray(P[:3, 3])
out = {}
for l, iters in ((0, 200), (1, 200), (2, 200)):
    w, h = W >> l, H >> l
    vg = np.ascontiguousarray(ol.arr(od.vmap_g[l], (3, h, w), np.float32)); ng = np.ascontiguousarray(ol.arr(od.nmap_g[l], (3, h, w), np.float32))
    ms = ref.ref_icp_step_time_ms(ol.ptr(Rc), ol.ptr(tc), ol.ptr(fa[f"vmap{l}"]), ol.ptr(fa[f"nmap{l}"]), ol.ptr(Rpi), ol.ptr(tc),
                                  C.c_float(528 / (1 << l)), C.c_float(528 / (1 << l)), C.c_float(320 / (1 << l)), C.c_float(240 / (1 << l)),
                                  ol.ptr(vg), ol.ptr(ng), C.c_float(0.1), C.c_float(np.float32(np.sin(20.0 * 3.14159254 / 180.0))), w, h, 128, 112, iters)
    out[f"icpStep_L{l}_us"] = round(float(ms) * 1e3, 2)
out["schedule_10_5_4_us"] = round(10 * out["icpStep_L0_us"] + 5 * out["icpStep_L1_us"] + 4 * out["icpStep_L2_us"], 1)
print(json.dumps(out))
