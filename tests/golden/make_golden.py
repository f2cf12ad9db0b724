"""Generates tests/golden/static_icp_160x120.npz from the CPU oracle (oracle/).
The reference ships no golden vectors and cannot be executed here
(OpenGL), so this vector pins the ORACLE against regressions; the oracle itself is pinned
against the reference's recompiled CUDA kernels on the GPU box (tests/test_gpu_ref.py).
Run:  python tests/golden/make_golden.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from maskfusion_b200.synth import SynthScene
from tests import oracle_lib as ol

w, h, n = 160, 120, 5
sc = SynthScene(w, h, n_objects=0, seed=3)
p = ol.OraclePipeline(ol.default_config(w, h, capacityGlobal=60000, icpWeight=100.0, so3=0))
for t in range(n):
    rgb, depth, *_ = sc.render(t)
    p.process_frame(rgb, depth, t)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "static_icp_160x120.npz"), nframes=n, count=p.count(0), pose=p.pose(0),
                    idx=p.tex(0, "idx").copy(), surfels_97=p.surfels(0)[::97].copy())
print("count", p.count(0), "pose t", p.pose(0)[:3, 3])
