"""One rank of the sharding tests; launched by torch.distributed.run (see test_cpu_sharding.py / test_gpu_sharded.py).

    mode "cpu": gloo, CPU tensors only: packet broadcast, unsigned key merge (random + oracle projections), pose rows
    mode "gpu": the object-sharded CUDA pipeline over `world` ranks (NCCL with one GPU per rank when the box has them,
                otherwise gloo with all ranks on cuda:0) on the multi-model replay; results -> out_dir/rank{r}.npz
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


# replay used by the GPU sharding test: objects spawn every 6 frames so that three of them exist after 20 frames; capacities chosen so
# that the placement rule puts stores on both ranks.  Tracked objects need the photometric term (GUI default icpWeight=20): with ICP
# alone the normal equations of a 3k-surfel sphere/capsule are near singular, the first step jumps > 0.2 m and the reference rule
# (MaskFusion.cpp:268-272) removes the model one frame after its spawn -- in the oracle and in the CUDA path alike.
def shard_kw(track_all):
    return dict(capacityGlobal=1000000, capacityObject=600000, enableMultipleModels=1, icpWeight=20.0 if track_all else 100.0, so3=0,
                trackAllModels=int(track_all), modelSpawnOffset=6)


def cpu_mode(out_dir):
    import torch
    import torch.distributed as dist
    from maskfusion_b200 import sharding as sh
    from tests import oracle_lib as ol
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    W, H = 64, 48
    P = W * H
    res = {}
    # 1. frame packet: every rank ends up with the loader's bytes, header and class ids
    rng = np.random.default_rng(5)
    rgb = rng.integers(0, 256, (H, W, 3), dtype=np.uint8); depth = rng.random((H, W), dtype=np.float32) * 4
    mask = rng.integers(0, 4, (H, W), dtype=np.uint8); cls = np.array([0, 40, 41, 77], np.int32)
    buf = torch.zeros(sh.frame_packet_bytes(W, H), dtype=torch.uint8)
    if rank == 0:
        sh.pack_frame(buf, W, H, rgb, depth, mask, 123456789012, cls)
    dist.broadcast(buf, src=0)
    ts, got_cls = sh.unpack_header(buf, W, H)
    b = buf.numpy()
    res["packet_ok"] = bool(ts == 123456789012 and np.array_equal(got_cls, cls) and np.array_equal(b[:3 * P], rgb.reshape(-1))
                            and np.array_equal(b[3 * P:7 * P].view(np.float32), depth.reshape(-1)) and np.array_equal(b[7 * P:8 * P], mask.reshape(-1)))
    # 2. unsigned MIN over uint64 patterns, including the empty key and values with the top bit set
    def keys_of(r):
        g = np.random.default_rng(100 + r)
        k = g.integers(0, 2 ** 64, P, dtype=np.uint64)
        k[g.random(P) < 0.3] = np.uint64(0xFFFFFFFFFFFFFFFF)
        return k
    mine = torch.from_numpy(keys_of(rank).view(np.int64).copy())
    sh.allreduce_min_u64(mine)
    want = keys_of(0)
    for r in range(1, world):
        want = np.minimum(want, keys_of(r))
    res["keys_ok"] = bool(np.array_equal(mine.numpy().view(np.uint64), want))
    # 3. the ID projection of models spread over ranks == the projection of all models in one process (oracle kernels)
    L = ol.lib()
    L.orc_global_projection_begin.argtypes = [C.c_int, C.c_int, C.c_void_p]
    L.orc_global_projection_add.argtypes = [C.c_void_p, C.c_int, C.c_void_p, ol.Cam, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p]
    cam = ol.cam(52.8, 52.8, 32, 24)
    def cloud(seed, n=4000):
        g = np.random.default_rng(seed)
        s = np.zeros((n, 12), np.float32)
        s[:, 0] = g.uniform(-1, 1, n); s[:, 1] = g.uniform(-0.8, 0.8, n); s[:, 2] = g.uniform(0.8, 3.0, n)
        s[:, 3] = 20.0; s[:, 4] = 0; s[:, 6] = 1; s[:, 7] = 1
        s[:, 8:11] = (0, 0, -1); s[:, 11] = 0.02
        return s
    nmodels = 5
    pose = np.eye(4, dtype=np.float32)
    def project(indices):
        keys = np.zeros(P, np.uint64)
        L.orc_global_projection_begin(W, H, ol.ptr(keys))
        for i in indices:
            s = cloud(i)
            L.orc_global_projection_add(ol.ptr(s), s.shape[0], ol.ptr(pose), cam, W, H, 4.0, 12.0, 2, 2, 1 << 30, np.uint32(i << 26), ol.ptr(keys))
        return keys
    owners = [i % world for i in range(nmodels)]
    local = torch.from_numpy(project([i for i in range(nmodels) if owners[i] == rank]).view(np.int64).copy())
    sh.allreduce_min_u64(local)
    full = project(range(nmodels))
    res["proj_ok"] = bool(np.array_equal(local.numpy().view(np.uint64), full))
    res["proj_hit"] = int((full != np.uint64(0xFFFFFFFFFFFFFFFF)).sum())
    # 4. pose rows: bit patterns survive the gather (negative zero, denormals)
    rows = np.zeros((3, 32), np.float32)
    rows[rank % 3] = np.float32(-0.0); rows[rank % 3, 5] = np.float32(1e-42); rows[rank % 3, 7] = rank + 0.25
    g = sh.gather_rows(torch.from_numpy(rows)).numpy()
    ok = g.shape == (world, 3, 32)
    for r in range(world):
        e = np.zeros((3, 32), np.float32); e[r % 3] = np.float32(-0.0); e[r % 3, 5] = np.float32(1e-42); e[r % 3, 7] = r + 0.25
        ok = ok and np.array_equal(g[r].view(np.uint32), e.view(np.uint32))
    res["rows_ok"] = bool(ok)
    np.savez(os.path.join(out_dir, f"cpu_rank{rank}.npz"), **res)
    dist.destroy_process_group()


def gpu_mode(out_dir, nframes, track_all):
    import torch
    import torch.distributed as dist
    import maskfusion_b200 as mfb
    from maskfusion_b200.sharding import ShardedMaskFusion
    from maskfusion_b200.synth import SynthScene
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    one_gpu_each = torch.cuda.device_count() >= world
    dev = int(os.environ.get("LOCAL_RANK", 0)) if one_gpu_each else 0
    torch.cuda.set_device(dev)
    if one_gpu_each:
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo")
    W, H = 640, 480
    smf = ShardedMaskFusion(mfb.default_config(W, H, **shard_kw(track_all)), device=dev)
    sc = SynthScene(W, H, n_objects=3, seed=0)
    cls = np.array([0] + [o.class_id for o in sc.objects], np.int32)
    out = {"backend": np.array(dist.get_backend())}
    for t in range(nframes):
        if rank == 0:
            rgb, depth, mask, *_ = sc.render(t)
            smf.processFrame(rgb, depth, t * 33333, mask=np.ascontiguousarray(mask), classIDs=cls)
        else:
            smf.processFrame()
        models = smf.models()
        seg, proj = smf.mf.segmentation()
        out[f"ids{t}"] = np.array([m.getID() for m in models]); out[f"cls{t}"] = np.array([m.getClassID() for m in models])
        out[f"own{t}"] = np.array([smf.owner(i) for i in range(len(models))])
        out[f"pose{t}"] = np.stack([m.getPose() for m in models])
        out[f"cnt{t}"] = np.array([m.lastCount() if smf.owner(i) == rank else -1 for i, m in enumerate(models)])
        out[f"seg{t}"] = np.packbits(seg == 0); out[f"segsum{t}"] = np.array([int(seg.astype(np.int64).sum()), int(proj.astype(np.int64).sum())])
    models = smf.models()
    for i, m in enumerate(models):
        if smf.owner(i) == rank:
            out[f"map{i}"] = m.downloadMap()
    st = smf.stats()
    out["bytes_collective"] = np.array(st["bytes"]); out["transport"] = np.array(st["transport"]); out["nranks"] = np.array(st["nranks"])
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **out)
    smf.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    mode, out_dir = sys.argv[1], sys.argv[2]
    os.makedirs(out_dir, exist_ok=True)
    if mode == "cpu":
        cpu_mode(out_dir)
    else:
        gpu_mode(out_dir, int(sys.argv[3]), int(sys.argv[4]))
