#!/usr/bin/env python
"""Per-rank stage table of the object-sharded frame (configs[3] replay): where a shard's frame time goes.
torchrun --nproc-per-node N scripts/shard_profile.py  ->  gpurun_out/shard_prof_rank<r>.json (in-stream CUDA-event stage clock, overlap off)
plus the un-instrumented frame rate of the same frames."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import bench
    import maskfusion_b200 as mfb
    from maskfusion_b200.sharding import ShardedMaskFusion
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    n, t_prof, t_free = 84, 64, 44
    frames = cls = None
    if rank == 0:                       # before CUDA / NCCL exist in this process: the renderer forks worker processes
        frames, cls = bench.multi_frames(8, n)
    torch.cuda.set_device(local)
    import datetime
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(minutes=10))
    dist.barrier()
    cfg = mfb.default_config(bench.W, bench.H, **bench.MULTI_KW)
    smf = ShardedMaskFusion(cfg, device=local)
    clsp = None
    if rank == 0:
        dev = [(torch.from_numpy(f[0]).cuda(), torch.from_numpy(f[1]).cuda(), torch.from_numpy(np.ascontiguousarray(f[2])).cuda()) for f in frames]
        clsp = np.ascontiguousarray(cls, np.int32)
    torch.cuda.synchronize()

    def step(t):
        if rank == 0:
            smf.processFramePtr(dev[t][0].data_ptr(), dev[t][1].data_ptr(), t * 33333, dev[t][2].data_ptr(), clsp.ctypes.data, len(clsp), True)
        else:
            smf.processFramePtr(0, 0, 0, 0, 0, 0, False)
    for t in range(t_free):
        step(t)
    smf.mf.sync(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(smf.stream)
    for t in range(t_free, t_prof):
        step(t)
    e1.record(smf.stream)
    smf.mf.sync(); dist.barrier(); torch.cuda.synchronize()
    ms_free = e0.elapsed_time(e1) / (t_prof - t_free)
    smf.mf.setProfiling(True)
    for t in range(t_prof, n):
        step(t)
    smf.mf.sync()
    st = smf.mf.stageTimes()
    smf.mf.setProfiling(False)
    K = n - t_prof
    models = smf.models()
    owners = [smf.owner(i) for i in range(len(models))]
    tab = sorted(((k, c / K, ms / K * 1e3) for k, (c, ms) in st.items()), key=lambda x: -x[2])
    out = {"rank": rank, "world": world, "ms_per_frame_uninstrumented": round(ms_free, 4), "owners": owners,
           "us_per_frame_instrumented_sum": round(sum(x[2] for x in tab), 1),
           "stages_us_per_frame": {k: [round(c, 2), round(us, 1)] for k, c, us in tab}}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"shard_prof_w{world}_rank{rank}.json"), "w") as f:
        json.dump(out, f, indent=1)
    if rank in (0, world - 1):
        print(json.dumps(out)[:3000])
    dist.barrier()
    smf.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
