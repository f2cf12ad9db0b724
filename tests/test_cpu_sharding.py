"""Host side of the object-sharded mode on CPU: world_size-2 gloo run of the exchange helpers in
maskfusion_b200/sharding.py (frame packet broadcast, unsigned 64-bit key MIN merge checked against the oracle's
GlobalProjection on a partition of the models, pose-row gather) + the placement rule exported by the library."""
from __future__ import annotations

import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def launch(world, args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), os.path.join(ROOT, "tests", "shard_worker.py")] + [str(a) for a in args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def test_gloo_world2_exchange(tmp_path, oracle, product_lib):
    oracle.lib()                                     # build the oracle once, before two ranks race for it
    launch(2, ["cpu", tmp_path])
    for r in range(2):
        z = np.load(tmp_path / f"cpu_rank{r}.npz")
        assert bool(z["packet_ok"]) and bool(z["keys_ok"]) and bool(z["rows_ok"]), {k: z[k] for k in z.files}
        assert bool(z["proj_ok"]) and int(z["proj_hit"]) > 500, {k: z[k] for k in z.files}


def test_pick_owner_rule(product_lib):
    from maskfusion_b200 import sharding as sh
    # the background (rank 0) outweighs objects: new objects go elsewhere until the capacities balance; ties -> highest rank
    assert sh.pick_owner([9437184, 0]) == 1
    assert sh.pick_owner([9437184, 0, 0, 0]) == 3
    assert sh.pick_owner([9437184, 1048576, 1048576, 0]) == 3
    assert sh.pick_owner([9437184, 2097152, 1048576, 1048576]) == 3
    assert sh.pick_owner([9437184, 2097152, 1048576, 2097152]) == 2
    assert sh.pick_owner([1048576, 9 * 1048576]) == 0
    rng = np.random.default_rng(0)
    for _ in range(200):
        w = int(rng.integers(1, 9)); loads = rng.integers(0, 5, w) * 1048576
        want = max(r for r in range(w) if loads[r] == loads.min())
        assert sh.pick_owner(loads) == want
    # simulated spawn sequence on 8 ranks: one object per rank before any rank takes a second one
    loads = np.zeros(8, np.int64); loads[0] = 4734976
    placed = []
    for _ in range(7):
        r = sh.pick_owner(loads); placed.append(r); loads[r] += 1048576
    assert sorted(placed) == [1, 2, 3, 4, 5, 6, 7]
