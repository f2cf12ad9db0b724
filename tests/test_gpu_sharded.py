"""Object-sharded pipeline against the single-process pipeline on the same replay: the shards must
reproduce it BIT FOR BIT (same kernels on the same inputs; the only cross-rank arithmetic is an integer MIN).
Two ranks: with two GPUs the in-library NCCL exchange (mf_shard_process_frame: broadcast / all-gather / 64-bit MIN all-reduce issued
by the library on its stream), else gloo with both ranks on cuda:0 (host-staged transport over the phase-split ABI, same kernels) --
so the round-end single-GPU run covers the sharded code path as well."""
from __future__ import annotations

import numpy as np
import pytest

from tests.shard_worker import shard_kw
from tests.test_cpu_sharding import launch

pytestmark = pytest.mark.gpu
W, H = 640, 480


def single(nframes, track_all):
    import maskfusion_b200 as mfb
    from maskfusion_b200.synth import SynthScene
    sc = SynthScene(W, H, n_objects=3, seed=0)
    mf = mfb.MaskFusion(mfb.default_config(W, H, **shard_kw(track_all)))
    cls = np.array([0] + [o.class_id for o in sc.objects], np.int32)
    out = {}
    for t in range(nframes):
        rgb, depth, mask, *_ = sc.render(t)
        mf.processFrame(rgb, depth, t * 33333, mask=np.ascontiguousarray(mask), classIDs=cls)
        models = mf.getModels()
        seg, proj = mf.segmentation()
        out[f"ids{t}"] = np.array([m.getID() for m in models]); out[f"cls{t}"] = np.array([m.getClassID() for m in models])
        out[f"pose{t}"] = np.stack([m.getPose() for m in models]); out[f"cnt{t}"] = np.array([m.lastCount() for m in models])
        out[f"seg{t}"] = np.packbits(seg == 0); out[f"segsum{t}"] = np.array([int(seg.astype(np.int64).sum()), int(proj.astype(np.int64).sum())])
    for i, m in enumerate(mf.getModels()):
        out[f"map{i}"] = m.downloadMap()
    mf.close()
    return out


@pytest.mark.parametrize("track_all", [0, 1])
def test_two_shards_equal_one_process(tmp_path, track_all):
    nframes = 24
    ref = single(nframes, track_all)
    launch(2, ["gpu", tmp_path, nframes, track_all], timeout=900)
    ranks = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
    nmodels_final = len(ref[f"ids{nframes - 1}"])
    assert max(len(ref[f"ids{t}"]) for t in range(nframes)) >= 3, "fewer than two object models were spawned: the test would not exercise sharding"
    owners = ranks[0][f"own{nframes - 1}"]
    seen = set()
    for t in range(nframes):
        seen |= set(ranks[0][f"own{t}"].tolist())
    assert owners[0] == 0 and seen == {0, 1}, (owners, seen)                      # both ranks held stores
    for t in range(nframes):
        for k in ("ids", "cls", "pose", "seg", "segsum", "own"):                  # replicated state is identical on every rank
            assert np.array_equal(ranks[0][f"{k}{t}"], ranks[1][f"{k}{t}"]), (t, k)
    # fp64 sums rounded to float do not depend on the launch shape (one process batches all tracked models into one launch with the SMs
    # split between them, a shard tracks only its own models with the whole GPU): the shards reproduce the single process BIT FOR BIT,
    # tracked objects included -- poses, ids, segmentation, counts on every frame and the surfel stores at the end.
    for t in range(nframes):
        z = ranks[0]
        for k in ("ids", "cls", "pose", "seg", "segsum"):
            assert np.array_equal(z[f"{k}{t}"], ref[f"{k}{t}"]), (t, k)
        cnt = np.where(ranks[0][f"cnt{t}"] >= 0, ranks[0][f"cnt{t}"], ranks[1][f"cnt{t}"])
        assert np.array_equal(cnt, ref[f"cnt{t}"]), (t, cnt, ref[f"cnt{t}"])
    for i in range(nmodels_final):                                            # the surfel stores themselves
        m = ranks[int(owners[i])][f"map{i}"]
        assert m.shape == ref[f"map{i}"].shape and np.array_equal(m.view(np.uint32), ref[f"map{i}"].view(np.uint32)), i
