"""The N > 1 host logic on CPU: world-size-2 `gloo` run of the clip sharding and the 64-byte metric
all-reduce (pipeline.shard_range / pipeline.allreduce_metrics), with the oracle standing in for the CUDA
evaluation kernel (the oracle is the checker here; the product reduction code is what is under test)."""
import importlib
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import eval_oracle, synth

PKG = "audio-to-motion-generation_b200"
N_CLIPS = 11            # odd on purpose: ragged shards (6 + 5)
FIELDS = ("pck_hits", "n_keypoints", "n_frames", "n_pose", "n_motion")


def _pack(partials):
    buf = torch.zeros(8, dtype=torch.int64)
    for i, k in enumerate(FIELDS):
        buf[i] = partials[k]
    buf[5:7] = torch.tensor([partials["abs_pose"], partials["abs_motion"]], dtype=torch.float64).view(torch.int64)
    return buf


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pipeline = importlib.import_module(PKG + ".pipeline")
        me = importlib.import_module(PKG + ".motion_evaluation")
        lo, hi = pipeline.shard_range(N_CLIPS, rank, world)
        assert (lo, hi) == synth.shard_range(N_CLIPS, rank, world)
        gt = synth.gt_pose_batch(lo, hi - lo)
        pred = synth.noisy_pred_batch(lo, hi - lo)
        accum = _pack(eval_oracle.metric_partials(pred, gt))
        pipeline.allreduce_metrics(accum)                      # gloo path of the product reduction
        m = me.read_metrics(accum)
        m.update(me.finalize_metrics(m))
        np.save(os.path.join(out_dir, "rank%d.npy" % rank),
                np.array([m["pck_hits"], m["n_keypoints"], m["n_frames"], m["abs_pose"], m["abs_motion"], m["pck"],
                          m["l1_pose"], m["l1_motion"], lo, hi], dtype=np.float64))
    finally:
        dist.destroy_process_group()


def test_world2_sharded_metrics_equal_single_rank(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (np.load(str(tmp_path / ("rank%d.npy" % r))) for r in (0, 1))
    np.testing.assert_array_equal(r0[:8], r1[:8])              # every rank holds the same reduced result
    assert (r0[8], r0[9], r1[8], r1[9]) == (0, 6, 6, 11)
    whole = eval_oracle.metric_partials(synth.noisy_pred_batch(0, N_CLIPS), synth.gt_pose_batch(0, N_CLIPS))
    fin = eval_oracle.finalize(whole)
    assert int(r0[0]) == whole["pck_hits"] and int(r0[1]) == whole["n_keypoints"] and int(r0[2]) == whole["n_frames"]
    np.testing.assert_allclose(r0[3], whole["abs_pose"], rtol=1e-12)
    np.testing.assert_allclose(r0[4], whole["abs_motion"], rtol=1e-12)
    assert r0[5] == fin["pck"]                                 # integer hits: bit-identical to one rank
    np.testing.assert_allclose([r0[6], r0[7]], [fin["l1_pose"], fin["l1_motion"]], rtol=1e-12)


def test_shard_ranges_cover_everything_once():
    pipeline = importlib.import_module(PKG + ".pipeline")
    for n in (0, 1, 7, 256, 100000):
        for world in (1, 2, 4, 8):
            spans = [pipeline.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
