"""CPU: the multi-GPU host logic of the joint pipeline -- pair sharding and the single all-gather of key-point
records -- with world_size 2 over gloo, plus the synthetic workload generator."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flowtrack.pytorch_b200.tracking import pipeline as pl
from flowtrack.pytorch_b200.tracking import synthetic
from oracle import decode_oracle


@pytest.mark.parametrize("n,world", [(16, 1), (16, 2), (16, 8), (7, 2), (5, 8), (0, 4)])
def test_shard_pairs_partitions_every_pair_once(n, world):
    spans = [pl.shard_pairs(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c and a <= b and c <= d
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, pairs, people, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = pl.shard_pairs(pairs, rank, world)
        d = synthetic.frame_pair_batch(pairs, people, (64, 128), (64, 48), seed=5)       # same global batch on every rank
        # the rank's share of the per-person records: here the oracle decode of seeded heat-maps stands in for the GPU step
        g = torch.Generator().manual_seed(11)
        hm = torch.rand(pairs * people, 17, 16, 12, generator=g)
        sl = slice(lo * people, hi * people)
        xy, sc = decode_oracle.final_preds(hm[sl].numpy(), d["centers"][sl].numpy(), d["scales"][sl].numpy(), True)
        local = torch.from_numpy(np.concatenate((xy, sc), axis=2).astype(np.float32))
        allk = pl.gather_keypoints(local)
        assert allk.shape == (pairs * people, 17, 3)
        xy_all, sc_all = decode_oracle.final_preds(hm.numpy(), d["centers"].numpy(), d["scales"].numpy(), True)
        np.testing.assert_array_equal(allk.numpy(), np.concatenate((xy_all, sc_all), axis=2).astype(np.float32))
        open(os.path.join(out_dir, "ok%d" % rank), "w").close()
    finally:
        dist.destroy_process_group()


def test_gather_keypoints_world2_gloo(tmp_path):
    world, pairs, people = 2, 4, 3
    mp.start_processes(_worker, args=(world, _free_port(), pairs, people, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


def test_gather_is_identity_without_process_group():
    t = torch.arange(12.0).view(2, 2, 3)
    assert pl.gather_keypoints(t) is t


def test_synthetic_batch_is_seeded_and_consistent():
    a = synthetic.frame_pair_batch(2, 3, (64, 128), (64, 48), seed=9)
    b = synthetic.frame_pair_batch(2, 3, (64, 128), (64, 48), seed=9)
    for k in a:
        assert torch.equal(a[k], b[k])
    assert a["pair"].shape == (2, 3, 2, 64, 128) and a["crops"].shape == (6, 3, 64, 48) and a["prev_keypoints"].shape == (6, 17, 3)
    c, s = a["centers"].numpy(), a["scales"].numpy()
    bx = a["boxes"].numpy()
    np.testing.assert_allclose(c[:, 0], bx[:, [0, 2]].mean(1), rtol=1e-6)
    np.testing.assert_allclose(s, np.maximum(bx[:, 3] - bx[:, 1], (bx[:, 2] - bx[:, 0]) / 48 * 64), rtol=1e-6)   # net_utils.py:49-50
    kp = a["prev_keypoints"].numpy()
    assert (kp[:, :, 0] >= bx[:, None, 0]).all() and (kp[:, :, 0] <= bx[:, None, 2]).all()


def test_flo_writer_bytes_and_roundtrip(tmp_path):
    """Middlebury .flo layout (reference lib/flownet/utils/flow_utils.py:5-55): magic float 202021.25, int32 w, int32 h,
    interleaved (u, v) float32 rows -- checked against hand-packed bytes, then read back."""
    import struct

    import numpy as np
    import torch

    from flowtrack.pytorch_b200.flownet.utils import flow_utils
    u = np.array([[1.5, -2.0, 3.25], [0.0, 7.0, -8.5]], np.float32)
    v = np.array([[10.0, 20.0, 30.0], [-1.0, -2.0, -3.0]], np.float32)
    want = struct.pack("<f", 202021.25) + struct.pack("<ii", 3, 2) + b"".join(struct.pack("<ff", u[r, c], v[r, c]) for r in range(2) for c in range(3))
    assert want[:4] == b"PIEH"
    p = tmp_path / "a.flo"
    flow_utils.writeFlow(str(p), np.stack((u, v), axis=2))
    assert p.read_bytes() == want
    flow_utils.writeFlow(str(p), u, v)
    assert p.read_bytes() == want
    flow_utils.writeFlow(str(p), torch.from_numpy(np.stack((u, v))))          # the network's [2,H,W] output
    assert p.read_bytes() == want
    back = flow_utils.readFlow(str(p))
    assert back.shape == (2, 3, 2) and np.array_equal(back[:, :, 0], u) and np.array_equal(back[:, :, 1], v)
    (tmp_path / "bad.flo").write_bytes(b"\x00" * 16)
    assert flow_utils.readFlow(str(tmp_path / "bad.flo")) is None
