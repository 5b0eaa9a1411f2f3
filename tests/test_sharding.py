"""CPU tests of the multi-GPU host logic: block partition of a batch and the
variable-length gather of point clouds, run with torch.distributed/gloo at world_size 2, plus the
PLY writer of 8(f) n2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stereo_reconstruction_cv_b200 import pointcloud, sharding


def test_shard_ranges_cover_batch():
    for n in (0, 1, 7, 8, 512, 513):
        for world in (1, 2, 3, 4, 8):
            r = sharding.shard_ranges(n, world)
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))                # contiguous, ordered, no overlap
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.shard_ranges(512, 8) == [(64 * g, 64 * (g + 1)) for g in range(8)]
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


class _FakeStereo:
    """Stands in for StereoSGBM on the CPU: 'disparity' = left - right, so the test sees which frames ran."""

    def compute(self, l, r, out=None):
        return l.astype(np.int16) - r.astype(np.int16)


def test_compute_shard_empty_block_and_batched_host_call():
    """More ranks than frames: the surplus ranks get an empty block on BOTH paths (they must reach the
    gather's collectives instead of raising); host blocks go through ONE compute_batch call."""
    calls = []

    class _Batched(_FakeStereo):
        def compute_batch(self, lb, rb, out=None):
            calls.append(lb.shape)
            return lb.astype(np.int16) - rb.astype(np.int16)

    rng = np.random.default_rng(1)
    lefts = [rng.integers(0, 255, (4, 6), dtype=np.uint8) for _ in range(3)]
    rights = [rng.integers(0, 255, (4, 6), dtype=np.uint8) for _ in range(3)]
    got = {}
    for rank in range(5):
        s, e, d = sharding.compute_shard(_Batched(), lefts, rights, 5, rank)
        assert len(d) == e - s
        for i in range(s, e):
            got[i] = d[i - s]
    assert sorted(got) == [0, 1, 2] and len(calls) == 3 and all(c == (1, 4, 6) for c in calls)
    for i in range(3):
        assert np.array_equal(got[i], lefts[i].astype(np.int16) - rights[i].astype(np.int16))
    s, e, d = sharding.compute_shard(_Batched(), lefts, rights, 1, 0)
    assert (s, e) == (0, 3) and calls[-1] == (3, 4, 6)

    class _Dev:                                             # shaped like a CUDA tensor batch, never touched when the block is empty
        is_cuda = True
        shape = (1, 4, 6)
        device = "cpu"

        def dim(self): return 3
        def __len__(self): return 1

    s, e, d = sharding.compute_shard(_FakeStereo(), _Dev(), _Dev(), 4, 0)
    assert s == e == 0 and tuple(d.shape) == (0, 4, 6) and d.dtype == torch.int16


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        lefts = [rng.integers(0, 255, (4, 6), dtype=np.uint8) for _ in range(5)]
        rights = [rng.integers(0, 255, (4, 6), dtype=np.uint8) for _ in range(5)]
        start, stop, disps = sharding.compute_shard(_FakeStereo(), lefts, rights, world, rank)
        assert (start, stop) == sharding.shard_range(5, world, rank) and len(disps) == stop - start
        # ragged "point clouds": one point per frame pixel with positive fake disparity
        pts = [np.argwhere(d > 0).astype(np.float32) for d in disps]
        xyz = torch.from_numpy(np.concatenate([np.c_[p, np.full(len(p), start + i, np.float32)] for i, p in enumerate(pts)]))
        rgb = torch.full((xyz.shape[0], 3), rank + 1, dtype=torch.uint8)
        gx, gc = sharding.gather_point_cloud(xyz, rgb, dst=0)
        if rank == 0:
            ref = []
            for i in range(5):
                d = lefts[i].astype(np.int16) - rights[i].astype(np.int16)
                p = np.argwhere(d > 0).astype(np.float32)
                ref.append(np.c_[p, np.full(len(p), i, np.float32)])
            ref = np.concatenate(ref)
            assert np.array_equal(gx.numpy(), ref), "gathered cloud differs from the single-process result"
            assert gc.shape[0] == ref.shape[0] and set(np.unique(gc.numpy())) <= {1, 2}
            open(os.path.join(tmp, "ok"), "w").write("%d" % ref.shape[0])
        else:
            assert gx is None and gc is None
    finally:
        dist.destroy_process_group()


def test_gloo_world2_shard_and_gather(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert int(open(tmp_path / "ok").read()) > 0


def test_gather_without_process_group():
    t = torch.arange(6, dtype=torch.float32).reshape(2, 3)
    assert sharding.gather_varlen(t)[0] is t


def test_ply_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    xyz = rng.normal(size=(4, 5, 3)).astype(np.float32)
    xyz[0, 0, 0] = np.inf
    xyz[1, 2, 2] = -np.inf
    rgb = rng.integers(0, 256, (4, 5, 3), dtype=np.uint8)
    n = pointcloud.write_ply(tmp_path / "a.ply", xyz, rgb)
    assert n == 18
    x, c = pointcloud.read_ply(tmp_path / "a.ply")
    ok = np.isfinite(xyz.reshape(-1, 3)).all(1)
    assert x.dtype == np.float64 and np.array_equal(x, xyz.reshape(-1, 3)[ok].astype(np.float64))
    assert np.array_equal(c, rgb.reshape(-1, 3)[ok])
    # the reference's second call site writes every pixel, non-finite included (main.ipynb:796)
    assert pointcloud.write_ply(tmp_path / "b.ply", xyz, keep_nonfinite=True, dtype=np.float32) == 20
    x2, c2 = pointcloud.read_ply(tmp_path / "b.ply")
    assert c2 is None and x2.dtype == np.float32 and np.array_equal(np.isfinite(x2), np.isfinite(xyz.reshape(-1, 3)))
    pts, col = pointcloud.open3d_arrays(xyz, rgb)
    assert pts.dtype == np.float64 and col.dtype == np.float64 and col.max() <= 1.0
    hdr = open(tmp_path / "a.ply", "rb").read(200).decode("ascii", "ignore")
    assert hdr.startswith("ply\nformat binary_little_endian 1.0") and "property double x" in hdr


# ------------------------------------------------------------------------------------------------
# The optional NCCL exchange of SURVEY 8(e) on real GPUs: runs when >= 2 CUDA devices are visible
# ------------------------------------------------------------------------------------------------
def _nccl_worker(rank, world, port, tmp):
    import stereo_reconstruction_cv_b200 as sg
    from synth import make_pair
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        W, H, D, n_frames = 640, 200, 64, 5
        Q = np.array([[1, 0, 0, -W / 2], [0, 1, 0, -H / 2], [0, 0, 0, 500.0], [0, 0, -1, 0]], np.float64)
        pairs = [make_pair(W, H, D, seed=70 + i)[:2] for i in range(n_frames)]
        lefts = torch.from_numpy(np.stack([p[0] for p in pairs])).cuda()
        rights = torch.from_numpy(np.stack([p[1] for p in pairs])).cuda()
        st = sg.StereoSGBM_create(minDisparity=0, numDisparities=D, blockSize=5, P1=200, P2=800, disp12MaxDiff=1, preFilterCap=63,
                                  uniquenessRatio=10, speckleWindowSize=100, speckleRange=32, mode=2)
        start, stop, disp = sharding.compute_shard(st, lefts, rights, world, rank)       # this rank's block, on this rank's GPU
        clouds = [sg.reprojectCompact(disp[i], Q, None)[0] for i in range(stop - start)]
        xyz = torch.cat(clouds, 0) if clouds else torch.empty((0, 3), dtype=torch.float32, device="cuda")
        gx, _ = sharding.gather_point_cloud(xyz, None, dst=0)                              # NCCL: counts, then padded gather
        if rank == 0:
            ref = torch.cat([sg.reprojectCompact(st.compute(lefts[i], rights[i]), Q, None)[0] for i in range(n_frames)], 0)
            assert gx.is_cuda and gx.shape == ref.shape and bool((gx.view(torch.int32) == ref.view(torch.int32)).all()), \
                "gathered cloud differs from the single-GPU cloud in global frame order"
            open(os.path.join(tmp, "ok"), "w").write("%d" % ref.shape[0])
        else:
            assert gx is None
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_nccl_world2_shard_and_gather(tmp_path):
    """Frames block-sharded over two GPUs, clouds gathered over NCCL: equal to the one-GPU result (main.ipynb:726-737
    for what is gathered).  Needs two visible devices; on a one-GPU box the gloo test above covers the host logic."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices (run with gpurun --gpus 2)")
    port = _free_port()
    mp.spawn(_nccl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert int(open(tmp_path / "ok").read()) > 1000


@pytest.mark.gpu
def test_one_process_multi_gpu_batch():
    """sharding.compute_batch_devices: one process, one host thread and one handle per GPU, block-partitioned batch, no
    collective.  Every frame equals the oracle; with two visible devices both are used (and the single-device call agrees)."""
    import oracle
    from oracle import OracleParams
    import stereo_reconstruction_cv_b200 as sg  # noqa: F401
    from synth import make_pair
    W, H, D, B = 900, 120, 64, 7
    pairs = [make_pair(W, H, D, seed=300 + i)[:2] for i in range(B)]
    ls, rs = np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs])
    kw = dict(minDisparity=0, numDisparities=D, blockSize=5, P1=200, P2=800, disp12MaxDiff=1, preFilterCap=63, uniquenessRatio=10,
              speckleWindowSize=100, speckleRange=32, mode=1)
    ref = np.stack([oracle.compute(OracleParams(**kw), l, r) for l, r in pairs])
    ndev = torch.cuda.device_count()
    assert ndev >= 1
    got = sharding.compute_batch_devices(kw, ls, rs)                       # every visible device
    assert got.shape == (B, H, W) and got.dtype == np.int16 and np.array_equal(got, ref)
    assert np.array_equal(sharding.compute_batch_devices(kw, ls, rs, devices=[ndev - 1]), ref)
    if ndev >= 2:                                                           # more devices than frames: empty blocks are fine
        assert np.array_equal(sharding.compute_batch_devices(kw, ls[:1], rs[:1], devices=[0, 1]), ref[:1])
    with pytest.raises(sg.error):                                           # errors of a worker thread reach the caller
        sharding.compute_batch_devices(dict(kw, numDisparities=2048), ls, rs)
