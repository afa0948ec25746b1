"""N > 1 host logic on CPU (gloo, world_size 2): frame-chunk partitioning with the global
keyframe parity, the all-gather of the 40-byte-per-pair records, and the sequential trajectory
over the gathered table must reproduce the single-process result exactly."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_measurement(f):
    """Deterministic stand-in for AlignNextFrame's result of frame f (no GPU here)."""
    rng = np.random.default_rng(1000 + f)
    ok = 0.0 if (f == 0 or f % 41 == 40) else 1.0
    m = rng.normal(0, 1, 4) * np.array([0.002, 0.002, 5.0, 5.0])
    return np.array([*(m if f > 0 else np.zeros(4)), ok])


def _worker(rank, world, port, n_frames, out_dir):
    import torch.distributed as dist
    from video_stabilizer_b200 import partition
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    chunks = partition.frame_chunks(n_frames, world)
    first, last = chunks[rank]
    local = np.stack([_fake_measurement(f) for f in range(first, last)]) if last > first else np.zeros((0, 5))
    table = partition.gather_measurements(local, chunks[rank], n_frames)
    corr = partition.corrections_for_video(table, 1920, 1080)
    np.save(os.path.join(out_dir, "corr_%d.npy" % rank), corr)
    np.save(os.path.join(out_dir, "table_%d.npy" % rank), table)
    dist.destroy_process_group()


def test_frame_chunks_preserve_parity_and_cover_the_video():
    from video_stabilizer_b200 import partition
    for n in (1, 2, 7, 300, 301, 2400):
        for world in (1, 2, 4, 8):
            chunks = partition.frame_chunks(n, world)
            assert chunks[0][0] == 0 and chunks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(chunks, chunks[1:]))
            assert all(c[0] % 2 == 0 for c in chunks)
            # the union of per-chunk pairs is exactly the single-process pair list, in order
            allp = []
            for r, c in enumerate(chunks):
                up0, _ = partition.chunk_upload_range(c, r)
                pairs, keys = partition.chunk_pairs(c, r)
                for (t, k, inv) in pairs:
                    allp.append((t + up0, k + up0, inv))
                    assert (k + up0) % 2 == 1 and (t + up0) % 2 == 0      # keyframes are the odd frames
                assert all((s + up0) % 2 == 1 for s in keys)
            single, _ = partition.chunk_pairs((0, n), 0)
            assert allp == single


def test_pairs_match_clip_module():
    from video_stabilizer_b200 import partition
    from video_stabilizer_b200.clip import pairs_for_frames
    pairs, keys = pairs_for_frames(0, 9)
    mine, mykeys = partition.chunk_pairs((0, 9), 0)
    assert [(p.template_slot, p.keyframe_slot, p.invert) for p in pairs] == mine and keys == mykeys


def test_two_rank_gather_and_trajectory_equal_single_process(tmp_path):
    import torch.multiprocessing as mp
    from video_stabilizer_b200 import partition
    n_frames, world = 123, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_frames, str(tmp_path)), nprocs=world, join=True)
    table = np.stack([_fake_measurement(f) for f in range(n_frames)])
    want = partition.corrections_for_video(table, 1920, 1080)
    assert want.shape == (n_frames - 10, 4)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("table_%d.npy" % r)), table)
        assert np.array_equal(np.load(tmp_path / ("corr_%d.npy" % r)), want)


# ---------------------------------------------------------------- the C++ partitioned trajectory (host/partitioned.hpp)
def _sequential_corrections(table, w, h, params=None):
    """frame-by-frame StabilizerTrajectory (stabilizer.cpp:19-88) over a measurement table -> {frame: correction}"""
    from video_stabilizer_b200 import host
    tr = host.StabilizerTrajectory(params)
    lag = tr.params.lag
    out = {}
    for n in range(len(table)):
        due, corr = tr.push(table[n, :4], table[n, 4] != 0, w, h)
        if due:
            out[n - lag] = corr
    return out


def _part_worker(rank, world, name, n_frames, sub, block, videos, threads, out_dir):
    from video_stabilizer_b200 import host
    host.load()
    pt = host.PartitionedTrajectory(rank, world, 1920, 1080, n_frames, sub, block, None, name, threads)
    for v in range(videos):
        table = np.stack([_fake_measurement(f + 1000 * v) for f in range(n_frames)])
        table[0] = 0.0
        frames, corr = pt.run(table[:, :4], table[:, 4] != 0)
        np.save(os.path.join(out_dir, "pf_%d_%d.npy" % (rank, v)), frames)
        np.save(os.path.join(out_dir, "pc_%d_%d.npy" % (rank, v)), corr)


@pytest.mark.parametrize("n_frames,sub,block", [(300, 100, 3), (300, 32, 1), (123, 10, 1), (57, 20, 2), (11, 12, 1)])
def test_partitioned_trajectory_single_worker_equals_frame_by_frame(tmp_path, n_frames, sub, block):
    """world = 1: smoothing on a worker pool + the chain == StabilizerTrajectory::push frame by frame, bit for bit."""
    _part_worker(0, 1, "", n_frames, sub, block, 2, 3, str(tmp_path))
    for v in range(2):
        table = np.stack([_fake_measurement(f + 1000 * v) for f in range(n_frames)])
        table[0] = 0.0
        want = _sequential_corrections(table, 1920, 1080)
        frames, corr = np.load(tmp_path / ("pf_0_%d.npy" % v)), np.load(tmp_path / ("pc_0_%d.npy" % v))
        assert list(frames) == sorted(want)
        for f, c in zip(frames, corr):
            assert np.array_equal(c, want[int(f)]), f


@pytest.mark.parametrize("n_frames,sub,block,world", [(600, 100, 3, 2), (600, 32, 1, 2), (250, 10, 1, 3), (64, 32, 1, 2)])
def test_partitioned_trajectory_processes_equal_frame_by_frame(tmp_path, n_frames, sub, block, world):
    """Several worker PROCESSES sharing the table in POSIX shared memory (no collective): every output frame's correction
    equals the frame-by-frame trajectory bit for bit, over three videos in a row (the table alternates between two copies)."""
    import torch.multiprocessing as mp
    name = "/vstab_test_%d_%d" % (os.getpid(), n_frames * 100 + sub)
    videos = 3
    mp.spawn(_part_worker, args=(world, name, n_frames, sub, block, videos, 2, str(tmp_path)), nprocs=world, join=True)
    for v in range(videos):
        table = np.stack([_fake_measurement(f + 1000 * v) for f in range(n_frames)])
        table[0] = 0.0
        want = _sequential_corrections(table, 1920, 1080)
        got = {}
        for r in range(world):
            frames, corr = np.load(tmp_path / ("pf_%d_%d.npy" % (r, v))), np.load(tmp_path / ("pc_%d_%d.npy" % (r, v)))
            for f, c in zip(frames, corr):
                assert int(f) not in got
                got[int(f)] = c
        assert sorted(got) == sorted(want)
        for f in want:
            assert np.array_equal(got[f], want[f]), (v, f)
