"""Frame-parallel sharding (vaw_shard_range) and the N > 1 host logic on CPU with gloo, world_size 2.

Frames are independent once each has its rotation (FrameSourceWarp.cpp:272-314), so ranks take
contiguous frame ranges and exchange nothing on the data path; the only collectives are the
barrier and the max-over-ranks of the timing that bench.py uses.  Here two CPU ranks warp their
ranges with the oracle and the gathered checksums must equal a single-process pass."""
import os
import socket
import zlib

import numpy as np
import pytest


def test_shard_ranges_partition_the_clip():
    import video_annotator_b200 as V
    for n in (0, 1, 7, 64, 600, 601):
        for parts in (1, 2, 3, 4, 8):
            nxt = 0
            sizes = []
            for p in range(parts):
                first, count = V.shard_range(n, parts, p)
                assert first == nxt and count >= 0
                nxt = first + count
                sizes.append(count)
            assert nxt == n and max(sizes) - min(sizes) <= 1
            assert sizes == sorted(sizes, reverse=True)       # the remainder goes to the first parts
    assert V.shard_range(600, 8, 3) == (225, 75)               # C4: 600 frames over 8 GPUs
    with pytest.raises(V.VawError):
        V.shard_range(10, 2, 2)


def test_rank_rotations_are_slices_of_the_clip():
    from video_annotator_b200 import configs
    w = configs.workload("C2")
    whole = w.rotations(40, first=10, total=50)
    for parts in (2, 4):
        import video_annotator_b200 as V
        got = [w.rotations(c, first=10 + f, total=50) for f, c in (V.shard_range(40, parts, p) for p in range(parts))]
        assert np.array_equal(np.concatenate(got), whole)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_main(rank, world, port, n_frames, out_queue):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import video_annotator_b200 as V
    from oracle import oracle as O
    from tests.test_shard import _clip_setup, _warp_frame
    k, rots, (sw, sh), (ow, oh) = _clip_setup(O, n_frames)
    first, count = V.shard_range(n_frames, world, rank)
    sums = torch.zeros(n_frames, dtype=torch.int64)
    for i in range(first, first + count):
        sums[i] = _warp_frame(O, k, rots[i], i, sw, sh, ow, oh)
    dist.all_reduce(sums)                       # test-only gather of the per-frame checksums
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)    # the max-over-ranks bench.py takes of the timing
    dist.barrier()
    if rank == 0:
        out_queue.put((sums.tolist(), float(t.item()), (first, count)))
    dist.destroy_process_group()


def _clip_setup(O, n_frames):
    from video_annotator_b200 import configs
    sw, sh, ow, oh = 96, 64, 80, 48
    K_in = np.array([[48.0, 0, 47.3], [0, 47.5, 31.6], [0, 0, 1]])
    K_out = np.array([[30.0, 0, 39.5], [0, 30.0, 23.5], [0, 0, 1]])
    k = O.intrinsics(K_in, K_out)
    rots = configs.make_rotations(n_frames, 1.0, radius=4)
    return k, rots, (sw, sh), (ow, oh)


def _warp_frame(O, k, rot, index, sw, sh, ow, oh):
    src = O.synth_nv12(sw, sh, index)
    return zlib.crc32(O.warp_nv12(src, sw, sh, ow, oh, k, rot).tobytes())


def test_two_ranks_gloo_equal_single_process(oracle):
    import torch.multiprocessing as mp
    n_frames, world = 9, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    sums, tmax, r0 = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    k, rots, (sw, sh), (ow, oh) = _clip_setup(oracle, n_frames)
    want = [_warp_frame(oracle, k, rots[i], i, sw, sh, ow, oh) for i in range(n_frames)]
    assert sums == want
    assert tmax == 2.0 and r0 == (0, 5)
