"""N > 1 host path on CPU: segment sharding (`seq_idx % n_gpus == select_idx`, script :338, applied to segments) and the
single all-gather that stitches the clip, with 2 gloo ranks."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _segment_frames(s, T=5, H=8, W=12):
    g = torch.Generator().manual_seed(100 + s)
    return torch.randint(0, 256, (T, 3, H, W), generator=g, dtype=torch.uint8)


def _worker(rank, world, port, num_segments, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from mgld_vsr_b200.pipeline import gather_clip, shard_segments
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_segments(num_segments, world, rank)
    local = {s: _segment_frames(s) for s in mine}
    clip = gather_clip(local, num_segments, 5, world, rank, frame_shape=(3, 8, 12), device="cpu")   # every rank enters
    q.put((rank, mine, clip))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather_matches_single_process():
    num_segments, world = 7, 2           # BASELINE config 3: 32 frames -> 7 segments; odd count exercises the padding slot
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, num_segments, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        rank, mine, clip = q.get(timeout=120)
        res[rank] = (mine, clip)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][0] == [0, 2, 4, 6] and res[1][0] == [1, 3, 5]
    ref = torch.cat([_segment_frames(s) for s in range(num_segments)], 0)
    for r in range(world):
        assert torch.equal(res[r][1], ref)


def test_fewer_segments_than_ranks():
    """2 segments on 3 ranks (BASELINE config 2 on more GPUs than segments): the rank without work still contributes its
    zero slot to the collective instead of hanging the others."""
    num_segments, world = 2, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, num_segments, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        rank, mine, clip = q.get(timeout=120)
        res[rank] = (mine, clip)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[2][0] == []
    ref = torch.cat([_segment_frames(s) for s in range(num_segments)], 0)
    for r in range(world):
        assert torch.equal(res[r][1], ref)


def test_shard_segments_partition():
    from mgld_vsr_b200.pipeline import shard_segments
    for n, w in [(2, 8), (13, 8), (4, 4), (7, 2), (1, 1)]:
        parts = [shard_segments(n, w, r) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
