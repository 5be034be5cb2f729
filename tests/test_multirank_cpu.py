"""N > 1 host path on CPU with gloo ranks: sequence / segment sharding (`seq_idx % n_gpus == select_idx`, script :338) with
the single all-gather that stitches the clip, and ONE clip sharded over ranks at unit granularity (segment x VAE tile,
SURVEY.md §8e) through `VSRPipeline(..., world_size, rank)` — the N-rank clip must equal the single-process clip."""
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


def _unit_tile(u, shape):
    g = torch.Generator().manual_seed(500 + u)
    return torch.randn(*shape, generator=g)


def _gather_units_worker(rank, world, port, shapes, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from mgld_vsr_b200.pipeline import gather_units, shard_units
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_units(len(shapes), world, rank)
    tiles = gather_units([_unit_tile(u, shapes[u]) for u in range(lo, hi)], shapes, world, rank, "cpu")
    q.put((rank, [t.numpy().copy() for t in tiles]))        # by value: the worker may exit before the parent reads
    dist.barrier()
    dist.destroy_process_group()


def test_gather_units_blocks_and_mixed_shapes():
    """5 units of two tile shapes on 3 ranks (blocks of 1/2/2), and 2 units on 3 ranks (a rank without work)"""
    for shapes in ([(2, 3, 4, 6)] * 3 + [(2, 3, 4, 5)] * 2, [(2, 3, 4, 6)] * 2):
        world = 3
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_gather_units_worker, args=(r, world, port, shapes, q)) for r in range(world)]
        for p in procs:
            p.start()
        res = dict(q.get(timeout=120) for _ in range(world))
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        for r in range(world):
            assert len(res[r]) == len(shapes)
            for u, t in enumerate(res[r]):
                assert torch.equal(torch.from_numpy(t), _unit_tile(u, shapes[u]))


def _clip_worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p_ in (root, os.path.join(root, "tests")):
        if p_ not in sys.path:
            sys.path.insert(0, p_)
    torch.set_num_threads(2)
    import test_batching_cpu as B
    from common import det_tensor
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
    pipe = B._tiny_pipeline(2)
    ctx = det_tensor("ctx", (1, 77, 128))
    T, H, W = B.T, 160, 256                            # 2 VAE tiles per segment (vqgantile 160 / stride 96)
    segs = [det_tensor(f"mr_seg{k}", (T, 3, H, W)).clamp(-1, 1) for k in range(3)]
    flows = [tuple(f[0] for f in B._flows(f"mr{k}", H // 8, W // 8)[:2]) for k in range(3)]
    pipe.upsample_scale = pipe.upscale
    pipe.segments = lambda frames: (segs, 3 * T - 1)    # 3 segments (last frame is padding) -> 6 units
    out = pipe(torch.zeros(3 * T - 1, 3, H // 4, W // 4), context=ctx, flows_override=flows, world_size=world, rank=rank)
    q.put((rank, out.numpy().copy()))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def test_one_clip_sharded_over_ranks_equals_single_process():
    """6 units (3 segments x 2 VAE tiles) on 1, 2 and 4 ranks (4 ranks: blocks of 1/2/1/2 units, segments split across
    ranks): every rank ends up with the whole clip, equal to the single-process clip up to the fp16 rounding of a different
    lock-step batch composition (units are batched in pairs)."""
    results = {}
    for world in (1, 2, 4):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_clip_worker, args=(r, world, port, q)) for r in range(world)]
        for p in procs:
            p.start()
        res = {r_: torch.from_numpy(o) for r_, o in (q.get(timeout=600) for _ in range(world))}
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        for r in range(1, world):
            assert torch.equal(res[r], res[0])                                    # every rank holds the same clip
        results[world] = res[0]
    assert results[1].shape == (5, 3, 160, 256)
    for world in (2, 4):
        d = (results[world] - results[1]).abs()         # units are batched in pairs through the nets AND the VAE passes: the
        assert d.max() < 1e-2 and d.mean() < 3e-4, (world, d.max(), d.mean())   # partner of a unit changes with the sharding
