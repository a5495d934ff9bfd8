"""CPU test of the N>1 path (world_size 2, gloo): every rank takes its contiguous x-slab of a region
(underwaterworld_b200.region.shard_region) with no data-path collective; the slabs tile the region in
order, and the bench's max-over-ranks reduction works over torch.distributed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from underwaterworld_b200 import region


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    box = ((-8, 9), (-3, 3), (-4, 4))                      # 17 x-columns: an uneven split
    mine = region.shard_region(*box, rank, world)
    # sizes are exchanged only to check the partition -- the compute path itself needs no collective
    n = torch.tensor([len(mine)], dtype=torch.int64)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, n)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)   # bench.py: ms_per_step = MAX over ranks
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    np.save(os.path.join(out_dir, f"slab{rank}.npy"), mine)
    if rank == 0:
        np.save(os.path.join(out_dir, "meta.npy"), np.array([int(s.item()) for s in sizes] + [int(t.item())]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_slab_partition_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    slabs = [np.load(tmp_path / f"slab{r}.npy") for r in range(world)]
    meta = np.load(tmp_path / "meta.npy")
    whole = region.box_region((-8, 9), (-3, 3), (-4, 4))
    assert np.array_equal(np.concatenate(slabs), whole)            # contiguous, ordered, nothing lost or doubled
    assert [len(s) for s in slabs] == meta[:2].tolist() == [9 * 6 * 8, 8 * 6 * 8]
    assert meta[2] == world                                        # max over ranks
    assert slabs[0][:, 0].max() < slabs[1][:, 0].min()             # cut along x, never along z
    for s in slabs:
        assert set(np.unique(s[:, 2]).tolist()) == set(range(-4, 4))


def test_weak_scaling_regions_are_disjoint():
    a, b = region.weak_region(16, (-8, 8), (-4, 4), 0), region.weak_region(16, (-8, 8), (-4, 4), 1)
    assert len(a) == len(b) == 2048
    assert not set(map(tuple, a)) & set(map(tuple, b))


def _info_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from underwaterworld_b200 import _ffi, gather
    info = None
    if rank == 1:                                           # the rendering rank need not be rank 0
        info = _ffi.UwGatherInfo()
        info.abi_version, info.n_segments, info.device, info.index_bytes = 2, world, 1, 2
        info.owner_pid, info.base, info.bytes = 4242, 0x7F0000000000, 1 << 30
        info.off_descs, info.off_verts, info.off_inds = 256, 1 << 20, 1 << 29
        info.n_chunks, info.seg_vcap, info.seg_icap = 524288, 12_000_000, 42_000_000
        for k in range(64):
            info.ipc_handle[k] = (7 * k + 1) & 0xFF
    got = gather.broadcast_info(info, src=1)
    # every rank derives its slab of the request from (n, world, rank) alone: no exchange on the data path
    first, count = gather.slab_bounds(int(got.n_chunks), world, rank)
    np.save(os.path.join(out_dir, f"info{rank}.npy"), np.frombuffer(gather.info_to_bytes(got), dtype=np.uint8))
    np.save(os.path.join(out_dir, f"slab{rank}.npy"), np.array([first, count]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_info_reaches_every_rank_gloo(tmp_path):
    """Setup of the one-sided mesh gather: the rendering rank's uw_gather_info (arena addresses, capacities,
    CUDA IPC handle -- plain bytes) is broadcast once; afterwards every rank knows its segment and its slab of the
    request.  Here over gloo; bench.py does the same under NCCL."""
    world = 3
    mp.spawn(_info_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    blobs = [np.load(tmp_path / f"info{r}.npy") for r in range(world)]
    assert all(np.array_equal(b, blobs[1]) for b in blobs)
    from underwaterworld_b200 import _ffi, gather
    info = gather.info_from_bytes(blobs[0].tobytes())
    assert (info.n_segments, info.device, info.owner_pid, info.n_chunks) == (3, 1, 4242, 524288)
    assert bytes(info.ipc_handle) == bytes((7 * k + 1) & 0xFF for k in range(64))
    slabs = [np.load(tmp_path / f"slab{r}.npy") for r in range(world)]
    assert slabs[0][0] == 0 and all(slabs[r][0] + slabs[r][1] == slabs[r + 1][0] for r in range(world - 1))
    assert slabs[-1][0] + slabs[-1][1] == 524288


def test_gather_structs_match_the_header():
    """ctypes mirrors of uw_gather_info / uw_gather_result have the C layout (no GPU needed: sizes follow from the
    header's field list under the C ABI's natural alignment)."""
    import ctypes as C
    from underwaterworld_b200 import _ffi
    assert C.sizeof(_ffi.UwGatherInfo) == 16 + 8 * 11 + 64
    assert C.sizeof(_ffi.UwGatherSegment) == 48
    assert C.sizeof(_ffi.UwGatherResult) == 8 + 24 + 24 + 16 + 8 + 24 + 48 * _ffi.UW_MAX_SEGMENTS
    assert _ffi.UwGatherInfo.ipc_handle.offset == 104


def _tune_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from underwaterworld_b200 import gather
    rg = gather.RegionGather.__new__(gather.RegionGather)      # control-plane logic only: no builder, no GPU
    rg.world, rank_, rg.dst, rg.group, rg.render_share = 8, rank, 0, None, 0.0
    rg.rank = rank_
    calls = []

    def run_step():
        # a step is as slow as its slowest rank: the rendering rank's compute grows with its share, the producers'
        # ingress-bound stores shrink with it; rank 1 plays "slowest producer"
        calls.append(rg.render_share)
        return 3.25 * rg.render_share if rank == 0 else 1.03 * (1.0 - rg.render_share)

    best = rg.tune(run_step)
    np.save(os.path.join(out_dir, f"tune{rank}.npy"), np.array([best, len(calls)] + rg.candidate_shares()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_render_share_search_picks_the_fastest_candidate_gloo(tmp_path):
    """RegionGather.tune (bench.py's warm-up): every rank measures whole steps under each candidate share, the step time
    is the MAX over ranks, the fastest share wins on every rank alike."""
    world = 2
    mp.spawn(_tune_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a, b = np.load(tmp_path / "tune0.npy"), np.load(tmp_path / "tune1.npy")
    cands = a[2:].tolist()
    assert cands == b[2:].tolist() and cands[0] == 0.125 and max(cands) <= 0.5 and len(cands) == 7
    assert a[0] == b[0]                                            # same decision everywhere
    model = lambda s: max(3.25 * s, 1.03 * (1.0 - s))
    coarse = cands[int(np.argmin([model(s) for s in cands]))]      # 0.21875 for these constants (optimum 0.2407)
    half = 0.5 * (cands[1] - cands[0])
    assert a[0] == min([coarse, coarse - half, coarse + half], key=model) == 0.234375          # refined towards it
    assert a[1] == b[1] == 2 + 2 * (len(cands) + 2)                # two cold steps, two per candidate, two refinements
