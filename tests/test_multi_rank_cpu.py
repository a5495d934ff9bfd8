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


def _gather_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from underwaterworld_b200.gather import gather_meshes
    rng = np.random.default_rng(rank)
    n_chunks, n_verts, n_inds = 5 + rank, 100 * (rank + 1), 0 if rank == 1 else 300     # ragged, one rank has no indices
    d = torch.from_numpy(rng.integers(0, 255, n_chunks * 32, dtype=np.uint8))
    v = torch.from_numpy(rng.integers(0, 255, n_verts * 24, dtype=np.uint8))
    i = torch.from_numpy(rng.integers(0, 255, n_inds * 2, dtype=np.uint8))
    got = gather_meshes(d, v, i, dst=0)
    if rank == 0:
        assert got is not None and len(got) == world
        for r, (gd, gv, gi) in enumerate(got):
            rr = np.random.default_rng(r)
            nc, nv, ni = 5 + r, 100 * (r + 1), 0 if r == 1 else 300
            assert np.array_equal(gd.numpy(), rr.integers(0, 255, nc * 32, dtype=np.uint8))
            assert np.array_equal(gv.numpy(), rr.integers(0, 255, nv * 24, dtype=np.uint8))
            assert np.array_equal(gi.numpy(), rr.integers(0, 255, ni * 2, dtype=np.uint8))
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_mesh_gather_to_render_rank_gloo(tmp_path):
    """Optional gather of finished meshes to the rendering rank: sizes by all_gather, ragged
    payloads by send/recv; here over gloo with host tensors, on the GPU box over NCCL / NVLink."""
    mp.spawn(_gather_worker, args=(3, _free_port(), str(tmp_path)), nprocs=3, join=True)
    assert (tmp_path / "ok").exists()
