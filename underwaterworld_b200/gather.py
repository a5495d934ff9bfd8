"""Optional gather of finished meshes to the rendering GPU.

Chunks are independent, so the compute path has NO collective; the only cross-GPU traffic a renderer may
want is "bring every rank's packed mesh to the GPU that draws".  With the NCCL backend the point-to-point
transfers below run over NVLink / NVSwitch (peer copies measured at ~770 GB/s per direction on this pool);
with gloo the same code moves host tensors (CPU tests).

    sizes  = all_gather([n_chunks, n_verts, n_inds])            24 bytes per rank
    rank r = send(descs), send(verts), send(inds)  ->  dst      variable length, no padding

`device_batch_tensors` wraps the library's device arenas as torch uint8 tensors without copying
(`__cuda_array_interface__`), so the sends read the kernel's output buffers directly.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from ._ffi import DESC_DTYPE, VERT_DTYPE


class _DevMem:
    """Zero-copy view of raw device memory for torch.as_tensor."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}


def device_batch_tensors(builder) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(descs, verts, inds) of the builder's last device-resident build as uint8 CUDA tensors (no copy).
    Call after builder.sync(); valid until the next build on that builder."""
    v = builder.device_view()
    isz = 4 if v.d_inds32 else 2
    iptr = v.d_inds32 or v.d_inds16

    def wrap(ptr, nbytes):
        if nbytes == 0 or not ptr:
            return torch.empty(0, dtype=torch.uint8, device="cuda")
        return torch.as_tensor(_DevMem(ptr, nbytes), device="cuda")

    return (wrap(v.d_descs, v.n_chunks * DESC_DTYPE.itemsize), wrap(v.d_verts, v.n_verts * VERT_DTYPE.itemsize),
            wrap(iptr, v.n_inds * isz))


def gather_meshes(descs: torch.Tensor, verts: torch.Tensor, inds: torch.Tensor, dst: int = 0,
                  group: Optional[dist.ProcessGroup] = None) -> Optional[List[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]]:
    """Bring every rank's (descs, verts, inds) byte tensors to rank `dst`.

    Returns, on `dst`, a list indexed by source rank (its own entry aliases the inputs); None elsewhere.
    Descriptor offsets stay relative to the source rank's own vertex / index arrays."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = descs.device
    mine = torch.tensor([descs.numel(), verts.numel(), inds.numel()], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(3, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, mine, group=group)
    ops = []
    out = None
    if rank != dst:
        ops = [dist.P2POp(dist.isend, t.contiguous(), dst, group) for t in (descs, verts, inds) if t.numel()]
    else:
        out = []
        for r in range(world):
            if r == dst:
                out.append((descs, verts, inds))
                continue
            bufs = [torch.empty(int(sizes[r][k].item()), dtype=torch.uint8, device=dev) for k in range(3)]
            ops += [dist.P2POp(dist.irecv, b, r, group) for b in bufs if b.numel()]
            out.append(tuple(bufs))
    if ops:                                  # one batched group: the transfers run concurrently over NVLink
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out


def as_numpy_batch(descs: torch.Tensor, verts: torch.Tensor, inds: torch.Tensor, index32: bool = False):
    """Host numpy views (structured dtypes) of a gathered (descs, verts, inds) byte triple."""
    d = descs.cpu().numpy().view(DESC_DTYPE)
    v = verts.cpu().numpy().view(VERT_DTYPE)
    i = inds.cpu().numpy().view(np.uint32 if index32 else np.uint16)
    return d, v, i
