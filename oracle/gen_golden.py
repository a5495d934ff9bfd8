#!/usr/bin/env python3
"""Generate tests/golden/*.npz -- run in the BUILD CONTAINER ONLY (needs /root/reference).

TEST INFRASTRUCTURE.  Two kinds of vectors are produced:

``ref_wasm_*.npz`` -- outputs of the REFERENCE ITSELF: functions of the reference's shipped
    binary (builds/web_build.zip : pkg/underwater_world_bg.wasm) executed by
    oracle/wasm_forensics.py.  These pin the oracle.
      * perlin3   : wasm func 466 = noise-0.8.2 core::perlin::perlin_3d  (f64, bit-exact)
      * perm      : the PermutationTable construction inlined in State::new
                    (func 381, instructions 25632..25966)                (u8, bit-exact)
      * chunks    : wasm func 397 = Chunk::build_partial, driven to completion on a Chunk
                    laid out exactly as the inlined Chunk::new does (func 842 @3773-3836);
                    execution is stopped at the first wgpu buffer-creation call, where
                    build.isos / build.verts / build.inds are read from linear memory.
                    NOTE the shipped binary was compiled with INTERNAL_SIZE = 10
                    (SURVEY.md §8c), so these vectors are S=10; the oracle is parametric
                    in S and is compared at S=10.
``oracle_kat_s12.npz`` -- known-answer vectors of the (pinned) oracle at HEAD's S=12, frozen so
    that any later drift of the oracle, or of the GPU path, is caught without the reference.

Usage:  python oracle/gen_golden.py [--skip-chunks]
"""
from __future__ import annotations

import argparse
import hashlib
import os
import random
import struct
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import Oracle, MODE_FAITHFUL  # noqa: E402
from oracle import wasm_forensics as wf  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# function indices inside the shipped wasm
F_PERLIN3 = 466
F_BUILD_PARTIAL = 397
F_STATE_NEW = 381
F_RUST_ALLOC = 2713
PERM_REGION = (25632, 25966)   # instruction indices inside func 381; local 21 = seed, local 4 = Vec<u8> ptr
BUFFER_CREATE_FUNCS = (978, 532, 2009, 1722)   # first calls on the create_buffer_init path in func 397
HASHBROWN_EMPTY_CTRL = 2636960   # static empty control group stored by the inlined Chunk::new


class _Stop(Exception):
    pass


def wasm_perlin3(m, perm, pts):
    inst = wf.Instance(m)
    base = m.mem_min * 65536 + 4096
    inst.write(base, bytes(np.asarray(perm, dtype=np.uint8).tolist()))
    out = np.empty(len(pts), dtype=np.float64)
    for i, (x, y, z) in enumerate(pts):
        inst.write(base + 512, struct.pack("<3d", x, y, z))
        out[i] = inst.invoke(F_PERLIN3, [base, base + 512])[0]
    return out


def wasm_perm(m, seed):
    inst = wf.Instance(m)
    heap = m.mem_min * 65536 + 8192
    inst.call_hook = lambda fidx, args: [heap] if fidx == F_RUST_ALLOC else None
    inst.globals[0] = m.mem_min * 65536 + 4096
    loc = inst.run_region(F_STATE_NEW, PERM_REGION[0], PERM_REGION[1], {21: seed & 0xFFFFFFFF})
    return np.frombuffer(inst.read(loc[4], 256), dtype=np.uint8).copy()


def wasm_build_chunk(m, perm, pos):
    """Drive Chunk::build_partial (func 397) like Chunk::build_full does (chunk.rs:266-268)."""
    S = 10
    n_iso = (S + 1) ** 3
    inst = wf.Instance(m)
    base = m.mem_min * 65536 + 65536
    inst.globals[0] = base - 1024
    perlin_ptr, chunk_ptr = base, base + 1024
    inst.write(perlin_ptr, bytes(np.asarray(perm, dtype=np.uint8).tolist()) + struct.pack("<I", 0))
    isos_ptr = inst.invoke(F_RUST_ALLOC, [n_iso * 4, 4])[0]
    ch = bytearray(512)
    struct.pack_into("<I", ch, 0, HASHBROWN_EMPTY_CTRL)   # tris: empty HashMap (+4..+12 zero, keys +16/+24 zero)
    struct.pack_into("<Q", ch, 32, 4)                     # verts      {ptr=4 (dangling), cap=0}, len@40=0
    struct.pack_into("<Q", ch, 40, 4 << 32)               # vert_pairs {ptr@44=4}, cap@48=0, len@52=0
    struct.pack_into("<Q", ch, 56, 2)                     # inds       {ptr=2, cap=0}, len@64=0
    struct.pack_into("<I", ch, 68, isos_ptr)              # isos       {ptr, cap=1331, len=0}
    struct.pack_into("<Q", ch, 72, n_iso)
    struct.pack_into("<3i", ch, 80, pos[0] * 16, pos[1] * 16, pos[2] * 16)   # chunk_offset
    ch[280] = 1                                           # BuildState::Iso
    inst.write(chunk_ptr, bytes(ch))

    def hook(fidx, args):
        if fidx in BUFFER_CREATE_FUNCS:
            raise _Stop()
        return None

    inst.call_hook = hook
    res = dict(calls=0, finished=False, at_buffer=False)
    try:
        for it in range(32):
            r = inst.invoke(F_BUILD_PARTIAL, [chunk_ptr, perlin_ptr, 0])[0]
            res["calls"] += 1
            if it == 0:
                ip, _, il = struct.unpack_from("<3I", inst.mem, chunk_ptr + 68)
                if il:   # a blank chunk has already dropped its isos (Build::finish)
                    res["isos"] = np.frombuffer(inst.read(ip, il * 4), dtype=np.float32).copy()
            if r:
                res["finished"] = True
                break
    except _Stop:
        res["at_buffer"] = True
    vp, _, vl = struct.unpack_from("<3I", inst.mem, chunk_ptr + 32)
    ip2, _, il2 = struct.unpack_from("<3I", inst.mem, chunk_ptr + 56)
    res["num_inds"] = struct.unpack_from("<I", inst.mem, chunk_ptr + 92)[0]
    if res["at_buffer"]:
        res["verts"] = np.frombuffer(inst.read(vp, vl * 24), dtype=np.float32).reshape(-1, 6).copy()
        res["inds"] = np.frombuffer(inst.read(ip2, il2 * 2), dtype=np.uint16).copy()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-chunks", action="store_true")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    m = wf.load_reference_module()
    o12 = Oracle(12)

    # ---- perm tables --------------------------------------------------------------------
    seeds = [0, 1, 2, 42, 0xDEADBEEF, 0xFFFFFFFF, 123456789, 1697500000]
    perms = np.stack([wasm_perm(m, s) for s in seeds])
    np.savez_compressed(os.path.join(GOLD, "ref_wasm_perm.npz"), seeds=np.array(seeds, dtype=np.uint64), perms=perms)
    print("perm tables:", perms.shape)

    # ---- perlin3 ------------------------------------------------------------------------
    rng = random.Random(20261017)
    pts, pseed = [], []
    out = []
    for si, s in enumerate([0, 1, 42, 0xDEADBEEF]):
        p = []
        for k in range(1000):
            sc = rng.choice([1, 4, 16, 100, 1000, 60000])
            x, y, z = [rng.uniform(-sc, sc) for _ in range(3)]
            if k % 10 == 0:
                x = float(round(x))
            if k % 17 == 0:
                z = float(round(z)) + 2.98e-8
            if k % 23 == 0:
                y = float(round(y)) - 1.19e-7
            p.append((x, y, z))
        out.append(wasm_perlin3(m, perms[seeds.index(s)], p))
        pts.extend(p)
        pseed.extend([s] * len(p))
    np.savez_compressed(os.path.join(GOLD, "ref_wasm_perlin3.npz"), seeds=np.array(pseed, dtype=np.uint64),
                        points=np.array(pts, dtype=np.float64), values=np.concatenate(out))
    print("perlin3 vectors:", len(pts))

    # ---- chunks (S=10, from the reference binary) -----------------------------------------
    if not args.skip_chunks:
        cases = [(0, (0, 0, -1)), (0, (0, 0, 0)), (0, (3, -2, -2)), (42, (-1, 5, -1)), (42, (7, 7, 0)),
                 (0xDEADBEEF, (-4, 2, -1)), (1, (0, 0, 3)), (1, (0, 0, -6)), (1, (100, -100, -2))]
        pack = {}
        meta = []
        for ci, (s, pos) in enumerate(cases):
            t0 = time.time()
            perm = wasm_perm(m, s)
            r = wasm_build_chunk(m, perm, pos)
            meta.append((s, pos[0], pos[1], pos[2], int(r["finished"]), int(r["at_buffer"]), r["num_inds"], r["calls"]))
            if "isos" in r:
                pack[f"isos_{ci}"] = r["isos"]
            if r["at_buffer"]:
                pack[f"verts_{ci}"] = r["verts"]
                pack[f"inds_{ci}"] = r["inds"]
            print(f"chunk seed={s} pos={pos}: calls={r['calls']} finished={r['finished']} at_buffer={r['at_buffer']} "
                  f"num_inds={r['num_inds']} ({time.time() - t0:.1f}s)")
        pack["meta"] = np.array(meta, dtype=np.int64)
        np.savez_compressed(os.path.join(GOLD, "ref_wasm_chunks_s10.npz"), **pack)

    # ---- oracle KATs at S=12 --------------------------------------------------------------
    kat = {}
    kmeta = []
    kcases = [(0, (0, 0, 0)), (0, (0, 0, -1)), (0, (-8, 7, -2)), (1, (5, 5, -3)), (42, (-3, 2, 1)),
              (0xDEADBEEF, (63, -64, -1)), (42, (0, 0, 2)), (42, (0, 0, -4)), (7, (1000, -2000, -1))]
    for ci, (s, pos) in enumerate(kcases):
        perm = o12.perm_table(s)
        r = o12.build_chunk(perm, pos, MODE_FAITHFUL)
        kat[f"isos_{ci}"] = r["isos"]
        kat[f"cases_{ci}"] = r["cases"]
        kat[f"verts_{ci}"] = np.concatenate([r["verts"]["pos"], r["verts"]["color"]], axis=1)
        kat[f"inds_{ci}"] = r["inds"].astype(np.uint16)
        kmeta.append((s, pos[0], pos[1], pos[2], r["flags"], len(r["verts"]), len(r["inds"])))
    kat["meta"] = np.array(kmeta, dtype=np.int64)
    np.savez_compressed(os.path.join(GOLD, "oracle_kat_s12.npz"), **kat)
    print("oracle KATs:", kmeta)

    for fn in sorted(os.listdir(GOLD)):
        p = os.path.join(GOLD, fn)
        print(f"{fn}: {os.path.getsize(p)} B sha256={hashlib.sha256(open(p, 'rb').read()).hexdigest()[:16]}")


if __name__ == "__main__":
    main()
