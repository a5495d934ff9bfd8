"""The C++ host mirror of the reference's Chunk API (include/uw_chunk.hpp): compiles against the C ABI;
without a GPU it must fail loudly (no CPU fallback); on a GPU its output must match the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "underwaterworld_b200", "lib")
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "test_chunk_api")


def _fnv(b: bytes) -> int:
    h = 1469598103934665603
    for x in b:
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.fixture(scope="module")
def exe():
    from underwaterworld_b200.build import build_library
    build_library()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "test_chunk_api.cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                           "-L", LIBDIR, "-luwcuda", f"-Wl,-rpath,{LIBDIR}"])
    return EXE


def test_cpp_mirror_compiles_and_refuses_to_run_without_a_gpu(exe):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 42 and "no CPU fallback" in p.stdout      # UW_ERR_NO_DEVICE surfaced as uw::Error


@pytest.mark.gpu
def test_cpp_mirror_matches_oracle(exe, oracle12):
    from oracle import MODE_FAST
    p = subprocess.run([exe, "0"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "OK" in p.stdout and "blank_slice_throws=1" in p.stdout
    assert re.search(r"index32 meshes=[1-9]\d* bad=0", p.stdout)        # u32 builders: no NULL inds16 dereference
    perm = oracle12.perm_table(0)
    n = 0
    for line in p.stdout.splitlines():
        m = re.match(r"chunk (-?\d+) (-?\d+) (-?\d+) not_blank=(\d) num_inds=(\d+)(.*)", line)
        if not m:
            continue
        n += 1
        pos = tuple(int(m.group(i)) for i in (1, 2, 3))
        r = oracle12.build_chunk(perm, pos, MODE_FAST)
        assert int(m.group(5)) == len(r["inds"]) and int(m.group(4)) == (1 if len(r["inds"]) else 0)
        if len(r["inds"]):
            mm = re.search(r"verts=(\d+) inds_hash=([0-9a-f]+) pos0=(\S+)", m.group(6))
            assert int(mm.group(1)) == len(r["verts"])
            assert int(mm.group(2), 16) == _fnv(r["inds"].astype(np.uint16).tobytes())
            vp = oracle12.vertex_pairs(perm, pos)[0]                     # per-edge position bound, as in test_gpu_parity._pos_bound
            gap = abs(float(r["isos"][vp[1]]) - float(r["isos"][vp[0]]))
            bound = 1e-5 + (16.0 / 12.0) * min(1.0, 2e-6 / max(gap - 4e-6, 1e-300))
            assert abs(float.fromhex(mm.group(3)) - float(r["verts"]["pos"][0][0])) <= bound
        else:
            assert f"blank_early={r['flags'] & 1}" in m.group(6)
    assert n == 5
