#!/usr/bin/env python3
"""Per-region table of the fused kernel from an .ncu-rep: derives the source line ranges of uw_kernels.cuh from
marker strings, then calls tools/ncu_regions.py.   usage: python tools/fused_regions.py rep.ncu-rep"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "underwaterworld_b200/csrc/uw_kernels.cuh")).read().split("\n")


def line_of(marker, start=0):
    for i in range(start, len(src)):
        if marker in src[i]:
            return i + 1
    raise SystemExit(f"marker not found: {marker}")


spec0 = line_of("__device__ __forceinline__ uint32_t noise_chunk_spec")
h = line_of("// ---- stage H", spec0); x = line_of("// ---- stage X", spec0); yz = line_of("// ---- stage YZ", spec0)
spec_end = line_of("__global__ void __launch_bounds__(SpecDims<ST, NOCT>::NT, 5)", spec0)
exact0 = line_of("__device__ __forceinline__ double x_grad3"); exact1 = line_of("k_noise_exact(DevCfg cfg") - 2
scan0 = line_of("__device__ __forceinline__ void block_scan2"); scan1 = line_of("k_classify_small(const") - 5
color0 = line_of("__device__ __forceinline__ float srgb_of"); color1 = line_of("// util::Tri::new") - 1
mkv0 = line_of("__device__ __forceinline__ void make_vertex("); mkv1 = line_of("// owner (first cell in scan order") - 1
prep0 = line_of("__device__ __forceinline__ ChunkShape emit_prepare"); fill0 = line_of("__device__ __forceinline__ void emit_fill")
verts0 = line_of("__device__ __forceinline__ void emit_verts"); inds0 = line_of("__device__ __forceinline__ void emit_indices")
tris0 = line_of("__device__ __forceinline__ void emit_tris"); rest0 = line_of("__device__ __forceinline__ void emit_rest")
hand0 = line_of("struct Ticket {"); hand1 = line_of("// K1 (fast path, compile-time specialised)") - 2
fused0 = line_of("k_build_fused(const __grid_constant__ DevCfg cfg"); fused1 = line_of("// Measurement aid (not on the product path)") - 2
regions = [
    f"K1_stage_H:{h}-{x - 1}", f"K1_stage_X:{x}-{yz - 1}", f"K1_stage_YZ:{yz}-{spec_end - 3}",
    f"exact_f64_guard:{exact0}-{exact1}", f"block_scan:{scan0}-{scan1}",
    f"vertex_lerp_colour:{color0}-{color1}", f"vertex_lerp_colour:{mkv0}-{mkv1}",
    f"K2_prepare:{prep0}-{fill0 - 1}", f"K4_fill:{fill0}-{verts0 - 1}", f"K4_verts:{verts0}-{inds0 - 1}",
    f"K4_indices:{inds0}-{tris0 - 1}", f"K4_tris:{tris0}-{rest0 - 1}", f"hand_out:{hand0}-{hand1}", f"loop_body:{fused0}-{fused1}",
]
kern = sys.argv[2] if len(sys.argv) > 2 else "k_build_fused"
mangled = sys.argv[3] if len(sys.argv) > 3 else "k_build_fusedILi12ELi3EtLb0"
subprocess.check_call([sys.executable, os.path.join(ROOT, "tools/ncu_regions.py"), sys.argv[1], kern, mangled] + regions)
