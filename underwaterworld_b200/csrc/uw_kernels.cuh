// uw_kernels.cuh -- sm_100a kernels of the chunk-build hot path (K1 noise, K2 classify,
// K3 scan, K4 emit).  Included once by uwcuda.cu.  See DESIGN.md for layout and rooflines.
//
// Reference semantics reproduced here (file:line under /root/reference/underwater_world/src):
//   K1  chunk.rs:105-129 (build_iso) + perlin_util.rs:6-29 + noise-0.8.2 perlin_3d
//   K2  chunk.rs:131-133 (early_blank_check), chunk.rs:141-164 (triangulation_idx)
//   K3  implicit order of the sequential pushes, chunk.rs:233-243
//   K4  chunk.rs:178-243 (edge lerp, colour, ordered-pair dedup, index emission)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "mc_tables.h"
#include "../../include/uwcuda.h"

#ifdef UW_PHASE_TIMING
__device__ unsigned long long g_phase[16];
__device__ unsigned long long g_cta[1024][4];     // per CTA: globaltimer start, end, chunks, heavy chunks
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define PHASE_MARK(idx) do { if (threadIdx.x == 0) { const long long now_ = clock64(); atomicAdd(&g_phase[idx], (unsigned long long)(now_ - t_phase)); t_phase = now_; } } while (0)
#define PHASE_ARG , long long& t_phase
#define PHASE_PASS , t_phase
#else
#define PHASE_MARK(idx) do { } while (0)
#define PHASE_ARG
#define PHASE_PASS
#endif

#define UW_MAX_OCT 4
#define UW_SMALL_MAX_L 16        // small path: whole chunk per CTA, L = S+1 <= 16
#define UW_AXIS_PAD 20

// ---------------------------------------------------------------------------------------
// Parameter blocks (passed by value -> constant bank, uniform access)
// ---------------------------------------------------------------------------------------
struct DevCfg {
    int S, L, L2, L3, chunk_size, octaves;
    float iso_level, max_height, adj_z_mod, size_scale;
    float min_hue, max_hue, min_z, max_z;
    float guard_eps;
    float hsv_c[3], hsv_m[3];          // c = value*saturation, m = value - c per value level (util.rs:129,132)
    float srgb_hi[3], srgb_lo[3];      // to_srgb((c+m)*255), to_srgb((0+m)*255) per value level (host powf)
    uint32_t dens_stride;              // floats per chunk in the density array (multiple of 4)
    int cs_pow2, mod_pow2;             // chunk_size / adj_z_mod are powers of two: exact reciprocal multiplies
    double inv_chunk_size;             // 1 / chunk_size (exact when cs_pow2)
    float inv_adj_z_mod;               // 1 / adj_z_mod  (exact when mod_pow2)
    const float* terr_tab;             // [terr_nz][16]: terrace term - 1 per (chunk z layer, lattice index), see k_terrace_table
    int terr_z0, terr_nz;              // layers terr_z0 .. terr_z0 + terr_nz - 1 are tabulated (others are computed in place)
    int G[UW_MAX_OCT];                 // lattice points per axis per octave = 2^o + 2
    int lat_base[UW_MAX_OCT + 1];      // prefix of G^3
    int x_base[UW_MAX_OCT + 1];        // prefix of L*G^2
};

// Chunk-independent per-axis tables: for octave o and lattice index i,
// p = 2^o * u_i with u_i = (i * f64(size_scale_f32)) / chunk_size; c = floor(p), d = p - c,
// d1 = d - 1, w = fade(d).  Computed on the host in f64 in the reference's operation order.
struct AxisTables {
    float d[UW_MAX_OCT][UW_AXIS_PAD];
    float d1[UW_MAX_OCT][UW_AXIS_PAD];
    float w[UW_MAX_OCT][UW_AXIS_PAD];
    int   c[UW_MAX_OCT][UW_AXIS_PAD];
};

// MC tables in device global memory (L1/L2 resident, divergent lookups)
struct McTables {
    uint64_t rows[256];        // nibble-packed edge rows, 0xF terminated
    uint16_t crossed[256];     // edges present in the row
    uint16_t before[256][12];  // edges first-appearing before edge e in the row
    uint8_t  ninds[256];       // indices per case
    uint32_t lut[256];         // by "natural" corner pattern (natural_of): case | ninds << 8 | crossed << 12
    float powtab[48];          // pow24_tab: [i] = 1/c_i, [16+i] = hi(log2 c_i), [32+i] = lo(log2 c_i), c_i ~ 1 + (i + 1/2)/16
};

__device__ __constant__ uint8_t c_edge_a[12] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3};
__device__ __constant__ uint8_t c_edge_b[12] = {1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7};

// corner offsets, chunk.rs:144-153: bit0 = dx, bit1 = dy, bit2 = dz packed per corner
// corners: 0(0,0,0) 1(1,0,0) 2(1,0,1) 3(0,0,1) 4(0,1,0) 5(1,1,0) 6(1,1,1) 7(0,1,1)
__device__ __forceinline__ void corner_off(int c, int& dx, int& dy, int& dz) {
    const uint32_t DX = 0x66u, DY = 0xF0u, DZ = 0xCCu;   // bit c of each = offset of corner c
    dx = (DX >> c) & 1; dy = (DY >> c) & 1; dz = (DZ >> c) & 1;
}

// ---------------------------------------------------------------------------------------
// Exact f64 density: the reference's operation order, no fused operations.
//   coords chunk.rs:107-116; iso_at perlin_util.rs:24-29; octaves perlin_util.rs:6-22;
//   perlin_3d SURVEY App. A.4.  Bit-identical to the reference arithmetic by construction
//   (IEEE add/sub/mul/div in the same order; __d*_rn intrinsics are never contracted).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double x_grad3(uint32_t h, double x, double y, double z) {
    switch (h & 15u) {
        case 0: case 12: return __dadd_rn(x, y);
        case 1: case 13: return __dsub_rn(y, x);
        case 2:          return __dsub_rn(x, y);
        case 3:          return __dsub_rn(-x, y);
        case 4:          return __dadd_rn(x, z);
        case 5:          return __dsub_rn(z, x);
        case 6:          return __dsub_rn(x, z);
        case 7:          return __dsub_rn(-x, z);
        case 8:          return __dadd_rn(y, z);
        case 9: case 14: return __dsub_rn(z, y);
        case 10:         return __dsub_rn(y, z);
        default:         return __dsub_rn(-y, z);
    }
}

__device__ __forceinline__ double x_fade(double t) {
    double c = t < 0.0 ? 0.0 : t;
    c = c > 1.0 ? 1.0 : c;
    const double c3 = __dmul_rn(__dmul_rn(c, c), c);
    const double in = __dadd_rn(__dmul_rn(c, __dadd_rn(__dmul_rn(c, 6.0), -15.0)), 10.0);
    return __dmul_rn(c3, in);
}

__device__ __forceinline__ uint32_t x_hash(const uint8_t* perm, int ix, int iy, int iz) {
    return perm[perm[perm[ix & 255] ^ (iy & 255)] ^ (iz & 255)];
}

__device__ __noinline__ double x_perlin3(const uint8_t* perm, double px, double py, double pz) {
    const double fx = floor(px), fy = floor(py), fz = floor(pz);
    const double dx = __dsub_rn(px, fx), dy = __dsub_rn(py, fy), dz = __dsub_rn(pz, fz);
    const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    const double dx1 = __dadd_rn(dx, -1.0), dy1 = __dadd_rn(dy, -1.0), dz1 = __dadd_rn(dz, -1.0);
    const double g000 = x_grad3(x_hash(perm, ix,     iy,     iz    ), dx,  dy,  dz );
    const double g100 = x_grad3(x_hash(perm, ix + 1, iy,     iz    ), dx1, dy,  dz );
    const double g010 = x_grad3(x_hash(perm, ix,     iy + 1, iz    ), dx,  dy1, dz );
    const double g110 = x_grad3(x_hash(perm, ix + 1, iy + 1, iz    ), dx1, dy1, dz );
    const double g001 = x_grad3(x_hash(perm, ix,     iy,     iz + 1), dx,  dy,  dz1);
    const double g101 = x_grad3(x_hash(perm, ix + 1, iy,     iz + 1), dx1, dy,  dz1);
    const double g011 = x_grad3(x_hash(perm, ix,     iy + 1, iz + 1), dx,  dy1, dz1);
    const double g111 = x_grad3(x_hash(perm, ix + 1, iy + 1, iz + 1), dx1, dy1, dz1);
    const double a = x_fade(dx), b = x_fade(dy), c = x_fade(dz);
    const double k0 = g000;
    const double k1 = __dsub_rn(g100, g000);
    const double k2 = __dsub_rn(g010, g000);
    const double k3 = __dsub_rn(g001, g000);
    const double k4 = __dsub_rn(__dsub_rn(__dadd_rn(g000, g110), g100), g010);
    const double k5 = __dsub_rn(__dsub_rn(__dadd_rn(g000, g101), g100), g001);
    const double k6 = __dsub_rn(__dsub_rn(__dadd_rn(g000, g011), g010), g001);
    const double k7 = __dsub_rn(__dsub_rn(__dsub_rn(__dsub_rn(
                        __dadd_rn(__dadd_rn(__dadd_rn(g100, g010), g001), g111), g000), g110), g101), g011);
    double r = __dadd_rn(k0, __dmul_rn(k1, a));
    r = __dadd_rn(r, __dmul_rn(k2, b));
    r = __dadd_rn(r, __dmul_rn(k3, c));
    r = __dadd_rn(r, __dmul_rn(__dmul_rn(k4, a), b));
    r = __dadd_rn(r, __dmul_rn(__dmul_rn(k5, a), c));
    r = __dadd_rn(r, __dmul_rn(__dmul_rn(k6, b), c));
    r = __dadd_rn(r, __dmul_rn(__dmul_rn(__dmul_rn(k7, a), b), c));
    r = __dmul_rn(r, 1.1547005383792515);
    return r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
}

// perlin_util.rs:24-29 on an arbitrary f64 point
__device__ __forceinline__ float x_iso_at(const DevCfg& cfg, const uint8_t* perm, double x, double y, double z) {
    double total = 0.0, freq = 1.0, amp = 1.0, maxv = 0.0;
    for (int o = 0; o < cfg.octaves; ++o) {
        const double v = x_perlin3(perm, __dmul_rn(x, freq), __dmul_rn(y, freq), __dmul_rn(z, freq));
        total = __dadd_rn(total, __dmul_rn(v, amp));
        maxv = __dadd_rn(maxv, amp);
        amp = __dmul_rn(amp, 0.5);
        freq = __dmul_rn(freq, 2.0);
    }
    const float p = __double2float_rn(__ddiv_rn(total, maxv));
    const float adj_z = __fdiv_rn(__fmul_rn(__double2float_rn(z), (float)cfg.chunk_size), cfg.max_height);
    return __fsub_rn(__fadd_rn(adj_z, p), fmodf(adj_z, cfg.adj_z_mod));
}

// perlin_util.rs:27-28: adj_z and adj_z % ADJ_Z_MOD for lattice index k of a chunk at pz.  Bit-identical
// to the reference's f64 coordinate -> f32 arithmetic; power-of-two divisors use exact reciprocal
// multiplies (x / 2^n == x * 2^-n, fmod(a, 2^n) == a - 2^n * trunc(a * 2^-n) while |a * 2^-n| < 2^22).
__device__ __forceinline__ void terrace_terms(const DevCfg& cfg, int k, int pz, float& adj, float& fm) {
    const double local = __dmul_rn((double)k, (double)cfg.size_scale);
    const double sum = __dadd_rn(local, (double)(pz * cfg.chunk_size));
    const double zc = cfg.cs_pow2 ? __dmul_rn(sum, cfg.inv_chunk_size) : __ddiv_rn(sum, (double)cfg.chunk_size);
    const float zf = __double2float_rn(zc);
    adj = __fdiv_rn(__fmul_rn(zf, (float)cfg.chunk_size), cfg.max_height);
    const float q = __fmul_rn(adj, cfg.inv_adj_z_mod);
    fm = (cfg.mod_pow2 && fabsf(q) < 4194304.0f) ? __fsub_rn(adj, __fmul_rn(cfg.adj_z_mod, truncf(q))) : fmodf(adj, cfg.adj_z_mod);
}

// chunk.rs:107-116: lattice index -> f64 sample coordinate
__device__ __forceinline__ double x_coord(const DevCfg& cfg, int i, int chunk_pos) {
    const double local = __dmul_rn((double)i, (double)cfg.size_scale);
    const int off = chunk_pos * cfg.chunk_size;                       // chunk.rs:90-94
    return __ddiv_rn(__dadd_rn(local, (double)off), (double)cfg.chunk_size);
}

__device__ __noinline__ float x_iso_lattice(const DevCfg& cfg, const uint8_t* perm,
                                            int px, int py, int pz, int i, int j, int k) {
    return x_iso_at(cfg, perm, x_coord(cfg, i, px), x_coord(cfg, j, py), x_coord(cfg, k, pz));
}

// ---------------------------------------------------------------------------------------
// K1 (exact mode): one thread per sample, f64 everywhere.  UW_FLAG_EXACT_F64.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_noise_exact(DevCfg cfg, const uint8_t* __restrict__ g_perm,
                                                     const int32_t* __restrict__ pos, uint32_t n,
                                                     float* __restrict__ dens) {
    __shared__ uint8_t s_perm[256];
    for (int t = threadIdx.x; t < 256; t += blockDim.x) s_perm[t] = g_perm[t];
    __syncthreads();
    const uint64_t total = (uint64_t)n * cfg.L3;
    for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t chunk = (uint32_t)(idx / cfg.L3);
        const int r = (int)(idx - (uint64_t)chunk * cfg.L3);
        const int i = r / cfg.L2, j = (r / cfg.L) % cfg.L, k = r % cfg.L;
        const int px = pos[3 * chunk], py = pos[3 * chunk + 1], pz = pos[3 * chunk + 2];
        dens[(size_t)chunk * cfg.dens_stride + r] = x_iso_lattice(cfg, s_perm, px, py, pz, i, j, k);
    }
}

// Batched point queries, perlin_util.rs:24-29 (boid.rs:132,324 callers).  SURVEY §8f-3.
__global__ void __launch_bounds__(256) k_iso_points(DevCfg cfg, const uint8_t* __restrict__ g_perm,
                                                    const double* __restrict__ pts, uint32_t n,
                                                    float* __restrict__ out) {
    __shared__ uint8_t s_perm[256];
    for (int t = threadIdx.x; t < 256; t += blockDim.x) s_perm[t] = g_perm[t];
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[i] = x_iso_at(cfg, s_perm, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
}

// ---------------------------------------------------------------------------------------
// K1 (fast path, small chunks): one CTA builds one chunk's L^3 lattice.
//
// Tensor-product factorisation of gradient noise.  Inside one noise cell the value is
//   n = lerp_z( lerp_y( lerp_x( g.(d - corner) ) ) )
// and the sample lattice is a tensor product (x index i, y index j, z index k) whose
// fractional offsets and fade weights depend only on the per-axis index (App. A.6).  So:
//   stage H : hash every noise-lattice point the chunk touches once  (sum_o G_o^3  <= 307)
//   stage X : lerp the two x-neighbours for every (i, cy, cz)         (sum_o L*G_o^2 = 793)
//             -> (Q, Py, Pz) with value-at-corner = Q + Py*dyc + Pz*dzc
//   stage YZ: thread (i,j) walks its z column; per lattice cz it lerps in y -> (R, Sz),
//             per sample it lerps in z: 4 FP32 ops per octave-sample.
// FP32 error vs the f64 reference is ~1e-7; samples that land within guard_eps of the
// isovalue are re-evaluated by the exact f64 path so that classification is bit-exact.
// ---------------------------------------------------------------------------------------
__device__ __constant__ float c_grad_vec[16][4] = {
    { 1,  1,  0, 0}, {-1,  1,  0, 0}, { 1, -1,  0, 0}, {-1, -1,  0, 0},
    { 1,  0,  1, 0}, {-1,  0,  1, 0}, { 1,  0, -1, 0}, {-1,  0, -1, 0},
    { 0,  1,  1, 0}, { 0, -1,  1, 0}, { 0,  1, -1, 0}, { 0, -1, -1, 0},
    { 1,  1,  0, 0}, {-1,  1,  0, 0}, { 0, -1,  1, 0}, { 0, -1, -1, 0}};

struct NoiseSmem {   // dynamic shared memory carve-up (offsets in bytes computed by host+device identically)
    uint8_t* perm; float4* grad; float4* lat; float4* X; float* dens; float* adjz; float* fm; int* red;
};

__host__ __device__ inline size_t noise_smem_bytes(const DevCfg& cfg) {
    size_t b = 256 + 16 * 16;
    b += (size_t)cfg.lat_base[cfg.octaves] * 16;
    b += (size_t)cfg.x_base[cfg.octaves] * 16;
    b += (size_t)cfg.dens_stride * 4;
    b += (size_t)UW_AXIS_PAD * 4 * 2;
    b += 64 * 4;
    return b;
}

__device__ __forceinline__ NoiseSmem noise_smem_carve(const DevCfg& cfg, unsigned char* base) {
    NoiseSmem s;
    s.perm = base;                          base += 256;
    s.grad = (float4*)base;                 base += 16 * 16;
    s.lat  = (float4*)base;                 base += (size_t)cfg.lat_base[cfg.octaves] * 16;
    s.X    = (float4*)base;                 base += (size_t)cfg.x_base[cfg.octaves] * 16;
    s.dens = (float*)base;                  base += (size_t)cfg.dens_stride * 4;
    s.adjz = (float*)base;                  base += UW_AXIS_PAD * 4;
    s.fm   = (float*)base;                  base += UW_AXIS_PAD * 4;
    s.red  = (int*)base;
    return s;
}

// per-chunk flags written by K1/K2
#define CF_ALL_GT   1u   // every sample >  iso_level  -> early blank (chunk.rs:131-133)
#define CF_ANY_LT   2u   // some sample  <  iso_level  -> mesh stage may emit
#define CF_ALL_LT   4u   // every sample <  iso_level  -> all-solid: every case is 255, nothing to emit (fused kernel only)

template <int LT, int NOCT>
__global__ void __launch_bounds__(256) k_noise_small(const __grid_constant__ DevCfg cfg,
                                                     const __grid_constant__ AxisTables tab,
                                                     const uint8_t* __restrict__ g_perm,
                                                     const int32_t* __restrict__ pos, uint32_t n,
                                                     float* __restrict__ dens,
                                                     unsigned long long* __restrict__ guard_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const NoiseSmem s = noise_smem_carve(cfg, smem_raw);
    const int L = LT > 0 ? LT : cfg.L;
    const int noct = NOCT > 0 ? NOCT : cfg.octaves;
    const int tid = threadIdx.x, NT = blockDim.x;

    for (int t = tid; t < 256; t += NT) s.perm[t] = g_perm[t];
    if (tid < 16) s.grad[tid] = make_float4(c_grad_vec[tid][0], c_grad_vec[tid][1], c_grad_vec[tid][2], 0.f);
    __syncthreads();

    const int lat_total = cfg.lat_base[noct], x_total = cfg.x_base[noct];

    for (uint32_t chunk = blockIdx.x; chunk < n; chunk += gridDim.x) {
        const int px = pos[3 * chunk], py = pos[3 * chunk + 1], pz = pos[3 * chunk + 2];

        // ---- stage H: gradient vector of every touched noise-lattice point -----------------
        for (int t = tid; t < lat_total; t += NT) {
            int o = 0;
#pragma unroll
            for (int q = 1; q < UW_MAX_OCT; ++q) if (q < noct && t >= cfg.lat_base[q]) o = q;
            const int G = cfg.G[o];
            const int r = t - cfg.lat_base[o];
            const int cx = r / (G * G), cy = (r / G) % G, cz = r % G;
            const int F = 1 << o;
            const uint32_t h = s.perm[s.perm[s.perm[(F * px + cx) & 255] ^ ((F * py + cy) & 255)] ^ ((F * pz + cz) & 255)];
            s.lat[t] = s.grad[h & 15u];
        }
        // terrace term, perlin_util.rs:27-28: exact (f64 coordinate -> f32), per (chunk, k)
        if (tid < L) {
            float adj, fm;
            terrace_terms(cfg, tid, pz, adj, fm);
            s.adjz[tid] = adj;
            s.fm[tid] = fm;
        }
        __syncthreads();

        // ---- stage X: x-lerp for every (octave, i, cy, cz) -----------------------------------
        for (int t = tid; t < x_total; t += NT) {
            int o = 0;
#pragma unroll
            for (int q = 1; q < UW_MAX_OCT; ++q) if (q < noct && t >= cfg.x_base[q]) o = q;
            const int G = cfg.G[o];
            const int r = t - cfg.x_base[o];
            const int i = r / (G * G), cy = (r / G) % G, cz = r % G;
            const int c = tab.c[o][i];
            const float4 g0 = s.lat[cfg.lat_base[o] + (c * G + cy) * G + cz];
            const float4 g1 = s.lat[cfg.lat_base[o] + ((c + 1) * G + cy) * G + cz];
            const float d = tab.d[o][i], d1 = tab.d1[o][i], w = tab.w[o][i];
            const float q0 = g0.x * d, q1 = g1.x * d1;
            float4 e;
            e.x = fmaf(w, q1 - q0, q0);
            e.y = fmaf(w, g1.y - g0.y, g0.y);
            e.z = fmaf(w, g1.z - g0.z, g0.z);
            e.w = 0.f;
            s.X[t] = e;
        }
        __syncthreads();

        // ---- stage YZ: one thread per (i, j) column ------------------------------------------
        uint32_t my_flags_allgt = 1u, my_flags_anylt = 0u;
        if (tid < L * L) {
            const int i = tid / L, j = tid - i * L;
            float R0[UW_MAX_OCT], S0[UW_MAX_OCT], R1[UW_MAX_OCT], S1[UW_MAX_OCT];
            const float inv_max = 1.0f / (2.0f - ldexpf(1.0f, 1 - noct));   // 1 / sum_{o<noct} 2^-o
            float* out = s.dens + (size_t)tid * L;

            auto ystage = [&](int o, int cz, float& R, float& Sz) {
                const int G = cfg.G[o];
                const int cyj = tab.c[o][j];
                const float4 E0 = s.X[cfg.x_base[o] + (i * G + cyj) * G + cz];
                const float4 E1 = s.X[cfg.x_base[o] + (i * G + cyj + 1) * G + cz];
                const float dy = tab.d[o][j], dy1 = tab.d1[o][j], wy = tab.w[o][j];
                const float A0 = fmaf(E0.y, dy, E0.x);
                const float A1 = fmaf(E1.y, dy1, E1.x);
                R = fmaf(wy, A1 - A0, A0);
                Sz = fmaf(wy, E1.z - E0.z, E0.z);
            };

#pragma unroll
            for (int k = 0; k < (LT > 0 ? LT : UW_SMALL_MAX_L); ++k) {
                if (LT == 0 && k >= L) break;
                float total = 0.f;
#pragma unroll
                for (int o = 0; o < UW_MAX_OCT; ++o) {
                    if (o >= noct) break;
                    const int c = tab.c[o][k];
                    if (k == 0) {
                        ystage(o, c, R0[o], S0[o]);
                        ystage(o, c + 1, R1[o], S1[o]);
                    } else if (c != tab.c[o][k - 1]) {      // warp-uniform: depends on k and tables only
                        R0[o] = R1[o]; S0[o] = S1[o];
                        ystage(o, c + 1, R1[o], S1[o]);
                    }
                    const float a0 = fmaf(S0[o], tab.d[o][k], R0[o]);
                    const float a1 = fmaf(S1[o], tab.d1[o][k], R1[o]);
                    float v = fmaf(tab.w[o][k], a1 - a0, a0) * 1.1547005383792515f;
                    v = fminf(fmaxf(v, -1.0f), 1.0f);
                    total = fmaf(v, ldexpf(1.0f, -o), total);
                }
                const float pf = total * inv_max;
                float iso = (s.adjz[k] + pf) - s.fm[k];
                if (fabsf(iso - cfg.iso_level) < cfg.guard_eps) {
                    iso = x_iso_lattice(cfg, s.perm, px, py, pz, i, j, k);
                    atomicAdd(guard_count, 1ull);
                }
                out[k] = iso;
                my_flags_allgt &= (iso > cfg.iso_level) ? 1u : 0u;
                my_flags_anylt |= (iso < cfg.iso_level) ? 1u : 0u;
            }
        }
        __syncthreads();

        // ---- coalesced write-out (float4; per-chunk stride is a multiple of 4 floats) ---------
        {
            float4* dst = reinterpret_cast<float4*>(dens + (size_t)chunk * cfg.dens_stride);
            const float4* src = reinterpret_cast<const float4*>(s.dens);
            const int n4 = (int)(cfg.dens_stride >> 2);
            for (int t = tid; t < n4; t += NT) dst[t] = src[t];
        }
        __syncthreads();   // s.dens / s.lat / s.X reused by the next chunk
        (void)my_flags_allgt; (void)my_flags_anylt;
    }
}

// Control block of one fused launch.  Two blocks alternate between launches: the last CTA of launch k
// zeroes the block of launch k+1, so no memset sits on the launch path.
struct BatchTotals {
    unsigned long long n_verts, n_inds;
    uint32_t n_active, overflow, n_blank, n_mesh;
};
#define UW_NCLS 8                    // hand-out classes: 0 = expected heaviest z layer ... UW_NCLS - 1 = provably trivial
struct FusedControl {
    uint32_t ticket, done;
    uint32_t classified, cls_ticket; // cost-ordered hand-out: chunks filed into their class list so far, filing work claimed
    uint32_t cls_n[UW_NCLS];         // entries per class list
    unsigned long long alloc;        // n_verts << 32 | n_inds (completion-order packing)
    unsigned long long guard;        // f64 guard-band re-evaluations
    BatchTotals totals;
};

// What the host reads back after a fused launch; written by the last CTA out to a per-batch slot, so that a
// later launch (which zeroes and reuses the control blocks) cannot disturb it.
struct FusedSummary {
    unsigned long long alloc, guard;
    BatchTotals totals;
};

// Multi-GPU gather (uw_gather_*): every segment of the render GPU's arenas has one 64-byte head.  The LAST CTA
// of the fused kernel that filled the segment -- running on any GPU of the box, writing through NVLink peer
// addresses -- stores its summary there, fences system-wide, then publishes the launch epoch: a consumer that
// has seen head.epoch >= e may read everything the e-th build of that segment wrote.
struct GatherHead {
    FusedSummary sum;
    uint32_t epoch, n_chunks;
    uint32_t first_chunk_lo, first_chunk_hi;
};
static_assert(sizeof(GatherHead) == 64, "GatherHead is one 64-byte line");

// Where a fused launch writes.  Default: the context's own arenas.  Attached to a gather segment: the render
// GPU's arenas (peer or local addresses) with the segment's element offsets added to the descriptors.
struct FusedOut {
    uint32_t desc_vbase, desc_ibase;   // added to uw_chunk_desc::vert_offset / index_offset (segment base in the arena)
    GatherHead* head;                  // nullable
    uw_chunk_desc* drawlist;           // nullable: compact list of the descriptors of chunks that END WITH A MESH (completion order)
    uint32_t epoch, first_chunk_lo, first_chunk_hi;
};

// Chunk hand-out for the persistent kernel.
//
// Request order (order == nullptr): tickets 0..n-1 walk the request list.
//
// Cost order (order != nullptr; used when a CTA gets only a few chunks, so the tail decides the run time): a
// chunk's cost is set by how much surface it holds, and that is mostly a function of its z layer -- iso =
// terrace(z) + noise -- so the host ranks the layers by the likelihood of surface (class 0 = heaviest; layers
// that provably hold none, |noise| <= 1, are the last class).  Every CTA first files a few chunks of the request
// into per-class lists (order[class][slot] = chunk index + position); tickets then walk class 0, class 1, ...:
// the expensive chunks all start in the first wave and the cheap ones fill the gaps at the end.  Results do not
// depend on the order.
struct Ticket { uint32_t chunk; int px, py, pz; };
#define TICKET_DONE 0xFFFFFFFFu
#define UW_POS_LIMIT (1 << 24)       // |chunk position| <= 2^24: SURVEY App. A.6 (exact lattice offsets), pos * 16 inside i32
#define UW_OVERFLOW_BAD_POS 3u       // BatchTotals::overflow code: a position of the request is out of range

struct Handout {
    FusedControl* ctr; const int32_t* pos; uint32_t n;
    uint4* order;                      // [UW_NCLS][n] or nullptr
    uint32_t* state;                   // shared memory, UW_NCLS + 1 words: class-list prefix sums once known, [UW_NCLS] = ready
    int z_lo, z_hi; unsigned long long zcls;   // class of layer z_lo + i in bits 4i..4i+3
    uw_chunk_desc* skip;               // UW_FLAG_ANALYTIC_SKIP: provably trivial chunks are answered without being handed out
};

// Requests handed over in pinned memory are not scanned on the host (6 MB for config 3: 0.3-0.8 ms of one core);
// the thread that fetches a ticket checks its position instead and the batch fails with UW_ERR_INVALID at its wait.
__device__ __forceinline__ void check_position(FusedControl* ctr, int px, int py, int pz) {
    const unsigned lim = 2u * UW_POS_LIMIT;
    if ((unsigned)(px + UW_POS_LIMIT) > lim || (unsigned)(py + UW_POS_LIMIT) > lim || (unsigned)(pz + UW_POS_LIMIT) > lim)
        atomicMax(&ctr->totals.overflow, UW_OVERFLOW_BAD_POS);
}

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ bool answer_trivial(const Handout& h, uint32_t c, int px, int py, int pz) {
    if (!h.skip || (pz >= h.z_lo && pz <= h.z_hi)) return false;
    uw_chunk_desc d;
    d.pos[0] = px; d.pos[1] = py; d.pos[2] = pz;
    d.flags = pz > h.z_hi ? UW_CHUNK_BLANK_EARLY : 0u;      // blank-early above the surface layers, solid (no mesh) below
    d.vert_offset = 0; d.vert_count = 0; d.index_offset = 0; d.index_count = 0;
    h.skip[c] = d;
    if (pz > h.z_hi) atomicAdd(&h.ctr->totals.n_blank, 1u);
    return true;
}

// cost order, step 1 (all threads of a CTA, before its first ticket): file the request into the class lists.
// CTA b files the chunks [b NT, b NT + NT) (and every grid-stride image of that range): no ticket for the filing work,
// so a CTA with nothing to file goes straight on and a filing CTA starts with its position loads -- one or two global
// round trips fewer at the head of a 44 us launch.  The wait in take_ticket depends only on the lowest-numbered CTAs,
// which are dispatched first.
__device__ __forceinline__ void handout_classify(const Handout& h) {
    const uint32_t lane = threadIdx.x & 31u;
#ifdef UW_FILE_BY_TICKET
    for (;;) {
        if (threadIdx.x == 0) h.state[0] = atomicAdd(&h.ctr->cls_ticket, blockDim.x);
        __syncthreads();
        const uint32_t c0 = h.state[0];
        __syncthreads();
        if (c0 >= h.n) break;
#else
    for (uint32_t c0 = blockIdx.x * blockDim.x; c0 < h.n; c0 += gridDim.x * blockDim.x) {
#endif
        const uint32_t c = c0 + threadIdx.x;
        int px = 0, py = 0, pz = 0;
        bool valid = c < h.n;
        if (valid) { px = h.pos[3 * c]; py = h.pos[3 * c + 1]; pz = h.pos[3 * c + 2]; check_position(h.ctr, px, py, pz); }
        if (valid && answer_trivial(h, c, px, py, pz)) valid = false;
        const int dz = pz - h.z_lo;
        const uint32_t cls = (pz < h.z_lo || pz > h.z_hi) ? UW_NCLS - 1 : dz < 16 ? (uint32_t)((h.zcls >> (4 * dz)) & 15ull) : UW_NCLS - 2;
        const uint32_t act = __ballot_sync(0xFFFFFFFFu, valid);
        if (valid) {                                       // one atomic per (warp, class)
            const uint32_t peers = __match_any_sync(act, cls);
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(&h.ctr->cls_n[cls], (uint32_t)__popc(peers));
            base = __shfl_sync(peers, base, leader);
            const uint32_t slot = base + __popc(peers & ((1u << lane) - 1u));
            h.order[(size_t)cls * h.n + slot] = make_uint4(c, (uint32_t)px, (uint32_t)py, (uint32_t)pz);
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(&h.ctr->classified, min((uint32_t)blockDim.x, h.n - c0));
    }
}

// cost order, step 2 (ONE thread, once per CTA): wait until the whole request is filed, keep the lists' prefix sums
__device__ __forceinline__ void handout_ready(const Handout& h) {
    if (h.state[UW_NCLS]) return;
    while (ld_volatile_u32(&h.ctr->classified) < h.n) { }
    __threadfence();
    uint32_t run = 0;
    for (int k = 0; k < UW_NCLS; ++k) { run += ld_volatile_u32(&h.ctr->cls_n[k]); h.state[k] = run; }
    h.state[UW_NCLS] = 1u;
}

// executed by ONE thread
__device__ __noinline__ Ticket take_ticket(const Handout& h) {
    Ticket tk;
    tk.chunk = TICKET_DONE; tk.px = tk.py = tk.pz = 0;
    if (h.order) {
        const uint32_t t = atomicAdd(&h.ctr->ticket, 1u);
        handout_ready(h);
        if (t >= h.state[UW_NCLS - 1]) return tk;
        int cls = 0;
        while (t >= h.state[cls]) ++cls;
        const uint32_t before = cls ? h.state[cls - 1] : 0u;
        const uint4 e = __ldcg(&h.order[(size_t)cls * h.n + (t - before)]);
        tk.chunk = e.x; tk.px = (int)e.y; tk.py = (int)e.z; tk.pz = (int)e.w;
        return tk;
    }
    while (true) {
        const uint32_t c = atomicAdd(&h.ctr->ticket, 1u);
        if (c >= h.n) return tk;
        tk.px = h.pos[3 * c]; tk.py = h.pos[3 * c + 1]; tk.pz = h.pos[3 * c + 2];
        check_position(h.ctr, tk.px, tk.py, tk.pz);
        if (answer_trivial(h, c, tk.px, tk.py, tk.pz)) continue;
        tk.chunk = c;
        return tk;
    }
}

// Split-phase hand-out for the steady state (the CTA's first ticket goes through take_ticket, which also sets up
// h.state): ticket_begin issues the atomic, ticket_fetch turns its result into (chunk, position) loads, and
// nothing waits until the values are first USED -- the caller keeps whole stages of K1 between the three steps,
// so both global round trips overlap compute without relying on how divergent paths of a warp are scheduled.
__device__ __forceinline__ uint32_t ticket_begin(const Handout& h) { return atomicAdd(&h.ctr->ticket, 1u); }

__device__ __forceinline__ Ticket ticket_fetch(const Handout& h, uint32_t t) {
    Ticket tk;
    tk.chunk = TICKET_DONE; tk.px = tk.py = tk.pz = 0;
    if (h.order) {
        if (t < h.state[UW_NCLS - 1]) {
            int cls = 0;
            while (t >= h.state[cls]) ++cls;
            const uint32_t before = cls ? h.state[cls - 1] : 0u;
            const uint4 e = __ldcg(&h.order[(size_t)cls * h.n + (t - before)]);
            tk.chunk = e.x; tk.px = (int)e.y; tk.py = (int)e.z; tk.pz = (int)e.w;
        }
        return tk;
    }
    if (t < h.n) {
        tk.px = h.pos[3 * t]; tk.py = h.pos[3 * t + 1]; tk.pz = h.pos[3 * t + 2];
        check_position(h.ctr, tk.px, tk.py, tk.pz);
        if (h.skip) {                                  // UW_FLAG_ANALYTIC_SKIP: may have to move on to the next ticket (blocking)
            if (answer_trivial(h, t, tk.px, tk.py, tk.pz)) return take_ticket(h);
        }
        tk.chunk = t;
    }
    return tk;
}

// ---------------------------------------------------------------------------------------
// K1 (fast path, compile-time specialised): same algorithm as k_noise_small with S and the
// octave count as template parameters, so that every loop bound, lattice size and -- crucially --
// every "did the z column enter the next noise cell" test folds at compile time.  Requires
// (host-verified) that the axis cell table equals (i << o) / S, true for the reference's
// constants (SIZE_SCALE = f32(16/S) rounds up or is exact for S = 10, 12, 64).
//
// Arithmetic per octave-sample (z stage): with (R0,S0) / (R1,S1) the y-lerped value/z-slope at
// the low/high noise-lattice plane,  v = a0 + w (a1 - a0),  a0 = R0 + S0 d,  a1 = R1 + S1 (d-1)
//   =>  v = fma(w, fma(d, D, C), fma(d, S0, R0)),   C = (R1 - R0) - S1,  D = S1 - S0
// The reference's clamp of every octave to [-1, 1] costs nothing: stage X carries the octave normalised to its
// clamp range and shifted, u = v / sqrt(3) + 1/2 (the shift rides on the value component through the y and z
// lerps, whose weights sum to 1), so the clamp is the .sat of the last FFMA, and the octave weight
// 2^-o / sum(2^-o) is the multiplier of the FFMA that accumulates the octaves onto (terrace term - 1):
// 3 FFMA (one .SAT) + 1 FFMA per octave-sample (it was 3 FFMA + 2 FMNMX + 1 FADD).
// ---------------------------------------------------------------------------------------
template <int ST, int NOCT>
struct SpecDims {
    static constexpr int S = ST, L = ST + 1;
    // Noise-lattice planes per axis and octave.  The chunk spans 2^o noise cells -> 2^o + 1 lattice planes, plus
    // one more ONLY for the chunk's last sample (k = S), which sits a hair (2^o * 3e-8) inside the next cell when
    // SIZE_SCALE = f32(16/S) rounds up (chunk.rs:7).  Its weight on that extra plane is fade(3e-8) ~ 1e-22 -- far
    // below half an ulp of anything it is added to -- so when that holds for every octave (PRUNE; true for
    // S = 10, 12 and, with weight exactly 0, 64) the plane is never hashed, lerped or read: 160 instead of 307
    // hash chains and 494 instead of 793 x-lerps per 13^3 chunk.  (The exact f64 guard-band path keeps the term.)
    __host__ __device__ static constexpr bool prune() {
        for (int o = 0; o < NOCT; ++o) {
            if (!(tab_w(o, ST) < 1e-15f) || cell(o, ST) != (1 << o) || cell(o, ST - 1) >= (1 << o)) return false;
        }
        return true;
    }
    static constexpr bool PRUNE = prune();
    __host__ __device__ static constexpr int G(int o) { return (1 << o) + (PRUNE ? 1 : 2); }
    __host__ __device__ static constexpr int lat_base(int o) { int b = 0; for (int q = 0; q < o; ++q) b += G(q) * G(q) * G(q); return b; }   // prefix of G^3
    __host__ __device__ static constexpr int x_base(int o) { int b = 0; for (int q = 0; q < o; ++q) b += G(q) * G(q); return L * b; }       // L * prefix of G^2
    static constexpr int DSTRIDE = (L * L * L + 3) & ~3;
    static constexpr int NT = ((L * L + 31) / 32) * 32;          // one thread per (x, y) column: the stand-alone noise kernel
    // The fused kernel runs one warp more than the columns need (7 instead of 6 at L = 13): stage YZ leaves it idle,
    // but the stages whose item counts are not tied to the columns (H, X, the K4 passes) and the latency hiding of
    // every stage gain more than 72 instead of 80 registers cost: -1.8 % at 2048 chunks, -2.6 % at 524 288 (it was
    // -4 % / +3 % before the K1 rework); an eighth warp (64 registers) gives it back.  The stand-alone noise kernel
    // is faster with 5 x 6 warps than with 4 x 7 (209 vs 222 us at 32 768 chunks).
#ifndef UW_FUSED_EXTRA_WARPS
#define UW_FUSED_EXTRA_WARPS 1
#endif
    static constexpr int NTF = NT + 32 * UW_FUSED_EXTRA_WARPS;
    __host__ __device__ static constexpr int cell(int o, int k) { return (k << o) / ST; }
    // compile-time axis tables for CHUNK_SIZE = 16 (chunk.rs:5): same f64 arithmetic as setup_tables();
    // the host selects this kernel only if its runtime tables match these bit for bit.
    __host__ __device__ static constexpr float size_scale() { return 16.0f / (float)ST; }
    __host__ __device__ static constexpr double axis_p(int o, int k) {
        return ((((double)k * (double)size_scale()) + 0.0) / 16.0) * (double)(1 << o);
    }
    __host__ __device__ static constexpr double cfloor(double x) { return (double)(long long)x; }   // x >= 0 here
    __host__ __device__ static constexpr double axis_frac(int o, int k) { return axis_p(o, k) - cfloor(axis_p(o, k)); }
    __host__ __device__ static constexpr double cfade(double c) { return (c * c * c) * (c * (c * 6.0 + (-15.0)) + 10.0); }
    __host__ __device__ static constexpr float tab_d(int o, int k) { return (float)axis_frac(o, k); }
    __host__ __device__ static constexpr float tab_w(int o, int k) { return (float)cfade(axis_frac(o, k)); }
    __host__ __device__ static constexpr int tab_c(int o, int k) { return (int)cfloor(axis_p(o, k)); }
    static constexpr int LAT = lat_base(NOCT), XN = x_base(NOCT), GTOP = G(NOCT - 1);
    // Stage X by cell groups (see noise_chunk_spec): when the finest octave's noise cells hold a whole number GS of
    // samples (S = 12, 3 octaves: 4 cells of 3), the samples i = q GS .. q GS + GS - 1 lie in ONE cell of every octave,
    // c = (q << o) >> (NOCT - 1), so one thread can lerp them all from a single pair of lattice points.
    static constexpr int NGRP = 1 << (NOCT - 1), GS = ST / NGRP;
    __host__ __device__ static constexpr int g2_base(int o) { int b = 0; for (int q = 0; q < o; ++q) b += G(q) * G(q); return b; }   // prefix of G^2
    static constexpr int SG2 = g2_base(NOCT);
    __host__ __device__ static constexpr int h_rounds() { int r = 0; for (int o = 0; o < NOCT; ++o) r += (G(o) * G(o) * G(o) + 31) / 32; return r; }
    static constexpr bool XGROUPED = PRUNE && (ST % NGRP == 0) && GS >= 2;
};

template <int ST, int NOCT>
struct SpecSmem {
    using D = SpecDims<ST, NOCT>;
    float4 X[D::XN];
    // X (+ this pad) is dead after the noise stages; the fused kernel reuses the region for the vertex-id table of K4
    // (L^3 * 5 u16).  lat is NOT part of it: it is dead after stage X already, and the fused kernel's spare warp hashes
    // the NEXT chunk's lattice into it while this chunk's columns are walked (see noise_chunk_spec, PF) -- those
    // gradients must survive this chunk's K2..K4.
    static constexpr int VID_BYTES = D::L * D::L * D::L * 5 * 2;
    static constexpr int PAD0 = VID_BYTES > D::XN * 16 ? ((VID_BYTES - D::XN * 16 + 15) / 16) * 16 : 16;
    static constexpr int PAD = PAD0 > D::GTOP * 16 ? PAD0 : D::GTOP * 16;   // >= one X row: see stage X (PRUNE)
    unsigned char xpad[PAD];
    float4 lat[D::LAT];
    float dens[D::DSTRIDE];
    float4 grad[16];
    float4 axis[NOCT][D::L + 1];   // (d, d - 1, fade(d), -) per octave and lattice index
    float terr[2][16];         // adj_z - fmod(adj_z, mod) - 1 of this chunk / the prefetched next chunk
    uint32_t mask[D::L * D::L + 3];
    uint8_t perm[256];
    int red[2][3];             // block votes of the chunk [tb]: all samples > iso_level, any inside, all inside
    // fused kernel: (chunk index, chunk position) of the current ticket [tb] and of the next one [tb ^ 1].  Kept here,
    // not in registers: nothing in K1 needs them on its fast path (the guard band, the descriptor and K4 read them when
    // they get there), and the registers they would hold across the whole iteration are what ptxas otherwise spills --
    // the reloads sat at the head of stages X and YZ (4 % of the stall samples at config 3).
    int ticket[2][4];
#ifdef UW_PHASE_TIMING
    uint32_t arrive[7];        // clock() of every warp's arrival at the barrier that ends stage YZ
#endif
};
static_assert(SpecDims<12, 3>::L <= 16 && SpecDims<10, 3>::L <= 16, "terr rows hold 16 entries");

// terrace term of perlin_util.rs:27-28, minus 1: every octave value is carried as u = v / (2 lim) + 1/2 in [0, 1]
// (see stage YZ) and sum_o lim_o = 1.  Tabulated per z layer at context creation by the SAME device function
// (k_terrace_table), so a table hit is bit-identical to the in-place evaluation.
__device__ __forceinline__ float terrace_minus_one(const DevCfg& cfg, int k, int pz) {
    float adj, fm;
    terrace_terms(cfg, k, pz, adj, fm);
    return __fsub_rn(__fsub_rn(adj, fm), 1.0f);
}
__device__ __forceinline__ float terrace_lookup(const DevCfg& cfg, int k, int pz) {
    const unsigned dz = (unsigned)(pz - cfg.terr_z0);
    if (cfg.terr_tab != nullptr && dz < (unsigned)cfg.terr_nz) return __ldg(cfg.terr_tab + dz * 16u + (unsigned)k);
    return terrace_minus_one(cfg, k, pz);
}
__global__ void __launch_bounds__(256) k_terrace_table(DevCfg cfg, float* __restrict__ tab, int z0, int nz) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nz * 16; t += gridDim.x * blockDim.x) {
        const int k = t & 15;
        tab[t] = k < cfg.L ? terrace_minus_one(cfg, k, z0 + (t >> 4)) : 0.f;
    }
}

// stage H of one chunk for the items t = first, first + step, ...: the gradient vector of every touched noise-lattice point
template <int ST, int NOCT>
__device__ __forceinline__ void noise_stage_h(SpecSmem<ST, NOCT>& sm, int px, int py, int pz, int first, int step) {
    using D = SpecDims<ST, NOCT>;
#pragma unroll
    for (int o = 0; o < NOCT; ++o) {
        const int G = D::G(o), F = 1 << o, base = D::lat_base(o);
        for (int t = first; t < G * G * G; t += step) {
            const int cx = t / (G * G), r = t - cx * G * G, cy = r / G, cz = r - cy * G;
            const uint32_t h = sm.perm[sm.perm[sm.perm[(F * px + cx) & 255] ^ ((F * py + cy) & 255)] ^ ((F * pz + cz) & 255)];
            sm.lat[base + t] = sm.grad[h & 15u];
        }
    }
}

// stage H of one chunk by ONE warp (the fused kernel's spare warp).  tools/phase_timing.py: with the loop above the
// spare warp is the LAST to reach the barrier that ends stage YZ in ~87 % of the chunks, by ~600 cycles: 6 rounds of
// (3 dependent table loads + gradient load + store), one after the other.  Same hash, factorised: perm[perm[X] ^ Y]
// takes only G^2 values per octave (38 in all) -- one lane each, handed round by shuffle -- and the G^3 last-level
// steps (one table load, the gradient load, the store) are branch-free (lanes past the end repeat the last point), so
// the rounds are independent instruction streams the scheduler can overlap.
template <int ST, int NOCT>
__device__ __forceinline__ void noise_stage_h_warp(SpecSmem<ST, NOCT>& sm, int px, int py, int pz, int lane_in) {
    using D = SpecDims<ST, NOCT>;
    constexpr int OT = NOCT - 1, GT = D::G(OT), NLOW = D::g2_base(OT);
    static_assert(GT * GT <= 32 && NLOW <= 32, "the second-level hashes of the top octave / of the others fit one warp each");
    int lane = lane_in;
    asm volatile("" : "+r"(lane));      // keep the index arithmetic here: hoisted out of the chunk loop it is spilled
    uint32_t hb_top, hb_low;
    {
        const int l = lane < GT * GT ? lane : GT * GT - 1;
        const int cx = l / GT, cy = l - cx * GT;
        hb_top = sm.perm[sm.perm[((px << OT) + cx) & 255] ^ (((py << OT) + cy) & 255)];
    }
    {
        int o = 0, cx = 0, cy = 0;
#pragma unroll
        for (int p = 0; p < OT; ++p) {
            const int r = lane - D::g2_base(p);
            if (r >= 0 && r < D::G(p) * D::G(p)) { o = p; cx = r / D::G(p); cy = r - cx * D::G(p); }
        }
        hb_low = sm.perm[sm.perm[((px << o) + cx) & 255] ^ (((py << o) + cy) & 255)];
    }
    // three passes over the rounds (second-level value + last table load, gradient load, store): written apart so that
    // the loads of all rounds are in flight together instead of one round waiting for the previous round's store
    constexpr int NR = D::h_rounds();
    uint32_t hh[NR];
    int dst[NR];
    {
        int k = 0;
#pragma unroll
        for (int o = 0; o < NOCT; ++o) {
            const int G = D::G(o), base = D::lat_base(o);
#pragma unroll
            for (int t0 = 0; t0 < G * G * G; t0 += 32, ++k) {
                const int t = t0 + lane;
                const int tt = t < G * G * G ? t : G * G * G - 1;
                const int cxy = tt / G, cz = tt - cxy * G;                   // cxy = cx G + cy
                const uint32_t hb = __shfl_sync(0xFFFFFFFFu, o == OT ? hb_top : hb_low, (o == OT ? 0 : D::g2_base(o)) + cxy);
                hh[k] = sm.perm[hb ^ (((pz << o) + cz) & 255)];
                dst[k] = base + tt;
            }
        }
    }
    float4 gg[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) gg[k] = sm.grad[hh[k] & 15u];
#pragma unroll
    for (int k = 0; k < NR; ++k) sm.lat[dst[k]] = gg[k];
}

// stages H, X, YZ for one chunk; leaves densities in sm.dens and column sign masks in sm.mask.
// Returns (block-uniform) CF_ALL_GT | CF_ANY_LT.  All threads must call; ends with a barrier.
//
// PF (fused kernel, NT = NTF: one warp more than the columns need): while warps 0..5 walk the columns of THIS chunk
// (stage YZ, the longest stage), the spare warp hashes the NEXT chunk's lattice (stage H) and looks up its terrace
// terms -- lat is dead after stage X, the next ticket is known by then.  Every call then starts at stage X (the
// caller hashes the lattice of its first chunk itself): one block barrier and the whole hash-chain latency (3 dependent shared-memory loads per lattice point,
// 14 % of the kernel's stall samples at config 3, 58 % of them at the barrier) leave the per-chunk critical path.
// tb = which half of sm.terr belongs to this chunk (the prefetch writes the other one).
template <int ST, int NOCT, int NT /*threads of the CTA: SpecDims::NT or ::NTF*/, bool PF = false>
__device__ __forceinline__ uint32_t noise_chunk_spec(const DevCfg& cfg, const AxisTables& tab, SpecSmem<ST, NOCT>& sm,
                                                     int px_arg, int py_arg, int pz_arg, unsigned long long* guard_count PHASE_ARG,
                                                     const Handout* hand = nullptr,
                                                     int tb = 0) {
    using D = SpecDims<ST, NOCT>;
    constexpr int L = D::L;
    const int tid = threadIdx.x;
    // with a hand-out (fused kernel) the chunk position is read from sm.ticket[tb] where it is needed
    const volatile int* const cur = sm.ticket[tb];
    auto PX = [&]() { return hand ? cur[1] : px_arg; };
    auto PY = [&]() { return hand ? cur[2] : py_arg; };
    auto PZ = [&]() { return hand ? cur[3] : pz_arg; };
    // the NEXT chunk's ticket: atomic issued here, (chunk, position) loads after stage X, values first touched by
    // the caller at the end of the iteration -- see ticket_begin / ticket_fetch
    uint32_t tk_t = 0;
    if (hand && tid == NT - 1) tk_t = ticket_begin(*hand);

    // ---- stage H (skipped when the previous call's spare warp has done it) ------------------------------
    if (!PF) {
        noise_stage_h<ST, NOCT>(sm, PX(), PY(), PZ(), tid, NT);
        if (tid >= NT - 32 && tid - (NT - 32) < L)   // terrace term perlin_util.rs:27-28 (last warp: it has idle lanes later)
            sm.terr[tb][tid - (NT - 32)] = terrace_lookup(cfg, tid - (NT - 32), PZ());
        __syncthreads();
    }
    PHASE_MARK(11);

    // ---- stage X --------------------------------------------------------------------------------
    constexpr float inv_max = 1.0f / (2.0f - 1.0f / (float)(1 << (NOCT - 1)));
    // the octave value is carried normalised to its clamp range and shifted: u = v * (2/sqrt(3)) / 2 + 1/2, so
    // that the reference's clamp to [-1, 1] becomes the .sat of the last FFMA of stage YZ; the shift rides on
    // the value component through the y and z lerps (weights sum to 1), the slopes are only scaled
    const float sc = 1.1547005383792515f * 0.5f;
    auto xlerp = [&](const float4 g0, const float4 g1, const float4 ax) {
        const float d = ax.x, d1 = ax.y, w = ax.z;
        const float q0 = g0.x * d, q1 = g1.x * d1;
        float4 e;
        e.x = fmaf(fmaf(w, q1 - q0, q0), sc, 0.5f);
        e.y = fmaf(w, g1.y - g0.y, g0.y) * sc;
        e.z = fmaf(w, g1.z - g0.z, g0.z) * sc;
        e.w = 0.f;
        return e;
    };
#ifndef UW_NO_X_GROUPED
    constexpr int XITEMS = D::NGRP * D::SG2;                   // (cell group q, octave, lattice (cy, cz)) items, q-major
    constexpr int XTOP0 = NT >= ((XITEMS + 31) & ~31) + D::SG2 ? ((XITEMS + 31) & ~31) : XITEMS;   // tops start a warp if room
    constexpr bool XG = D::XGROUPED && XTOP0 + D::SG2 <= NT;
#else
    constexpr int XITEMS = 0, XTOP0 = 0;
    constexpr bool XG = false;
#endif
    if constexpr (XG) {
        // ONE round: thread (q, o, r) loads its two lattice points once and writes the GS x-samples of cell group q;
        // SG2 more threads write the chunk's last sample plane (i = S: both points are the plane G - 1, weight ~ 0 on
        // the pruned one).  Same operations per output as the per-(o, i, r) loop below, a third of its loads and index
        // arithmetic, and no warp takes more than one round.
        int u = -1, q = 0;
        int xt = tid;
#ifndef UW_X_HOIST
        // keep the (chunk-independent) item decode inside the chunk loop: hoisted, its results are spilled to local
        // memory by the 72-register budget and reloaded at the head of this stage's dependent chain
        asm volatile("" : "+r"(xt));
#endif
        const bool top = xt >= XTOP0;
        if (xt < XITEMS) { q = xt / D::SG2; u = xt - q * D::SG2; }
        else if (top && xt < XTOP0 + D::SG2) u = xt - XTOP0;
        if (u >= 0) {
            int o = 0, r = u, G2 = D::G(0) * D::G(0), lb = D::lat_base(0), xb = D::x_base(0);
#pragma unroll
            for (int p = 1; p < NOCT; ++p)
                if (u >= D::g2_base(p)) { o = p; r = u - D::g2_base(p); G2 = D::G(p) * D::G(p); lb = D::lat_base(p); xb = D::x_base(p); }
            const float4* axo = sm.axis[o];
            float4* xo = sm.X + xb + r;
            if (!top) {
                const int c = (q << o) >> (NOCT - 1);
                const float4 g0 = sm.lat[lb + c * G2 + r];
                const float4 g1 = sm.lat[lb + (c + 1) * G2 + r];
                // all loads first, then the arithmetic, then the stores: the GS outputs are independent chains
                float4 ax[D::GS], e[D::GS];
#pragma unroll
                for (int m = 0; m < D::GS; ++m) ax[m] = axo[q * D::GS + m];
#pragma unroll
                for (int m = 0; m < D::GS; ++m) e[m] = xlerp(g0, g1, ax[m]);
#pragma unroll
                for (int m = 0; m < D::GS; ++m) xo[(q * D::GS + m) * G2] = e[m];
            } else {
                const float4 g = sm.lat[lb + (G2 << o) + r];              // plane G - 1 = 2^o
                xo[ST * G2] = xlerp(g, g, axo[ST]);
            }
        }
    } else {
#pragma unroll
    for (int o = 0; o < NOCT; ++o) {
        const int G = D::G(o), lb = D::lat_base(o), xb = D::x_base(o);
        // Items are dealt round-robin ACROSS the octaves (thread rotation = items of the octaves before this one, mod NT):
        // no warp takes more than ceil(all items / NT) rounds -- 3 instead of 4 for warps 0-1 at S = 12 -- without a
        // per-item octave selection (measured slower in round 1).  The slowest warp sets the time of the barrier below.
#ifndef UW_NO_X_ROTATION
        const int rot = D::x_base(o) % NT;                      // folds: the octave loop is unrolled
#else
        const int rot = 0;
#endif
#ifndef UW_K1_HOIST
        int xtid = tid;
        asm volatile("" : "+r"(xtid));          // see stage YZ: keep the item index arithmetic inside the chunk loop
#else
        const int xtid = tid;
#endif
        for (int t = xtid >= rot ? xtid - rot : xtid - rot + NT; t < L * G * G; t += NT) {
            const int i = t / (G * G), r = t - i * G * G;
            const int c = (i << o) / ST;
            const int c1 = D::PRUNE ? min(c + 1, G - 1) : c + 1;         // i = S: weight ~ 1e-22 on a plane that is not kept
            sm.X[xb + t] = xlerp(sm.lat[lb + c * G * G + r], sm.lat[lb + c1 * G * G + r], sm.axis[o][i]);
        }
    }
    }
    // PRUNE: a column with j = S reads the X row "one past" its last kept row with weight ~ 1e-22; that is row 0 of
    // the next x-plane / the next octave (finite values) or, for the very last one, this pad row: keep it finite
    if (D::PRUNE && tid < D::GTOP) sm.X[D::XN + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) { sm.red[tb][0] = 1; sm.red[tb][1] = 0; sm.red[tb][2] = 1; }
    __syncthreads();
    PHASE_MARK(12);
    // the next ticket goes to sm.ticket[tb ^ 1]: by the last thread here, or (PF) by the spare warp's last lane below
    if (!PF && hand && tid == NT - 1) {
        const Ticket nx = ticket_fetch(*hand, tk_t);
        sm.ticket[tb ^ 1][0] = (int)nx.chunk; sm.ticket[tb ^ 1][1] = nx.px; sm.ticket[tb ^ 1][2] = nx.py; sm.ticket[tb ^ 1][3] = nx.pz;
    }

    // ---- stage YZ -------------------------------------------------------------------------------
    // lanes of this warp that own a column: taken while the warp is still converged, so that the votes at the end
    // of the branch name their participants explicitly (the guard-band loop in between diverges)
    const unsigned col_lanes = __ballot_sync(0xFFFFFFFFu, tid < L * L);
    if (tid < L * L) {
        // the column's indices and table pointers do not depend on the chunk: hoisted out of the chunk loop by the
        // compiler they do not fit the 72-register budget and come back as local-memory reloads at the head of this
        // stage (37 % L1 misses) -- recomputing them here is ~15 instructions per column
#ifndef UW_K1_HOIST
        int ytid = tid;
        asm volatile("" : "+r"(ytid));
#else
        const int ytid = tid;
#endif
        const int i = ytid / L, j = ytid - i * L;
        float R0[NOCT], S0[NOCT], R1[NOCT], S1[NOCT], Cc[NOCT], Dd[NOCT];
        const float4* xrow[NOCT];
        float dy[NOCT], dy1[NOCT], wy[NOCT];
#pragma unroll
        for (int o = 0; o < NOCT; ++o) {
            const int G = D::G(o);
            const int cyj = (j << o) / ST;
            xrow[o] = sm.X + D::x_base(o) + (i * G + cyj) * G;
            const float4 ay = sm.axis[o][j];
            dy[o] = ay.x; dy1[o] = ay.y; wy[o] = ay.z;
        }
        auto ystage = [&](int o, int cz, float& R, float& Sz) {
            const int G = D::G(o);
            const float4 E0 = xrow[o][cz];
            const float4 E1 = xrow[o][G + cz];
            const float A0 = fmaf(E0.y, dy[o], E0.x);
            const float A1 = fmaf(E1.y, dy1[o], E1.x);
            R = fmaf(wy[o], A1 - A0, A0);
            Sz = fmaf(wy[o], E1.z - E0.z, E0.z);
        };
        float* out = sm.dens + ytid * L;
        const float isl = cfg.iso_level, eps = cfg.guard_eps;
        uint32_t signs = 0;                       // bit (L-1-k) <- (iso_k < iso_level), shifted in MSB-first
        float nearest = 3.0e38f;                  // min |iso - iso_level| of the column: one FMNMX per sample
        float4 terr4 = make_float4(0.f, 0.f, 0.f, 0.f); (void)terr4;
#pragma unroll
        for (int k = 0; k < L; ++k) {
#ifndef UW_NO_TERR_VEC4               // one 16-byte load per four samples (-0.4 % at config 3 once the spare warp is off the critical path)
            if ((k & 3) == 0) terr4 = reinterpret_cast<const float4*>(sm.terr[tb])[k >> 2];
            float total = (k & 3) == 0 ? terr4.x : (k & 3) == 1 ? terr4.y : (k & 3) == 2 ? terr4.z : terr4.w;
#else
            float total = sm.terr[tb][k];                                // terrace term - 1
#endif
#pragma unroll
            for (int o = 0; o < NOCT; ++o) {
                const int c = D::cell(o, k);
                const bool first = (k == 0), step = (k > 0) && (c != D::cell(o, k > 0 ? k - 1 : 0));
                // PRUNE: the last sample's weight on the far plane is ~ 1e-22 (or 0): v = a0, from the plane already held
                const bool top = D::PRUNE && k == L - 1;
                if (first) { ystage(o, c, R0[o], S0[o]); ystage(o, c + 1, R1[o], S1[o]); }
                else if (step) { R0[o] = R1[o]; S0[o] = S1[o]; if (!top) ystage(o, c + 1, R1[o], S1[o]); }
                if ((first || step) && !top) { Cc[o] = (R1[o] - R0[o]) - S1[o]; Dd[o] = S1[o] - S0[o]; }
                const float d = D::tab_d(o, k), w = D::tab_w(o, k);      // immediates after unrolling
                // FFMA.SAT = the reference's clamp to [-1, 1] in the normalised, shifted scale
                const float u = top ? __saturatef(fmaf(d, S0[o], R0[o]))
                                    : __saturatef(fmaf(w, fmaf(d, Dd[o], Cc[o]), fmaf(d, S0[o], R0[o])));
                total = fmaf(u, 2.0f * inv_max / (float)(1 << o), total);   // octave weight 2^-o / sum(2^-o), times 2 lim
            }
            const float iso = total;
            const float diff = iso - isl;
            out[k] = iso;
            nearest = fminf(nearest, fabsf(diff));
            signs = __funnelshift_l(__float_as_uint(diff), signs, 1);   // (signs << 1) | sign bit of diff
        }
        uint32_t inside = __brev(signs) >> (32 - L);                     // bit k <- (iso_k < iso_level)
        bool any_eq = false;
        uint32_t near = 0;                        // bit k <- sample k fell inside the guard band
        if (nearest < eps) {                      // rare: find which samples (the column is still in shared memory)
#pragma unroll 1
            for (int k = 0; k < L; ++k)
                if (fabsf(out[k] - isl) < eps) near |= 1u << k;
        }
        while (near) {                                                   // rare: exact f64 re-evaluation
            const int k = __ffs(near) - 1;
            near &= near - 1;
            const float iso = x_iso_lattice(cfg, sm.perm, PX(), PY(), PZ(), i, j, k);
            out[k] = iso;
            inside = (inside & ~(1u << k)) | ((iso < isl) ? (1u << k) : 0u);
            any_eq |= (iso == isl);
            atomicAdd(guard_count, 1ull);
        }
        sm.mask[ytid] = inside;
        // outside the guard band |iso - isl| >= eps > 0, so "all > isl" <=> no inside bit and no exact tie
        const bool all_gt = (inside == 0u) && !any_eq, any_lt = inside != 0u;
        const bool w_all = __all_sync(col_lanes, all_gt), w_any = __any_sync(col_lanes, any_lt);
        const bool w_solid = __all_sync(col_lanes, inside == (1u << L) - 1u);           // the all-full vote
        if ((tid & 31) == 0 || tid == (L * L / 32) * 32) {
            if (!w_all) sm.red[tb][0] = 0;
            if (w_any) sm.red[tb][1] = 1;
            if (!w_solid) sm.red[tb][2] = 0;
        }
    } else if (PF && tid >= NT - 32) {
        // the spare warp: the NEXT chunk's ticket (last lane -> sm.ticket[tb ^ 1]), then its stage H + terrace terms
        const int lane = tid - (NT - 32);
        volatile int* const nxt = sm.ticket[tb ^ 1];
        if (lane == 31) {
            const Ticket nx = ticket_fetch(*hand, tk_t);
            nxt[0] = (int)nx.chunk; nxt[1] = nx.px; nxt[2] = nx.py; nxt[3] = nx.pz;
        }
        __syncwarp();
        const uint32_t nchunk = (uint32_t)nxt[0];
        const int npx = nxt[1], npy = nxt[2], npz = nxt[3];
        if (nchunk != TICKET_DONE) {
#ifndef UW_NO_TERR_FIRST
            // global table load first: its latency hides under the hash rounds (the barrier that ends stage YZ waits for
            // this warp: -2.6 % at config 3, profiles/r02_ab_terr_first.txt; writing the six hash rounds out for more
            // loads in flight costs registers and was slower, +4 %)
            float tv = 0.f;
            if (lane < L) tv = terrace_lookup(cfg, lane, npz);
#ifndef UW_NO_H_WARP
            noise_stage_h_warp<ST, NOCT>(sm, npx, npy, npz, lane);
#else
            noise_stage_h<ST, NOCT>(sm, npx, npy, npz, lane, 32);
#endif
            if (lane < L) sm.terr[tb ^ 1][lane] = tv;
#else
            noise_stage_h<ST, NOCT>(sm, npx, npy, npz, lane, 32);
            if (lane < L) sm.terr[tb ^ 1][lane] = terrace_lookup(cfg, lane, npz);
#endif
        }
    }
#ifdef UW_PHASE_TIMING
    if (PF && (tid & 31) == 0 && (tid >> 5) < 7) sm.arrive[tid >> 5] = (uint32_t)clock();
#endif
    __syncthreads();
#ifdef UW_PHASE_TIMING
    if (PF && tid == 0 && NT == 224) {         // who is last at the barrier: the column warps (0..5) or the spare warp (6)?
        int32_t col = 0;
        const uint32_t base = sm.arrive[0];
        for (int w = 1; w < 6; ++w) col = max(col, (int32_t)(sm.arrive[w] - base));
        const int32_t d = (int32_t)(sm.arrive[6] - base) - col;
        if (d > 0) { atomicAdd(&g_phase[13], (unsigned long long)d); atomicAdd(&g_phase[15], 1ull); }
        else atomicAdd(&g_phase[14], (unsigned long long)(-d));
    }
#endif
    // (the flags alternate with tb: a caller that skips its end-of-chunk barrier may be resetting the next chunk's word already)
    return (sm.red[tb][0] ? CF_ALL_GT : 0u) | (sm.red[tb][1] ? CF_ANY_LT : 0u) | (sm.red[tb][2] ? CF_ALL_LT : 0u);
}

template <int ST, int NOCT>
__global__ void __launch_bounds__(SpecDims<ST, NOCT>::NT, 5)
k_noise_spec(const __grid_constant__ DevCfg cfg, const __grid_constant__ AxisTables tab,
             const uint8_t* __restrict__ g_perm, const int32_t* __restrict__ pos, uint32_t n,
             float* __restrict__ dens, unsigned long long* __restrict__ guard_count) {
    using D = SpecDims<ST, NOCT>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SpecSmem<ST, NOCT>& sm = *reinterpret_cast<SpecSmem<ST, NOCT>*>(smem_raw);
    const int tid = threadIdx.x;
    for (int t = tid; t < 256; t += D::NT) sm.perm[t] = g_perm[t];
    if (tid < 16) sm.grad[tid] = make_float4(c_grad_vec[tid][0], c_grad_vec[tid][1], c_grad_vec[tid][2], 0.f);
    for (int t = tid; t < NOCT * D::L; t += D::NT) {
        const int o = t / D::L, i = t - o * D::L;
        sm.axis[o][i] = make_float4(tab.d[o][i], tab.d1[o][i], tab.w[o][i], 0.f);
    }
    __syncthreads();
    for (uint32_t chunk = blockIdx.x; chunk < n; chunk += gridDim.x) {
        const int px = pos[3 * chunk], py = pos[3 * chunk + 1], pz = pos[3 * chunk + 2];
#ifdef UW_PHASE_TIMING
        long long t_phase = 0;
#endif
        noise_chunk_spec<ST, NOCT, D::NT>(cfg, tab, sm, px, py, pz, guard_count PHASE_PASS);
        float4* dst = reinterpret_cast<float4*>(dens + (size_t)chunk * D::DSTRIDE);
        const float4* src = reinterpret_cast<const float4*>(sm.dens);
        for (int t = tid; t < D::DSTRIDE / 4; t += D::NT) dst[t] = src[t];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// Shared extraction helpers (small path): sign bits -> per-column masks -> cases
// ---------------------------------------------------------------------------------------
struct ChunkCounts { uint32_t n_verts, n_inds, flags, pad; };

// Every chunk's range of the packed arenas starts on a 16-byte boundary: vertex allocations are rounded up to an
// even count (2 x 24 B), index allocations to a multiple of 16 bytes.  That is what lets the emit stage write whole
// 16-byte vectors (shared-memory staged, warp-contiguous) -- to HBM or, in the gather path, over NVLink into another
// GPU's memory, where partial-sector stores cost a full packet each.  Descriptors carry the exact counts.
#define UW_VSTAGE_VERTS 16                               // vertices per warp and staging round
#define UW_VSTAGE_BYTES (UW_VSTAGE_VERTS * 24)
__host__ __device__ __forceinline__ uint32_t pad_verts(uint32_t nv) { return (nv + 1u) & ~1u; }
template <typename IndexT>
__host__ __device__ __forceinline__ uint32_t pad_inds(uint32_t ni) { return (ni + (16u / sizeof(IndexT)) - 1u) & ~((16u / (uint32_t)sizeof(IndexT)) - 1u); }

__device__ __forceinline__ uint32_t own_mask_of(int x, int y, int z) {
    // SURVEY App. B.4 ownership table: edges this cell is the FIRST (scan order) holder of
    uint32_t m = 0x4F0u;                       // 4,5,6,7,10 always
    if (y == 0) m |= 0x00Fu;                   // 0,1,2,3
    if (z == 0) m |= 0x200u;                   // 9
    if (x == 0) m |= 0x800u;                   // 11
    if (x == 0 && z == 0) m |= 0x100u;         // 8
    return m;
}

// column (x,y) sign mask: bit z = (iso[x][y][z] < iso_level)
__device__ __forceinline__ uint32_t col_mask(const uint32_t* bits, int col, int L) {
    const int b = col * L;
    const uint32_t lo = bits[b >> 5], hi = bits[(b >> 5) + 1];
    return __funnelshift_r(lo, hi, b & 31) & ((1u << L) - 1u);
}

__device__ __forceinline__ uint32_t case_of(uint32_t m00, uint32_t m10, uint32_t m01, uint32_t m11, int z) {
    const uint32_t a = (m00 >> z) & 3u, b = (m10 >> z) & 3u, c = (m01 >> z) & 3u, d = (m11 >> z) & 3u;
    return (a & 1u) | ((b & 1u) << 1) | ((b >> 1) << 2) | ((a >> 1) << 3)
         | ((c & 1u) << 4) | ((d & 1u) << 5) | ((d >> 1) << 6) | ((c >> 1) << 7);
}

// Loads one chunk's densities into shared memory (optional) and builds the sign bit array.
// Returns block-uniform flags (CF_ALL_GT, CF_ANY_LT).  All threads must call.
template <bool KEEP>
__device__ __forceinline__ uint32_t load_signs(const DevCfg& cfg, const float* __restrict__ src,
                                               float* s_dens, uint32_t* s_bits, int* s_red) {
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31;
    const int L3 = cfg.L3;
    if (tid < 2) s_red[tid] = (tid == 0) ? 1 : 0;   // [0] = all_gt (and), [1] = any_lt (or)
    __syncthreads();
    bool all_gt = true, any_lt = false;
    for (int base = tid - lane; base < L3; base += NT) {
        const int idx = base + lane;
        float v = 0.f;
        const bool ok = idx < L3;
        if (ok) { v = __ldg(src + idx); if (KEEP) s_dens[idx] = v; }
        const bool lt = ok && (v < cfg.iso_level);
        const bool gt = !ok || (v > cfg.iso_level);
        const uint32_t w = __ballot_sync(0xFFFFFFFFu, lt);
        if (lane == 0) s_bits[base >> 5] = w;
        all_gt &= gt; any_lt |= lt;
    }
    if (tid == 0) s_bits[(L3 + 31) >> 5] = 0;       // padding word read by col_mask's funnel shift
    const bool w_all = __all_sync(0xFFFFFFFFu, all_gt);
    const bool w_any = __any_sync(0xFFFFFFFFu, any_lt);
    if (lane == 0) { if (!w_all) atomicAnd(&s_red[0], 0); if (w_any) atomicOr(&s_red[1], 1); }
    __syncthreads();
    return (s_red[0] ? CF_ALL_GT : 0u) | (s_red[1] ? CF_ANY_LT : 0u);
}

// block-wide exclusive scan of two 32-bit counters packed per thread; returns this thread's
// exclusive prefix and the block totals.  NT <= 1024.
__device__ __forceinline__ void block_scan2(uint32_t v, uint32_t i, uint32_t& ev, uint32_t& ei,
                                            uint32_t& tv, uint32_t& ti, uint32_t* s_w /*[64]*/) {
    // ONE barrier: warp scans by shuffle, warp totals to shared memory, then every thread sums the totals of
    // the warps before its own.  The caller must have a barrier between two calls that reuse s_w (all do).
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = (blockDim.x + 31) >> 5;
    uint32_t sv = v, si = i;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, sv, d), b = __shfl_up_sync(0xFFFFFFFFu, si, d);
        if (lane >= d) { sv += a; si += b; }
    }
    if (lane == 31) { s_w[warp] = sv; s_w[32 + warp] = si; }
    __syncthreads();
    uint32_t wv = 0, wi = 0, av = 0, ai = 0;
    if (nw > 8) {
        // many warps: every warp scans the warp totals itself (lane l holds warp l's total) instead of looping
        uint32_t a = lane < nw ? s_w[lane] : 0u, b = lane < nw ? s_w[32 + lane] : 0u;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, a, d), y = __shfl_up_sync(0xFFFFFFFFu, b, d);
            if (lane >= d) { a += x; b += y; }
        }
        av = __shfl_sync(0xFFFFFFFFu, a, nw - 1); ai = __shfl_sync(0xFFFFFFFFu, b, nw - 1);
        const uint32_t pa = __shfl_sync(0xFFFFFFFFu, a, warp > 0 ? warp - 1 : 0), pb = __shfl_sync(0xFFFFFFFFu, b, warp > 0 ? warp - 1 : 0);
        if (warp > 0) { wv = pa; wi = pb; }
    } else {
        for (int w = 0; w < nw; ++w) {
            const uint32_t a = s_w[w], b = s_w[32 + w];
            av += a; ai += b;
            if (w < warp) { wv += a; wi += b; }
        }
    }
    ev = wv + sv - v; ei = wi + si - i;
    tv = av; ti = ai;
}

// ---------------------------------------------------------------------------------------
// K2: classify + count.  One CTA per chunk.  Ballot-based blank (all > iso) / no-surface
// skip; optional coalesced case-byte output (debug tap).  Writes per-chunk (V, I, flags).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_classify_small(const __grid_constant__ DevCfg cfg,
                                                        const McTables* __restrict__ mc,
                                                        const float* __restrict__ dens, uint32_t n,
                                                        ChunkCounts* __restrict__ counts,
                                                        uint8_t* __restrict__ cases_out /*nullable*/) {
    __shared__ uint32_t s_bits[(UW_SMALL_MAX_L * UW_SMALL_MAX_L * UW_SMALL_MAX_L + 31) / 32 + 2];
    __shared__ int s_red[2];
    __shared__ uint32_t s_w[64];
    __shared__ __align__(16) uint8_t s_cases[(UW_SMALL_MAX_L - 1) * (UW_SMALL_MAX_L - 1) * (UW_SMALL_MAX_L - 1) + 16];
    const int tid = threadIdx.x, NT = blockDim.x;
    const int S = cfg.S, L = cfg.L, S3 = S * S * S;

    for (uint32_t chunk = blockIdx.x; chunk < n; chunk += gridDim.x) {
        const uint32_t fl = load_signs<false>(cfg, dens + (size_t)chunk * cfg.dens_stride, nullptr, s_bits, s_red);
        uint32_t nv = 0, ni = 0;
        if (fl & CF_ANY_LT) {             // block-uniform: otherwise every case is 0 -> nothing to count
            for (int col = tid; col < S * S; col += NT) {
                const int x = col / S, y = col - x * S;
                const uint32_t m00 = col_mask(s_bits, x * L + y, L), m10 = col_mask(s_bits, (x + 1) * L + y, L);
                const uint32_t m01 = col_mask(s_bits, x * L + y + 1, L), m11 = col_mask(s_bits, (x + 1) * L + y + 1, L);
                const uint32_t any = m00 | m10 | m01 | m11, all = m00 & m10 & m01 & m11;
                const bool trivial = (any == 0u) || (all == ((1u << L) - 1u));
                for (int z = 0; z < S; ++z) {
                    uint32_t cs = 0;
                    if (!trivial) {
                        cs = case_of(m00, m10, m01, m11, z);
                        if (cs != 0u && cs != 255u) {
                            ni += mc->ninds[cs];
                            nv += __popc((uint32_t)mc->crossed[cs] & own_mask_of(x, y, z));
                        }
                    } else if (any) cs = 255u;
                    if (cases_out) s_cases[col * S + z] = (uint8_t)cs;
                }
            }
        } else if (cases_out) {
            for (int t = tid; t < S3; t += NT) s_cases[t] = 0;
        }
        uint32_t ev, ei, tv, ti;
        block_scan2(nv, ni, ev, ei, tv, ti, s_w);
        if (tid == 0) {
            ChunkCounts c;
            c.n_verts = tv; c.n_inds = ti; c.flags = fl; c.pad = 0;
            counts[chunk] = c;
        }
        if (cases_out) {
            __syncthreads();
            uint8_t* dst = cases_out + (size_t)chunk * S3;
            if ((S3 & 15) == 0 && ((size_t)dst & 15) == 0) {
                const uint4* s4 = reinterpret_cast<const uint4*>(s_cases);
                uint4* d4 = reinterpret_cast<uint4*>(dst);
                for (int t = tid; t < (S3 >> 4); t += NT) d4[t] = s4[t];
            } else {
                for (int t = tid; t < S3; t += NT) dst[t] = s_cases[t];
            }
        }
        __syncthreads();
    }
}

// Compile-time-sized variant of load_signs (lattice L = ST + 1, NT threads): the loop is fully unrolled, so all
// of a thread's loads of the chunk are in flight together and the memory latency is paid once per chunk
// instead of once per 32-sample word.  Per-warp flags go to s_flag[NT / 32]; ends with a barrier.
template <int ST, int NT, bool KEEP>
__device__ __forceinline__ uint32_t load_signs_spec(float iso, const float* __restrict__ src, float* s_dens,
                                                    uint32_t* s_bits, uint32_t* s_flag) {
    constexpr int L = ST + 1, L3 = L * L * L, NLD = (L3 + NT - 1) / NT, NW = NT / 32, NWORD = (L3 + 31) / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float v[NLD];
#pragma unroll
    for (int k = 0; k < NLD; ++k) {
        const int idx = tid + NT * k;
        v[k] = ((k + 1) * NT <= L3 || idx < L3) ? __ldg(src + idx) : 0.f;
    }
    bool all_gt = true, any_lt = false;
#pragma unroll
    for (int k = 0; k < NLD; ++k) {
        const int idx = tid + NT * k;
        const bool ok = (k + 1) * NT <= L3 || idx < L3;
        if (KEEP && ok) s_dens[idx] = v[k];
        const bool lt = ok && (v[k] < iso);
        all_gt &= !ok || (v[k] > iso); any_lt |= lt;
        const uint32_t w = __ballot_sync(0xFFFFFFFFu, lt);
        if (lane == 0 && NT * k + 32 * warp < L3) s_bits[(NT * k >> 5) + warp] = w;
    }
    if (tid == 0) { s_bits[NWORD] = 0; s_bits[NWORD + 1] = 0; }      // padding read by col_mask's funnel shift
    const bool w_all = __all_sync(0xFFFFFFFFu, all_gt);
    const bool w_any = __any_sync(0xFFFFFFFFu, any_lt);
    if (lane == 0) s_flag[warp] = (w_all ? CF_ALL_GT : 0u) | (w_any ? CF_ANY_LT : 0u);
    __syncthreads();
    uint32_t a = CF_ALL_GT, o = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) { const uint32_t f = s_flag[w]; a &= f; o |= f; }
    return a | (o & CF_ANY_LT);
}

// 8-bit corner pattern of cell z of a column: bits (m00 z, m00 z+1, m10 z, m10 z+1, m01 z, m01 z+1, m11 z, m11 z+1)
__device__ __forceinline__ uint32_t natural_of(uint32_t q0 /*m00 | m10<<16*/, uint32_t q1 /*m01 | m11<<16*/, int z) {
    const uint32_t a = (q0 >> z) & 0x00030003u, b = (q1 >> z) & 0x00030003u;
    return ((a | (a >> 14)) & 0xFu) | (((b | (b >> 14)) & 0xFu) << 4);
}

// ---- bulk async copy (TMA engine, no tensor map) + mbarrier helpers --------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-B aligned; completion is signalled on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

// K2, compile-time sizes (internal_size 12 / 10), built to be HBM-bound: one CTA per chunk, two threads per
// (x, y) column.  Densities arrive through a 3-stage ring of bulk async copies (one elected thread, mbarrier
// completion), so up to three chunks per CTA are in flight while the current one is counted; the 256-entry
// pattern table lives in shared memory, and there is ONE block barrier per chunk (sign words, warp flags and
// warp partials rotate through three slots; thread 0 writes chunk k's counts during chunk k+1).
#ifndef UW_CLS_MINB
#define UW_CLS_MINB 7
#endif
template <int ST>
struct ClsDims {
    static constexpr int S = ST, L = ST + 1, L3 = L * L * L, NCOL = ST * ST;
    static constexpr int NT = ((2 * NCOL + 31) / 32) * 32, NW = NT / 32;
    static constexpr int NLD = (L3 + NT - 1) / NT, NWORD = (L3 + 31) / 32;
    static constexpr int HALF = (ST + 1) / 2;
    static constexpr int NSTAGE = 3;
    static constexpr int BYTES = ((L3 * 4 + 15) / 16) * 16;          // <= dens_stride * 4 (stride is padded to 4 floats)
};

template <int ST>
__global__ void __launch_bounds__(ClsDims<ST>::NT, UW_CLS_MINB) k_classify_spec(const __grid_constant__ DevCfg cfg,
                                                                                const McTables* __restrict__ mc,
                                                                                const float* __restrict__ dens, uint32_t n,
                                                                                ChunkCounts* __restrict__ counts) {
    using D = ClsDims<ST>;
    constexpr int S = D::S, L = D::L, L3 = D::L3, NT = D::NT, NW = D::NW, NLD = D::NLD, NSTAGE = D::NSTAGE;
    __shared__ __align__(16) float s_dens[NSTAGE][D::BYTES / 4];
    __shared__ __align__(8) uint64_t s_bar[NSTAGE];
    __shared__ uint32_t s_bits[3][D::NWORD + 2];
    __shared__ uint32_t s_flag[3][NW];
    __shared__ uint32_t s_part[3][NW];
    __shared__ uint32_t s_lut[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int t = tid; t < 256; t += NT) s_lut[t] = mc->lut[t];
    if (tid < 3) { s_bits[tid][D::NWORD] = 0; s_bits[tid][D::NWORD + 1] = 0; }
    const float iso = cfg.iso_level;
    const size_t stride = cfg.dens_stride;
    const uint32_t grid = gridDim.x;

    if (tid == 0) {
        for (int st = 0; st < NSTAGE; ++st) mbar_init(&s_bar[st], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int st = 0; st < NSTAGE; ++st) {
            const uint64_t ch = (uint64_t)blockIdx.x + (uint64_t)st * grid;
            if (ch < n) { mbar_expect_tx(&s_bar[st], D::BYTES); bulk_g2s(s_dens[st], dens + ch * stride, D::BYTES, &s_bar[st]); }
        }
    }

    const bool has_col = tid < 2 * D::NCOL;
    const int col = has_col ? (tid >= D::NCOL ? tid - D::NCOL : tid) : 0;
    const int x = col / S, y = col - x * S;
    const int z0 = tid >= D::NCOL ? D::HALF : 0, z1 = tid >= D::NCOL ? S : D::HALF;
    const uint32_t ownn = 0x4F0u | (y == 0 ? 0x00Fu : 0u) | (x == 0 ? 0x800u : 0u);      // SURVEY App. B.4, z > 0
    const uint32_t own0 = ownn | 0x200u | (x == 0 ? 0x100u : 0u);

    auto finalize = [&](uint32_t chunk, int slot) {
        uint32_t a = CF_ALL_GT, o = 0, acc = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) { const uint32_t f = s_flag[slot][w]; a &= f; o |= f; acc += s_part[slot][w]; }
        uint4 c;
        c.x = acc >> 16; c.y = acc & 0xFFFFu; c.z = a | (o & CF_ANY_LT); c.w = 0;
        *reinterpret_cast<uint4*>(counts + chunk) = c;
    };

    __syncthreads();                                     // barriers initialised, table loaded
    uint32_t prev = 0xFFFFFFFFu, it = 0;
    int slot = 0, pslot = 2;                             // slot == it % 3 (ring stage and scratch slot alike)
    for (uint32_t chunk = blockIdx.x; chunk < n; chunk += grid, ++it) {
        mbar_wait(&s_bar[slot], (it / NSTAGE) & 1u);
        const float* src = s_dens[slot];
        bool all_gt = true, any_lt = false;
#pragma unroll
        for (int k = 0; k < NLD; ++k) {
            const bool ok = (k + 1) * NT <= L3 || tid + NT * k < L3;
            const float v = ok ? src[tid + NT * k] : 0.f;
            const bool lt = ok && (v < iso);
            all_gt &= !ok || (v > iso); any_lt |= lt;
            const uint32_t w = __ballot_sync(0xFFFFFFFFu, lt);
            if (lane == 0 && NT * k + 32 * warp < L3) s_bits[slot][(NT * k >> 5) + warp] = w;
        }
        const bool w_all = __all_sync(0xFFFFFFFFu, all_gt);
        const bool w_any = __any_sync(0xFFFFFFFFu, any_lt);
        if (lane == 0) s_flag[slot][warp] = (w_all ? CF_ALL_GT : 0u) | (w_any ? CF_ANY_LT : 0u);
        const int has_lt = __syncthreads_or(w_any);      // sign words visible; every thread is done with this stage
        if (tid == 0) {
            const uint64_t nx = (uint64_t)chunk + (uint64_t)NSTAGE * grid;      // refill the stage just drained
            if (nx < n) { mbar_expect_tx(&s_bar[slot], D::BYTES); bulk_g2s(s_dens[slot], dens + nx * stride, D::BYTES, &s_bar[slot]); }
            if (prev != 0xFFFFFFFFu) finalize(prev, pslot);
        }
        uint32_t acc = 0;                                  // n_inds | n_verts << 16
        if (has_lt && has_col) {
            const uint32_t* bits = s_bits[slot];
            const uint32_t q0 = col_mask(bits, x * L + y, L) | (col_mask(bits, (x + 1) * L + y, L) << 16);
            const uint32_t q1 = col_mask(bits, x * L + y + 1, L) | (col_mask(bits, (x + 1) * L + y + 1, L) << 16);
            // this thread's surface cells straight from the sign masks (see emit_prepare): only they are looked up
            const uint32_t m00 = q0 & 0xFFFFu, dis = (m00 ^ (q0 >> 16)) | (m00 ^ (q1 & 0xFFFFu)) | (m00 ^ (q1 >> 16));
            uint32_t rest = (dis | (dis >> 1) | (m00 ^ (m00 >> 1))) & ((1u << z1) - (1u << z0));
            for (; rest; rest &= rest - 1u) {
                const int z = __ffs(rest) - 1;
                const uint32_t t = s_lut[natural_of(q0, q1, z)];                // case | ninds << 8 | crossed << 12
                acc += ((t >> 8) & 15u) + (__popc((t >> 12) & (z == 0 ? own0 : ownn)) << 16);
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
        if (lane == 0) s_part[slot][warp] = acc;
        prev = chunk; pslot = slot; slot = slot == 2 ? 0 : slot + 1;
    }
    __syncthreads();
    if (tid == 0 && prev != 0xFFFFFFFFu) finalize(prev, pslot);
}

// ---------------------------------------------------------------------------------------
// K3: chunk-level exclusive scan of (V, I) -> descriptors, totals, active-chunk list.
// One CTA per tile of 1024 chunks (one chunk per thread, coalesced 16-B loads / 32-B stores).  Tiles are taken
// by ticket, so every tile's predecessors are already running: a tile publishes its totals (data, fence, epoch
// flag), then sums the published totals of ALL earlier tiles -- one per thread, no chain -- and writes its chunks.
// The flags carry the launch epoch, so nothing is cleared between launches; the last CTA out resets the tickets.
// ---------------------------------------------------------------------------------------
struct ScanPart { unsigned long long v, i; uint32_t a, blank; };
struct ScanCtl { uint32_t ticket, done, emit_ticket, pad1; };   // emit_ticket: work hand-out of the following emit kernel

__global__ void __launch_bounds__(1024) k_scan_chunks(const ChunkCounts* __restrict__ counts,
                                                      const int32_t* __restrict__ pos, uint32_t n,
                                                      uw_chunk_desc* __restrict__ descs,
                                                      uint32_t* __restrict__ active,
                                                      BatchTotals* __restrict__ totals,
                                                      unsigned long long vcap, unsigned long long icap,
                                                      ScanPart* __restrict__ part, uint32_t* __restrict__ flag,
                                                      ScanCtl* __restrict__ ctl, uint32_t epoch, uint32_t ipad_mask) {
    __shared__ uint32_t s_w[4][32];
    __shared__ unsigned long long s_c[2][32];
    __shared__ uint32_t s_ca[2][32];
    __shared__ uint32_t s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(&ctl->ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile, ntiles = gridDim.x;
    const uint32_t idx = tile * 1024u + tid;
    uint4 c = make_uint4(0u, 0u, 0u, 0u);                    // ChunkCounts {n_verts, n_inds, flags, pad}
    if (idx < n) c = *reinterpret_cast<const uint4*>(counts + idx);
    int32_t px = 0, py = 0, pz = 0;
    if (idx < n) { px = pos[3 * idx]; py = pos[3 * idx + 1]; pz = pos[3 * idx + 2]; }
    const uint32_t act = c.y > 0 ? 1u : 0u, blank = (idx < n && (c.z & CF_ALL_GT)) ? 1u : 0u;

    // tile-local inclusive scans of (verts, inds, active, blank); a chunk OCCUPIES padded counts (16-byte aligned
    // allocations, see pad_verts / pad_inds), its descriptor carries the exact ones
    const uint32_t pv = pad_verts(c.x), pi = (c.y + ipad_mask) & ~ipad_mask;
    uint32_t xv = pv, xi = pi, xa = act | (blank << 16);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, xv, d), b = __shfl_up_sync(0xFFFFFFFFu, xi, d),
                       e = __shfl_up_sync(0xFFFFFFFFu, xa, d);
        if (lane >= d) { xv += a; xi += b; xa += e; }
    }
    if (lane == 31) { s_w[0][warp] = xv; s_w[1][warp] = xi; s_w[2][warp] = xa; }
    __syncthreads();
    if (warp == 0) {
        uint32_t a = s_w[0][lane], b = s_w[1][lane], e = s_w[2][lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, a, d), y = __shfl_up_sync(0xFFFFFFFFu, b, d),
                           z = __shfl_up_sync(0xFFFFFFFFu, e, d);
            if (lane >= d) { a += x; b += y; e += z; }
        }
        s_w[0][lane] = a; s_w[1][lane] = b; s_w[2][lane] = e;
        if (lane == 31) {                                    // tile totals: publish, then the epoch flag
            ScanPart p;
            p.v = a; p.i = b; p.a = e & 0xFFFFu; p.blank = e >> 16;
            part[tile] = p;
            __threadfence();
            asm volatile("st.volatile.global.u32 [%0], %1;" :: "l"(flag + tile), "r"(epoch) : "memory");
        }
    }
    __syncthreads();
    const uint32_t lv = (warp ? s_w[0][warp - 1] : 0u) + xv - pv;           // exclusive, tile-local
    const uint32_t li = (warp ? s_w[1][warp - 1] : 0u) + xi - pi;
    const uint32_t la = ((warp ? s_w[2][warp - 1] : 0u) + xa - (act | (blank << 16))) & 0xFFFFu;
    const uint32_t tile_v = s_w[0][31], tile_i = s_w[1][31], tile_a = s_w[2][31] & 0xFFFFu, tile_b = s_w[2][31] >> 16;

    // carry = sum of all earlier tiles' totals
    unsigned long long cv = 0, ci = 0;
    uint32_t ca = 0, cb = 0;
    for (uint32_t t = tid; t < tile; t += 1024u) {
        while (ld_volatile_u32(flag + t) != epoch) { }
        __threadfence();
        const volatile ScanPart* p = part + t;
        cv += p->v; ci += p->i; ca += p->a; cb += p->blank;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        cv += __shfl_xor_sync(0xFFFFFFFFu, cv, d); ci += __shfl_xor_sync(0xFFFFFFFFu, ci, d);
        ca += __shfl_xor_sync(0xFFFFFFFFu, ca, d); cb += __shfl_xor_sync(0xFFFFFFFFu, cb, d);
    }
    if (lane == 0) { s_c[0][warp] = cv; s_c[1][warp] = ci; s_ca[0][warp] = ca; s_ca[1][warp] = cb; }
    __syncthreads();
    cv = 0; ci = 0; ca = 0; cb = 0;
#pragma unroll 8
    for (int w = 0; w < 32; ++w) { cv += s_c[0][w]; ci += s_c[1][w]; ca += s_ca[0][w]; cb += s_ca[1][w]; }

    if (idx < n) {
        const unsigned long long ov = cv + lv, oi = ci + li;
        uint4 d0, d1;
        d0.x = (uint32_t)px; d0.y = (uint32_t)py; d0.z = (uint32_t)pz;
        d0.w = ((c.z & CF_ALL_GT) ? UW_CHUNK_BLANK_EARLY : 0u) | (c.y > 0 ? UW_CHUNK_HAS_MESH : 0u)
             | (c.x > 65536u ? UW_CHUNK_U16_OVERFLOW : 0u);
        d1.x = (uint32_t)ov; d1.y = c.x; d1.z = (uint32_t)oi; d1.w = c.y;
        uint4* dst = reinterpret_cast<uint4*>(descs + idx);  // uw_chunk_desc: pos[3], flags, vert_offset, vert_count, index_offset, index_count
        dst[0] = d0; dst[1] = d1;
        if (act) active[ca + la] = idx;
    }
    if (tile == ntiles - 1 && tid == 0) {
        BatchTotals t;
        t.n_verts = cv + tile_v; t.n_inds = ci + tile_i; t.n_active = ca + tile_a;
        t.overflow = (t.n_verts > vcap || t.n_inds > icap || t.n_verts > 0xFFFFFFFFull || t.n_inds > 0xFFFFFFFFull) ? 1u : 0u;
        t.n_blank = cb + tile_b; t.n_mesh = t.n_active;
        *totals = t;
    }
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&ctl->done, 1u) == ntiles - 1) { ctl->ticket = 0; ctl->done = 0; ctl->emit_ticket = 0; }
    }
}

// ---------------------------------------------------------------------------------------
// colour, chunk.rs:215-222 + util.rs:93-95,106-112,122-153.  f32, unfused, reference order.
// Two of the three channels only depend on the value level -> host-precomputed constants
// (identical to the oracle's); the hue-dependent channel needs one powf.
// ---------------------------------------------------------------------------------------
// powf(x, 2.4f) (util.rs:110) in 36 FP32 instructions, no branches: x = 2^e m, m in [1, 2) falls into one of 16
// intervals with midpoint c_i; r = m / c_i - 1 exactly (one FMA against the tabulated 1/c_i, |r| < 1/32),
// log2 x = (e + hi_i) + (lo_i + r P4(r)) as a two-float sum, times 2.4f with the rounding error of the product
// recovered by FMA, 2^f by a degree-6 polynomial on [-1/2, 1/2], 2^n through the exponent field.  Measured over every
// f32 in the colour's range against f64 pow (tests/test_gpu_parity.py): <= 2 ulp, as CUDA's powf here (which costs ~100
// instructions with branches).  tab = McTables::powtab in shared memory.  Domain: x in [2^-20, 2^20]; else powf.
__device__ __forceinline__ float pow24_tab(float x, const float* __restrict__ tab) {
    const uint32_t bits = __float_as_uint(x);
    const int e = (int)(bits >> 23) - 127;
    const float m = __uint_as_float((bits & 0x007FFFFFu) | 0x3F800000u);
    const uint32_t i = (bits >> 19) & 15u;
    const float r = fmaf(m, tab[i], -1.0f);
    float p = fmaf(0.2888079881668091f, r, -0.3609875738620758f);
    p = fmaf(p, r, 0.48089829087257385f); p = fmaf(p, r, -0.721347451210022f); p = fmaf(p, r, 1.4426950216293335f);
    const float Lh = (float)e + tab[16 + i];                     // exact: hi_i has 16 fractional bits
    const float Ll = fmaf(p, r, tab[32 + i]);
    const float yh = 2.4f * Lh;
    const float yl = fmaf(2.4f, Ll, fmaf(2.4f, Lh, -yh));
    const float n = rintf(yh);
    const float f = (yh - n) + yl;
    float q = fmaf(0.00015461444854736328f, f, 0.0013400427997112274f);
    q = fmaf(q, f, 0.009618056938052177f); q = fmaf(q, f, 0.05550327152013779f); q = fmaf(q, f, 0.24022650718688965f);
    q = fmaf(q, f, 0.6931471824645996f);   q = fmaf(q, f, 1.0f);
    return __uint_as_float(__float_as_uint(q) + ((uint32_t)(int)n << 23));
}

__device__ __forceinline__ float srgb_of(float ch_plus_m, const float* __restrict__ powtab) {
    const float c255 = __fmul_rn(ch_plus_m, 255.0f);
    const float b = __fdiv_rn(__fadd_rn(__fdiv_rn(c255, 255.0f), 0.055f), 1.055f);
#ifndef UW_POWF_LIBM
    if (powtab != nullptr && b >= 9.5367431640625e-07f && b <= 1048576.0f) return pow24_tab(b, powtab);
#endif
    return powf(b, 2.4f);
}

__device__ __forceinline__ void vertex_color(const DevCfg& cfg, float world_z, int vi, float out[3], const float* __restrict__ powtab = nullptr) {
    // x / 2^n == x * 2^-n exactly (no subnormals in reach): the two power-of-two divisions of the reference's
    // constants (CHUNK_SIZE = 16, MAX_Z - MIN_Z = 4) are multiplies; anything else divides
    const float ratio = cfg.cs_pow2 ? __fmul_rn(world_z, (float)cfg.inv_chunk_size) : __fdiv_rn(world_z, (float)cfg.chunk_size);
    const float zr = __fsub_rn(cfg.max_z, cfg.min_z);
    const bool zr_pow2 = (__float_as_uint(zr) & 0x807FFFFFu) == 0u && zr >= 1.0f / 1024.0f && zr <= 1024.0f;
    const float num = __fsub_rn(ratio, cfg.min_z);
    const float mix = zr_pow2 ? __fmul_rn(num, __uint_as_float(0x7F000000u - __float_as_uint(zr))) : __fdiv_rn(num, zr);
    float hue = __fadd_rn(cfg.min_hue, __fmul_rn(__fsub_rn(cfg.max_hue, cfg.min_hue), mix));
    float r = fabsf(hue) < 360.0f ? hue : fmodf(hue, 360.0f);   // f32::rem_euclid, util.rs:123 (fmod is exact)
    if (r < 0.0f) r = __fadd_rn(r, 360.0f);
    hue = r;
    const float c = cfg.hsv_c[vi], m = cfg.hsv_m[vi];
    const float h = __fdiv_rn(hue, 60.0f);
    // h in [0, 6]: fmod(h, 2) == h - 2*floor(h/2) exactly (every step is exact in f32)
    const float hm2 = (h >= 0.0f && h < 16.0f) ? __fsub_rn(h, __fmul_rn(2.0f, floorf(__fmul_rn(h, 0.5f)))) : fmodf(h, 2.0f);
    const float x = __fmul_rn(c, __fsub_rn(1.0f, fabsf(__fsub_rn(hm2, 1.0f))));
    const float X = srgb_of(__fadd_rn(x, m), powtab);
    const float HI = cfg.srgb_hi[vi], LO = cfg.srgb_lo[vi];
    if      (0.0f <= h && h < 1.0f) { out[0] = HI; out[1] = X;  out[2] = LO; }
    else if (1.0f <= h && h < 2.0f) { out[0] = X;  out[1] = HI; out[2] = LO; }
    else if (2.0f <= h && h < 3.0f) { out[0] = LO; out[1] = HI; out[2] = X;  }
    else if (3.0f <= h && h < 4.0f) { out[0] = LO; out[1] = X;  out[2] = HI; }
    else if (4.0f <= h && h < 5.0f) { out[0] = X;  out[1] = LO; out[2] = HI; }
    else                            { out[0] = HI; out[1] = LO; out[2] = X;  }
}

// world position of the vertex on the directed edge `e` of cell (x,y,z): chunk.rs:178-213,224-229.
// dens_at(ax, ay, az) -> density.  Returns corner_b's cube-local index (chunk.rs:219 needs it).
template <class DensAt>
__device__ __forceinline__ int edge_position(const DevCfg& cfg, DensAt dens_at, int x, int y, int z, int e,
                                             int offx, int offy, int offz, float p[3]) {
    // EDGE_VERTEX_INDICES (marching_table.rs:1-14) as nibble-packed immediates: a per-lane index into constant
    // memory would replay once per distinct edge in the warp
    const int ca = (int)((0x321076543210ull >> (4 * e)) & 15ull), cb = (int)((0x765447650321ull >> (4 * e)) & 15ull);
    int ax, ay, az, bx, by, bz;
    corner_off(ca, ax, ay, az); corner_off(cb, bx, by, bz);
    ax += x; ay += y; az += z; bx += x; by += y; bz += z;
    const float iso_a = dens_at(ax, ay, az), iso_b = dens_at(bx, by, bz);
    const float t = __fdiv_rn(__fsub_rn(cfg.iso_level, iso_a), __fsub_rn(iso_b, iso_a));
    const float sax = __fmul_rn((float)ax, cfg.size_scale), say = __fmul_rn((float)ay, cfg.size_scale), saz = __fmul_rn((float)az, cfg.size_scale);
    const float sbx = __fmul_rn((float)bx, cfg.size_scale), sby = __fmul_rn((float)by, cfg.size_scale), sbz = __fmul_rn((float)bz, cfg.size_scale);
    const float mx = __fadd_rn(sax, __fmul_rn(t, __fsub_rn(sbx, sax)));
    const float my = __fadd_rn(say, __fmul_rn(t, __fsub_rn(sby, say)));
    const float mz = __fadd_rn(saz, __fmul_rn(t, __fsub_rn(sbz, saz)));
    p[0] = __fadd_rn(mx, (float)offx); p[1] = __fadd_rn(my, (float)offy); p[2] = __fadd_rn(mz, (float)offz);
    return cb;
}

// vertex (position + colour) of the directed edge `e` of cell (x,y,z): chunk.rs:178-231
template <class DensAt>
__device__ __forceinline__ void make_vertex_from(const DevCfg& cfg, DensAt dens_at, int x, int y, int z, int e,
                                                 int offx, int offy, int offz, float v[6], const float* __restrict__ powtab = nullptr) {
    const int cb = edge_position(cfg, dens_at, x, y, z, e, offx, offy, offz, v);
    vertex_color(cfg, v[2], cb % 3, v + 3, powtab);
}

// parity tap: the product's vertex colour (table-driven pow included) on arbitrary (world z, value level) pairs
__global__ void __launch_bounds__(256) k_vertex_colors(const __grid_constant__ DevCfg cfg, const McTables* __restrict__ mc,
                                                       const float* __restrict__ world_z, const uint32_t* __restrict__ level,
                                                       uint32_t n, float* __restrict__ rgb) {
    __shared__ float s_powtab[48];
    if (threadIdx.x < 48) s_powtab[threadIdx.x] = mc->powtab[threadIdx.x];
    __syncthreads();
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        float c[3];
        vertex_color(cfg, world_z[t], (int)(level[t] % 3u), c, s_powtab);
        rgb[3 * t] = c[0]; rgb[3 * t + 1] = c[1]; rgb[3 * t + 2] = c[2];
    }
}

// util::Tri::new (util.rs:12-21) with cgmath's cross / magnitude / div and safe_normalize (util.rs:61-64); f32, unfused
__device__ __forceinline__ void tri_normal(const float a[3], const float b[3], const float c[3], float n[3]) {
    const float e1x = __fsub_rn(b[0], a[0]), e1y = __fsub_rn(b[1], a[1]), e1z = __fsub_rn(b[2], a[2]);
    const float e2x = __fsub_rn(c[0], a[0]), e2y = __fsub_rn(c[1], a[1]), e2z = __fsub_rn(c[2], a[2]);
    const float nx = __fsub_rn(__fmul_rn(e1y, e2z), __fmul_rn(e1z, e2y));
    const float ny = __fsub_rn(__fmul_rn(e1z, e2x), __fmul_rn(e1x, e2z));
    const float nz = __fsub_rn(__fmul_rn(e1x, e2y), __fmul_rn(e1y, e2x));
    const float mag = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz)));
    if (mag == 0.0f) { n[0] = nx; n[1] = ny; n[2] = nz; }
    else { n[0] = __fdiv_rn(nx, mag); n[1] = __fdiv_rn(ny, mag); n[2] = __fdiv_rn(nz, mag); }
}

__device__ __forceinline__ void make_vertex(const DevCfg& cfg, const float* s_dens, int x, int y, int z, int e,
                                            int offx, int offy, int offz, float v[6], const float* __restrict__ powtab = nullptr) {
    const int L = cfg.L;
    make_vertex_from(cfg, [=](int ax, int ay, int az) { return s_dens[(ax * L + ay) * L + az]; }, x, y, z, e, offx, offy, offz, v, powtab);
}

// owner (first cell in scan order holding the same ORDERED corner pair), SURVEY App. B.4
__device__ __forceinline__ void owner_of(int e, int x, int y, int z, int& ox, int& oy, int& oz, int& oe) {
    ox = x; oy = y; oz = z; oe = e;
    if (e < 4) { if (y > 0) { oy = y - 1; oe = e + 4; } }
    else if (e == 9)  { if (z > 0) { oz = z - 1; oe = 10; } }
    else if (e == 11) { if (x > 0) { ox = x - 1; oe = 10; } }
    else if (e == 8) {
        if (x > 0 && z > 0) { ox = x - 1; oz = z - 1; oe = 10; }
        else if (x > 0)     { ox = x - 1; oe = 9; }
        else if (z > 0)     { oz = z - 1; oe = 11; }
    }
}

// ---------------------------------------------------------------------------------------
// K2..K4 on one chunk whose densities and column sign masks sit in shared memory.
//   prepare  B  per cell column (x,y): "natural" 8-bit corner pattern from the 4 column masks ->
//               one LUT load gives (case, index count, crossed-edge mask); block scan in scan order
//            C  surface cells only: per-cell vertex base / index base + compact surface-cell list
//   write    D1 one thread per surface cell walks its row once: owned edges, in first-appearance
//               order, get vertex ids vbase+0,1,..; ids go into vid[(corner_a, direction)] (the
//               ORDERED lattice pair, i.e. the reference's dedup key chunk.rs:233) and the compact
//               vertex list
//            D2 one thread per vertex: edge lerp + colour, consecutive vertices by consecutive threads
//            E  one thread per surface cell: index = vid[(corner_a, direction) of the slot's edge]
// ---------------------------------------------------------------------------------------
#define UW_SMALL_MAX_CELLS ((UW_SMALL_MAX_L - 1) * (UW_SMALL_MAX_L - 1) * (UW_SMALL_MAX_L - 1))
#ifndef UW_VLIST_CAP
#define UW_VLIST_CAP 2816    // (3072 before the lattice table moved out of the region K4's vertex-id table aliases)
#endif
#define UW_EDGE_KINDS 5      // +x, -x, +y, +z, -z  (-y never occurs: edges 8..11 all run +y)

struct EmitSmem {
    float* dens; uint32_t* bits; uint32_t* mask; uint16_t* vbase; uint16_t* ibase; uint16_t* alist;
    uint16_t* vlist; uint16_t* vid; uint8_t* cs; uint32_t* lut; uint16_t* eoff;
    const float* powtab;            // McTables::powtab in shared memory
    const uint64_t* rows;           // McTables::rows in shared memory (the L1 left beside 4 x 55 KB of shared memory does not keep it)
    // Output staging (see emit_verts / emit_indices): both alias tables that are dead by then.
    float* vstage;                  // = vbase region (dead after D1): UW_VSTAGE_BYTES per warp
    void* istage;                   // = vlist region (dead after D2): UW_VLIST_CAP * 2 bytes
};



// per edge: (corner_a lattice offset) * 5 + direction kind, for lattice size L
__device__ __forceinline__ void fill_edge_offsets(uint16_t* eoff, int L) {
    const int e = threadIdx.x;
    if (e < 12) {
        int ax, ay, az, bx, by, bz;
        corner_off(c_edge_a[e], ax, ay, az); corner_off(c_edge_b[e], bx, by, bz);
        const int kind = bx > ax ? 0 : bx < ax ? 1 : by > ay ? 2 : bz > az ? 3 : 4;
        eoff[e] = (uint16_t)(((ax * L + ay) * L + az) * UW_EDGE_KINDS + kind);
    }
}

__host__ __device__ inline size_t emit_smem_bytes(const DevCfg& cfg) {
    const size_t cells = (size_t)cfg.S * cfg.S * cfg.S, cells2 = (cells + 7) & ~(size_t)7;
    size_t b = (size_t)cfg.dens_stride * 4;
    b += ((size_t)(cfg.L3 + 31) / 32 + 2) * 4;
    b += ((size_t)cfg.L2 + 3) / 4 * 16;
    b += cells2 * 2 * 2;
    b += (cells2 * 2 > 8 * UW_VSTAGE_BYTES ? cells2 * 2 : 8 * UW_VSTAGE_BYTES);      // vbase, reused as the vertex staging of 8 warps
    b += UW_VLIST_CAP * 2;
    b += (((size_t)cfg.L3 * UW_EDGE_KINDS + 7) & ~(size_t)7) * 2;
    b += 256 * 4 + 16 * 2;
    b += 256 * 8 + 48 * 4;
    b += (cells + 15) & ~(size_t)15;
    return b + 16;                                       // alignment slack of the staging regions
}

__device__ __forceinline__ EmitSmem emit_smem_carve(const DevCfg& cfg, unsigned char* base) {
    // offsets are kept as integers relative to `base` (16-byte aligned dynamic shared memory): rounding a POINTER through
    // uintptr_t would hide the address space from the compiler and turn every access behind it into a generic load
    const size_t cells = (size_t)cfg.S * cfg.S * cfg.S, cells2 = (cells + 7) & ~(size_t)7;
    EmitSmem s;
    size_t o = 0;
    s.dens = (float*)(base + o);          o += (size_t)cfg.dens_stride * 4;       // multiple of 16 bytes
    s.rows = (const uint64_t*)(base + o); o += 256 * 8;
    s.powtab = (const float*)(base + o);  o += 48 * 4;
    s.bits = (uint32_t*)(base + o);       o += ((size_t)(cfg.L3 + 31) / 32 + 2) * 4;
    s.mask = (uint32_t*)(base + o);       o += ((size_t)cfg.L2 + 3) / 4 * 16;
    o = (o + 15) & ~(size_t)15;                                                    // the staging regions hold 16-byte vectors
    s.vbase = (uint16_t*)(base + o);      s.vstage = (float*)(base + o);
    o += (cells2 * 2 > 8 * UW_VSTAGE_BYTES ? cells2 * 2 : 8 * UW_VSTAGE_BYTES);
    s.ibase = (uint16_t*)(base + o);      o += cells2 * 2;
    s.alist = (uint16_t*)(base + o);      o += cells2 * 2;
    s.vlist = (uint16_t*)(base + o);      s.istage = (void*)(base + o);    o += UW_VLIST_CAP * 2;
    s.vid = (uint16_t*)(base + o);        o += (((size_t)cfg.L3 * UW_EDGE_KINDS + 7) & ~(size_t)7) * 2;
    s.lut = (uint32_t*)(base + o);        o += 256 * 4;
    s.eoff = (uint16_t*)(base + o);       o += 16 * 2;
    s.cs = (uint8_t*)(base + o);
    return s;
}

struct ChunkShape { uint32_t n_vert, n_ind, n_act; };

// tri_cell (nullable): global u16[S^3 + 1], first triangle of every cell (chunk-local, scan order)
template <int ST>
__device__ __forceinline__ ChunkShape emit_prepare(const DevCfg& cfg, const EmitSmem& s, uint32_t* s_w,
                                                   uint16_t* __restrict__ tri_cell = nullptr,
                                                   unsigned long long* alloc_ctr = nullptr, unsigned long long* packed = nullptr,
                                                   int index_pad_mask = 7 /* 16 / sizeof(IndexT) - 1 */) {
    const int tid = threadIdx.x, NT = blockDim.x;
    const int S = ST > 0 ? ST : cfg.S, L = S + 1, ncol = S * S;

    // ---- B -----------------------------------------------------------------------------------------
    uint32_t nva = 0, ni = 0, smask = 0, own0 = 0, ownn = 0, q0 = 0, q1 = 0;
    const int col = tid;                                 // one column per thread (NT >= S*S on this path)
    const int x = col / S, y = col - x * S;
    if (col < ncol) {
        q0 = s.mask[x * L + y] | (s.mask[(x + 1) * L + y] << 16);
        q1 = s.mask[x * L + y + 1] | (s.mask[(x + 1) * L + y + 1] << 16);
        ownn = 0x4F0u | (y == 0 ? 0x00Fu : 0u) | (x == 0 ? 0x800u : 0u);      // SURVEY App. B.4 ownership, z > 0
        own0 = ownn | 0x200u | (x == 0 ? 0x100u : 0u);                          // z == 0 also owns 9 (and 8 if x == 0)
        // surface cells of the column straight from the four sign masks: cell z is mixed iff the columns disagree at
        // bit z or z + 1, or column 00 changes sign between them -- only those cells (2.2 of 12 on average) are
        // looked up.  (The case bytes are only ever read at surface cells: phase C files them.)
        const uint32_t m00 = q0 & 0xFFFFu, dis = (m00 ^ (q0 >> 16)) | (m00 ^ (q1 & 0xFFFFu)) | (m00 ^ (q1 >> 16));
        smask = (dis | (dis >> 1) | (m00 ^ (m00 >> 1))) & ((1u << S) - 1u);
        for (uint32_t rest = smask; rest; rest &= rest - 1u) {
            const int z = __ffs(rest) - 1;
            const uint32_t t = s.lut[natural_of(q0, q1, z)];                // case | ninds << 8 | crossed << 12
            ni += (t >> 8) & 15u;
            nva += __popc((t >> 12) & (z == 0 ? own0 : ownn)) + 0x10000u;
        }
    }
    uint32_t eva, ei, tva, ti;
    block_scan2(nva, ni, eva, ei, tva, ti, s_w);
    // completion-order packing: claim this chunk's range of the arenas now (one 64-bit atomic); the
    // round trip hides under phases C and D1
    if (alloc_ctr && tid == 0 && ti > 0)
        *packed = atomicAdd(alloc_ctr, ((unsigned long long)pad_verts(tva & 0xFFFFu) << 32) | (ti + (uint32_t)index_pad_mask & ~(uint32_t)index_pad_mask));

    if (tri_cell && col < ncol) {        // per-cell triangle offsets (the reference's per-cell Vec<Tri>, chunk.rs:167-174)
        uint32_t rt = ei;
        for (int z = 0; z < S; ++z) {
            tri_cell[col * S + z] = (uint16_t)(rt / 3u);
            if ((smask >> z) & 1u) rt += (s.lut[natural_of(q0, q1, z)] >> 8) & 15u;
        }
        if (col == ncol - 1) tri_cell[ncol * S] = (uint16_t)(ti / 3u);
    }
    // ---- C: surface cells only ------------------------------------------------------------------------
    uint32_t rv = eva & 0xFFFFu, ra = eva >> 16, ri = ei;
    while (smask) {
        const int z = __ffs(smask) - 1;
        smask &= smask - 1;
        const int cell = col * S + z;
        const uint32_t t = s.lut[natural_of(q0, q1, z)];
        s.vbase[ra] = (uint16_t)rv; s.ibase[ra] = (uint16_t)ri;        // indexed by surface-cell rank, like alist
        s.alist[ra++] = (uint16_t)cell;
        s.cs[cell] = (uint8_t)t;                                       // case index, chunk.rs:155-162
        ri += (t >> 8) & 15u;
        rv += __popc((t >> 12) & (z == 0 ? own0 : ownn));
    }
    __syncthreads();
    ChunkShape sh;
    sh.n_vert = tva & 0xFFFFu; sh.n_ind = ti; sh.n_act = tva >> 16;
    return sh;
}

// D1 for the vertex tile [v0, v0 + UW_VLIST_CAP): vertex ids (v0 == 0 only) + compact vertex list.
// Caller must barrier before emit_verts / emit_indices.
template <int ST>
__device__ __forceinline__ void emit_fill(const DevCfg& cfg, const McTables* __restrict__ mc, const EmitSmem& s,
                                          const ChunkShape sh, uint32_t v0) {
    const int tid = threadIdx.x, NT = blockDim.x;
    const int S = ST > 0 ? ST : cfg.S, L = S + 1;
    for (uint32_t a = tid; a < sh.n_act; a += NT) {
        const int cell = s.alist[a];
        const int x = cell / (S * S), r = cell - x * S * S, y = r / S, z = r - y * S;
        const uint32_t own = 0x4F0u | (y == 0 ? 0x00Fu : 0u) | (x == 0 ? 0x800u : 0u)
                           | (z == 0 ? (0x200u | (x == 0 ? 0x100u : 0u)) : 0u);
        const uint64_t row = s.rows[s.cs[cell]];
        const uint32_t rlo = (uint32_t)row, rhi = (uint32_t)(row >> 32);
        const int lbase = ((x * L + y) * L + z) * UW_EDGE_KINDS;
        uint32_t vnext = s.vbase[a], todo = own;        // owned edges not yet numbered
#pragma unroll
        for (int k = 0; k < 15; ++k) {
            const uint32_t e = ((k < 8 ? rlo : rhi) >> (4 * (k & 7))) & 15u;
            if (e == 15u) break;
            if ((todo >> e) & 1u) {
                todo &= ~(1u << e);
                if (v0 == 0) s.vid[lbase + s.eoff[e]] = (uint16_t)vnext;
                const uint32_t slot = vnext - v0;
                if (slot < UW_VLIST_CAP) s.vlist[slot] = (uint16_t)(cell | (e << 12));
                ++vnext;
            }
        }
    }
}

// D2: one thread per vertex of the tile.  The vertices leave through shared memory: a warp parks 16 of its 32
// vertices (384 B) in its staging slot and writes them as 24 consecutive 16-byte vectors, twice -- whole sectors,
// warp-contiguous, instead of 3 x 8-byte stores per lane at a 24-byte stride (which touch every 32-byte sector of
// the run three times with a third of it each: measured 331 GB/s over NVLink against ~2x that for full vectors).
// vout + v0 is 16-byte aligned (allocations start on even vertex counts); a tile's odd tail writes one pad vertex.
template <int ST, bool STAGED>
__device__ __forceinline__ void emit_verts(const DevCfg& cfg, const EmitSmem& s, const ChunkShape sh, uint32_t v0,
                                           int px, int py, int pz, uw_vert* __restrict__ vout) {
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int S = ST > 0 ? ST : cfg.S;
    const int offx = px * cfg.chunk_size, offy = py * cfg.chunk_size, offz = pz * cfg.chunk_size;
    const uint32_t cnt = min((uint32_t)UW_VLIST_CAP, sh.n_vert - v0);
    float* stage = s.vstage + warp * (UW_VSTAGE_BYTES / 4);
    for (uint32_t base = 0; base < cnt; base += NT) {
        const uint32_t w0 = base + warp * 32u;                       // first vertex of this warp's group
        if (w0 >= cnt) break;                                        // warp-uniform
        const uint32_t t = w0 + lane;
        float v[6];
        if (t < cnt) {
            const uint32_t ent = s.vlist[t];
            const int cell = ent & 0xFFF, e = ent >> 12;
            const int x = cell / (S * S), r = cell - x * S * S, y = r / S, z = r - y * S;
            make_vertex(cfg, s.dens, x, y, z, e, offx, offy, offz, v, s.powtab);
        }
        // Straight from registers (3 x 8-byte stores per lane) when the arena is this GPU's own HBM -- the L2 merges the
        // partial sectors and the staging costs more issue slots than it saves (measured, profiles/r02_ab_staged_stores.txt:
        // +7 % at 32 768 chunks) -- and for multi-tile chunks (> UW_VLIST_CAP vertices, worst-case fields only), whose
        // staging slots alias vbase, which the fill pass of the NEXT tile still reads.
        if (!STAGED || sh.n_vert > (uint32_t)UW_VLIST_CAP) {
            // t == cnt on an odd LAST tile: the allocation's pad vertex, zeroed so that the used extent of the arena is defined
            if (t < cnt || (t == cnt && (cnt & 1u) && v0 + cnt == sh.n_vert)) {
                if (t == cnt) { v[0] = v[1] = v[2] = v[3] = v[4] = v[5] = 0.f; }
                float2* dst = reinterpret_cast<float2*>(vout + v0 + t);
                dst[0] = make_float2(v[0], v[1]); dst[1] = make_float2(v[2], v[3]); dst[2] = make_float2(v[4], v[5]);
            }
            continue;
        }
        const uint32_t wcnt = min(32u, cnt - w0);
        float4* gdst = reinterpret_cast<float4*>(vout + v0 + w0);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            if (wcnt <= 16u * half) break;                           // warp-uniform
            const uint32_t hcnt = min(16u, wcnt - 16u * half);
            if ((lane >> 4) == half && t <= cnt) {                    // t == cnt: the pad vertex of an odd tail, zeroed
                if (t == cnt) { v[0] = v[1] = v[2] = v[3] = v[4] = v[5] = 0.f; }
                float2* d = reinterpret_cast<float2*>(stage + (lane & 15) * 6);
                d[0] = make_float2(v[0], v[1]); d[1] = make_float2(v[2], v[3]); d[2] = make_float2(v[4], v[5]);
            }
            __syncwarp();
            const uint32_t n16 = (pad_verts(hcnt) * 3u) >> 1;        // 24 B per vertex = 1.5 vectors
            if ((uint32_t)lane < n16) gdst[half * 24 + lane] = reinterpret_cast<const float4*>(stage)[lane];
            __syncwarp();
        }
    }
}

// E: indices; every slot is one table lookup keyed by the ordered lattice pair.  The chunk's index buffer is
// assembled in shared memory (s.istage aliases the vertex list, dead after D2) and leaves as whole 16-byte vectors in
// one linear sweep; chunks with more indices than the staging holds take several tiles.  iout is 16-byte aligned
// (index allocations are padded), the last vector may carry up to 7 pad entries inside the chunk's own allocation.
// ALL threads must call; contains block barriers; the caller has a barrier between D2 and this.
template <int ST, typename IndexT, bool STAGED>
__device__ __forceinline__ void emit_indices(const DevCfg& cfg, const McTables* __restrict__ mc, const EmitSmem& s,
                                             const ChunkShape sh, IndexT* __restrict__ iout) {
    const int tid = threadIdx.x, NT = blockDim.x;
    const int S = ST > 0 ? ST : cfg.S, L = S + 1;
    constexpr uint32_t TILE = UW_VLIST_CAP * 2 / sizeof(IndexT), PER16 = 16 / sizeof(IndexT);
    IndexT* stage = reinterpret_cast<IndexT*>(s.istage);
    if constexpr (!STAGED) {                                         // own HBM: per-cell runs straight from registers
        for (uint32_t a = tid; a < sh.n_act; a += NT) {
            const int cell = s.alist[a];
            const int x = cell / (S * S), r = cell - x * S * S, y = r / S, z = r - y * S;
            const uint64_t row = s.rows[s.cs[cell]];
            const uint32_t rlo = (uint32_t)row, rhi = (uint32_t)(row >> 32);
            const int lbase = ((x * L + y) * L + z) * UW_EDGE_KINDS;
            IndexT* dst = iout + s.ibase[a];
            const uint16_t* vid = s.vid + lbase;
#pragma unroll
            for (int k = 0; k < 15; k += 3) {                        // a triangle at a time: three independent lookups in flight
                const uint32_t e0 = ((k < 8 ? rlo : rhi) >> (4 * (k & 7))) & 15u;
                if (e0 == 15u) break;
                const uint32_t e1 = ((k + 1 < 8 ? rlo : rhi) >> (4 * ((k + 1) & 7))) & 15u;
                const uint32_t e2 = ((k + 2 < 8 ? rlo : rhi) >> (4 * ((k + 2) & 7))) & 15u;
                const uint32_t o0 = s.eoff[e0], o1 = s.eoff[e1], o2 = s.eoff[e2];
                const IndexT i0 = (IndexT)vid[o0], i1 = (IndexT)vid[o1], i2 = (IndexT)vid[o2];   // `ind as u16`, chunk.rs:243
                dst[k] = i0; dst[k + 1] = i1; dst[k + 2] = i2;
            }
        }
        if ((uint32_t)tid < PER16 && sh.n_ind + tid < pad_inds<IndexT>(sh.n_ind)) iout[sh.n_ind + tid] = (IndexT)0;   // pad entries, defined
        return;
    }
    for (uint32_t lo = 0; lo < sh.n_ind; lo += TILE) {
        const uint32_t hi = min(lo + TILE, sh.n_ind);
        if (lo) __syncthreads();                                     // previous tile fully copied out
        for (uint32_t a = tid; a < sh.n_act; a += NT) {
            const uint32_t ib = s.ibase[a];
            if (ib >= hi || ib + 15u <= lo) continue;                // the cell's run (<= 15 indices) misses this tile
            const int cell = s.alist[a];
            const int x = cell / (S * S), r = cell - x * S * S, y = r / S, z = r - y * S;
            const uint64_t row = s.rows[s.cs[cell]];
            const uint32_t rlo = (uint32_t)row, rhi = (uint32_t)(row >> 32);
            const int lbase = ((x * L + y) * L + z) * UW_EDGE_KINDS;
            const uint16_t* vid = s.vid + lbase;
            const bool inside = ib >= lo && ib + 15u <= hi;          // common case: no per-entry range test
#pragma unroll
            for (int k = 0; k < 15; k += 3) {                        // a triangle at a time: three independent lookups in flight
                const uint32_t e0 = ((k < 8 ? rlo : rhi) >> (4 * (k & 7))) & 15u;
                if (e0 == 15u) break;
                const uint32_t e1 = ((k + 1 < 8 ? rlo : rhi) >> (4 * ((k + 1) & 7))) & 15u;
                const uint32_t e2 = ((k + 2 < 8 ? rlo : rhi) >> (4 * ((k + 2) & 7))) & 15u;
                const uint32_t o0 = s.eoff[e0], o1 = s.eoff[e1], o2 = s.eoff[e2];
                const IndexT i0 = (IndexT)vid[o0], i1 = (IndexT)vid[o1], i2 = (IndexT)vid[o2];   // `ind as u16`, chunk.rs:243
                const uint32_t q = ib + k - lo;                      // may wrap below lo: the range tests catch it
                if (inside) { stage[q] = i0; stage[q + 1] = i1; stage[q + 2] = i2; }
                else {
                    if (ib + k >= lo && ib + k < hi) stage[q] = i0;
                    if (ib + k + 1 >= lo && ib + k + 1 < hi) stage[q + 1] = i1;
                    if (ib + k + 2 >= lo && ib + k + 2 < hi) stage[q + 2] = i2;
                }
            }
        }
        const uint32_t n16 = (hi - lo + PER16 - 1) / PER16;
        if ((uint32_t)tid < PER16 && hi - lo + tid < n16 * PER16) stage[hi - lo + tid] = (IndexT)0;   // pad entries of the last vector
        __syncthreads();
        uint4* gdst = reinterpret_cast<uint4*>(iout + lo);
        const uint4* src = reinterpret_cast<const uint4*>(stage);
        for (uint32_t t = tid; t < n16; t += NT) gdst[t] = src[t];
    }
}

// F (UW_FLAG_TRIS): the reference's collision triangles, one thread per TRIANGLE (a thread per surface cell would
// run every warp for its busiest cell's 5 triangles while the average cell has 1.5).  Triangle j of the chunk is
// (indices 3j..3j+2); its cell is found by a binary search over the surface cells' index bases, its corners are
// recomputed from the cell's own edges (same ordered corner pairs and densities as the vertex buffer ->
// bit-identical positions).
template <int ST>
__device__ __forceinline__ void emit_tris(const DevCfg& cfg, const McTables* __restrict__ mc, const EmitSmem& s,
                                          const ChunkShape sh, int px, int py, int pz, uw_tri* __restrict__ tout) {
    const int tid = threadIdx.x, NT = blockDim.x;
    const int S = ST > 0 ? ST : cfg.S, L = S + 1;
    const int offx = px * cfg.chunk_size, offy = py * cfg.chunk_size, offz = pz * cfg.chunk_size;
    const float* dens = s.dens;
    auto dens_at = [=](int ax, int ay, int az) { return dens[(ax * L + ay) * L + az]; };
    const uint32_t ntri = sh.n_ind / 3u;
    for (uint32_t j = tid; j < ntri; j += NT) {
        uint32_t lo = 0, hi = sh.n_act;                    // largest a with ibase[a] <= 3j (ibase is ascending; every surface cell has >= 1 triangle)
        while (hi - lo > 1u) {
            const uint32_t mid = (lo + hi) >> 1;
            if ((uint32_t)s.ibase[mid] <= 3u * j) lo = mid; else hi = mid;
        }
        const int cell = s.alist[lo];
        const int t = (int)((3u * j - s.ibase[lo]) / 3u);
        const int x = cell / (S * S), r = cell - x * S * S, y = r / S, z = r - y * S;
        const uint64_t row = s.rows[s.cs[cell]];
        const int e0 = (int)((row >> (12 * t)) & 0xFull), e1 = (int)((row >> (12 * t + 4)) & 0xFull), e2 = (int)((row >> (12 * t + 8)) & 0xFull);
        float v[12];
        edge_position(cfg, dens_at, x, y, z, e0, offx, offy, offz, v);
        edge_position(cfg, dens_at, x, y, z, e1, offx, offy, offz, v + 3);
        edge_position(cfg, dens_at, x, y, z, e2, offx, offy, offz, v + 6);
        tri_normal(v, v + 3, v + 6, v + 9);
        float4* d4 = reinterpret_cast<float4*>(tout + j);                   // uw_tri is 48 B; the array base is 16-B aligned
        d4[0] = make_float4(v[0], v[1], v[2], v[3]);
        d4[1] = make_float4(v[4], v[5], v[6], v[7]);
        d4[2] = make_float4(v[8], v[9], v[10], v[11]);
    }
}

// everything after the first fill + barrier.  ALL threads must call (block barriers inside).
template <int ST, typename IndexT, bool STAGED>
__device__ __forceinline__ void emit_rest(const DevCfg& cfg, const McTables* __restrict__ mc, const EmitSmem& s,
                                          const ChunkShape sh, int px, int py, int pz,
                                          uw_vert* __restrict__ vout, IndexT* __restrict__ iout) {
    emit_verts<ST, STAGED>(cfg, s, sh, 0, px, py, pz, vout);
    for (uint32_t v0 = UW_VLIST_CAP; v0 < sh.n_vert; v0 += UW_VLIST_CAP) {
        __syncthreads();
        emit_fill<ST>(cfg, mc, s, sh, v0);
        __syncthreads();
        emit_verts<ST, STAGED>(cfg, s, sh, v0, px, py, pz, vout);
    }
    if (STAGED) __syncthreads();                       // the vertex list is dead: its memory becomes the index staging
    emit_indices<ST, IndexT, STAGED>(cfg, mc, s, sh, iout);
}

template <int ST, typename IndexT>
__global__ void __launch_bounds__(256) k_emit_small(const __grid_constant__ DevCfg cfg,
                                                    const McTables* __restrict__ mc,
                                                    const float* __restrict__ dens,
                                                    const uw_chunk_desc* __restrict__ descs,
                                                    const uint32_t* __restrict__ active,
                                                    const BatchTotals* __restrict__ totals,
                                                    uw_vert* __restrict__ verts, IndexT* __restrict__ inds,
                                                    uw_tri* __restrict__ tris, uint16_t* __restrict__ tri_cell) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const EmitSmem s = emit_smem_carve(cfg, smem_raw);
    __shared__ int s_red[2];
    __shared__ uint32_t s_w[64];
    const int tid = threadIdx.x, NT = blockDim.x;
    const int L = (ST > 0 ? ST : cfg.S) + 1;
    const uint32_t n_active = totals->n_active;
    if (totals->overflow) return;                       // host grows the arenas and relaunches
    for (int t = tid; t < 256; t += NT) { s.lut[t] = mc->lut[t]; const_cast<uint64_t*>(s.rows)[t] = mc->rows[t]; }
    if (tid < 48) const_cast<float*>(s.powtab)[tid] = mc->powtab[tid];
    fill_edge_offsets(s.eoff, L);

    for (uint32_t a = blockIdx.x; a < n_active; a += gridDim.x) {
        const uint32_t chunk = active[a];
        const uw_chunk_desc d = descs[chunk];
        if constexpr (ST > 0) load_signs_spec<ST, 256, true>(cfg.iso_level, dens + (size_t)chunk * cfg.dens_stride, s.dens, s.bits, s_w);
        else                  load_signs<true>(cfg, dens + (size_t)chunk * cfg.dens_stride, s.dens, s.bits, s_red);
        for (int c = tid; c < L * L; c += NT) s.mask[c] = col_mask(s.bits, c, L);
        __syncthreads();
        const int ncell = (L - 1) * (L - 1) * (L - 1);
        const ChunkShape sh = emit_prepare<ST>(cfg, s, s_w, tri_cell ? tri_cell + (size_t)chunk * (ncell + 1) : nullptr);
        emit_fill<ST>(cfg, mc, s, sh, 0);
        __syncthreads();
        emit_rest<ST, IndexT, false>(cfg, mc, s, sh, d.pos[0], d.pos[1], d.pos[2], verts + d.vert_offset, inds + d.index_offset);
        if (tris) emit_tris<ST>(cfg, mc, s, sh, d.pos[0], d.pos[1], d.pos[2], tris + d.index_offset / 3u);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// K1 for large chunks (e.g. 65^3 samples): the same tensor-product factorisation as noise_chunk_spec,
// tiled over the chunk's (i, j) columns.  A unit is 256 consecutive columns of one chunk (they touch at
// most PLMAX x-planes, so stage X is tiny); every thread walks its whole z column with the per-sample
// weights as compile-time immediates, parks the results in a per-warp shared-memory tile and the warp
// flushes the tile with contiguous stores (a column is 260 contiguous bytes in HBM).
// ---------------------------------------------------------------------------------------
template <int ST, int NOCT>
struct BigNoiseSmem {
    using D = SpecDims<ST, NOCT>;
    static constexpr int NT = 256, NW = NT / 32;
    static constexpr int PLMAX = (NT + D::L - 2) / D::L + 1;           // x-planes 256 consecutive columns can touch
    static constexpr int G2SUM = D::x_base(NOCT) / D::L;                 // sum over octaves of G^2
    static constexpr int HALF = (D::L + 1) / 2;                         // z samples per tile flush
    float4 lat[D::LAT];
    float4 X[PLMAX * G2SUM];
    float4 axis[NOCT][D::L + 1];
    float4 grad[16];
    float terr[D::L + 3];
    uint8_t perm[256];
    float tile[NW][32][HALF + 1 - (HALF & 1)];                          // odd row stride: conflict-free column writes
};

template <int ST, int NOCT>
__global__ void __launch_bounds__(256, 3)
k_noise_big(const __grid_constant__ DevCfg cfg, const float4* __restrict__ g_axis /*[NOCT][L]: d, d-1, fade, -*/,
            const uint8_t* __restrict__ g_perm, const int32_t* __restrict__ pos, uint32_t n,
            float* __restrict__ dens, unsigned long long* __restrict__ guard_count) {
    using D = SpecDims<ST, NOCT>;
    using SM = BigNoiseSmem<ST, NOCT>;
    constexpr int L = D::L, L2 = L * L, NT = SM::NT, HALF = SM::HALF, TS = HALF + 1 - (HALF & 1);
    constexpr int UPC = (L2 + NT - 1) / NT;                             // units per chunk
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SM& sm = *reinterpret_cast<SM*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int t = tid; t < 256; t += NT) sm.perm[t] = g_perm[t];
    if (tid < 16) sm.grad[tid] = make_float4(c_grad_vec[tid][0], c_grad_vec[tid][1], c_grad_vec[tid][2], 0.f);
    for (int t = tid; t < NOCT * L; t += NT) sm.axis[t / L][t % L] = g_axis[t];
    // PRUNE: columns with j = S read one X row past their last kept row (weight ~ 0): that row must hold finite
    // values from the first unit on -- X only ever holds finite values afterwards, and `axis` follows it
    for (int t = tid; t < SM::PLMAX * SM::G2SUM; t += NT) sm.X[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    constexpr float inv_max = 1.0f / (2.0f - 1.0f / (float)(1 << (NOCT - 1)));
    const unsigned long long n_units = (unsigned long long)n * UPC;
    for (unsigned long long u = blockIdx.x; u < n_units; u += gridDim.x) {
        const uint32_t chunk = (uint32_t)(u / UPC);
        const int col0 = (int)(u - (unsigned long long)chunk * UPC) * NT;
        const int ncols = min(NT, L2 - col0);
        const int i_min = col0 / L, npl = (col0 + ncols - 1) / L - i_min + 1;
        const int px = pos[3 * chunk], py = pos[3 * chunk + 1], pz = pos[3 * chunk + 2];

        // ---- stage H ----------------------------------------------------------------------------
        int htid = tid;
        asm volatile("" : "+r"(htid));       // lattice indices recomputed per unit: hoisted, they are spilled (cf. noise_chunk_spec)
#pragma unroll
        for (int o = 0; o < NOCT; ++o) {
            const int G = D::G(o), F = 1 << o, base = D::lat_base(o);
            for (int t = htid; t < G * G * G; t += NT) {
                const int cx = t / (G * G), r = t - cx * G * G, cy = r / G, cz = r - cy * G;
                const uint32_t h = sm.perm[sm.perm[sm.perm[(F * px + cx) & 255] ^ ((F * py + cy) & 255)] ^ ((F * pz + cz) & 255)];
                sm.lat[base + t] = sm.grad[h & 15u];
            }
        }
        if (tid < L) {
            float adj, fm;
            terrace_terms(cfg, tid, pz, adj, fm);
            sm.terr[tid] = __fsub_rn(__fsub_rn(adj, fm), 1.0f);       // minus 1: octave values are carried shifted, see noise_chunk_spec
        }
        __syncthreads();
        // ---- stage X (only the planes this unit touches) ------------------------------------------
        {
            int xb = 0;
#pragma unroll
            for (int o = 0; o < NOCT; ++o) {
                const int G = D::G(o), lb = D::lat_base(o);
                const float sc = 1.1547005383792515f * 0.5f;         // normalised to the clamp range, see noise_chunk_spec
                for (int t = tid; t < npl * G * G; t += NT) {
                    const int pl = t / (G * G), r = t - pl * G * G;
                    const int i = i_min + pl, c = (i << o) / ST;
                    const int c1 = D::PRUNE ? min(c + 1, G - 1) : c + 1;  // see SpecDims::prune
                    const float4 g0 = sm.lat[lb + c * G * G + r], g1 = sm.lat[lb + c1 * G * G + r];
                    const float4 ax = sm.axis[o][i];
                    const float q0 = g0.x * ax.x, q1 = g1.x * ax.y;
                    float4 e;
                    e.x = fmaf(fmaf(ax.z, q1 - q0, q0), sc, 0.5f);
                    e.y = fmaf(ax.z, g1.y - g0.y, g0.y) * sc;
                    e.z = fmaf(ax.z, g1.z - g0.z, g0.z) * sc;
                    e.w = 0.f;
                    sm.X[xb + t] = e;
                }
                xb += SM::PLMAX * G * G;
            }
        }
        __syncthreads();
        // ---- stage YZ ---------------------------------------------------------------------------------
        const int col = col0 + tid;
        const bool live = tid < ncols;
        const int i = live ? col / L : i_min, j = live ? col - i * L : 0;
        float R0[NOCT], S0[NOCT], R1[NOCT], S1[NOCT], Cc[NOCT], Dd[NOCT];
        const float4* xrow[NOCT];
        float dy[NOCT], dy1[NOCT], wy[NOCT];
        {
            int xb = 0;
#pragma unroll
            for (int o = 0; o < NOCT; ++o) {
                const int G = D::G(o);
                xrow[o] = sm.X + xb + ((i - i_min) * G + ((j << o) / ST)) * G;
                const float4 ay = sm.axis[o][j];
                dy[o] = ay.x; dy1[o] = ay.y; wy[o] = ay.z;
                xb += SM::PLMAX * G * G;
            }
        }
        auto ystage = [&](int o, int cz, float& R, float& Sz) {
            const int G = D::G(o);
            const float4 E0 = xrow[o][cz];
            const float4 E1 = xrow[o][G + cz];
            const float A0 = fmaf(E0.y, dy[o], E0.x);
            const float A1 = fmaf(E1.y, dy1[o], E1.x);
            R = fmaf(wy[o], A1 - A0, A0);
            Sz = fmaf(wy[o], E1.z - E0.z, E0.z);
        };
        const float isl = cfg.iso_level, eps = cfg.guard_eps;
        float* trow = &sm.tile[warp][lane][0];
        float* gout = dens + (size_t)chunk * cfg.dens_stride;
        const int wcol0 = col0 + warp * 32;
        unsigned long long near = 0;
        // flush the warp's tile: `cnt` z samples starting at kbase of up to 32 consecutive columns
        auto flush = [&](int kbase, int cnt) {
            while (near) {                                               // rare: exact f64 re-evaluation
                const int kk = __ffsll((long long)near) - 1;
                near &= near - 1;
                trow[kk] = x_iso_lattice(cfg, sm.perm, px, py, pz, i, j, kbase + kk);
                atomicAdd(guard_count, 1ull);
            }
            __syncwarp();
            for (int e = lane; e < 32 * cnt; e += 32) {
                const int c = e / cnt, kk = e - c * cnt;
                if (wcol0 + c < L2) gout[(size_t)(wcol0 + c) * L + kbase + kk] = sm.tile[warp][c][kk];
            }
            __syncwarp();
        };
#pragma unroll
        for (int k = 0; k < L; ++k) {
            float total = sm.terr[k];
#pragma unroll
            for (int o = 0; o < NOCT; ++o) {
                const int c = D::cell(o, k);
                const bool first = (k == 0), step = (k > 0) && (c != D::cell(o, k > 0 ? k - 1 : 0));
                // PRUNE: the last sample's weight on the far plane is ~ 1e-22 (or 0): v = a0, from the plane already held
                const bool top = D::PRUNE && k == L - 1;
                if (first) { ystage(o, c, R0[o], S0[o]); ystage(o, c + 1, R1[o], S1[o]); }
                else if (step) { R0[o] = R1[o]; S0[o] = S1[o]; if (!top) ystage(o, c + 1, R1[o], S1[o]); }
                if ((first || step) && !top) { Cc[o] = (R1[o] - R0[o]) - S1[o]; Dd[o] = S1[o] - S0[o]; }
                const float d = D::tab_d(o, k), w = D::tab_w(o, k);
                const float u = top ? __saturatef(fmaf(d, S0[o], R0[o]))
                                    : __saturatef(fmaf(w, fmaf(d, Dd[o], Cc[o]), fmaf(d, S0[o], R0[o])));   // FFMA.SAT = the clamp
                total = fmaf(u, 2.0f * inv_max / (float)(1 << o), total);
            }
            const float iso = total;
            const int kk = k < HALF ? k : k - HALF;
            trow[kk] = iso;
            if (live && fabsf(iso - isl) < eps) near |= 1ull << kk;
            if (k == HALF - 1) flush(0, HALF);
        }
        flush(HALF, L - HALF);
        __syncthreads();                                                 // lat / X are rewritten by the next unit
        (void)TS;
    }
}

// ---------------------------------------------------------------------------------------
// LARGE-CHUNK PATH (internal_size 16..64, e.g. BASELINE config 4: 64^3 cells, 65^3 samples = 1.1 MB of
// densities per chunk -- far beyond shared memory).  Densities are materialised in HBM by the noise
// kernel; one CTA then walks a chunk's x-slabs IN ORDER, holding two density planes (x, x+1), the sign
// bits of both, and the (case, vertex base) arrays of the previous and current slab -- everything the
// reference's scan-order numbering needs, since an edge's owner is never more than
// one slab back.  COUNT pass -> k_scan_chunks -> EMIT pass (same walk, now writing vertices/indices).
// ---------------------------------------------------------------------------------------
#define UW_BIG_VCAP 4096     // slab vertex-list tile (entries)
struct BigSmem {
    float* plane[3]; uint32_t* bits[3]; uint16_t* vb[2]; uint8_t* cs[2]; uint16_t* ib; uint16_t* alist; uint16_t* vlist;   // planes / sign bits: ring of 3
    uint32_t* lut; uint64_t* rows; uint16_t* before; uint16_t* crossed; uint8_t* nind;
};

__host__ __device__ inline size_t big_smem_bytes(const DevCfg& cfg) {
    const size_t L2p = ((size_t)cfg.L2 + 3) & ~(size_t)3, nw = (((size_t)cfg.L2 + 31) / 32 + 3) & ~(size_t)1;
    const size_t cells2 = ((size_t)cfg.S * cfg.S + 7) & ~(size_t)7;
    return 3 * L2p * 4 + 256 * 8 + 256 * 4 + 3 * nw * 4 + 256 * 12 * 2 + 256 * 2 + 256
         + 4 * cells2 * 2 + UW_BIG_VCAP * 2 + 2 * ((cells2 + 15) & ~(size_t)15);
}

__device__ __forceinline__ BigSmem big_smem_carve(const DevCfg& cfg, unsigned char* base) {
    const size_t L2p = ((size_t)cfg.L2 + 3) & ~(size_t)3, nw = (((size_t)cfg.L2 + 31) / 32 + 3) & ~(size_t)1;
    const size_t cells2 = ((size_t)cfg.S * cfg.S + 7) & ~(size_t)7;
    BigSmem s;
    s.plane[0] = (float*)base; base += L2p * 4;
    s.plane[1] = (float*)base; base += L2p * 4;
    s.plane[2] = (float*)base; base += L2p * 4;
    s.rows = (uint64_t*)base; base += 256 * 8;
    s.lut = (uint32_t*)base; base += 256 * 4;
    s.bits[0] = (uint32_t*)base; base += nw * 4;
    s.bits[1] = (uint32_t*)base; base += nw * 4;
    s.bits[2] = (uint32_t*)base; base += nw * 4;
    s.before = (uint16_t*)base; base += 256 * 12 * 2;
    s.crossed = (uint16_t*)base; base += 256 * 2;
    s.nind = (uint8_t*)base; base += 256;
    s.vb[0] = (uint16_t*)base; base += cells2 * 2;
    s.vb[1] = (uint16_t*)base; base += cells2 * 2;
    s.ib = (uint16_t*)base; base += cells2 * 2;
    s.alist = (uint16_t*)base; base += cells2 * 2;
    s.vlist = (uint16_t*)base; base += UW_BIG_VCAP * 2;
    s.cs[0] = (uint8_t*)base; base += (cells2 + 15) & ~(size_t)15;
    s.cs[1] = (uint8_t*)base;
    return s;
}

// two consecutive bits (b, b+1) of a flat bit array
__device__ __forceinline__ uint32_t two_bits(const uint32_t* bits, int b) {
    return __funnelshift_r(bits[b >> 5], bits[(b >> 5) + 1], b & 31) & 3u;
}

// bits b .. b+15 of a flat bit array
__device__ __forceinline__ uint32_t bits16(const uint32_t* bits, int b) {
    return __funnelshift_r(bits[b >> 5], bits[(b >> 5) + 1], b & 31) & 0xFFFFu;
}

#define UW_BIG_NT 512
#ifndef UW_BIG_COUNT_MINB
#define UW_BIG_COUNT_MINB 3
#endif
#define UW_BIG_NLD 9          // ceil(65 * 65 / 512): plane elements per thread

// COUNT pass: per-chunk vertex / index totals and the blank / no-surface vote.  Totals do not depend on
// the scan order, so no state is carried between slabs: two sign-bit planes in shared memory, per-thread
// counters, one block reduction per chunk.  The next plane's loads are issued before the current slab is
// classified so that their DRAM latency hides behind the bit work.
template <int ST /*compile-time internal_size, 0 = runtime*/>
__global__ void __launch_bounds__(UW_BIG_NT, UW_BIG_COUNT_MINB) k_count_big(const __grid_constant__ DevCfg cfg, const McTables* __restrict__ mc,
                                                         const float* __restrict__ dens, uint32_t n,
                                                         ChunkCounts* __restrict__ counts,
                                                         uint2* __restrict__ quarters /*[n][4]: (verts, inds) of each x-quarter*/) {
    __shared__ uint32_t s_bits[2][(65 * 65 + 31) / 32 + 2];
    __shared__ uint32_t s_lut[256];
    __shared__ uint32_t s_acc[4];
    __shared__ uint32_t s_q[4][2];
    const int tid = threadIdx.x, lane = tid & 31;
    const int S = ST > 0 ? ST : cfg.S, L = S + 1, L2 = L * L, ncell = S * S;
    const int CPT = (ncell + UW_BIG_NT - 1) / UW_BIG_NT;
    const bool row_runs = CPT <= 15 && S % CPT == 0;
    for (int t = tid; t < 256; t += UW_BIG_NT) s_lut[t] = mc->lut[t];

    for (uint32_t chunk = blockIdx.x; chunk < n; chunk += gridDim.x) {
        const float* D = dens + (size_t)chunk * cfg.dens_stride;
        if (tid < 4) s_acc[tid] = tid == 2 ? 1u : 0u;       // [0] verts, [1] inds, [2] all_gt, [3] any_lt
        if (tid < 8) s_q[tid >> 1][tid & 1] = 0u;
        const int qlen = (S & 3) == 0 ? S >> 2 : S;         // x-quarters (the emit pass may split a chunk there)
        float pre[UW_BIG_NLD];
        bool all_gt = true, any_lt = false;
        uint32_t nv = 0, ni = 0;
        auto issue = [&](int x) {
            const float* src = D + (size_t)x * L2;
#pragma unroll
            for (int q = 0; q < UW_BIG_NLD; ++q) { const int idx = q * UW_BIG_NT + tid; pre[q] = idx < L2 ? __ldg(src + idx) : 0.f; }
        };
        auto commit = [&](int x) {
            uint32_t* bt = s_bits[x & 1];
#pragma unroll
            for (int q = 0; q < UW_BIG_NLD; ++q) {
                const int base = q * UW_BIG_NT + tid - lane;
                if (base < L2) {                                          // warp-uniform
                    const bool ok = base + lane < L2;
                    const bool lt = ok && (pre[q] < cfg.iso_level);
                    const uint32_t word = __ballot_sync(0xFFFFFFFFu, lt);
                    if (lane == 0) bt[base >> 5] = word;
                    all_gt &= !ok || (pre[q] > cfg.iso_level);
                    any_lt |= lt;
                }
            }
            if (tid == 0) { bt[(L2 + 31) >> 5] = 0; bt[((L2 + 31) >> 5) + 1] = 0; }
        };
        issue(0); commit(0); issue(1);
        for (int cx = 0; cx < S; ++cx) {
            commit(cx + 1);
            if (cx + 2 <= S) issue(cx + 2);
            __syncthreads();
            const uint32_t* A = s_bits[cx & 1];
            const uint32_t* B = s_bits[(cx + 1) & 1];
            if (row_runs) {
                // a thread's CPT consecutive cells lie in one (y) row: one 4-column mask fetch per thread and slab,
                // and a run with no sign change (the common case) costs nothing more
                const int c0 = tid * CPT;
                if (c0 < ncell) {
                    const int y = c0 / S, z0 = c0 - y * S, b = y * L + z0;
                    const uint32_t keep = (2u << CPT) - 1u;                    // CPT + 1 samples
                    const uint32_t q0 = (bits16(A, b) & keep) | ((bits16(B, b) & keep) << 16);
                    const uint32_t q1 = (bits16(A, b + L) & keep) | ((bits16(B, b + L) & keep) << 16);
                    const uint32_t any = q0 | q1, all = q0 & q1;
                    if (any != 0u && (all & (all >> 16) & keep) != keep) {
                        for (int q = 0; q < CPT; ++q) {
                            const uint32_t t = s_lut[natural_of(q0, q1, q)];
                            if (t >> 8) { ni += (t >> 8) & 15u; nv += __popc((t >> 12) & own_mask_of(cx, y, z0 + q)); }
                        }
                    }
                }
            } else {
                for (int cell = tid; cell < ncell; cell += UW_BIG_NT) {
                    const int y = cell / S, z = cell - y * S;
                    const uint32_t nat = two_bits(A, y * L + z) | (two_bits(B, y * L + z) << 2)
                                       | (two_bits(A, (y + 1) * L + z) << 4) | (two_bits(B, (y + 1) * L + z) << 6);
                    const uint32_t t = s_lut[nat];
                    if (t >> 8) { ni += (t >> 8) & 15u; nv += __popc((t >> 12) & own_mask_of(cx, y, z)); }
                }
            }
            if ((cx + 1) % qlen == 0) {                     // quarter boundary: flush the per-thread counters
                nv = __reduce_add_sync(0xFFFFFFFFu, nv); ni = __reduce_add_sync(0xFFFFFFFFu, ni);
                if (lane == 0 && (nv | ni)) { atomicAdd(&s_q[cx / qlen][0], nv); atomicAdd(&s_q[cx / qlen][1], ni); }
                nv = 0; ni = 0;
            }
            __syncthreads();
        }
        const bool w_all = __all_sync(0xFFFFFFFFu, all_gt), w_any = __any_sync(0xFFFFFFFFu, any_lt);
        if (lane == 0) {
            if (!w_all) atomicAnd(&s_acc[2], 0u);
            if (w_any) atomicOr(&s_acc[3], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            ChunkCounts c;
            c.n_verts = s_q[0][0] + s_q[1][0] + s_q[2][0] + s_q[3][0]; c.n_inds = s_q[0][1] + s_q[1][1] + s_q[2][1] + s_q[3][1];
            c.flags = (s_acc[2] ? CF_ALL_GT : 0u) | (s_acc[3] ? CF_ANY_LT : 0u); c.pad = 0;
            counts[chunk] = c;
        }
        if (tid < 4) quarters[(size_t)chunk * 4 + tid] = make_uint2(s_q[tid][0], s_q[tid][1]);
        __syncthreads();
    }
}

// EMIT pass over the active chunks (see the section header).  Chunks are taken by ticket (their costs differ by
// the amount of surface they hold).  Per slab: classify + block scan -> per-cell vertex bases and the slab's
// surface-cell list; then one thread per SURFACE CELL writes its indices (owner lookups against the current /
// previous slab, every table in shared memory) and lists its owned edges; then one thread per VERTEX does the
// edge lerp + colour.  Slab-relative u16 bases keep the footprint at two CTAs per SM.
template <typename IndexT, int ST /*compile-time internal_size, 0 = runtime*/>
__global__ void __launch_bounds__(UW_BIG_NT, 2) k_emit_big(const __grid_constant__ DevCfg cfg, const McTables* __restrict__ mc,
                                                           const float* __restrict__ dens,
                                                           const uw_chunk_desc* __restrict__ descs, const uint32_t* __restrict__ active,
                                                           const BatchTotals* __restrict__ totals,
                                                           uw_vert* __restrict__ verts, IndexT* __restrict__ inds,
                                                           uint32_t* __restrict__ ticket, const uint2* __restrict__ quarters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const BigSmem s = big_smem_carve(cfg, smem_raw);
    __shared__ uint32_t s_w[64];
    __shared__ uint32_t s_next;
    const int tid = threadIdx.x, NT = UW_BIG_NT, lane = tid & 31;
    const int S = ST > 0 ? ST : cfg.S, L = S + 1, L2 = L * L, ncell = S * S;
    const int CPT = (ncell + NT - 1) / NT;                 // consecutive cells per thread (scan order y, z)
    const bool row_runs = CPT <= 15 && S % CPT == 0;       // a thread's run never straddles two rows
    for (int t = tid; t < 256; t += NT) {
        s.lut[t] = mc->lut[t]; s.rows[t] = mc->rows[t]; s.crossed[t] = mc->crossed[t]; s.nind[t] = mc->ninds[t];
    }
    for (int t = tid; t < 256 * 12; t += NT) s.before[t] = mc->before[t / 12][t % 12];
    if (totals->overflow) return;                          // host grows the arenas and relaunches
    // Work unit = (active chunk, x-part): with only a few chunks per CTA the last wave would sit half empty, so a
    // chunk is cut into 2 or 4 x-ranges when that shortens the schedule.  A part that starts inside the chunk
    // re-classifies the slab before its range (no output) to rebuild the owner state, and takes its running
    // vertex / index bases from the count pass's per-quarter totals.
    const uint32_t n_active = totals->n_active;
    uint32_t parts = 1;
    if ((S & 3) == 0 && n_active > 0) {
        float best = 1e30f;
        for (uint32_t p = 1; p <= 4; p <<= 1) {
            const float waves = (float)((n_active * p + gridDim.x - 1) / gridDim.x);
            const float cost = waves * ((float)S / (float)p + (p > 1 ? 1.f : 0.f));
            if (cost < best * 0.97f) { best = cost; parts = p; }
        }
    }
    const uint32_t n_work = n_active * parts;
    const int xlen = S / (int)parts;

    for (;;) {
        if (tid == 0) s_next = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t w = s_next;
        if (w >= n_work) break;
        const uint32_t chunk = active[w / parts];
        const int part = (int)(w % parts), x0 = part * xlen, x1 = x0 + xlen, xs = part ? x0 - 1 : 0;
        uint32_t pv = 0, pi = 0;                           // vertices / indices of the chunk before slab x0
        for (int q = 0; q < part * (4 / (int)parts); ++q) { const uint2 t = quarters[(size_t)chunk * 4 + q]; pv += t.x; pi += t.y; }
        const float* D = dens + (size_t)chunk * cfg.dens_stride;
        const uw_chunk_desc d = descs[chunk];
        const int offx = d.pos[0] * cfg.chunk_size, offy = d.pos[1] * cfg.chunk_size, offz = d.pos[2] * cfg.chunk_size;
        uw_vert* vout = verts + d.vert_offset;
        IndexT* iout = inds + d.index_offset;
        uint32_t vrun = 0, irun = 0, vrun_prev = 0;        // chunk-local running bases: this slab's, the previous slab's
        float pre[UW_BIG_NLD];
        auto issue = [&](int x) {
            const float* src = D + (size_t)x * L2;
#pragma unroll
            for (int q = 0; q < UW_BIG_NLD; ++q) { const int idx = q * UW_BIG_NT + tid; pre[q] = idx < L2 ? __ldg(src + idx) : 0.f; }
        };
        // plane x -> ring slot x % 3: floats and sign bits (warp ballots).  Three slots: the vertex pass of slab cx
        // still reads planes cx and cx+1 while the next iteration already commits plane cx+2, so no barrier is
        // needed at the end of a slab.
        auto commit = [&](int x) {
            float* pl = s.plane[x % 3];
            uint32_t* bt = s.bits[x % 3];
#pragma unroll
            for (int q = 0; q < UW_BIG_NLD; ++q) {
                const int base = q * UW_BIG_NT + tid - lane;
                if (base < L2) {                                          // warp-uniform
                    const bool ok = base + lane < L2;
                    if (ok) pl[base + lane] = pre[q];
                    const uint32_t word = __ballot_sync(0xFFFFFFFFu, ok && (pre[q] < cfg.iso_level));
                    if (lane == 0) bt[base >> 5] = word;
                }
            }
            if (tid == 0) { bt[(L2 + 31) >> 5] = 0; bt[((L2 + 31) >> 5) + 1] = 0; }
        };
        issue(xs); commit(xs); issue(xs + 1);

        for (int cx = xs; cx < x1; ++cx) {
            const bool warm = cx < x0;                     // owner state only, nothing is written
            commit(cx + 1);
            if (cx + 2 <= x1) issue(cx + 2);               // in flight while this slab is classified and emitted
            __syncthreads();
            const uint32_t* A = s.bits[cx % 3];            // plane x = cx
            const uint32_t* B = s.bits[(cx + 1) % 3];      // plane x = cx + 1
            uint8_t* cs_cur = s.cs[cx & 1];
            uint16_t* vb_cur = s.vb[cx & 1];
            const uint8_t* cs_prev = s.cs[(cx + 1) & 1];
            const uint16_t* vb_prev = s.vb[(cx + 1) & 1];

            // ---- classify this slab; per-thread counts over its CPT consecutive cells --------------------
            const int c0 = tid * CPT;
            uint32_t nva = 0, ni = 0;
            if (row_runs) {
                if (c0 < ncell) {
                    const int y = c0 / S, z0 = c0 - y * S, b = y * L + z0;
                    const uint32_t keep = (2u << CPT) - 1u;
                    const uint32_t q0 = (bits16(A, b) & keep) | ((bits16(B, b) & keep) << 16);
                    const uint32_t q1 = (bits16(A, b + L) & keep) | ((bits16(B, b + L) & keep) << 16);
                    const uint32_t any = q0 | q1, all = q0 & q1;
                    if constexpr (ST == 64) {                // CPT == 8: the run's case bytes leave as one 8-byte store
                        uint32_t w0 = 0, w1 = 0;
                        if (any == 0u || (all & (all >> 16) & keep) == keep) {
                            w0 = w1 = any ? 0xFFFFFFFFu : 0u;
                        } else {
                            const uint32_t ownn = own_mask_of(cx, y, 1), own0 = own_mask_of(cx, y, 0);
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const uint32_t t = s.lut[natural_of(q0, q1, q)];
                                if (q < 4) w0 |= (t & 0xFFu) << (8 * q); else w1 |= (t & 0xFFu) << (8 * (q - 4));
                                if (t >> 8) {
                                    ni += (t >> 8) & 15u;
                                    nva += __popc((t >> 12) & (z0 + q == 0 ? own0 : ownn)) + 0x10000u;
                                }
                            }
                        }
                        *reinterpret_cast<uint2*>(cs_cur + c0) = make_uint2(w0, w1);
                    } else if (any == 0u || (all & (all >> 16) & keep) == keep) {
                        const uint8_t fill = any ? 255 : 0;
                        for (int q = 0; q < CPT; ++q) cs_cur[c0 + q] = fill;
                    } else {
                        for (int q = 0; q < CPT; ++q) {
                            const uint32_t t = s.lut[natural_of(q0, q1, q)];
                            cs_cur[c0 + q] = (uint8_t)t;
                            if (t >> 8) {
                                ni += (t >> 8) & 15u;
                                nva += __popc((t >> 12) & own_mask_of(cx, y, z0 + q)) + 0x10000u;
                            }
                        }
                    }
                }
            } else {
                for (int q = 0; q < CPT; ++q) {
                    const int cell = c0 + q;
                    if (cell >= ncell) break;
                    const int y = cell / S, z = cell - y * S;
                    const uint32_t nat = two_bits(A, y * L + z) | (two_bits(B, y * L + z) << 2)
                                       | (two_bits(A, (y + 1) * L + z) << 4) | (two_bits(B, (y + 1) * L + z) << 6);
                    const uint32_t t = s.lut[nat];
                    cs_cur[cell] = (uint8_t)t;
                    if (t >> 8) {
                        ni += (t >> 8) & 15u;
                        nva += __popc((t >> 12) & own_mask_of(cx, y, z)) + 0x10000u;        // verts <= 12 * S^2 < 2^16 | surface cells << 16
                    }
                }
            }
            uint32_t eva, ei, tva, ti;
            block_scan2(nva, ni, eva, ei, tva, ti, s_w);
            const uint32_t slab_nv = tva & 0xFFFFu, slab_na = tva >> 16;
            if (warm) { vrun = pv - slab_nv; irun = pi - ti; }
            if (slab_na) {                                  // block-uniform
                if (nva) {                                  // slab-relative bases of this thread's surface cells
                    uint32_t rv = eva & 0xFFFFu, ra = eva >> 16, ri = ei;
                    for (int q = 0; q < CPT; ++q) {
                        const int cell = c0 + q;
                        if (cell >= ncell) break;
                        const uint32_t csv = cs_cur[cell];
                        if (csv != 0u && csv != 255u) {
                            const int y = cell / S, z = cell - y * S;
                            vb_cur[cell] = (uint16_t)rv; s.ib[ra] = (uint16_t)ri;
                            s.alist[ra++] = (uint16_t)cell;
                            ri += s.nind[csv];
                            rv += __popc((uint32_t)s.crossed[csv] & own_mask_of(cx, y, z));
                        }
                    }
                }
                __syncthreads();
                const float* P0 = s.plane[cx % 3];
                const float* P1 = s.plane[(cx + 1) % 3];
                auto dens_at = [=](int ax, int ay, int az) { return (ax == cx ? P0 : P1)[ay * L + az]; };
                for (uint32_t v0 = 0; !warm && v0 < (slab_nv ? slab_nv : 1u); v0 += UW_BIG_VCAP) {
                    if (v0) __syncthreads();                // the previous tile's vertex threads are done with vlist
                    // ---- one thread per surface cell: indices (first tile) + this tile of the owned-edge list ----
                    for (uint32_t a = tid; a < slab_na; a += NT) {
                        const int cell = s.alist[a];
                        const int y = cell / S, z = cell - y * S;
                        const uint32_t csv = cs_cur[cell];
                        const uint64_t row = s.rows[csv];
                        uint32_t vnext = vb_cur[cell], todo = own_mask_of(cx, y, z);
                        IndexT* dst = iout + irun + s.ib[a];
#pragma unroll 1
                        for (int k = 0; k < 15; ++k) {
                            const int e = (int)((row >> (4 * k)) & 0xFull);
                            if (e == 15) break;
                            if ((todo >> e) & 1u) {                    // owned edge, first appearance: its vertex
                                todo &= ~(1u << e);
                                const uint32_t slot = vnext - v0;
                                if (slot < UW_BIG_VCAP) s.vlist[slot] = (uint16_t)(cell | (e << 12));
                                ++vnext;
                            }
                            if (v0 == 0) {
                                int ox, oy, oz, oe;
                                owner_of(e, cx, y, z, ox, oy, oz, oe);
                                const int ocell = oy * S + oz;
                                const uint32_t ocs = ox == cx ? (uint32_t)cs_cur[ocell] : (uint32_t)cs_prev[ocell];
                                const uint32_t ovb = ox == cx ? vrun + vb_cur[ocell] : vrun_prev + vb_prev[ocell];
                                const uint32_t rank = __popc((uint32_t)s.before[ocs * 12 + oe] & own_mask_of(ox, oy, oz));
                                dst[k] = (IndexT)(ovb + rank);
                            }
                        }
                    }
                    __syncthreads();
                    // ---- one thread per vertex of the tile ------------------------------------------------------
                    const uint32_t cnt = min((uint32_t)UW_BIG_VCAP, slab_nv - v0);
                    for (uint32_t t = tid; t < cnt; t += NT) {
                        const uint32_t ent = s.vlist[t];
                        const int cell = ent & 0xFFF, e = ent >> 12;
                        const int y = cell / S, z = cell - y * S;
                        float v[6];
                        make_vertex_from(cfg, dens_at, cx, y, z, e, offx, offy, offz, v);
                        float2* vd = reinterpret_cast<float2*>(vout + vrun + v0 + t);
                        vd[0] = make_float2(v[0], v[1]); vd[1] = make_float2(v[2], v[3]); vd[2] = make_float2(v[4], v[5]);
                    }
                }
            }
            vrun_prev = vrun;
            vrun += slab_nv; irun += ti;
            // no barrier here: the next commit writes ring slot (cx + 2) % 3; case / vertex-base arrays, the surface-
            // cell list and the vertex list are only rewritten behind the next iteration's first barriers
        }
        __syncthreads();                                   // before the next work unit reuses every buffer
    }
}

// ---------------------------------------------------------------------------------------
// FUSED PATH: one persistent kernel does K1..K4 per chunk; densities never leave the SM.
//
//   ticket   chunks are handed out by an atomic counter, so a chunk's predecessors are always
//            owned by CTAs that are already running (single-pass scan precondition)
//   K1       noise_chunk_spec  -> densities + column sign masks in shared memory
//   K2       cases + per-chunk (vertex, index) counts, blank / no-surface vote
//   K3       decoupled look-back over the per-chunk aggregates -> this chunk's packed offsets
//            (one warp polls 32 predecessors at a time; value and status share one 64-bit word)
//   K4       emit_chunk        -> vertices + indices straight to the packed output arenas
//
// HBM traffic = 12 B position in, 32 B descriptor + 24 B/vertex + 2 B/index out.
// ---------------------------------------------------------------------------------------
struct ScanSlot { unsigned long long v, i; };     // bits 63..62: 0 = empty, 1 = aggregate, 2 = inclusive prefix
#define SCAN_AGG  (1ull << 62)
#define SCAN_PFX  (2ull << 62)
#define SCAN_VAL  ((1ull << 62) - 1ull)

template <int ST, int NOCT>
struct FusedSmem {
    SpecSmem<ST, NOCT> n;                             // n.lat / n.X are dead after K1 and reused by K4 (vid)
    static constexpr int VSTAGE_U16 = (SpecDims<ST, NOCT>::NTF / 32) * (UW_VSTAGE_BYTES / 2);
    alignas(16) uint16_t vlist[UW_VLIST_CAP];         // K4 vertex list; dead after D2 -> index staging (emit_indices)
    // vbase / ibase / alist: one entry per SURFACE cell (worst case: all).  vbase is dead after D1 -> vertex staging
    alignas(16) uint16_t vbase[ST * ST * ST + 8 > VSTAGE_U16 ? ST * ST * ST + 8 : VSTAGE_U16];
    uint16_t ibase[ST * ST * ST + 8];
    uint16_t alist[ST * ST * ST + 8];
    uint64_t rows[256];
    float powtab[48];
    uint32_t lut[256];
    uint16_t eoff[16];
    uint8_t cs[((ST * ST * ST + 15) / 16) * 16];
    uint32_t w[64];
    unsigned long long part[4 * 8];
    int last;                                         // set in the CTA that leaves the launch last
    uint32_t hand[UW_NCLS + 1];                       // cost-ordered hand-out: class-list prefix sums, ready flag
    Handout handout;
};

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// Executed by the WHOLE CTA (NT threads, NT a multiple of 32): every thread polls one predecessor,
// so one round covers NT chunks.  Returns (in every thread) the exclusive prefix of (v, i) over
// chunks < c.  s_part: 4 * (NT/32) u64 of shared scratch.
__device__ __forceinline__ void lookback_block(ScanSlot* st, uint32_t c, unsigned long long& ev, unsigned long long& ei,
                                               unsigned long long* s_part) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    unsigned long long sv = 0, si = 0;
    long long j = (long long)c - 1;
    while (true) {
        const long long idx = j - tid;
        unsigned long long a = SCAN_PFX, b = SCAN_PFX;        // virtual predecessor (idx < 0): prefix 0
        if (idx >= 0) {
            // value and status share one 64-bit word, so each word is self-consistent; a slot being
            // upgraded from aggregate to prefix can show mixed tags for a moment: re-read until equal
            do {
                a = ld_volatile_u64(&st[idx].v);
                b = ld_volatile_u64(&st[idx].i);
            } while ((a >> 62) == 0 || (a >> 62) != (b >> 62));
        }
        const bool pfx = (a >> 62) == 2;
        const uint32_t ball = __ballot_sync(0xFFFFFFFFu, pfx);
        const int first = ball ? (__ffs(ball) - 1) : 32;       // nearest predecessor (in this warp) holding a prefix
        unsigned long long cv = lane <= first ? (a & SCAN_VAL) : 0ull;
        unsigned long long ci = lane <= first ? (b & SCAN_VAL) : 0ull;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            cv += __shfl_xor_sync(0xFFFFFFFFu, cv, d);
            ci += __shfl_xor_sync(0xFFFFFFFFu, ci, d);
        }
        if (lane == 0) { s_part[warp * 4] = cv; s_part[warp * 4 + 1] = ci; s_part[warp * 4 + 2] = ball ? 1ull : 0ull; }
        __syncthreads();
        bool done = false;
        for (int w = 0; w < nw && !done; ++w) {                // warps are ordered nearest-first
            sv += s_part[w * 4]; si += s_part[w * 4 + 1];
            done = s_part[w * 4 + 2] != 0ull;
        }
        __syncthreads();
        if (done) break;
        j -= blockDim.x;
    }
    ev = sv; ei = si;
}

#ifndef UW_FUSED_MINB
#define UW_FUSED_MINB 4
#endif
#ifndef UW_FUSED_PREFETCH_H          // the spare warp hashes the next chunk's lattice during stage YZ (A/B switch)
#define UW_FUSED_PREFETCH_H (UW_FUSED_EXTRA_WARPS == 1)
#endif
// PEER: the output arenas are another GPU's memory (gather path): vertices and indices leave through shared memory
// as whole 16-byte vectors (see emit_verts / emit_indices).  A separate instantiation, so that the kernel that writes
// the GPU's own HBM is exactly the register-store code (both in one kernel cost the local path 7 %, measured).
template <int ST, int NOCT, typename IndexT, bool PEER>
__global__ void __launch_bounds__(SpecDims<ST, NOCT>::NTF, UW_FUSED_MINB)
k_build_fused(const __grid_constant__ DevCfg cfg, const __grid_constant__ AxisTables tab,
              const uint8_t* __restrict__ g_perm, const McTables* __restrict__ mc,
              const int32_t* __restrict__ pos, uint32_t n,
              ScanSlot* __restrict__ scan, FusedControl* __restrict__ ctr, FusedControl* __restrict__ ctr_next,
              uw_chunk_desc* __restrict__ descs,
              uw_vert* __restrict__ verts, IndexT* __restrict__ inds,
              unsigned long long vcap, unsigned long long icap,
              float* __restrict__ dens_out /*nullable: debug tap*/, int ordered,
              uw_tri* __restrict__ tris /*nullable: UW_FLAG_TRIS*/, uint16_t* __restrict__ tri_cell /*nullable*/,
              uint4* __restrict__ order /*nullable: cost-ordered hand-out, [UW_NCLS][n]*/, int z_lo, int z_hi,
              unsigned long long zcls, int analytic_skip, FusedSummary* __restrict__ sum_out,
              const __grid_constant__ FusedOut fo) {
    using D = SpecDims<ST, NOCT>;
    BatchTotals* const totals = &ctr->totals;
    unsigned long long* const guard_count = &ctr->guard;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FusedSmem<ST, NOCT>& sm = *reinterpret_cast<FusedSmem<ST, NOCT>*>(smem_raw);
    const int tid = threadIdx.x;
    constexpr int L = D::L;

    // the vertex-id table of K4 aliases the K1 lattice/X tables (dead once noise_chunk_spec has returned;
    // lat and X are adjacent members of SpecSmem)
    EmitSmem es;
    es.dens = sm.n.dens; es.bits = nullptr; es.mask = sm.n.mask;
    es.vid = reinterpret_cast<uint16_t*>(sm.n.X);
    es.vlist = sm.vlist; es.vbase = sm.vbase; es.ibase = sm.ibase; es.alist = sm.alist;
    es.cs = sm.cs; es.lut = sm.lut; es.eoff = sm.eoff; es.rows = sm.rows; es.powtab = sm.powtab;
    es.vstage = reinterpret_cast<float*>(sm.vbase); es.istage = sm.vlist;
    using NS = SpecSmem<ST, NOCT>;
    static_assert(offsetof(NS, xpad) == offsetof(NS, X) + sizeof(sm.n.X), "X and xpad must be adjacent");
    static_assert(sizeof(sm.n.X) + sizeof(sm.n.xpad) >= (size_t)L * L * L * UW_EDGE_KINDS * 2, "vid must fit");
    static_assert(sizeof(FusedSmem<ST, NOCT>) <= (233472 / UW_FUSED_MINB) - 1024, "shared memory budget of UW_FUSED_MINB CTAs per SM");

    // first ticket; later ones are requested inside K1 (see noise_chunk_spec) and published at the end of
    // the iteration, so chunks are still handed out on demand (committing a whole chunk ahead was measured
    // slower: with ~3.5 chunks per CTA the tail grows by up to one chunk)
    // hand-out parameters live in shared memory: the ticket code runs in ONE thread, deep inside K1, and must not
    // cost the other phases any registers
    Handout& hand = sm.handout;
    if (tid == 0) {
        hand.ctr = ctr; hand.pos = pos; hand.n = n; hand.order = order; hand.state = sm.hand;
        hand.z_lo = z_lo; hand.z_hi = z_hi; hand.zcls = zcls;
        hand.skip = (analytic_skip && !ordered && z_hi >= z_lo) ? descs : nullptr;
        sm.hand[UW_NCLS] = 0u;
    }
    __syncthreads();
    // start-up order matters at 2048 chunks (the prologue is ~10 % of the kernel): first the global round trips of
    // the hand-out (filing the request / the first ticket's atomic), then the table loads underneath them
    uint32_t t_first = 0;
    if (tid == D::NTF - 1) t_first = ticket_begin(hand);
    if (order) handout_classify(hand);

    for (int t = tid; t < 256; t += D::NTF) sm.n.perm[t] = g_perm[t];
    if (tid < 16) sm.n.grad[tid] = make_float4(c_grad_vec[tid][0], c_grad_vec[tid][1], c_grad_vec[tid][2], 0.f);
    for (int t = tid; t < NOCT * L; t += D::NTF) {
        const int o = t / L, i = t - o * L;
        sm.n.axis[o][i] = make_float4(tab.d[o][i], tab.d1[o][i], tab.w[o][i], 0.f);
    }
    for (int t = tid; t < 256; t += D::NTF) { sm.lut[t] = mc->lut[t]; sm.rows[t] = mc->rows[t]; }
    if (tid < 48) sm.powtab[tid] = mc->powtab[tid];
    fill_edge_offsets(sm.eoff, L);

    if (tid == D::NTF - 1) {
        if (order) handout_ready(hand);
        const Ticket t0 = ticket_fetch(hand, t_first);
        sm.n.ticket[0][0] = (int)t0.chunk; sm.n.ticket[0][1] = t0.px; sm.n.ticket[0][2] = t0.py; sm.n.ticket[0][3] = t0.pz;
    }
    __syncthreads();
#ifdef UW_PHASE_TIMING
    long long t_phase = clock64();
    if (tid == 0 && blockIdx.x < 1024) { g_cta[blockIdx.x][0] = gtimer(); g_cta[blockIdx.x][2] = 0; g_cta[blockIdx.x][3] = 0; }
#endif
    int tb = 0;                                            // half of sm.n.terr / sm.n.ticket that belongs to the current chunk
    // PF: every iteration finds its lattice hashed (by the previous iteration's spare warp); the first one by all threads here
    constexpr bool PFH = UW_FUSED_PREFETCH_H != 0;
    if (PFH && (uint32_t)sm.n.ticket[0][0] != TICKET_DONE) {           // block-uniform
        const int px = sm.n.ticket[0][1], py = sm.n.ticket[0][2], pz = sm.n.ticket[0][3];
        noise_stage_h<ST, NOCT>(sm.n, px, py, pz, tid, D::NTF);
        if (tid >= D::NTF - 32 && tid - (D::NTF - 32) < L) sm.n.terr[0][tid - (D::NTF - 32)] = terrace_lookup(cfg, tid - (D::NTF - 32), pz);
        __syncthreads();
    }
    while (true) {
        const volatile int* const cur = sm.n.ticket[tb];   // this chunk's ticket; K1 leaves the next one in ticket[tb ^ 1]
        const uint32_t chunk = (uint32_t)cur[0];
        if (chunk == TICKET_DONE) break;
        PHASE_MARK(0);                                     // ticket + position

        // ---- K1 ---------------------------------------------------------------------------------
        const uint32_t fl = noise_chunk_spec<ST, NOCT, D::NTF, PFH>(cfg, tab, sm.n, 0, 0, 0, guard_count PHASE_PASS, &hand, tb);
        tb ^= 1;
        PHASE_MARK(1);                                     // K1 noise
        if (dens_out) {
            float4* dst = reinterpret_cast<float4*>(dens_out + (size_t)chunk * D::DSTRIDE);
            const float4* src = reinterpret_cast<const float4*>(sm.n.dens);
            for (int t = tid; t < D::DSTRIDE / 4; t += D::NTF) dst[t] = src[t];
        }

        // ---- K2: cases, counts, per-cell bases (block-uniform skip when no sample is inside) -------------
        ChunkShape sh;
        sh.n_vert = 0; sh.n_ind = 0; sh.n_act = 0;
        unsigned long long packed = 0;
        if ((fl & (CF_ANY_LT | CF_ALL_LT)) == CF_ANY_LT)   // all-empty and all-full chunks skip K2..K4 (north_star's ballot skip)
            sh = emit_prepare<ST>(cfg, es, sm.w, tri_cell ? tri_cell + (size_t)chunk * (ST * ST * ST + 1) : nullptr,
                                  ordered ? nullptr : &ctr->alloc, &packed, (int)(16 / sizeof(IndexT)) - 1);
        const uint32_t nv = sh.n_vert, ni = sh.n_ind;
        const uint32_t nv_pad = pad_verts(nv), ni_pad = pad_inds<IndexT>(ni);   // what the chunk occupies in the arenas
        PHASE_MARK(2);                                     // K2 prepare

        // ---- K3: this chunk's offsets in the packed arenas ------------------------------------------------
        //  ordered  : decoupled look-back over per-chunk aggregates -> offsets follow request order
        //             (deterministic layout; a chunk waits for the slowest of its in-flight predecessors)
        //  unordered: one 64-bit atomic bump allocation -> no inter-CTA dependency; each chunk's OWN
        //             buffers are identical either way, only their placement in the arena differs.
        //             The atomic is issued here and its result consumed after the D1 fill.
        unsigned long long ev = 0, ei = 0;
        if (ordered) {
            if (tid == 0) {
                const unsigned long long tag = chunk == 0 ? SCAN_PFX : SCAN_AGG;
                atomicExch(&scan[chunk].v, tag | nv_pad);
                atomicExch(&scan[chunk].i, tag | ni_pad);
            }
            if (chunk > 0) lookback_block(scan, chunk, ev, ei, sm.part);
            if (tid == 0 && chunk > 0) {
                atomicExch(&scan[chunk].v, SCAN_PFX | (ev + nv_pad));
                atomicExch(&scan[chunk].i, SCAN_PFX | (ei + ni_pad));
            }
            packed = (ev << 32) | ei;
        }

        // ---- K4 ---------------------------------------------------------------------------------
        PHASE_MARK(3);                                     // K3 offsets
        if (ni > 0) {                                      // block-uniform
            emit_fill<ST>(cfg, mc, es, sh, 0);
            if (tid == 0) sm.part[0] = packed;
            __syncthreads();
            PHASE_MARK(4);                                 // K4 D1 fill
            ev = sm.part[0] >> 32; ei = sm.part[0] & 0xFFFFFFFFull;
            if (ev + nv_pad <= vcap && ei + ni_pad <= icap) {              // block-uniform
                const int px = cur[1], py = cur[2], pz = cur[3];
                emit_rest<ST, IndexT, PEER>(cfg, mc, es, sh, px, py, pz, verts + ev, inds + ei);
                if (tris) emit_tris<ST>(cfg, mc, es, sh, px, py, pz, tris + ei / 3u);
            }
            PHASE_MARK(5);                                 // K4 D2 + E (thread 0's own share)
        }
        if (tid == 0) {
            uw_chunk_desc d;
            d.pos[0] = cur[1]; d.pos[1] = cur[2]; d.pos[2] = cur[3];
            d.flags = ((fl & CF_ALL_GT) ? UW_CHUNK_BLANK_EARLY : 0u) | (ni > 0 ? UW_CHUNK_HAS_MESH : 0u)
                    | (nv > 65536u ? UW_CHUNK_U16_OVERFLOW : 0u);
            d.vert_offset = (uint32_t)ev + fo.desc_vbase; d.vert_count = nv;
            d.index_offset = (uint32_t)ei + fo.desc_ibase; d.index_count = ni;
            descs[chunk] = d;
            if (ni > 0) {
                const uint32_t slot = atomicAdd(&totals->n_active, 1u);
                // the draw list (world.rs:117-121 keeps only not_blank chunks for rendering): 32 bytes per meshed chunk
                if (fo.drawlist) fo.drawlist[slot] = d;
            }
            if (fl & CF_ALL_GT) atomicAdd(&totals->n_blank, 1u);
            if (ordered && chunk == n - 1) {
                totals->n_verts = ev + nv_pad; totals->n_inds = ei + ni_pad;
                if (ev + nv_pad > vcap || ei + ni_pad > icap || ev + nv_pad > 0xFFFFFFFFull || ei + ni_pad > 0xFFFFFFFFull)
                    totals->overflow = 1u;
            }
            if (!ordered && ni > 0 && (ev + nv_pad > vcap || ei + ni_pad > icap)) atomicMax(&totals->overflow, 1u);
            // completion-order packing keeps (vertices << 32 | indices) in ONE 64-bit counter: an index total beyond
            // 2^32 would carry into the vertex half.  The chunk whose claim crosses the boundary sees it here.
            if (!ordered && ni > 0 && (ei + ni_pad > 0xFFFFFFFFull || ev + nv_pad > 0xFFFFFFFFull)) atomicMax(&totals->overflow, 2u);
        }
        // chunk fully emitted, shared memory reusable.  A chunk without K2..K4 (all empty / all full: most of a region)
        // needs no barrier here: its next ticket became visible at the barrier that ended K1, the vote flags, the ticket
        // slots and the terrace terms alternate between chunks, and everything else K1 writes is ordered by K1's own
        // barriers -- the other warps start the next chunk's stage X while warp 0 writes the descriptor.
#ifndef UW_NO_LIGHT_SKIP
        if (dens_out != nullptr || (fl & (CF_ANY_LT | CF_ALL_LT)) == CF_ANY_LT)        // block-uniform
#endif
        __syncthreads();
        PHASE_MARK(6);                                     // tail: descriptor + waiting for the other warps
#ifdef UW_PHASE_TIMING
        if (tid == 0) { atomicAdd(&g_phase[8], 1ull); if (ni > 0) atomicAdd(&g_phase[9], 1ull); if (fl & CF_ANY_LT) atomicAdd(&g_phase[10], 1ull);
                        if (blockIdx.x < 1024) { g_cta[blockIdx.x][1] = gtimer(); g_cta[blockIdx.x][2] += 1; g_cta[blockIdx.x][3] += (ni > 0); } }
#endif
    }
    // last CTA out resets the other control block for the next launch
    __syncthreads();                                       // every thread has seen TICKET_DONE
    if (tid == 0) {
        if (fo.head) __threadfence_system();               // gather segment: the outputs may be another GPU's memory
        else __threadfence();
        sm.last = (atomicAdd(&ctr->done, 1u) == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (sm.last) {
        if (tid == 0) {
            FusedSummary sm_out;                     // every other CTA fenced its writes before bumping `done`
            sm_out.alloc = atomicAdd(&ctr->alloc, 0ull);
            sm_out.guard = atomicAdd(&ctr->guard, 0ull);
            sm_out.totals.n_verts = atomicAdd(&ctr->totals.n_verts, 0ull);
            sm_out.totals.n_inds = atomicAdd(&ctr->totals.n_inds, 0ull);
            sm_out.totals.n_active = atomicAdd(&ctr->totals.n_active, 0u);
            sm_out.totals.overflow = atomicAdd(&ctr->totals.overflow, 0u);
            sm_out.totals.n_blank = atomicAdd(&ctr->totals.n_blank, 0u);
            sm_out.totals.n_mesh = 0;
            *sum_out = sm_out;                       // host-mapped memory: visible to the host at kernel completion
            if (fo.head) {                           // gather segment: summary, system fence, then the epoch flag
                fo.head->sum = sm_out; fo.head->n_chunks = n;
                fo.head->first_chunk_lo = fo.first_chunk_lo; fo.head->first_chunk_hi = fo.first_chunk_hi;
                __threadfence_system();
                asm volatile("st.volatile.global.u32 [%0], %1;" :: "l"(&fo.head->epoch), "r"(fo.epoch) : "memory");
            }
            ctr_next->ticket = 0; ctr_next->done = 0; ctr_next->classified = 0; ctr_next->cls_ticket = 0;
            for (int k = 0; k < UW_NCLS; ++k) ctr_next->cls_n[k] = 0;
            ctr_next->alloc = 0; ctr_next->guard = 0;
            BatchTotals z; z.n_verts = 0; z.n_inds = 0; z.n_active = 0; z.overflow = 0; z.n_blank = 0; z.n_mesh = 0;
            ctr_next->totals = z;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Collision ray casts against the per-cell triangle lists (SURVEY 8f-1's consumer): util::Tri::intersects
// (util.rs:22-59) over the triangles Chunk::tris_around (chunk.rs:315-342) returns for the chunks a boid visits
// (boid.rs:175-208), for a whole flock of rays at once.  cgmath's operation order, f32, unfused.
//   k_chunk_table : chunk position -> chunk index of the last build (open addressing; the key is compared against
//                   the descriptor's position, so the table holds indices only)
//   k_raycast     : one warp per ray.  At most 2 x 2 x 2 chunks lie within +-wall_range world units (wall_range <
//                   chunk_size); lanes stride over the (<= (2 wall_range + 1)^3) cells of each, walk the cell's
//                   triangles, keep the smallest hit distance; warp min; -1 = every test returned None.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pos_hash(int x, int y, int z) {
    uint32_t h = (uint32_t)x * 0x9E3779B1u ^ (uint32_t)y * 0x85EBCA77u ^ (uint32_t)z * 0xC2B2AE3Du;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
    return h;
}

__global__ void __launch_bounds__(256) k_chunk_table(const uw_chunk_desc* __restrict__ descs, uint32_t n,
                                                     uint32_t* __restrict__ table, uint32_t mask) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uw_chunk_desc d = descs[i];
        uint32_t slot = pos_hash(d.pos[0], d.pos[1], d.pos[2]) & mask;
        while (atomicCAS(&table[slot], 0u, i + 1u) != 0u) slot = (slot + 1u) & mask;     // table is >= 2n slots: terminates
    }
}

__device__ __forceinline__ float dot3_rn(float ax, float ay, float az, float bx, float by, float bz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));     // cgmath: (x x' + y y') + z z'
}

// util.rs:22-59; returns t or -1 (None)
__device__ __forceinline__ float tri_intersects(const uw_tri& tr, float px, float py, float pz, float dx, float dy, float dz, float range) {
    const float nx = tr.normal[0], ny = tr.normal[1], nz = tr.normal[2];
    const float dnd = dot3_rn(nx, ny, nz, dx, dy, dz);
    if (fabsf(dnd) < 1e-5f) return -1.0f;
    const float t = __fdiv_rn(dot3_rn(nx, ny, nz, __fsub_rn(tr.verts[0][0], px), __fsub_rn(tr.verts[0][1], py), __fsub_rn(tr.verts[0][2], pz)), dnd);
    if (t < 0.0f || t > range) return -1.0f;
    const float ix = __fadd_rn(px, __fmul_rn(dx, t)), iy = __fadd_rn(py, __fmul_rn(dy, t)), iz = __fadd_rn(pz, __fmul_rn(dz, t));
    float q[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float* a = tr.verts[k];
        const float* b = tr.verts[k == 2 ? 0 : k + 1];
        const float ex = __fsub_rn(b[0], a[0]), ey = __fsub_rn(b[1], a[1]), ez = __fsub_rn(b[2], a[2]);
        const float wx = __fsub_rn(ix, a[0]), wy = __fsub_rn(iy, a[1]), wz = __fsub_rn(iz, a[2]);
        const float cx = __fsub_rn(__fmul_rn(ey, wz), __fmul_rn(ez, wy));
        const float cy = __fsub_rn(__fmul_rn(ez, wx), __fmul_rn(ex, wz));
        const float cz = __fsub_rn(__fmul_rn(ex, wy), __fmul_rn(ey, wx));
        q[k] = dot3_rn(cx, cy, cz, nx, ny, nz);
    }
    return (q[0] >= 0.0f && q[1] >= 0.0f && q[2] >= 0.0f) ? t : -1.0f;
}

__global__ void __launch_bounds__(256) k_raycast(const __grid_constant__ DevCfg cfg, const uw_chunk_desc* __restrict__ descs,
                                                 const uw_tri* __restrict__ tris, const uint16_t* __restrict__ tri_cell,
                                                 const uint32_t* __restrict__ table, uint32_t mask,
                                                 const float* __restrict__ origins, const float* __restrict__ dirs,
                                                 uint32_t n_rays, int wall_range, float* __restrict__ out_t) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    const int S = cfg.S, NC1 = S * S * S + 1;
    const float cs = (float)cfg.chunk_size, wr = (float)wall_range;
    for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rays; r += warps) {
        const float px = origins[3 * r], py = origins[3 * r + 1], pz = origins[3 * r + 2];
        const float dx = dirs[3 * r], dy = dirs[3 * r + 1], dz = dirs[3 * r + 2];
        const float p[3] = {px, py, pz};
        int ws[3], we[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {                                    // boid.rs:177-183
            ws[k] = (int)floorf(__fdiv_rn(__fsub_rn(p[k], wr), cs));
            we[k] = (int)floorf(__fdiv_rn(__fadd_rn(p[k], wr), cs));
        }
        float best = 3.0e38f;
        for (int a = ws[0]; a <= we[0]; ++a) for (int b = ws[1]; b <= we[1]; ++b) for (int c = ws[2]; c <= we[2]; ++c) {
            // world.get_chunk((a, b, c)): warp-uniform probe
            uint32_t slot = pos_hash(a, b, c) & mask, ci = 0xFFFFFFFFu;
            for (;;) {
                const uint32_t e = table[slot];
                if (e == 0u) break;
                const uw_chunk_desc& d = descs[e - 1u];
                if (d.pos[0] == a && d.pos[1] == b && d.pos[2] == c) { ci = e - 1u; break; }
                slot = (slot + 1u) & mask;
            }
            if (ci == 0xFFFFFFFFu) continue;
            const uint32_t nind = descs[ci].index_count;
            if (nind == 0u) continue;
            const uw_tri* ctris = tris + descs[ci].index_offset / 3u;
            const uint16_t* cstart = tri_cell + (size_t)ci * NC1;
            const int cp[3] = {a, b, c};
            int lo[3], hi[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {                                // boid.rs:186-203, chunk.rs:316-326
                const float local = __fsub_rn(p[k], __fmul_rn((float)cp[k], cs));
                const int mid = (int)floorf(__fmul_rn(__fdiv_rn(local, cs), (float)S));
                lo[k] = max(mid - wall_range, 0); hi[k] = min(min(mid + wall_range, S), S - 1);   // cells == S hold nothing
            }
            const int ny = hi[1] - lo[1] + 1, nz = hi[2] - lo[2] + 1, ncell = (hi[0] - lo[0] + 1) * ny * nz;
            if (hi[0] < lo[0] || ny <= 0 || nz <= 0) continue;
            for (int q = (int)lane; q < ncell; q += 32) {
                const int x = lo[0] + q / (ny * nz), rem = q % (ny * nz), y = lo[1] + rem / nz, z = lo[2] + rem % nz;
                const int cell = (x * S + y) * S + z;
                const uint32_t t0 = cstart[cell], t1 = cstart[cell + 1];
                for (uint32_t j = t0; j < t1; ++j) {
                    const float t = tri_intersects(ctris[j], px, py, pz, dx, dy, dz, wr);
                    if (t >= 0.0f) best = fminf(best, t);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(0xFFFFFFFFu, best, o));
        if (lane == 0) out_t[r] = best < 3.0e38f ? best : -1.0f;
    }
}

// ---------------------------------------------------------------------------------------
// Multi-GPU gather, consumer side (render GPU): wait until every segment's head carries the expected epoch.
// One warp; lane s polls segment s with volatile (system-coherent, L1-bypassing) loads.  The producers are
// fused-kernel launches on other GPUs of the box whose last CTA published the epoch after a system-wide
// fence, so once the flags are seen (and this kernel's own fence has run) everything they wrote into the
// arenas is visible to whatever follows on the stream.  status[0] = 0 ok, 1 timed out.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_gather_wait(const GatherHead* __restrict__ head, uint32_t n_segments,
                                                    uint32_t epoch, unsigned long long timeout_ns,
                                                    uint32_t* __restrict__ status) {
    const uint32_t lane = threadIdx.x;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    bool ok = true;
    for (uint32_t s = lane; s < n_segments; s += 32u) {
        // epochs only grow; signed distance keeps the comparison valid across a wrap
        while ((int32_t)(ld_volatile_u32(&head[s].epoch) - epoch) < 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
            if (now - t0 > timeout_ns) { ok = false; break; }
            __nanosleep(100);
        }
    }
    __threadfence_system();
    const bool all_ok = __all_sync(0xFFFFFFFFu, ok);
    if (lane == 0) status[0] = all_ok ? 0u : 1u;
}

// ---------------------------------------------------------------------------------------
// Measurement aid (not on the product path): FFMA-chain microbenchmark for the "measured FP32 peak" the
// noise-stage roofline is quoted against.  16 independent chains per thread.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ffma_peak(float* __restrict__ out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) x[q] = (float)(threadIdx.x + q) * 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 16; ++q) x[q] = fmaf(x[q], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 16; ++q) s += x[q];
    if (s == 123.456f) out[0] = s;                 // keeps the chains alive, never true in practice
}
