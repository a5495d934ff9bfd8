// uwcuda.cu -- host side of libuwcuda.so: context, buffers, launch sequence, C ABI.
// Declared in include/uwcuda.h.  No CPU fallback: every entry point that computes needs a
// CUDA device.  Nothing here includes, links or calls anything under oracle/.
#include <cuda.h>            // driver API TYPES only (VMM arenas); entry points come from cudaGetDriverEntryPoint
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <limits.h>
#include <string.h>
#include <chrono>
#include <string>
#include <vector>
#include <algorithm>
#include <utility>

#include "uw_kernels.cuh"

// ---------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------
static thread_local std::string g_create_error;

struct PinnedBlock { void* ptr; size_t bytes; };
struct VmmArena { void* ptr = nullptr; size_t bytes = 0; CUmemGenericAllocationHandle handle = 0; };
static void vmm_free(VmmArena* a);

struct uw_ctx {
    uw_config cfg;
    DevCfg dcfg;
    AxisTables tab;
    int device = 0, num_sms = 0;
    bool fast_path = true;          // FP32 factorised noise (+ f64 guard band) vs exact f64
    bool index32 = false;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    uint8_t perm[256];
    uint8_t* d_perm = nullptr;
    std::vector<float> ax_d, ax_d1, ax_w;   // [octaves][L] axis tables (f64-computed), any L
    std::vector<int> ax_c;
    float4* d_axis = nullptr;               // device copy [octaves][L] (d, d-1, fade, 0) for the large-chunk noise kernel
    bool big_fast_noise = false;            // FP32 plane-tiled noise available for this configuration
    size_t big_noise_smem = 0; int big_noise_blocks_per_sm = 1;
    McTables* d_mc = nullptr;
    float* d_terr_tab = nullptr;            // terrace terms per z layer (DevCfg::terr_tab)

    bool tris = false;              // UW_FLAG_TRIS: per-cell collision triangles
    bool exportable = false;        // UW_FLAG_EXPORTABLE: output arenas are VMM allocations with fd handles
    bool host_ptr_ok = false;       // the device can read pinned host allocations through their host address
    // grow-only per-batch buffers: TWO sets, so that a host build's D2H copies (set A, copy stream) overlap the
    // next batch's kernel (set B, compute stream).  B() is the set of the build being enqueued / last enqueued.
    struct BufSet {
        uint32_t cap_chunks = 0;
        int32_t* d_pos = nullptr;
        float* d_dens = nullptr;        // only the staged / large-chunk / debug paths materialise densities
        uint32_t cap_dens_chunks = 0;
        ChunkCounts* d_counts = nullptr;
        uint2* d_quarters = nullptr;    // large-chunk path: per-chunk (verts, inds) of each x-quarter
        uw_chunk_desc* d_descs = nullptr;
        uint32_t* d_active = nullptr;
        uint8_t* d_cases = nullptr;  uint32_t cap_cases_chunks = 0;
        ScanSlot* d_scan = nullptr;
        uint4* d_order = nullptr;       // cost-ordered hand-out lists of the fused kernel, [UW_NCLS][order_cap]
        unsigned long long vcap = 0, icap = 0;
        uw_vert* d_verts = nullptr;
        void* d_inds = nullptr;
        VmmArena vmm_verts, vmm_inds;   // UW_FLAG_EXPORTABLE: d_verts / d_inds alias these
        uw_tri* d_tris = nullptr;       // [icap / 3]
        uint16_t* d_tri_cell = nullptr; // [cap_chunks][S^3 + 1]
        int32_t* h_pos = nullptr; size_t h_pos_cap = 0;   // pinned staging of the request
        FusedSummary* h_sum = nullptr;  // pinned + mapped: the last CTA of the fused kernel writes it over PCIe
        cudaEvent_t done = nullptr;     // recorded after the set's kernels
        bool busy = false;              // an async batch that has not been collected owns this set
        uw_batch* owner = nullptr;      // that batch
        cudaEvent_t copied = nullptr;   // recorded on the copy stream after the owner's D2H copies
        // state of the set's last build
        uint32_t last_n = 0;
        const int32_t* last_pos_dev = nullptr;
        bool last_fused = false;
        bool last_gather = false;       // the set's last build wrote into a gather segment (no local arenas, no regrow)
        bool pending = false;           // kernels enqueued, totals not yet validated
        BatchTotals result = {};        // validated totals of the last finished build
    } sets[2];
    int cur = 0;
    BufSet& B() { return sets[cur]; }
    BatchTotals* d_totals = nullptr;
    unsigned long long* d_guard = nullptr;
    // chunk-level scan (staged / large-chunk paths): per-tile totals + epoch flags, see k_scan_chunks
    ScanPart* d_scan_part = nullptr; uint32_t* d_scan_flag = nullptr; ScanCtl* d_scan_ctl = nullptr;
    uint32_t scan_tiles_cap = 0, scan_epoch = 0;
    typedef void (*big_emit16_fn_t)(DevCfg, const McTables*, const float*, const uw_chunk_desc*, const uint32_t*, const BatchTotals*, uw_vert*, uint16_t*, uint32_t*, const uint2*);
    typedef void (*big_emit32_fn_t)(DevCfg, const McTables*, const float*, const uw_chunk_desc*, const uint32_t*, const BatchTotals*, uw_vert*, uint32_t*, uint32_t*, const uint2*);
    typedef void (*big_count_fn_t)(DevCfg, const McTables*, const float*, uint32_t, ChunkCounts*, uint2*);
    typedef void (*classify_fn_t)(DevCfg, const McTables*, const float*, uint32_t, ChunkCounts*);
    big_emit16_fn_t big_emit16_fn = nullptr;
    big_emit32_fn_t big_emit32_fn = nullptr;
    big_count_fn_t big_count_fn = nullptr;
    classify_fn_t classify_fn = nullptr;     // compile-time-sized classify (internal_size 12 / 10)
    int classify_spec_threads = 0, classify_spec_blocks_per_sm = 1;
    cudaStream_t copy_stream = nullptr;

    // pinned host staging
    BatchTotals* h_totals = nullptr;
    unsigned long long* h_guard = nullptr;
    std::vector<PinnedBlock> pool;

    // launch geometry / kernel selection
    typedef void (*noise_fn_t)(DevCfg, AxisTables, const uint8_t*, const int32_t*, uint32_t, float*, unsigned long long*);
    typedef void (*emit16_fn_t)(DevCfg, const McTables*, const float*, const uw_chunk_desc*, const uint32_t*, const BatchTotals*, uw_vert*, uint16_t*, uw_tri*, uint16_t*);
    typedef void (*emit32_fn_t)(DevCfg, const McTables*, const float*, const uw_chunk_desc*, const uint32_t*, const BatchTotals*, uw_vert*, uint32_t*, uw_tri*, uint16_t*);
    typedef void (*fused16_fn_t)(DevCfg, AxisTables, const uint8_t*, const McTables*, const int32_t*, uint32_t, ScanSlot*,
                                 FusedControl*, FusedControl*, uw_chunk_desc*, uw_vert*, uint16_t*, unsigned long long,
                                 unsigned long long, float*, int, uw_tri*, uint16_t*, uint4*, int, int, unsigned long long, int, FusedSummary*, FusedOut);
    typedef void (*fused32_fn_t)(DevCfg, AxisTables, const uint8_t*, const McTables*, const int32_t*, uint32_t, ScanSlot*,
                                 FusedControl*, FusedControl*, uw_chunk_desc*, uw_vert*, uint32_t*, unsigned long long,
                                 unsigned long long, float*, int, uw_tri*, uint16_t*, uint4*, int, int, unsigned long long, int, FusedSummary*, FusedOut);
    fused16_fn_t fused16_fn = nullptr, fused16_peer_fn = nullptr;     // *_peer_fn: staged 16-byte stores (another GPU's memory)
    fused32_fn_t fused32_fn = nullptr, fused32_peer_fn = nullptr;
    bool use_fused = false;
    int z_lo = 1, z_hi = 0;         // chunk z layers that can hold surface (empty range = unknown: request-order hand-out)
    unsigned long long zcls = 0;    // hand-out class of layer z_lo + i in bits 4i..4i+3 (0 = most likely to hold surface)
    size_t order_cap = 0;           // largest batch that is handed out in cost order
    bool big_path = false;          // internal_size > 15: slab-walking extraction, densities in HBM
    size_t big_smem = 0; int big_blocks_per_sm = 1, big_count_blocks_per_sm = 1;
    bool ordered = false;           // packed arenas follow request order (look-back) vs atomic bump allocation
    size_t fused_smem = 0; int fused_blocks_per_sm = 1;
    FusedControl* d_ctl = nullptr;  // two blocks, alternating per launch (each launch zeroes the other one)
    int ctl_parity = 0;
    noise_fn_t noise_fn = nullptr;
    emit16_fn_t emit16_fn = nullptr;
    emit32_fn_t emit32_fn = nullptr;
    bool spec_noise = false;        // compile-time specialised noise kernel in use
    int noise_threads = 192, fused_threads = 224, noise_blocks_per_sm = 1; size_t noise_smem = 0;
    int emit_blocks_per_sm = 1; size_t emit_smem = 0;
    int classify_blocks_per_sm = 1;

    // multi-GPU gather (uw_gather_*): the arena this context OWNS as the rendering side ...
    struct GatherArena {
        bool alive = false;
        uw_gather_info info = {};
        char* base = nullptr;
        GatherHead* h_head = nullptr;           // pinned
        uint32_t* d_status = nullptr; uint32_t* h_status = nullptr;
        uw_chunk_desc* h_descs = nullptr; size_t h_descs_cap = 0;   // pinned, UW_GATHER_DESCS_TO_HOST
        uw_chunk_desc* h_draw = nullptr; size_t h_draw_cap = 0;     // pinned, UW_GATHER_DRAW_TO_HOST
        uint32_t wait_epoch = 0;
    } arena;
    // ... and the segment this context WRITES as a producer (local, peer or IPC-mapped addresses)
    struct GatherTarget {
        bool active = false, ipc = false;
        char* base = nullptr;                   // mapping of the arena allocation in this process
        uw_gather_info info = {};
        uint32_t segment = 0, epoch = 0;
        uw_chunk_desc* descs = nullptr; uw_vert* verts = nullptr; char* inds = nullptr; GatherHead* head = nullptr;
        uw_chunk_desc* draw = nullptr;
    } gt;
    bool force_staged = false;          // UW_STAGED_STORES=1 in the environment: staged stores for local arenas too (A/B measurements)
    bool force_direct = false;          // UW_STAGED_STORES=0: register stores even into another GPU's memory (A/B measurements)
    uint64_t gather_first_chunk = 0;    // request index of the next gather build's first chunk
    bool gather_build = false;          // the build being enqueued writes into the attached segment

    // grow-only device scratch shared by the point-query style entry points (uw_iso_at, uw_raycast_tris,
    // uw_debug_vertex_colors): no cudaMalloc / cudaFree per call
    char* d_scratch = nullptr; size_t d_scratch_cap = 0;
    // uw_raycast_tris: chunk position -> chunk index of the last build
    uint32_t* d_ray_table = nullptr; uint32_t ray_table_cap = 0, ray_table_mask = 0;
    uint64_t build_serial = 0, ray_table_serial = ~0ull; int ray_table_set = -1;

    bool profiling = false;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    uw_stage_times times;
    uint32_t launches = 0;
    unsigned long long guard_total = 0;
    std::string err;
};

struct uw_batch {
    uw_ctx* ctx;
    uint32_t n;
    int set;            // buffer set the batch's kernels write (-1: empty batch)
    bool ready;
    bool copying;       // the D2H copies into `arena` have been issued on the copy stream (views are set)
    PinnedBlock arena;
    uw_batch_view view;
};

static uw_status fail(uw_ctx* c, uw_status s, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return s;
}

#define CU_TRY(ctx, call)                                                                        \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            char b_[512];                                                                        \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? UW_ERR_OOM : UW_ERR_CUDA, b_);    \
        }                                                                                        \
    } while (0)

// ---------------------------------------------------------------------------------------
// host restatements needed to configure the device path
// ---------------------------------------------------------------------------------------
// noise::Perlin::new(seed) permutation table (noise-0.8.2 + rand-0.7.3 + rand_xorshift;
// SURVEY App. A.1): XorShift128 {x=1,y=z=w=seed}, Fisher-Yates from i=255 down to 1 with the
// widening-multiply rejection sampler over u32.
static void make_perm_table(uint32_t seed, uint8_t out[256]) {
    uint32_t s[4] = {1u, seed, seed, seed};
    for (int i = 0; i < 256; ++i) out[i] = (uint8_t)i;
    for (uint32_t n = 256; n > 1; --n) {
        const uint32_t zone = (n << __builtin_clz(n)) - 1u;
        uint64_t wide;
        do {
            const uint32_t t = s[0] ^ (s[0] << 11);
            s[0] = s[1]; s[1] = s[2]; s[2] = s[3];
            s[3] = s[3] ^ (s[3] >> 19) ^ t ^ (t >> 8);
            wide = (uint64_t)s[3] * n;
        } while ((uint32_t)wide > zone);
        const uint32_t j = (uint32_t)(wide >> 32), i = n - 1;
        const uint8_t tmp = out[i]; out[i] = out[j]; out[j] = tmp;
    }
}

static double fade_f64(double t) {
    double c = t < 0.0 ? 0.0 : t;
    c = c > 1.0 ? 1.0 : c;
    return (c * c * c) * (c * (c * 6.0 + (-15.0)) + 10.0);
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

static uw_status setup_tables(uw_ctx* c) {
    const uw_config& cf = c->cfg;
    DevCfg& d = c->dcfg;
    memset(&d, 0, sizeof d);
    d.S = cf.internal_size; d.L = d.S + 1; d.L2 = d.L * d.L; d.L3 = d.L2 * d.L;
    d.chunk_size = cf.chunk_size; d.octaves = (int)cf.octaves;
    d.iso_level = cf.iso_level; d.max_height = cf.max_height; d.adj_z_mod = cf.adj_z_mod;
    d.size_scale = (float)cf.chunk_size / (float)cf.internal_size;            // chunk.rs:7 (f32 division)
    d.min_hue = cf.min_hue; d.max_hue = cf.max_hue; d.min_z = cf.min_z; d.max_z = cf.max_z;
    d.guard_eps = cf.guard_eps > 0.f ? cf.guard_eps : 1e-5f;
    d.dens_stride = (uint32_t)((d.L3 + 3) & ~3);
    d.cs_pow2 = is_pow2(cf.chunk_size) ? 1 : 0;
    d.inv_chunk_size = 1.0 / (double)cf.chunk_size;
    {
        int ex = 0;
        const float m = frexpf(fabsf(cf.adj_z_mod), &ex);
        d.mod_pow2 = (m == 0.5f && cf.adj_z_mod > 0.0f) ? 1 : 0;
        d.inv_adj_z_mod = 1.0f / cf.adj_z_mod;
    }
    for (int vi = 0; vi < 3; ++vi) {                                          // chunk.rs:219-221, util.rs:129-132,106-112
        const float value = cf.base_value + (float)vi / 9.0f;
        const float cc = value * cf.saturation;
        const float m = value - cc;
        d.hsv_c[vi] = cc; d.hsv_m[vi] = m;
        d.srgb_hi[vi] = powf((((cc + m) * 255.0f) / 255.0f + 0.055f) / 1.055f, 2.4f);
        d.srgb_lo[vi] = powf((((0.0f + m) * 255.0f) / 255.0f + 0.055f) / 1.055f, 2.4f);
    }
    d.lat_base[0] = 0; d.x_base[0] = 0;
    for (int o = 0; o < UW_MAX_OCT; ++o) {
        d.G[o] = (1 << o) + 2;
        d.lat_base[o + 1] = d.lat_base[o] + d.G[o] * d.G[o] * d.G[o];
        d.x_base[o + 1] = d.x_base[o] + d.L * d.G[o] * d.G[o];
    }
    // per-axis tables in f64, reference operation order (chunk.rs:107-108, perlin_util.rs:13)
    memset(&c->tab, 0, sizeof c->tab);
    c->ax_d.assign((size_t)d.octaves * d.L, 0.f); c->ax_d1 = c->ax_d; c->ax_w = c->ax_d;
    c->ax_c.assign((size_t)d.octaves * d.L, 0);
    for (int o = 0; o < d.octaves; ++o) {
        const double F = (double)(1 << o);
        for (int i = 0; i < d.L; ++i) {
            const double local = (double)i * (double)d.size_scale;
            const double u = (local + 0.0) / (double)cf.chunk_size;
            const double p = u * F;
            const double f = floor(p);
            const double dd = p - f;
            if (f < 0.0 || f > F) return fail(c, UW_ERR_INVALID, "axis table: lattice cell out of range");
            const size_t k = (size_t)o * d.L + i;
            c->ax_c[k] = (int)f; c->ax_d[k] = (float)dd; c->ax_d1[k] = (float)(dd + (-1.0)); c->ax_w[k] = (float)fade_f64(dd);
            if (d.L <= UW_AXIS_PAD) {
                c->tab.c[o][i] = c->ax_c[k]; c->tab.d[o][i] = c->ax_d[k]; c->tab.d1[o][i] = c->ax_d1[k]; c->tab.w[o][i] = c->ax_w[k];
            }
        }
    }
    // z layers that can hold surface: iso = terrace(z) + p with |p| <= 1 (every octave is clamped to [-1, 1]);
    // used only to hand out the expensive chunks first (take_ticket)
    {
        const float margin = 1e-3f;
        int lo = 1, hi = 0;
        bool found = false;
        for (int pz = -4096; pz <= 4096; ++pz) {
            float tmin = 3e38f, tmax = -3e38f;
            for (int k = 0; k < d.L; ++k) {
                const double local = (double)k * (double)d.size_scale;
                const float zf = (float)((local + (double)(pz * cf.chunk_size)) / (double)cf.chunk_size);
                const float adj = (zf * (float)cf.chunk_size) / cf.max_height;
                const float t = adj - fmodf(adj, cf.adj_z_mod);
                tmin = fminf(tmin, t); tmax = fmaxf(tmax, t);
            }
            const bool blank_certain = tmin - 1.0f > cf.iso_level + margin;
            const bool solid_certain = tmax + 1.0f < cf.iso_level - margin;
            if (!blank_certain && !solid_certain) { if (!found) { lo = pz; found = true; } hi = pz; }
        }
        // "blank above, solid below" (answer_trivial) holds only if the terrace term grows with z: positive
        // max_height and adj_z_mod.  Anything else keeps the range empty = unknown: no analytic skip, request-order hand-out.
        const bool rising = cf.max_height > 0.0f && cf.adj_z_mod > 0.0f;
        if (rising && found && lo > -4096 && hi < 4096) { c->z_lo = lo; c->z_hi = hi; } else { c->z_lo = 1; c->z_hi = 0; }
        // Rank those layers by how likely they are to hold surface: the noise term is roughly N(0, 0.25), so a
        // lattice level z contributes exp(-((iso - terrace(z)) / 0.25)^2 / 2).  Scheduling only (take_ticket).
        c->zcls = 0;
        const int nl = c->z_hi - c->z_lo + 1;
        if (nl > 0) {
            std::vector<std::pair<double, int>> w;
            for (int pz = c->z_lo; pz <= c->z_hi && pz < c->z_lo + 16; ++pz) {
                double acc = 0.0;
                for (int k = 0; k < d.L; ++k) {
                    const double local = (double)k * (double)d.size_scale;
                    const float zf = (float)((local + (double)(pz * cf.chunk_size)) / (double)cf.chunk_size);
                    const float adj = (zf * (float)cf.chunk_size) / cf.max_height;
                    const float t = adj - fmodf(adj, cf.adj_z_mod);
                    const double u = ((double)cf.iso_level - (double)t) / 0.25;
                    acc += exp(-0.5 * u * u);
                }
                w.push_back({-acc, pz - c->z_lo});
            }
            std::sort(w.begin(), w.end());
            for (size_t r = 0; r < w.size(); ++r) {
                const unsigned long long cls = r < (size_t)(UW_NCLS - 2) ? r : (size_t)(UW_NCLS - 2);
                c->zcls |= cls << (4 * w[r].second);
            }
        }
    }
    return UW_OK;
}

// The specialised kernels bake the axis tables in at compile time (SpecDims): usable only when this
// configuration's runtime f64 tables match them bit for bit.
template <class DD>
static bool spec_tables_match(const uw_ctx* c, DD) {
    const DevCfg& d = c->dcfg;
    if (d.S != DD::S || d.octaves != 3 || !c->fast_path) return false;
    bool ok = true;
    for (int o = 0; o < 3; ++o)
        for (int i = 0; i < DD::L; ++i) {
            const size_t k = (size_t)o * d.L + i;
            const float dd = DD::tab_d(o, i), ww = DD::tab_w(o, i);
            ok &= c->ax_c[k] == DD::cell(o, i) && c->ax_c[k] == DD::tab_c(o, i);
            ok &= memcmp(&dd, &c->ax_d[k], 4) == 0 && memcmp(&ww, &c->ax_w[k], 4) == 0;
        }
    return ok;
}

// ---------------------------------------------------------------------------------------
// lifecycle
// ---------------------------------------------------------------------------------------
extern "C" uint32_t uw_abi_version(void) { return UW_ABI_VERSION; }

extern "C" void uw_config_default(uw_config* c) {
    if (!c) return;
    memset(c, 0, sizeof *c);
    c->internal_size = 12; c->chunk_size = 16; c->octaves = 3;
    c->iso_level = -0.1f; c->max_height = 32.0f; c->adj_z_mod = 0.25f;
    c->min_hue = -150.0f; c->max_hue = 60.0f; c->saturation = 0.6f; c->base_value = 0.4f;
    c->min_z = -2.0f; c->max_z = 2.0f;
    c->seed = 0; c->device = -1; c->flags = 0; c->guard_eps = 0.0f;
}

extern "C" const char* uw_last_error(const uw_ctx* ctx) {
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

extern "C" uw_status uw_gather_detach(uw_ctx* c);
extern "C" uw_status uw_gather_destroy(uw_ctx* c);

extern "C" void uw_destroy(uw_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    uw_gather_detach(c);
    uw_gather_destroy(c);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    cudaFree(c->d_perm); cudaFree(c->d_mc); cudaFree(c->d_axis); cudaFree(c->d_totals); cudaFree(c->d_guard); cudaFree(c->d_ctl);
    cudaFree(c->d_scan_part); cudaFree(c->d_scan_flag); cudaFree(c->d_scan_ctl);
    cudaFree(c->d_scratch); cudaFree(c->d_ray_table); cudaFree(c->d_terr_tab);
    for (auto& b : c->sets) {
        cudaFree(b.d_pos); cudaFree(b.d_dens); cudaFree(b.d_counts); cudaFree(b.d_quarters); cudaFree(b.d_descs); cudaFree(b.d_active);
        cudaFree(b.d_cases);
        if (c->exportable) { vmm_free(&b.vmm_verts); vmm_free(&b.vmm_inds); }
        else { cudaFree(b.d_verts); cudaFree(b.d_inds); } cudaFree(b.d_tris); cudaFree(b.d_tri_cell);
        cudaFree(b.d_scan); cudaFree(b.d_order);
        if (b.h_pos) cudaFreeHost(b.h_pos);
        if (b.h_sum) cudaFreeHost(b.h_sum);
        if (b.done) cudaEventDestroy(b.done);
        if (b.copied) cudaEventDestroy(b.copied);
    }
    if (c->h_totals) cudaFreeHost(c->h_totals);
    if (c->h_guard) cudaFreeHost(c->h_guard);
    for (auto& b : c->pool) cudaFreeHost(b.ptr);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" uw_status uw_create(const uw_config* cfg, uw_ctx** out) {
    if (!cfg || !out) return fail(nullptr, UW_ERR_INVALID, "uw_create: null argument");
    *out = nullptr;
    if (cfg->internal_size < 1 || cfg->internal_size > 64)
        return fail(nullptr, UW_ERR_UNSUPPORTED, "uw_create: internal_size must be in 1..64");
    if (cfg->octaves < 1 || cfg->octaves > UW_MAX_OCT) return fail(nullptr, UW_ERR_INVALID, "uw_create: octaves must be 1..4");
    if (cfg->chunk_size < 1) return fail(nullptr, UW_ERR_INVALID, "uw_create: chunk_size must be positive");
    if (!(cfg->max_height != 0.0f) || !(cfg->adj_z_mod != 0.0f) || !(cfg->max_z != cfg->min_z))
        return fail(nullptr, UW_ERR_INVALID, "uw_create: max_height, adj_z_mod and (max_z - min_z) must be non-zero");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, UW_ERR_NO_DEVICE, std::string("uw_create: no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    int dev = cfg->device;
    if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) dev = 0; }
    if (dev >= ndev) return fail(nullptr, UW_ERR_NO_DEVICE, "uw_create: device ordinal out of range");

    uw_ctx* c = new uw_ctx();
    c->cfg = *cfg; c->device = dev;
    memset(&c->times, 0, sizeof c->times);
    auto bail = [&](uw_status s) { g_create_error = c->err; uw_destroy(c); return s; };
    if (cudaSetDevice(dev) != cudaSuccess) { c->err = "cudaSetDevice failed"; return bail(UW_ERR_CUDA); }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { c->err = "cudaGetDeviceProperties failed"; return bail(UW_ERR_CUDA); }
    if (prop.major != 10) {
        char b[256]; snprintf(b, sizeof b, "uw_create: device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);
        c->err = b; return bail(UW_ERR_NO_DEVICE);
    }
    c->num_sms = prop.multiProcessorCount;
    { const char* e = getenv("UW_STAGED_STORES"); c->force_staged = e && e[0] == '1'; c->force_direct = e && e[0] == '0'; }
    c->host_ptr_ok = prop.unifiedAddressing && prop.canMapHostMemory;
    c->index32 = (cfg->flags & UW_FLAG_INDEX32) != 0 || cfg->internal_size > 22;
    // FP32 factorisation needs chunk-independent fractional parts: chunk_size a power of two
    c->big_path = cfg->internal_size > UW_SMALL_MAX_L - 1;
    c->tris = (cfg->flags & UW_FLAG_TRIS) != 0;
    c->exportable = (cfg->flags & UW_FLAG_EXPORTABLE) != 0;
    if (c->tris && c->big_path) { c->err = "uw_create: UW_FLAG_TRIS is not available for internal_size > 15"; return bail(UW_ERR_UNSUPPORTED); }
    // FP32 factorised noise needs chunk-independent fractional parts: chunk_size a power of two.  Large chunks
    // have an FP32 kernel for internal_size 64 (BASELINE config 4); other large sizes use the exact f64 kernel.
    c->fast_path = !(cfg->flags & UW_FLAG_EXACT_F64) && is_pow2(cfg->chunk_size);

    uw_status st = setup_tables(c);
    if (st != UW_OK) return bail(st);
    make_perm_table(cfg->seed, c->perm);

    auto cu = [&](cudaError_t r, const char* what) {
        if (r != cudaSuccess) { c->err = std::string(what) + ": " + cudaGetErrorString(r); return false; }
        return true;
    };
    if (!cu(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "cudaStreamCreate")) return bail(UW_ERR_CUDA);
    c->own_stream = true;
    if (!cu(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking), "cudaStreamCreate copy")) return bail(UW_ERR_CUDA);
    for (auto& b : c->sets) {
        if (!cu(cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming), "cudaEventCreate done")) return bail(UW_ERR_CUDA);
        if (!cu(cudaEventCreateWithFlags(&b.copied, cudaEventDisableTiming), "cudaEventCreate copied")) return bail(UW_ERR_CUDA);
        if (!cu(cudaHostAlloc(&b.h_sum, sizeof(FusedSummary), cudaHostAllocMapped | cudaHostAllocPortable), "cudaHostAlloc summary")) return bail(UW_ERR_OOM);
    }
    if (!cu(cudaMalloc(&c->d_ctl, 2 * sizeof(FusedControl)), "cudaMalloc control")) return bail(UW_ERR_OOM);
    if (!cu(cudaMemset(c->d_ctl, 0, 2 * sizeof(FusedControl)), "memset control")) return bail(UW_ERR_CUDA);
    if (!cu(cudaMalloc(&c->d_perm, 256), "cudaMalloc perm")) return bail(UW_ERR_OOM);
    if (!cu(cudaMemcpy(c->d_perm, c->perm, 256, cudaMemcpyHostToDevice), "memcpy perm")) return bail(UW_ERR_CUDA);
    {
        static const uint64_t rows[256] = UW_MC_ROWS_INIT;
        static const uint8_t ninds[256] = UW_MC_NINDS_INIT;
        static const uint16_t crossed[256] = UW_MC_CROSSED_INIT;
        static const uint16_t before[256][12] = UW_MC_BEFORE_INIT;
        McTables* h = new McTables();
        memcpy(h->rows, rows, sizeof rows); memcpy(h->ninds, ninds, sizeof ninds);
        memcpy(h->crossed, crossed, sizeof crossed); memcpy(h->before, before, sizeof before);
        for (int nat = 0; nat < 256; ++nat) {
            // natural pattern bits (m00 z, m00 z+1, m10 z, m10 z+1, m01 z, m01 z+1, m11 z, m11 z+1) are the
            // corners (0, 3, 1, 2, 4, 7, 5, 6) of chunk.rs:144-153
            static const int corner_of_bit[8] = {0, 3, 1, 2, 4, 7, 5, 6};
            unsigned cs = 0;
            for (int b = 0; b < 8; ++b) if ((nat >> b) & 1) cs |= 1u << corner_of_bit[b];
            h->lut[nat] = cs | ((uint32_t)ninds[cs] << 8) | ((uint32_t)crossed[cs] << 12);
        }
        for (int i = 0; i < 16; ++i) {                       // pow24_tab (uw_kernels.cuh)
            const float ic = (float)(1.0 / (1.0 + (i + 0.5) / 16.0));
            const double l2 = -log2((double)ic);               // log2 of the c_i that 1/c_i = ic stands for
            const double hi = nearbyint(l2 * 65536.0) / 65536.0;
            h->powtab[i] = ic; h->powtab[16 + i] = (float)hi; h->powtab[32 + i] = (float)(l2 - hi);
        }
        bool ok = cu(cudaMalloc(&c->d_mc, sizeof(McTables)), "cudaMalloc mc") &&
                  cu(cudaMemcpy(c->d_mc, h, sizeof(McTables), cudaMemcpyHostToDevice), "memcpy mc");
        delete h;
        if (!ok) return bail(UW_ERR_CUDA);
    }
    if (!cu(cudaMalloc(&c->d_totals, sizeof(BatchTotals)), "cudaMalloc totals")) return bail(UW_ERR_OOM);
    if (!cu(cudaMalloc(&c->d_guard, sizeof(unsigned long long)), "cudaMalloc guard")) return bail(UW_ERR_OOM);
    if (!cu(cudaMemset(c->d_guard, 0, sizeof(unsigned long long)), "memset guard")) return bail(UW_ERR_CUDA);
    if (!cu(cudaHostAlloc(&c->h_totals, sizeof(BatchTotals), cudaHostAllocDefault), "cudaHostAlloc totals")) return bail(UW_ERR_OOM);
    if (!cu(cudaHostAlloc(&c->h_guard, sizeof(unsigned long long), cudaHostAllocDefault), "cudaHostAlloc guard")) return bail(UW_ERR_OOM);
    for (auto& ev : c->ev) if (!cu(cudaEventCreate(&ev), "cudaEventCreate")) return bail(UW_ERR_CUDA);

    // terrace terms of the z layers around the origin, tabulated by the device function the kernels would otherwise
    // evaluate per chunk (f64 coordinate, IEEE division, fmod): bit-identical by construction
    if (c->dcfg.L <= 16) {
        const int z0 = -256, nz = 512;
        if (!cu(cudaMalloc(&c->d_terr_tab, (size_t)nz * 16 * sizeof(float)), "cudaMalloc terrace table")) return bail(UW_ERR_OOM);
        k_terrace_table<<<32, 256, 0, c->stream>>>(c->dcfg, c->d_terr_tab, z0, nz);
        if (!cu(cudaGetLastError(), "k_terrace_table") || !cu(cudaStreamSynchronize(c->stream), "k_terrace_table")) return bail(UW_ERR_CUDA);
        c->dcfg.terr_tab = c->d_terr_tab; c->dcfg.terr_z0 = z0; c->dcfg.terr_nz = nz;
    }

    // launch geometry / kernel selection
    const DevCfg& d = c->dcfg;
    if (c->big_path) {
        if (spec_tables_match(c, SpecDims<64, 3>())) {
            std::vector<float4> h((size_t)3 * d.L);
            for (size_t k = 0; k < h.size(); ++k) h[k] = make_float4(c->ax_d[k], c->ax_d1[k], c->ax_w[k], 0.f);
            c->big_noise_smem = sizeof(BigNoiseSmem<64, 3>);
            bool okn = cu(cudaMalloc(&c->d_axis, h.size() * sizeof(float4)), "cudaMalloc axis") &&
                       cu(cudaMemcpy(c->d_axis, h.data(), h.size() * sizeof(float4), cudaMemcpyHostToDevice), "memcpy axis") &&
                       cu(cudaFuncSetAttribute((const void*)k_noise_big<64, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->big_noise_smem), "attr noise big");
            if (!okn) return bail(UW_ERR_CUDA);
            int nbn = 1;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nbn, (const void*)k_noise_big<64, 3>, 256, c->big_noise_smem) == cudaSuccess && nbn > 0)
                c->big_noise_blocks_per_sm = nbn;
            c->big_fast_noise = true;
        } else {
            c->fast_path = false;
        }
        c->big_smem = big_smem_bytes(d);
        auto set_attr = [&](const void* f) { return cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->big_smem); };
        if (d.S == 64) { c->big_emit16_fn = k_emit_big<uint16_t, 64>; c->big_emit32_fn = k_emit_big<uint32_t, 64>; c->big_count_fn = k_count_big<64>; }
        else           { c->big_emit16_fn = k_emit_big<uint16_t, 0>;  c->big_emit32_fn = k_emit_big<uint32_t, 0>;  c->big_count_fn = k_count_big<0>; }
        bool ok = cu(set_attr((const void*)c->big_emit16_fn), "attr big emit16") &&
                  cu(set_attr((const void*)c->big_emit32_fn), "attr big emit32");
        if (!ok) return bail(UW_ERR_CUDA);
        int nb = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)c->big_emit32_fn, UW_BIG_NT, c->big_smem) == cudaSuccess && nb > 0)
            c->big_blocks_per_sm = nb;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)c->big_count_fn, UW_BIG_NT, 0) == cudaSuccess && nb > 0)
            c->big_count_blocks_per_sm = nb;
    } else {
        // the specialised kernels bake the axis tables in at compile time (see SpecDims): usable only
        // when the runtime tables (from this configuration) match them bit for bit
        auto spec_ok = [&](auto dims) { return spec_tables_match(c, dims); };
        c->noise_threads = ((d.L2 + 31) / 32) * 32;
        c->noise_smem = noise_smem_bytes(d);
        c->noise_fn = k_noise_small<0, 0>;
        if (spec_ok(SpecDims<12, 3>())) {
            c->noise_fn = k_noise_spec<12, 3>; c->spec_noise = true;
            c->noise_threads = SpecDims<12, 3>::NT; c->fused_threads = SpecDims<12, 3>::NTF; c->noise_smem = sizeof(SpecSmem<12, 3>);
            c->fused16_fn = k_build_fused<12, 3, uint16_t, false>; c->fused32_fn = k_build_fused<12, 3, uint32_t, false>;
            c->fused16_peer_fn = k_build_fused<12, 3, uint16_t, true>; c->fused32_peer_fn = k_build_fused<12, 3, uint32_t, true>;
            c->fused_smem = sizeof(FusedSmem<12, 3>);
        } else if (spec_ok(SpecDims<10, 3>())) {
            c->noise_fn = k_noise_spec<10, 3>; c->spec_noise = true;
            c->noise_threads = SpecDims<10, 3>::NT; c->fused_threads = SpecDims<10, 3>::NTF; c->noise_smem = sizeof(SpecSmem<10, 3>);
            c->fused16_fn = k_build_fused<10, 3, uint16_t, false>; c->fused32_fn = k_build_fused<10, 3, uint32_t, false>;
            c->fused16_peer_fn = k_build_fused<10, 3, uint16_t, true>; c->fused32_peer_fn = k_build_fused<10, 3, uint32_t, true>;
            c->fused_smem = sizeof(FusedSmem<10, 3>);
        }
        c->use_fused = c->spec_noise && !(cfg->flags & UW_FLAG_STAGED);
        c->ordered = (cfg->flags & UW_FLAG_ORDERED) != 0;
        c->emit_smem = emit_smem_bytes(d);
        if (d.S == 12)      { c->emit16_fn = k_emit_small<12, uint16_t>; c->emit32_fn = k_emit_small<12, uint32_t>; }
        else if (d.S == 10) { c->emit16_fn = k_emit_small<10, uint16_t>; c->emit32_fn = k_emit_small<10, uint32_t>; }
        else                { c->emit16_fn = k_emit_small<0, uint16_t>;  c->emit32_fn = k_emit_small<0, uint32_t>; }
        auto set_attr = [&](const void* f, size_t smem) {
            return cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        };
        bool ok = cu(set_attr((const void*)c->noise_fn, c->noise_smem), "attr noise") &&
                  cu(set_attr((const void*)c->emit16_fn, c->emit_smem), "attr emit16") &&
                  cu(set_attr((const void*)c->emit32_fn, c->emit_smem), "attr emit32");
        if (ok && c->fused16_fn)
            ok = cu(set_attr((const void*)c->fused16_fn, c->fused_smem), "attr fused16") &&
                 cu(set_attr((const void*)c->fused32_fn, c->fused_smem), "attr fused32") &&
                 cu(set_attr((const void*)c->fused16_peer_fn, c->fused_smem), "attr fused16 peer") &&
                 cu(set_attr((const void*)c->fused32_peer_fn, c->fused_smem), "attr fused32 peer");
        if (!ok) return bail(UW_ERR_CUDA);
        int nb = 1;
        if (c->fused16_fn && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)c->fused16_fn, c->fused_threads, c->fused_smem) == cudaSuccess && nb > 0)
            c->fused_blocks_per_sm = nb;
        c->order_cap = (size_t)16 * c->num_sms * c->fused_blocks_per_sm;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)c->noise_fn, c->noise_threads, c->noise_smem) == cudaSuccess && nb > 0)
            c->noise_blocks_per_sm = nb;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)c->emit16_fn, 256, c->emit_smem) == cudaSuccess && nb > 0)
            c->emit_blocks_per_sm = nb;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)k_classify_small, 256, 0) == cudaSuccess && nb > 0)
            c->classify_blocks_per_sm = nb;
        if (d.S == 12)      { c->classify_fn = k_classify_spec<12>; c->classify_spec_threads = ClsDims<12>::NT; }
        else if (d.S == 10) { c->classify_fn = k_classify_spec<10>; c->classify_spec_threads = ClsDims<10>::NT; }
        if (c->classify_fn && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)c->classify_fn, c->classify_spec_threads, 0) == cudaSuccess && nb > 0)
            c->classify_spec_blocks_per_sm = nb;
    }
    *out = c;
    return UW_OK;
}

extern "C" uw_status uw_perm_table(const uw_ctx* ctx, uint8_t out[256]) {
    if (!ctx || !out) return UW_ERR_INVALID;
    memcpy(out, ctx->perm, 256);
    return UW_OK;
}

extern "C" uw_status uw_set_stream(uw_ctx* c, void* s) {
    if (!c) return UW_ERR_INVALID;
    CU_TRY(c, cudaSetDevice(c->device));
    if (c->stream) CU_TRY(c, cudaStreamSynchronize(c->stream));
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    c->stream = (cudaStream_t)s; c->own_stream = false;
    return UW_OK;
}

extern "C" uw_status uw_set_profiling(uw_ctx* c, int enabled) {
    if (!c) return UW_ERR_INVALID;
    c->profiling = enabled != 0;
    return UW_OK;
}

// ---------------------------------------------------------------------------------------
// exportable arenas (UW_FLAG_EXPORTABLE): CUDA virtual-memory-management allocations with a POSIX-fd handle
// type.  The driver entry points are fetched through the runtime, so the library does not link libcuda.
// ---------------------------------------------------------------------------------------
struct DriverApi {
    CUresult (*memCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*memRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*memAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*memAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*memMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*memUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*memSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*memGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*memExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    bool ok = false;
};

static const DriverApi& driver_api() {
    static DriverApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    auto get = [](const char* name, void** fn) {
        cudaDriverEntryPointQueryResult q;
        return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
    };
    api.ok = get("cuMemCreate", (void**)&api.memCreate) && get("cuMemRelease", (void**)&api.memRelease) &&
             get("cuMemAddressReserve", (void**)&api.memAddressReserve) && get("cuMemAddressFree", (void**)&api.memAddressFree) &&
             get("cuMemMap", (void**)&api.memMap) && get("cuMemUnmap", (void**)&api.memUnmap) &&
             get("cuMemSetAccess", (void**)&api.memSetAccess) &&
             get("cuMemGetAllocationGranularity", (void**)&api.memGetAllocationGranularity) &&
             get("cuMemExportToShareableHandle", (void**)&api.memExportToShareableHandle);
    return api;
}

static void vmm_free(VmmArena* a) {
    if (!a->ptr) return;
    const DriverApi& d = driver_api();
    d.memUnmap((CUdeviceptr)a->ptr, a->bytes);
    d.memAddressFree((CUdeviceptr)a->ptr, a->bytes);
    d.memRelease(a->handle);
    a->ptr = nullptr; a->bytes = 0; a->handle = 0;
}

static uw_status vmm_alloc(uw_ctx* c, VmmArena* a, size_t bytes) {
    const DriverApi& d = driver_api();
    if (!d.ok) return fail(c, UW_ERR_UNSUPPORTED, "UW_FLAG_EXPORTABLE: the CUDA driver does not expose the virtual memory management API");
    vmm_free(a);
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof prop);
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = c->device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t gran = 0;
    if (d.memGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0)
        return fail(c, UW_ERR_CUDA, "cuMemGetAllocationGranularity failed");
    const size_t size = (bytes + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle h = 0;
    CUresult r = d.memCreate(&h, size, &prop, 0);
    if (r != CUDA_SUCCESS) return fail(c, r == CUDA_ERROR_OUT_OF_MEMORY ? UW_ERR_OOM : UW_ERR_CUDA, "cuMemCreate failed (exportable arena)");
    CUdeviceptr p = 0;
    if (d.memAddressReserve(&p, size, 0, 0, 0) != CUDA_SUCCESS) { d.memRelease(h); return fail(c, UW_ERR_CUDA, "cuMemAddressReserve failed"); }
    if (d.memMap(p, size, 0, h, 0) != CUDA_SUCCESS) { d.memAddressFree(p, size); d.memRelease(h); return fail(c, UW_ERR_CUDA, "cuMemMap failed"); }
    CUmemAccessDesc acc;
    memset(&acc, 0, sizeof acc);
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = c->device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (d.memSetAccess(p, size, &acc, 1) != CUDA_SUCCESS) {
        d.memUnmap(p, size); d.memAddressFree(p, size); d.memRelease(h);
        return fail(c, UW_ERR_CUDA, "cuMemSetAccess failed");
    }
    a->ptr = (void*)p; a->bytes = size; a->handle = h;
    return UW_OK;
}

// ---------------------------------------------------------------------------------------
// buffers
// ---------------------------------------------------------------------------------------
template <typename T>
static cudaError_t regrow(T** p, size_t count) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    return cudaMalloc((void**)p, count * sizeof(T));
}

static uw_status ensure_chunks(uw_ctx* c, uint32_t n) {
    if (n <= c->B().cap_chunks) return UW_OK;
    uint32_t cap = c->B().cap_chunks ? c->B().cap_chunks : 256;
    while (cap < n) cap *= 2;
    if (cap > n && (uint64_t)cap * c->dcfg.dens_stride * 4 > (8ull << 30)) cap = n;   // no 2x slack on multi-GB batches
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    CU_TRY(c, regrow(&c->B().d_pos, (size_t)cap * 3));
    CU_TRY(c, regrow(&c->B().d_counts, cap));
    if (c->big_path) CU_TRY(c, regrow(&c->B().d_quarters, (size_t)cap * 4));
    CU_TRY(c, regrow(&c->B().d_descs, cap));
    CU_TRY(c, regrow(&c->B().d_active, cap));
    CU_TRY(c, regrow(&c->B().d_scan, cap));
    if (c->use_fused && !c->B().d_order)
        CU_TRY(c, cudaMalloc(&c->B().d_order, (size_t)UW_NCLS * c->order_cap * sizeof(uint4)));
    if (c->tris) CU_TRY(c, regrow(&c->B().d_tri_cell, (size_t)cap * ((size_t)c->dcfg.S * c->dcfg.S * c->dcfg.S + 1)));
    c->B().cap_chunks = cap;
    return UW_OK;
}

// The density field lives in HBM only on the staged / large-chunk / debug paths (the fused kernel keeps it in
// shared memory): allocated on first use, sized like the chunk arrays.
static uw_status ensure_dens(uw_ctx* c) {
    if (c->B().cap_dens_chunks >= c->B().cap_chunks) return UW_OK;
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    CU_TRY(c, regrow(&c->B().d_dens, (size_t)c->B().cap_chunks * c->dcfg.dens_stride));
    c->B().cap_dens_chunks = c->B().cap_chunks;
    return UW_OK;
}

static uw_status ensure_outputs(uw_ctx* c, unsigned long long nv, unsigned long long ni) {
    const size_t isz = c->index32 ? 4 : 2;
    if (nv > c->B().vcap) {
        unsigned long long cap = c->B().vcap ? c->B().vcap : 4096;
        while (cap < nv) cap *= 2;
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        if (c->exportable) {
            CU_TRY(c, cudaDeviceSynchronize());
            uw_status vst = vmm_alloc(c, &c->B().vmm_verts, (size_t)cap * sizeof(uw_vert));
            if (vst != UW_OK) return vst;
            c->B().d_verts = (uw_vert*)c->B().vmm_verts.ptr;
        } else {
            CU_TRY(c, regrow(&c->B().d_verts, (size_t)cap));
        }
        c->B().vcap = cap;
    }
    if (ni > c->B().icap) {
        unsigned long long cap = c->B().icap ? c->B().icap : 16384;
        while (cap < ni) cap *= 2;
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        if (c->exportable) {
            CU_TRY(c, cudaDeviceSynchronize());
            uw_status ist = vmm_alloc(c, &c->B().vmm_inds, (size_t)cap * isz);
            if (ist != UW_OK) return ist;
            c->B().d_inds = c->B().vmm_inds.ptr;
        } else {
            if (c->B().d_inds) { cudaFree(c->B().d_inds); c->B().d_inds = nullptr; }
            CU_TRY(c, cudaMalloc(&c->B().d_inds, (size_t)cap * isz));
        }
        if (c->tris) CU_TRY(c, regrow(&c->B().d_tris, (size_t)cap / 3 + 1));
        c->B().icap = cap;
    }
    return UW_OK;
}

static uw_status pinned_get(uw_ctx* c, size_t bytes, PinnedBlock* out) {
    size_t best = (size_t)-1;
    for (size_t i = 0; i < c->pool.size(); ++i)
        if (c->pool[i].bytes >= bytes && (best == (size_t)-1 || c->pool[i].bytes < c->pool[best].bytes)) best = i;
    if (best != (size_t)-1) { *out = c->pool[best]; c->pool.erase(c->pool.begin() + best); return UW_OK; }
    size_t cap = 1 << 16;
    while (cap < bytes) cap *= 2;
    void* p = nullptr;
    CU_TRY(c, cudaHostAlloc(&p, cap, cudaHostAllocDefault));
    out->ptr = p; out->bytes = cap;
    return UW_OK;
}

// ---------------------------------------------------------------------------------------
// launch sequence
// ---------------------------------------------------------------------------------------
static int persistent_grid(const uw_ctx* c, uint32_t n, int blocks_per_sm) {
    const long long full = (long long)c->num_sms * blocks_per_sm;
    return (int)((long long)n < full ? (long long)n : full);
}

static uw_status launch_noise(uw_ctx* c, const int32_t* d_pos, uint32_t n) {
    const DevCfg& d = c->dcfg;
    if (c->big_path && c->big_fast_noise && c->fast_path) {
        const unsigned long long units = (unsigned long long)n * ((d.L2 + 255) / 256);
        const unsigned long long full = (unsigned long long)c->num_sms * c->big_noise_blocks_per_sm;
        k_noise_big<64, 3><<<(int)(units < full ? units : full), 256, c->big_noise_smem, c->stream>>>(d, c->d_axis, c->d_perm, d_pos, n, c->B().d_dens, c->d_guard);
    } else if (c->fast_path && !c->big_path) {
        const int grid = persistent_grid(c, n, c->noise_blocks_per_sm);
        c->noise_fn<<<grid, c->noise_threads, c->noise_smem, c->stream>>>(d, c->tab, c->d_perm, d_pos, n, c->B().d_dens, c->d_guard);
    } else {
        const unsigned long long total = (unsigned long long)n * d.L3;
        unsigned long long blocks = (total + 255) / 256;
        const unsigned long long maxb = (unsigned long long)c->num_sms * 8;
        if (blocks > maxb) blocks = maxb;
        k_noise_exact<<<(int)blocks, 256, 0, c->stream>>>(d, c->d_perm, d_pos, n, c->B().d_dens);
    }
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    return UW_OK;
}

static uw_status launch_scan(uw_ctx* c, const int32_t* d_pos, uint32_t n) {
    const uint32_t tiles = (n + 1023u) / 1024u;
    if (tiles > c->scan_tiles_cap) {
        uint32_t cap = c->scan_tiles_cap ? c->scan_tiles_cap : 64;
        while (cap < tiles) cap *= 2;
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        CU_TRY(c, regrow(&c->d_scan_part, cap));
        CU_TRY(c, regrow(&c->d_scan_flag, cap));
        CU_TRY(c, cudaMemset(c->d_scan_flag, 0, (size_t)cap * sizeof(uint32_t)));
        if (!c->d_scan_ctl) {
            CU_TRY(c, cudaMalloc(&c->d_scan_ctl, sizeof(ScanCtl)));
            CU_TRY(c, cudaMemset(c->d_scan_ctl, 0, sizeof(ScanCtl)));
        }
        c->scan_tiles_cap = cap;
        c->scan_epoch = 0;
    }
    if (++c->scan_epoch == 0) {                       // epoch wrapped: flags of 2^32 launches ago would alias
        CU_TRY(c, cudaMemsetAsync(c->d_scan_flag, 0, (size_t)c->scan_tiles_cap * sizeof(uint32_t), c->stream));
        c->scan_epoch = 1;
    }
    k_scan_chunks<<<tiles, 1024, 0, c->stream>>>(c->B().d_counts, d_pos, n, c->B().d_descs, c->B().d_active, c->d_totals,
                                                 c->B().vcap, c->B().icap, c->d_scan_part, c->d_scan_flag, c->d_scan_ctl, c->scan_epoch,
                                                 c->index32 ? 3u : 7u);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    return UW_OK;
}

static uw_status launch_extract(uw_ctx* c, const int32_t* d_pos, uint32_t n, uint8_t* d_cases, bool only_emit) {
    const DevCfg& d = c->dcfg;
    if (c->big_path) {
        const int grid = persistent_grid(c, n, c->big_blocks_per_sm);
        if (!only_emit) {
            c->big_count_fn<<<persistent_grid(c, n, c->big_count_blocks_per_sm), UW_BIG_NT, 0, c->stream>>>(d, c->d_mc, c->B().d_dens, n, c->B().d_counts, c->B().d_quarters);
            c->launches++;
            CU_TRY(c, cudaGetLastError());
            if (c->profiling) CU_TRY(c, cudaEventRecord(c->ev[2], c->stream));
        }
        { uw_status sst = launch_scan(c, d_pos, n); if (sst != UW_OK) return sst; }
        if (c->profiling && !only_emit) CU_TRY(c, cudaEventRecord(c->ev[3], c->stream));
        if (c->index32)
            c->big_emit32_fn<<<grid, UW_BIG_NT, c->big_smem, c->stream>>>(d, c->d_mc, c->B().d_dens, c->B().d_descs, c->B().d_active,
                                                                             c->d_totals, c->B().d_verts, (uint32_t*)c->B().d_inds, &c->d_scan_ctl->emit_ticket, c->B().d_quarters);
        else
            c->big_emit16_fn<<<grid, UW_BIG_NT, c->big_smem, c->stream>>>(d, c->d_mc, c->B().d_dens, c->B().d_descs, c->B().d_active,
                                                                             c->d_totals, c->B().d_verts, (uint16_t*)c->B().d_inds, &c->d_scan_ctl->emit_ticket, c->B().d_quarters);
        c->launches++;
        CU_TRY(c, cudaGetLastError());
        return UW_OK;
    }
    if (!only_emit) {
        if (c->classify_fn && !d_cases)
            c->classify_fn<<<persistent_grid(c, n, c->classify_spec_blocks_per_sm), c->classify_spec_threads, 0, c->stream>>>(d, c->d_mc, c->B().d_dens, n, c->B().d_counts);
        else
            k_classify_small<<<persistent_grid(c, n, c->classify_blocks_per_sm), 256, 0, c->stream>>>(d, c->d_mc, c->B().d_dens, n, c->B().d_counts, d_cases);
        c->launches++;
        CU_TRY(c, cudaGetLastError());
        if (c->profiling) CU_TRY(c, cudaEventRecord(c->ev[2], c->stream));
    }
    { uw_status sst = launch_scan(c, d_pos, n); if (sst != UW_OK) return sst; }
    if (c->profiling && !only_emit) CU_TRY(c, cudaEventRecord(c->ev[3], c->stream));
    const int grid = persistent_grid(c, n, c->emit_blocks_per_sm);
    if (c->index32)
        c->emit32_fn<<<grid, 256, c->emit_smem, c->stream>>>(d, c->d_mc, c->B().d_dens, c->B().d_descs, c->B().d_active, c->d_totals,
                                                            c->B().d_verts, (uint32_t*)c->B().d_inds, c->B().d_tris, c->B().d_tri_cell);
    else
        c->emit16_fn<<<grid, 256, c->emit_smem, c->stream>>>(d, c->d_mc, c->B().d_dens, c->B().d_descs, c->B().d_active, c->d_totals,
                                                            c->B().d_verts, (uint16_t*)c->B().d_inds, c->B().d_tris, c->B().d_tri_cell);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    return UW_OK;
}

static uw_status launch_fused(uw_ctx* c, const int32_t* d_pos, uint32_t n, float* d_dens_out) {
    const DevCfg& d = c->dcfg;
    if (c->ordered) CU_TRY(c, cudaMemsetAsync(c->B().d_scan, 0, sizeof(ScanSlot) * (size_t)n, c->stream));
    FusedControl* ctl = c->d_ctl + c->ctl_parity;
    FusedControl* ctl_next = c->d_ctl + (c->ctl_parity ^ 1);
    c->ctl_parity ^= 1;
    const int grid = persistent_grid(c, n, c->fused_blocks_per_sm);
    // heavy-first hand-out (scheduling only): provably trivial z layers are deferred inside the kernel; the
    // ordered-packing mode needs tickets == request order.  Only worth it when CTAs get just a few chunks each.
    uint4* d_order = (!c->ordered && c->z_hi >= c->z_lo && n > (uint32_t)grid && n <= c->order_cap) ? c->B().d_order : nullptr;
    // where the outputs go: the context's own arenas, or the attached segment of the rendering GPU's arenas
    uw_chunk_desc* o_descs = c->B().d_descs; uw_vert* o_verts = c->B().d_verts; void* o_inds = c->B().d_inds;
    unsigned long long o_vcap = c->B().vcap, o_icap = c->B().icap;
    FusedOut fo = {};
    bool peer = c->force_staged;
    if (c->B().last_gather) {
        const uw_ctx::GatherTarget& g = c->gt;
        o_descs = g.descs + c->gather_first_chunk; o_verts = g.verts; o_inds = g.inds;
        o_vcap = g.info.seg_vcap; o_icap = g.info.seg_icap;
        fo.desc_vbase = (uint32_t)(g.segment * g.info.seg_vcap); fo.desc_ibase = (uint32_t)(g.segment * g.info.seg_icap);
        fo.head = g.head; fo.epoch = g.epoch; fo.drawlist = g.draw;
        peer = (g.ipc || g.info.device != c->device || c->force_staged) && !c->force_direct;     // NVLink: whole 16-byte vectors
        fo.first_chunk_lo = (uint32_t)c->gather_first_chunk; fo.first_chunk_hi = (uint32_t)(c->gather_first_chunk >> 32);
    }
    if (c->index32)
        (peer ? c->fused32_peer_fn : c->fused32_fn)<<<grid, c->fused_threads, c->fused_smem, c->stream>>>(d, c->tab, c->d_perm, c->d_mc, d_pos, n, c->B().d_scan, ctl, ctl_next,
            o_descs, o_verts, (uint32_t*)o_inds, o_vcap, o_icap, d_dens_out, c->ordered ? 1 : 0, c->B().d_tris, c->B().d_tri_cell, d_order, c->z_lo, c->z_hi, c->zcls, (c->cfg.flags & UW_FLAG_ANALYTIC_SKIP) ? 1 : 0, c->B().h_sum, fo);
    else
        (peer ? c->fused16_peer_fn : c->fused16_fn)<<<grid, c->fused_threads, c->fused_smem, c->stream>>>(d, c->tab, c->d_perm, c->d_mc, d_pos, n, c->B().d_scan, ctl, ctl_next,
            o_descs, o_verts, (uint16_t*)o_inds, o_vcap, o_icap, d_dens_out, c->ordered ? 1 : 0, c->B().d_tris, c->B().d_tri_cell, d_order, c->z_lo, c->z_hi, c->zcls, (c->cfg.flags & UW_FLAG_ANALYTIC_SKIP) ? 1 : 0, c->B().h_sum, fo);
    c->launches++;
    CU_TRY(c, cudaGetLastError());
    return UW_OK;
}

// enqueue the whole pipeline for n chunks whose positions are at d_pos (device), into the current buffer set
static uw_status enqueue_build(uw_ctx* c, const int32_t* d_pos, uint32_t n, bool from_densities) {
    // initial output capacity guess: grows (and the emit stage is re-run) on overflow
    uw_status st = UW_OK;
    c->B().last_gather = c->gather_build;
    if (!c->gather_build) st = ensure_outputs(c, (unsigned long long)n * 192 + 4096, (unsigned long long)n * 640 + 16384);
    if (st != UW_OK) return st;
    const bool fused = c->use_fused && !from_densities;
    const bool keep = (c->cfg.flags & UW_FLAG_KEEP_DENSITIES) != 0;
    if (!fused || keep) { st = ensure_dens(c); if (st != UW_OK) return st; }
    c->launches = 0;
    if (c->tris)      // cells of chunks that never reach the emit stage keep offset 0
        CU_TRY(c, cudaMemsetAsync(c->B().d_tri_cell, 0, (size_t)n * ((size_t)c->dcfg.S * c->dcfg.S * c->dcfg.S + 1) * 2, c->stream));
    if (!fused)       // the fused kernel counts guard re-evaluations in its control block
        CU_TRY(c, cudaMemsetAsync(c->d_guard, 0, sizeof(unsigned long long), c->stream));
    c->B().last_fused = fused;
    if (fused) {
        if (c->profiling) for (int e = 0; e < 4; ++e) CU_TRY(c, cudaEventRecord(c->ev[e], c->stream));
        st = launch_fused(c, d_pos, n, keep ? c->B().d_dens : nullptr);
        if (st != UW_OK) return st;
    } else {
        if (c->profiling) CU_TRY(c, cudaEventRecord(c->ev[0], c->stream));
        if (!from_densities) {
            st = launch_noise(c, d_pos, n);
            if (st != UW_OK) return st;
        }
        if (c->profiling) CU_TRY(c, cudaEventRecord(c->ev[1], c->stream));
        uint8_t* d_cases = nullptr;
        if (keep) {
            if (n > c->B().cap_cases_chunks) {
                CU_TRY(c, cudaStreamSynchronize(c->stream));
                CU_TRY(c, regrow(&c->B().d_cases, (size_t)n * c->dcfg.S * c->dcfg.S * c->dcfg.S));
                c->B().cap_cases_chunks = n;
            }
            d_cases = c->B().d_cases;
        }
        st = launch_extract(c, d_pos, n, d_cases, false);
        if (st != UW_OK) return st;
    }
    if (c->profiling) CU_TRY(c, cudaEventRecord(c->ev[4], c->stream));
    CU_TRY(c, cudaEventRecord(c->B().done, c->stream));
    c->B().last_n = n; c->B().last_pos_dev = d_pos; c->B().pending = true;
    c->build_serial++;
    return UW_OK;
}

static uw_status scratch_get(uw_ctx* c, size_t bytes, char** out) {
    if (bytes > c->d_scratch_cap) {
        size_t cap = c->d_scratch_cap ? c->d_scratch_cap : (1 << 16);
        while (cap < bytes) cap *= 2;
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        if (c->d_scratch) { cudaFree(c->d_scratch); c->d_scratch = nullptr; c->d_scratch_cap = 0; }
        CU_TRY(c, cudaMalloc((void**)&c->d_scratch, cap));
        c->d_scratch_cap = cap;
    }
    *out = c->d_scratch;
    return UW_OK;
}

// Wait for the current set's enqueued build and validate its totals; on output-arena overflow grow and re-run
// the emitting stage.  Read-backs go through the copy stream, so a later batch already running on the compute
// stream does not delay them.
static uw_status finish_build(uw_ctx* c) {
    uw_ctx::BufSet& B = c->B();
    if (!B.pending) return UW_OK;
    BatchTotals t = {};
    unsigned long long guard = 0;
    for (int attempt = 0; attempt < 3; ++attempt) {
        CU_TRY(c, cudaEventSynchronize(B.done));      // everything issued from here on is ordered after the set's kernels
        if (B.last_fused) {
            t = B.h_sum->totals;                      // written straight to host memory by the kernel
            guard = B.h_sum->guard;
            if (!c->ordered) {
                t.n_verts = B.h_sum->alloc >> 32; t.n_inds = B.h_sum->alloc & 0xFFFFFFFFull;
                const unsigned long long vc = B.last_gather ? c->gt.info.seg_vcap : B.vcap, ic = B.last_gather ? c->gt.info.seg_icap : B.icap;
                if ((t.n_verts > vc || t.n_inds > ic) && !t.overflow) t.overflow = 1;
            }
        } else {
            CU_TRY(c, cudaMemcpyAsync(c->h_totals, c->d_totals, sizeof(BatchTotals), cudaMemcpyDeviceToHost, c->copy_stream));
            CU_TRY(c, cudaMemcpyAsync(c->h_guard, c->d_guard, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->copy_stream));
            CU_TRY(c, cudaStreamSynchronize(c->copy_stream));
            t = *c->h_totals;
            guard = *c->h_guard;
        }
        if (!t.overflow) break;
        if (t.overflow >= 3) { B.pending = false; return fail(c, UW_ERR_INVALID, "chunk position out of supported range (|pos| <= 2^24)"); }
        if (t.overflow >= 2) { B.pending = false; return fail(c, UW_ERR_INVALID, "batch too large: packed vertex/index offsets exceed 32 bits; split the batch"); }
        if (B.last_gather) {
            // the segment belongs to the rendering side's arena: nothing to regrow here.  B.result keeps the sizes
            // the build needs, so the owner (uw_multi_build does) can recreate the arena and build again.
            if (t.n_verts <= c->gt.info.seg_vcap && t.n_inds <= c->gt.info.seg_icap) { t.n_verts = c->gt.info.seg_vcap + 1; }
            B.result = t; B.pending = false;
            char m[256];
            snprintf(m, sizeof m, "gather segment overflow: the build needs %llu vertices / %llu indices, the segment holds %llu / %llu",
                     t.n_verts, t.n_inds, (unsigned long long)c->gt.info.seg_vcap, (unsigned long long)c->gt.info.seg_icap);
            return fail(c, UW_ERR_OOM, m);
        }
        if (t.n_verts > 0xFFFFFFFFull || t.n_inds > 0xFFFFFFFFull)
            return fail(c, UW_ERR_INVALID, "batch too large: packed vertex/index offsets exceed 32 bits; split the batch");
        uw_status st = ensure_outputs(c, t.n_verts, t.n_inds);
        if (st != UW_OK) return st;
        if (B.last_fused) {
            st = launch_fused(c, B.last_pos_dev, B.last_n, (c->cfg.flags & UW_FLAG_KEEP_DENSITIES) ? B.d_dens : nullptr);
        } else {
            st = launch_extract(c, B.last_pos_dev, B.last_n, nullptr, true);
        }
        if (st != UW_OK) return st;
        CU_TRY(c, cudaEventRecord(B.done, c->stream));
    }
    if (t.overflow) return fail(c, UW_ERR_CUDA, "output arena overflow persisted");
    B.result = t;
    c->guard_total = guard;
    B.pending = false;
    if (c->profiling) {
        float a = 0, b = 0, d = 0, e = 0, tt = 0;
        cudaEventElapsedTime(&a, c->ev[0], c->ev[1]);
        cudaEventElapsedTime(&b, c->ev[1], c->ev[2]);
        cudaEventElapsedTime(&d, c->ev[2], c->ev[3]);
        cudaEventElapsedTime(&e, c->ev[3], c->ev[4]);
        cudaEventElapsedTime(&tt, c->ev[0], c->ev[4]);
        c->times.noise_ms = a; c->times.classify_ms = b; c->times.scan_ms = d; c->times.emit_ms = e; c->times.total_ms = tt;
    }
    c->times.launches = c->launches;
    return UW_OK;
}

// Returns (in *dev_pos) the pointer the kernels should read the positions from: normally the device copy; for a
// batch of the fused path that is small enough to be latency-bound, the pinned staging buffer itself -- it is
// mapped into the device's address space (UVA), every position is read exactly once (by the hand-out), and the
// H2D copy's issue + completion latency would sit in front of the kernel.
static uw_status stage_positions(uw_ctx* c, const int32_t* pos, uint32_t n, const int32_t** dev_pos = nullptr, bool zero_copy_ok = false,
                                 bool direct_ok = false) {
    // A request that already lives in pinned (page-locked) host memory is copied to the device straight from the
    // caller's buffer -- no staging memcpy, no host-side scan (the fused kernel validates the positions as it fetches
    // them and the batch fails at its wait).  Only for calls whose contract keeps the buffer alive until completion
    // (the blocking uw_build, uw_gather_build); uw_build_async copies, as its caller may reuse the array at once.
    if (direct_ok && n >= 4096 && (size_t)n > c->order_cap && c->use_fused) {   // (smaller batches: zero-copy hand-out below)
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, pos) == cudaSuccess && at.type == cudaMemoryTypeHost) {
            if (dev_pos) *dev_pos = c->B().d_pos;
            CU_TRY(c, cudaMemcpyAsync(c->B().d_pos, pos, (size_t)n * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
            return UW_OK;
        }
        cudaGetLastError();
    }
    {
        // fast path validity: |16*pos| must stay exactly representable next to
        // the 2^-23-granular lattice offsets; also keeps pos*chunk_size inside i32 (chunk.rs:90-94)
        int32_t mx = INT32_MIN, mn = INT32_MAX;            // min / max reductions vectorise
        const size_t m = (size_t)3 * n;
        for (size_t i = 0; i < m; ++i) { mx = pos[i] > mx ? pos[i] : mx; mn = pos[i] < mn ? pos[i] : mn; }
        if (n && (mx > (1 << 24) || mn < -(1 << 24))) return fail(c, UW_ERR_INVALID, "chunk position out of supported range (|pos| <= 2^24)");
    }
    if ((size_t)n * 3 > c->B().h_pos_cap) {
        if (c->B().h_pos) cudaFreeHost(c->B().h_pos);
        c->B().h_pos = nullptr;
        size_t cap = c->B().h_pos_cap ? c->B().h_pos_cap : 4096;
        while (cap < (size_t)n * 3) cap *= 2;
        CU_TRY(c, cudaHostAlloc(&c->B().h_pos, cap * sizeof(int32_t), cudaHostAllocDefault));
        c->B().h_pos_cap = cap;
    }
    memcpy(c->B().h_pos, pos, (size_t)n * 3 * sizeof(int32_t));
    if (dev_pos) *dev_pos = c->B().d_pos;
    // (only in the cost-ordered range: there a few CTAs read all positions in parallel while filing the request;
    // below it every CTA would fetch its own first position over PCIe on its critical path -- measured slower)
#ifdef UW_NO_ZC_POS
    zero_copy_ok = false;
#endif
    if (zero_copy_ok && dev_pos && c->use_fused && c->host_ptr_ok && !c->ordered && c->z_hi >= c->z_lo &&
        n > (uint32_t)(c->num_sms * c->fused_blocks_per_sm) && n <= c->order_cap) {
        *dev_pos = c->B().h_pos;
        return UW_OK;
    }
    CU_TRY(c, cudaMemcpyAsync(c->B().d_pos, c->B().h_pos, (size_t)n * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    return UW_OK;
}

// Collecting batch b (it owns buffer set b->set) has two halves, so that the copy engine never idles when two
// batches are in flight: collect_start waits for the set's kernels, validates the totals, and ISSUES the sized
// D2H copies into a pinned arena on the copy stream; collect_finish waits for them and releases the set.
static uw_status collect_start(uw_ctx* c, uw_batch* b) {
    if (b->copying) return UW_OK;
    const int saved = c->cur;
    c->cur = b->set;
    uw_ctx::BufSet& B = c->B();
    auto done = [&](uw_status st) { c->cur = saved; return st; };
    uw_status st = finish_build(c);
    if (st != UW_OK) return done(st);
    const BatchTotals t = B.result;
    const size_t isz = c->index32 ? 4 : 2;
    const size_t off_desc = 0;
    const size_t off_vert = (sizeof(uw_chunk_desc) * (size_t)b->n + 255) & ~(size_t)255;
    const size_t off_ind = (off_vert + sizeof(uw_vert) * (size_t)t.n_verts + 255) & ~(size_t)255;
    const size_t ncell1 = (size_t)c->dcfg.S * c->dcfg.S * c->dcfg.S + 1;
    const size_t off_tri = (off_ind + isz * (size_t)t.n_inds + 255) & ~(size_t)255;
    const size_t off_tcs = (off_tri + (c->tris ? sizeof(uw_tri) * (size_t)(t.n_inds / 3) : 0) + 255) & ~(size_t)255;
    const size_t bytes = off_tcs + (c->tris ? ncell1 * 2 * (size_t)b->n : 0) + 256;
    st = pinned_get(c, bytes, &b->arena);
    if (st != UW_OK) return done(st);
    char* base = (char*)b->arena.ptr;
    cudaStream_t cs = c->copy_stream;      // finish_build waited for the set's kernels
    cudaError_t e = cudaMemcpyAsync(base + off_desc, B.d_descs, sizeof(uw_chunk_desc) * (size_t)b->n, cudaMemcpyDeviceToHost, cs);
    if (e == cudaSuccess && t.n_verts) e = cudaMemcpyAsync(base + off_vert, B.d_verts, sizeof(uw_vert) * (size_t)t.n_verts, cudaMemcpyDeviceToHost, cs);
    if (e == cudaSuccess && t.n_inds) e = cudaMemcpyAsync(base + off_ind, B.d_inds, isz * (size_t)t.n_inds, cudaMemcpyDeviceToHost, cs);
    if (e == cudaSuccess && c->tris) {
        if (t.n_inds) e = cudaMemcpyAsync(base + off_tri, B.d_tris, sizeof(uw_tri) * (size_t)(t.n_inds / 3), cudaMemcpyDeviceToHost, cs);
        if (e == cudaSuccess) e = cudaMemcpyAsync(base + off_tcs, B.d_tri_cell, ncell1 * 2 * (size_t)b->n, cudaMemcpyDeviceToHost, cs);
    }
    if (e == cudaSuccess) e = cudaEventRecord(B.copied, cs);
    if (e != cudaSuccess) return done(fail(c, UW_ERR_CUDA, std::string("collect_start: ") + cudaGetErrorString(e)));
    memset(&b->view, 0, sizeof b->view);
    if (c->tris) { b->view.tris = (const uw_tri*)(base + off_tri); b->view.tri_cell_start = (const uint16_t*)(base + off_tcs); }
    b->view.n_chunks = b->n; b->view.n_verts = t.n_verts; b->view.n_inds = t.n_inds;
    b->view.descs = (const uw_chunk_desc*)(base + off_desc);
    b->view.verts = (const uw_vert*)(base + off_vert);
    if (c->index32) b->view.inds32 = (const uint32_t*)(base + off_ind);
    else b->view.inds16 = (const uint16_t*)(base + off_ind);
    b->copying = true;
    return done(UW_OK);
}

static uw_status collect_finish(uw_ctx* c, uw_batch* b) {
    uw_ctx::BufSet& B = c->sets[b->set];
    CU_TRY(c, cudaEventSynchronize(B.copied));
    b->ready = true;
    B.busy = false; B.owner = nullptr;
    return UW_OK;
}

static uw_status collect_batch(uw_ctx* c, uw_batch* b) {
    uw_status st = collect_start(c, b);
    if (st != UW_OK) return st;
    // Keep the copy engine busy: while b's copies drain, watch the OTHER in-flight batch; the moment its kernels
    // are done (they run underneath these copies) its own copies are queued right behind b's, instead of at its
    // uw_batch_wait -- no gap on the PCIe link, and the caller's next submit overlaps them.
    uw_ctx::BufSet& B = c->sets[b->set];
    uw_ctx::BufSet& O = c->sets[b->set ^ 1];
    for (;;) {
        const bool other_waiting = O.busy && O.owner && !O.owner->copying;
        if (!other_waiting) break;                               // nothing to watch: block in collect_finish
        const cudaError_t q = cudaEventQuery(B.copied);
        if (q == cudaSuccess) break;
        if (q != cudaErrorNotReady) return fail(c, UW_ERR_CUDA, std::string("collect_batch: ") + cudaGetErrorString(q));
        if (cudaEventQuery(O.done) == cudaSuccess) {
            st = collect_start(c, O.owner);
            if (st != UW_OK) return st;
        }
    }
    st = collect_finish(c, b);
    if (st != UW_OK) return st;
    if (O.busy && O.owner && !O.owner->copying && cudaEventQuery(O.done) == cudaSuccess) st = collect_start(c, O.owner);
    return st;
}

// Batches in flight: the fused small-chunk path keeps all per-batch state in its buffer set, so two batches may
// overlap (one computing, one draining over PCIe); the multi-kernel paths share the scan totals and run one at a time.
static int pick_set(uw_ctx* c, bool from_densities) {
    const int nbusy = (c->sets[0].busy ? 1 : 0) + (c->sets[1].busy ? 1 : 0);
    const int max_in_flight = (c->use_fused && !from_densities) ? 2 : 1;
    if (nbusy >= max_in_flight) return -1;
    if (nbusy == 0) return c->cur;          // keep reusing the warm set
    return c->sets[0].busy ? 1 : 0;
}

static uw_status build_common(uw_ctx* c, const int32_t* pos, const float* dens, uint32_t n, uw_batch** out, bool async) {
    if (!c) return UW_ERR_INVALID;
    if (!out || (!pos && n)) return fail(c, UW_ERR_INVALID, "uw_build: null argument");
    *out = nullptr;
    const int set = pick_set(c, dens != nullptr);
    if (set < 0) return fail(c, UW_ERR_NOT_READY, "uw_build: too many async batches in flight; wait on (or free) an earlier one");
    CU_TRY(c, cudaSetDevice(c->device));
    uw_batch* b = new uw_batch();
    b->ctx = c; b->n = n; b->set = -1; b->ready = false; b->copying = false; b->arena.ptr = nullptr; b->arena.bytes = 0;
    memset(&b->view, 0, sizeof b->view);
    if (n == 0) { b->ready = true; *out = b; return UW_OK; }
    c->cur = set;
    b->set = set;
    const int32_t* dev_pos = nullptr;
    uw_status st = ensure_chunks(c, n);
    if (st == UW_OK) st = stage_positions(c, pos, n, &dev_pos, dens == nullptr, !async && dens == nullptr);
    if (st == UW_OK && dens) {
        st = ensure_dens(c);
        const DevCfg& d = c->dcfg;
        cudaError_t e = st == UW_OK ? cudaMemcpy2DAsync(c->B().d_dens, (size_t)d.dens_stride * 4, dens, (size_t)d.L3 * 4, (size_t)d.L3 * 4, n,
                                                        cudaMemcpyHostToDevice, c->stream) : cudaSuccess;
        if (e != cudaSuccess) st = fail(c, UW_ERR_CUDA, std::string("cudaMemcpy2DAsync densities: ") + cudaGetErrorString(e));
    }
    if (st == UW_OK) st = enqueue_build(c, dev_pos, n, dens != nullptr);
    if (st == UW_OK) {
        c->B().busy = true; c->B().owner = b;
        if (!async) st = collect_batch(c, b);
    }
    if (st != UW_OK) {
        cudaStreamSynchronize(c->stream);
        cudaStreamSynchronize(c->copy_stream);
        c->sets[set].busy = false; c->sets[set].pending = false; c->sets[set].owner = nullptr;
        if (b->arena.ptr) c->pool.push_back(b->arena);
        delete b;
        return st;
    }
    *out = b;
    return UW_OK;
}

extern "C" uw_status uw_build(uw_ctx* c, const int32_t* pos, uint32_t n, uw_batch** out) {
    return build_common(c, pos, nullptr, n, out, false);
}
extern "C" uw_status uw_build_async(uw_ctx* c, const int32_t* pos, uint32_t n, uw_batch** out) {
    return build_common(c, pos, nullptr, n, out, true);
}
extern "C" uw_status uw_build_from_densities(uw_ctx* c, const int32_t* pos, const float* dens, uint32_t n, uw_batch** out) {
    if (!dens && n) return c ? fail(c, UW_ERR_INVALID, "uw_build_from_densities: null densities") : UW_ERR_INVALID;
    return build_common(c, pos, dens, n, out, false);
}

extern "C" uw_status uw_batch_wait(uw_batch* b) {
    if (!b) return UW_ERR_INVALID;
    if (b->ready) return UW_OK;
    CU_TRY(b->ctx, cudaSetDevice(b->ctx->device));
    return collect_batch(b->ctx, b);
}

extern "C" uw_status uw_batch_view_get(const uw_batch* b, uw_batch_view* out) {
    if (!b || !out) return UW_ERR_INVALID;
    if (!b->ready) return fail(b->ctx, UW_ERR_NOT_READY, "uw_batch_view_get: batch not complete; call uw_batch_wait");
    *out = b->view;
    return UW_OK;
}

extern "C" void uw_batch_free(uw_batch* b) {
    if (!b) return;
    if (!b->ready && b->ctx && b->set >= 0) {      // abandoned async batch: let its kernels and copies drain, release the set
        cudaSetDevice(b->ctx->device);
        cudaStreamSynchronize(b->ctx->stream);
        if (b->copying) cudaStreamSynchronize(b->ctx->copy_stream);
        uw_ctx::BufSet& B = b->ctx->sets[b->set];
        B.busy = false; B.pending = false; B.owner = nullptr;
    }
    if (b->arena.ptr) b->ctx->pool.push_back(b->arena);
    delete b;
}

extern "C" uw_status uw_build_device(uw_ctx* c, const int32_t* d_pos, uint32_t n) {
    if (!c) return UW_ERR_INVALID;
    if (!d_pos && n) return fail(c, UW_ERR_INVALID, "uw_build_device: null positions");
    if (c->sets[0].busy || c->sets[1].busy) return fail(c, UW_ERR_NOT_READY, "uw_build_device: an async batch is in flight");
    CU_TRY(c, cudaSetDevice(c->device));
    if (n == 0) { c->B().last_n = 0; c->B().pending = false; c->B().result = BatchTotals{}; return UW_OK; }
    uw_status st = ensure_chunks(c, n);
    if (st != UW_OK) return st;
    return enqueue_build(c, d_pos, n, false);
}

extern "C" uw_status uw_sync(uw_ctx* c) {
    if (!c) return UW_ERR_INVALID;
    CU_TRY(c, cudaSetDevice(c->device));
    if (c->B().pending) { uw_status st = finish_build(c); if (st != UW_OK) return st; }
    CU_TRY(c, cudaStreamSynchronize(c->stream));      // also covers whatever the caller put on the stream after the build
    return UW_OK;
}

extern "C" uw_status uw_device_view_get(uw_ctx* c, uw_device_view* out) {
    if (!c || !out) return UW_ERR_INVALID;
    if (c->B().pending) return fail(c, UW_ERR_NOT_READY, "uw_device_view_get: call uw_sync first");
    memset(out, 0, sizeof *out);
    out->n_chunks = c->B().last_n;
    out->n_verts = c->B().last_n ? c->B().result.n_verts : 0;
    out->n_inds = c->B().last_n ? c->B().result.n_inds : 0;
    out->d_descs = c->B().d_descs; out->d_verts = c->B().d_verts;
    if (c->index32) out->d_inds32 = c->B().d_inds; else out->d_inds16 = c->B().d_inds;
    out->d_densities = c->B().d_dens; out->density_stride = c->dcfg.dens_stride;
    return UW_OK;
}

extern "C" uw_status uw_export_arena_fd(uw_ctx* c, int which, int* fd, uint64_t* bytes) {
    if (!c) return UW_ERR_INVALID;
    if (!fd || !bytes || (which != 0 && which != 1)) return fail(c, UW_ERR_INVALID, "uw_export_arena_fd: bad argument");
    if (!c->exportable) return fail(c, UW_ERR_UNSUPPORTED, "uw_export_arena_fd: the context was not created with UW_FLAG_EXPORTABLE");
    if (c->B().pending) return fail(c, UW_ERR_NOT_READY, "uw_export_arena_fd: call uw_sync first");
    const VmmArena& a = which == 0 ? c->B().vmm_verts : c->B().vmm_inds;
    if (!a.ptr) return fail(c, UW_ERR_NOT_READY, "uw_export_arena_fd: nothing has been built yet");
    CU_TRY(c, cudaSetDevice(c->device));
    int out = -1;
    if (driver_api().memExportToShareableHandle(&out, a.handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS || out < 0)
        return fail(c, UW_ERR_CUDA, "cuMemExportToShareableHandle failed");
    *fd = out; *bytes = a.bytes;
    return UW_OK;
}

extern "C" uw_status uw_get_stage_times(uw_ctx* c, uw_stage_times* out) {
    if (!c || !out) return UW_ERR_INVALID;
    *out = c->times;
    return UW_OK;
}

extern "C" uw_status uw_get_guard_count(uw_ctx* c, uint64_t* out) {
    if (!c || !out) return UW_ERR_INVALID;
    *out = c->guard_total;
    return UW_OK;
}

// ---------------------------------------------------------------------------------------
// multi-GPU: gather arenas on the rendering GPU, producers that write into them over NVLink
// ---------------------------------------------------------------------------------------
#include <unistd.h>

extern "C" void uw_slab_bounds(uint32_t n, uint32_t parts, uint32_t part, uint32_t* first, uint32_t* count) {
    if (parts == 0) parts = 1;
    const uint32_t base = n / parts, rem = n % parts;
    const uint32_t lo = part * base + (part < rem ? part : rem);
    if (first) *first = lo;
    if (count) *count = base + (part < rem ? 1u : 0u);
}

extern "C" void uw_slab_bounds_weighted(uint32_t n, uint32_t parts, uint32_t part, uint32_t render_part, uint32_t render_permille,
                                        uint32_t* first, uint32_t* count) {
    if (parts <= 1 || render_part >= parts || (unsigned long long)render_permille * parts <= 1000ull) {
        uw_slab_bounds(n, parts, part, first, count);
        return;
    }
    if (render_permille > 1000) render_permille = 1000;
    const uint32_t cr = (uint32_t)(((unsigned long long)n * render_permille + 500ull) / 1000ull);
    const uint32_t rest = n - cr, others = parts - 1;
    const uint32_t base = rest / others, rem = rest % others;
    uint32_t lo = 0, cnt = 0;
    for (uint32_t p = 0, k = 0; p <= part; ++p) {           // k = index among the non-render slabs
        lo += cnt;
        if (p == render_part) cnt = cr;
        else { cnt = base + (k < rem ? 1u : 0u); ++k; }
    }
    if (first) *first = lo;
    if (count) *count = cnt;
}

static bool gather_supported(uw_ctx* c, const char* who) {
    if (c->use_fused && !c->tris && !(c->cfg.flags & UW_FLAG_KEEP_DENSITIES)) return true;
    fail(c, UW_ERR_UNSUPPORTED, std::string(who) + ": the gather path needs the fused kernel (internal_size 10 / 12, reference constants; "
                                                   "no UW_FLAG_STAGED / TRIS / KEEP_DENSITIES / EXACT_F64)");
    return false;
}

extern "C" uw_status uw_gather_destroy(uw_ctx* c) {
    if (!c) return UW_ERR_INVALID;
    uw_ctx::GatherArena& a = c->arena;
    if (!a.alive) return UW_OK;
    CU_TRY(c, cudaSetDevice(c->device));
    if (c->gt.active && !c->gt.ipc && c->gt.base == a.base) c->gt = uw_ctx::GatherTarget();
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    cudaFree(a.base); cudaFree(a.d_status);
    if (a.h_head) cudaFreeHost(a.h_head);
    if (a.h_status) cudaFreeHost(a.h_status);
    if (a.h_descs) cudaFreeHost(a.h_descs);
    if (a.h_draw) cudaFreeHost(a.h_draw);
    a = uw_ctx::GatherArena();
    return UW_OK;
}

extern "C" uw_status uw_gather_create(uw_ctx* c, uint32_t n_segments, uint64_t n_chunks, uint64_t seg_vcap, uint64_t seg_icap,
                                      uw_gather_info* out) {
    if (!c) return UW_ERR_INVALID;
    if (!out || n_segments == 0 || n_segments > UW_MAX_SEGMENTS || n_chunks == 0)
        return fail(c, UW_ERR_INVALID, "uw_gather_create: bad argument (1 <= n_segments <= UW_MAX_SEGMENTS, n_chunks > 0)");
    if (!gather_supported(c, "uw_gather_create")) return UW_ERR_UNSUPPORTED;
    if (c->arena.alive) { uw_status st = uw_gather_destroy(c); if (st != UW_OK) return st; }
    CU_TRY(c, cudaSetDevice(c->device));
    const uint64_t per = (n_chunks + n_segments - 1) / n_segments;
    if (seg_vcap == 0) seg_vcap = per * 192 + 4096;
    if (seg_icap == 0) seg_icap = per * 640 + 16384;
    seg_vcap = (seg_vcap + 1) & ~1ull;                     // segments start on 16-byte boundaries (2 x 24 B, 8 x 2 B)
    seg_icap = (seg_icap + 7) & ~7ull;
    if (seg_vcap * n_segments > 0xFFFFFFFFull || seg_icap * n_segments > 0xFFFFFFFFull)
        return fail(c, UW_ERR_INVALID, "uw_gather_create: arena too large for 32-bit descriptor offsets; use fewer chunks per arena");
    const size_t isz = c->index32 ? 4 : 2;
    auto up = [](uint64_t v) { return (v + 255) & ~255ull; };
    uw_gather_info g;
    memset(&g, 0, sizeof g);
    g.abi_version = UW_ABI_VERSION; g.n_segments = n_segments; g.device = c->device; g.index_bytes = (uint32_t)isz;
    g.owner_pid = (uint64_t)getpid();
    g.n_chunks = n_chunks; g.seg_vcap = seg_vcap; g.seg_icap = seg_icap;
    g.off_head = 0;
    g.off_descs = up(sizeof(GatherHead) * n_segments);
    g.off_verts = up(g.off_descs + sizeof(uw_chunk_desc) * n_chunks);
    g.off_inds = up(g.off_verts + sizeof(uw_vert) * seg_vcap * n_segments);
    g.off_draw = up(g.off_inds + isz * seg_icap * n_segments);
    g.bytes = up(g.off_draw + sizeof(uw_chunk_desc) * n_chunks * n_segments);
    uw_ctx::GatherArena& a = c->arena;
    CU_TRY(c, cudaMalloc((void**)&a.base, g.bytes));
    cudaError_t e = cudaMemsetAsync(a.base, 0, g.off_descs, c->stream);           // heads: epoch 0
    if (e == cudaSuccess) e = cudaMalloc((void**)&a.d_status, sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&a.h_head, sizeof(GatherHead) * n_segments, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&a.h_status, sizeof(uint32_t), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) {
        cudaIpcMemHandle_t h;
        // other PROCESSES map the arena through this handle; it is optional (a box whose driver refuses IPC can still
        // gather inside one process), so a failure only leaves the handle zeroed
        if (cudaIpcGetMemHandle(&h, a.base) == cudaSuccess) { static_assert(sizeof h == 64, "cudaIpcMemHandle_t"); memcpy(g.ipc_handle, &h, 64); }
        else cudaGetLastError();
    }
    if (e != cudaSuccess) {
        cudaFree(a.base); cudaFree(a.d_status);
        if (a.h_head) cudaFreeHost(a.h_head);
        if (a.h_status) cudaFreeHost(a.h_status);
        a = uw_ctx::GatherArena();
        return fail(c, e == cudaErrorMemoryAllocation ? UW_ERR_OOM : UW_ERR_CUDA, std::string("uw_gather_create: ") + cudaGetErrorString(e));
    }
    g.base = (uint64_t)(uintptr_t)a.base;
    a.info = g; a.alive = true; a.wait_epoch = 0;
    *out = g;
    return UW_OK;
}

extern "C" uw_status uw_gather_detach(uw_ctx* c) {
    if (!c) return UW_ERR_INVALID;
    if (!c->gt.active) return UW_OK;
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    if (c->gt.ipc && c->gt.base) cudaIpcCloseMemHandle(c->gt.base);
    c->gt = uw_ctx::GatherTarget();
    return UW_OK;
}

extern "C" uw_status uw_gather_attach(uw_ctx* c, const uw_gather_info* info, uint32_t segment) {
    if (!c) return UW_ERR_INVALID;
    if (!info || info->abi_version != UW_ABI_VERSION || segment >= info->n_segments || info->n_segments > UW_MAX_SEGMENTS)
        return fail(c, UW_ERR_INVALID, "uw_gather_attach: bad info / segment");
    if (!gather_supported(c, "uw_gather_attach")) return UW_ERR_UNSUPPORTED;
    if (info->index_bytes != (c->index32 ? 4u : 2u))
        return fail(c, UW_ERR_INVALID, "uw_gather_attach: index width differs from the arena's");
    if (c->sets[0].busy || c->sets[1].busy || c->B().pending) return fail(c, UW_ERR_NOT_READY, "uw_gather_attach: a build is in flight");
    { uw_status st = uw_gather_detach(c); if (st != UW_OK) return st; }
    CU_TRY(c, cudaSetDevice(c->device));
    uw_ctx::GatherTarget g;
    g.info = *info; g.segment = segment;
    if (info->owner_pid == (uint64_t)getpid()) {
        g.base = (char*)(uintptr_t)info->base;
        if (info->device != c->device) {
            int can = 0;
            CU_TRY(c, cudaDeviceCanAccessPeer(&can, c->device, info->device));
            if (!can) return fail(c, UW_ERR_UNSUPPORTED, "uw_gather_attach: no peer access between the producer and the render device");
            const cudaError_t e = cudaDeviceEnablePeerAccess(info->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(c, UW_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            cudaGetLastError();
        }
    } else {
        cudaIpcMemHandle_t h;
        memcpy(&h, info->ipc_handle, 64);
        void* p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { cudaGetLastError(); return fail(c, UW_ERR_CUDA, std::string("uw_gather_attach: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e)); }
        g.base = (char*)p; g.ipc = true;
    }
    const size_t isz = info->index_bytes;
    g.head = (GatherHead*)(g.base + info->off_head) + segment;
    g.descs = (uw_chunk_desc*)(g.base + info->off_descs);
    g.verts = (uw_vert*)(g.base + info->off_verts) + (size_t)segment * info->seg_vcap;
    g.inds = g.base + info->off_inds + (size_t)segment * info->seg_icap * isz;
    g.draw = (uw_chunk_desc*)(g.base + info->off_draw) + (size_t)segment * info->n_chunks;
    // continue the segment's epoch count (an arena outlives re-attachments)
    uint32_t ep = 0;
    const cudaError_t e = cudaMemcpy(&ep, &g.head->epoch, sizeof ep, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) {
        if (g.ipc) cudaIpcCloseMemHandle(g.base);
        return fail(c, UW_ERR_CUDA, std::string("uw_gather_attach: reading the segment head: ") + cudaGetErrorString(e));
    }
    g.epoch = ep; g.active = true;
    c->gt = g;
    return UW_OK;
}

static uw_status gather_build_common(uw_ctx* c, const int32_t* pos, bool host_pos, uint32_t n, uint64_t first_chunk) {
    if (!c) return UW_ERR_INVALID;
    if (!c->gt.active) return fail(c, UW_ERR_INVALID, "uw_gather_build: no gather segment attached (uw_gather_attach)");
    if (!pos && n) return fail(c, UW_ERR_INVALID, "uw_gather_build: null positions");
    if (first_chunk + n > c->gt.info.n_chunks) return fail(c, UW_ERR_INVALID, "uw_gather_build: first_chunk + n exceeds the arena's descriptor capacity");
    if (c->sets[0].busy || c->sets[1].busy) return fail(c, UW_ERR_NOT_READY, "uw_gather_build: an async host batch is in flight");
    CU_TRY(c, cudaSetDevice(c->device));
    // the previous build's totals must be validated before the set (and its pinned staging) is reused
    if (c->B().pending) { uw_status st = finish_build(c); if (st != UW_OK) return st; }
    uw_status st = ensure_chunks(c, n ? n : 1);
    if (st != UW_OK) return st;
    const int32_t* dev_pos = pos;
    if (host_pos && n) { st = stage_positions(c, pos, n, &dev_pos, true, true); if (st != UW_OK) return st; }
    c->gt.epoch += 1;
    c->gather_first_chunk = first_chunk;
    if (n == 0) {                                        // an empty slab still publishes its head
        GatherHead h;
        memset(&h, 0, sizeof h);
        h.epoch = c->gt.epoch; h.first_chunk_lo = (uint32_t)first_chunk; h.first_chunk_hi = (uint32_t)(first_chunk >> 32);
        CU_TRY(c, cudaMemcpyAsync(c->gt.head, &h, sizeof h, cudaMemcpyHostToDevice, c->stream));
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        c->B().last_n = 0; c->B().pending = false; c->B().result = BatchTotals{};
        return UW_OK;
    }
    c->gather_build = true;
    cudaEventRecord(c->ev[0], c->stream);                 // kernel time of this producer (uw_multi_build balances on it)
    st = enqueue_build(c, dev_pos, n, false);
    cudaEventRecord(c->ev[4], c->stream);
    c->gather_build = false;
    return st;
}

extern "C" uw_status uw_gather_build(uw_ctx* c, const int32_t* pos, uint32_t n, uint64_t first_chunk) {
    return gather_build_common(c, pos, true, n, first_chunk);
}
extern "C" uw_status uw_gather_build_device(uw_ctx* c, const int32_t* d_pos, uint32_t n, uint64_t first_chunk) {
    return gather_build_common(c, d_pos, false, n, first_chunk);
}

extern "C" uw_status uw_gather_wait(uw_ctx* c, uint32_t flags, uw_gather_result* out) {
    if (!c) return UW_ERR_INVALID;
    uw_ctx::GatherArena& a = c->arena;
    if (!a.alive) return fail(c, UW_ERR_INVALID, "uw_gather_wait: this context owns no gather arena (uw_gather_create)");
    if (!out) return fail(c, UW_ERR_INVALID, "uw_gather_wait: null result");
    CU_TRY(c, cudaSetDevice(c->device));
    // this context's own segment (if it produces one) is validated on the host like any other build
    if (c->B().pending) { uw_status st = finish_build(c); if (st != UW_OK) return st; }
    const uw_gather_info& g = a.info;
    const uint32_t epoch = ++a.wait_epoch;
    GatherHead* d_head = (GatherHead*)(a.base + g.off_head);
    k_gather_wait<<<1, 32, 0, c->stream>>>(d_head, g.n_segments, epoch, 10ull * 1000 * 1000 * 1000, a.d_status);
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaMemcpyAsync(a.h_status, a.d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaMemcpyAsync(a.h_head, d_head, sizeof(GatherHead) * g.n_segments, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    if (*a.h_status != 0) { --a.wait_epoch; return fail(c, UW_ERR_NOT_READY, "uw_gather_wait: timed out waiting for a segment (did every producer build?)"); }
    memset(out, 0, sizeof *out);
    out->n_segments = g.n_segments; out->epoch = epoch;
    out->d_descs = a.base + g.off_descs; out->d_verts = a.base + g.off_verts; out->d_inds = a.base + g.off_inds;
    out->seg_vcap = g.seg_vcap; out->seg_icap = g.seg_icap; out->d_draw = a.base + g.off_draw;
    uint64_t hi_chunk = 0;
    bool overflow = false;
    for (uint32_t s = 0; s < g.n_segments; ++s) {
        const GatherHead& h = a.h_head[s];
        uw_gather_segment& o = out->seg[s];
        o.first_chunk = (uint64_t)h.first_chunk_lo | ((uint64_t)h.first_chunk_hi << 32);
        o.n_chunks = h.n_chunks; o.n_mesh = h.sum.totals.n_active; o.n_blank = h.sum.totals.n_blank;
        o.guard = h.sum.guard;
        if (c->ordered) { o.n_verts = h.sum.totals.n_verts; o.n_inds = h.sum.totals.n_inds; }
        else { o.n_verts = h.sum.alloc >> 32; o.n_inds = h.sum.alloc & 0xFFFFFFFFull; }
        o.overflow = (h.sum.totals.overflow || o.n_verts > g.seg_vcap || o.n_inds > g.seg_icap) ? 1u : 0u;
        overflow |= o.overflow != 0;
        out->n_chunks += o.n_chunks; out->n_verts += o.n_verts; out->n_inds += o.n_inds; out->n_draw += o.n_mesh;
        if (o.n_chunks && o.first_chunk + o.n_chunks > hi_chunk) hi_chunk = o.first_chunk + o.n_chunks;
    }
    if (overflow) return fail(c, UW_ERR_OOM, "uw_gather_wait: a segment overflowed its capacity (see uw_gather_result.seg[].overflow); recreate the arena larger");
    if ((flags & UW_GATHER_DESCS_TO_HOST) && hi_chunk) {
        if (hi_chunk > a.h_descs_cap) {
            if (a.h_descs) cudaFreeHost(a.h_descs);
            a.h_descs = nullptr; a.h_descs_cap = 0;
            CU_TRY(c, cudaHostAlloc((void**)&a.h_descs, sizeof(uw_chunk_desc) * (size_t)g.n_chunks, cudaHostAllocDefault));
            a.h_descs_cap = (size_t)g.n_chunks;
        }
        CU_TRY(c, cudaMemcpyAsync(a.h_descs, a.base + g.off_descs, sizeof(uw_chunk_desc) * (size_t)hi_chunk, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        out->h_descs = a.h_descs;
    }
    if ((flags & UW_GATHER_DRAW_TO_HOST) && out->n_draw) {
        if (out->n_draw > a.h_draw_cap) {
            if (a.h_draw) cudaFreeHost(a.h_draw);
            a.h_draw = nullptr; a.h_draw_cap = 0;
            CU_TRY(c, cudaHostAlloc((void**)&a.h_draw, sizeof(uw_chunk_desc) * (size_t)g.n_chunks, cudaHostAllocDefault));
            a.h_draw_cap = (size_t)g.n_chunks;
        }
        size_t at = 0;
        for (uint32_t s = 0; s < g.n_segments; ++s) {
            const size_t cnt = out->seg[s].n_mesh;
            if (!cnt) continue;
            CU_TRY(c, cudaMemcpyAsync(a.h_draw + at, (const uw_chunk_desc*)(a.base + g.off_draw) + (size_t)s * g.n_chunks,
                                      sizeof(uw_chunk_desc) * cnt, cudaMemcpyDeviceToHost, c->stream));
            at += cnt;
        }
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        out->h_draw = a.h_draw;
    }
    return UW_OK;
}

extern "C" uw_status uw_debug_copy_to_host(const void* d_src, uint64_t bytes, void* h_dst) {
    if ((!d_src || !h_dst) && bytes) return UW_ERR_INVALID;
    if (bytes == 0) return UW_OK;
    if (cudaMemcpy(h_dst, d_src, (size_t)bytes, cudaMemcpyDefault) != cudaSuccess) { cudaGetLastError(); return UW_ERR_CUDA; }
    return UW_OK;
}

// ---- one process, G devices -----------------------------------------------------------------------------------
struct uw_multi {
    std::vector<uw_ctx*> ctx;            // ctx[0] renders
    uw_gather_info info = {};
    bool have_arena = false;
    double render_share = 0.0;           // fraction of a request the rendering GPU builds itself; 0 = even split
    uw_share_search search = {};         // state of the search that chooses it (uw_share_search_next)
    std::string err;
};
static thread_local std::string g_multi_error;

static uw_status mfail(uw_multi* m, uw_status st, const std::string& msg) {
    if (m) m->err = msg; else g_multi_error = msg;
    return st;
}

extern "C" const char* uw_multi_last_error(const uw_multi* m) { return m ? m->err.c_str() : g_multi_error.c_str(); }
extern "C" uint32_t uw_multi_render_share(const uw_multi* m) { return m ? (uint32_t)(m->render_share * 1000.0 + 0.5) : 0u; }

// include/uwcuda.h: damped hill climb on the measured cost of a whole request
extern "C" double uw_share_search_next(uw_share_search* s, uint32_t parts, double cost) {
    if (!s) return 0.0;
    if (parts < 2) { s->share = 0.0; return 0.0; }
    const double even = 1.0 / (double)parts, min_step = even / 64.0;
    auto clamp = [&](double x) { return x < even ? even : x > 0.5 ? 0.5 : x; };
    auto restart = [&](double from) {
        s->share = from; s->step = even / 4.0; s->dir = 1; s->last_cost = 0.0; s->best_share = from; s->best_cost = 0.0; s->settled = 0;
    };
    if (!(s->share > 0.0)) { restart(even); s->moves = 0; }
    if (!(cost > 0.0)) return s->share;
    if (++s->moves <= 2) return s->share;                  // the first two requests are cold (allocation, first touch): not evidence
    if (s->settled) {                                      // holding the cheapest share: only watch for another workload
        if (cost > s->best_cost * 1.10) { restart(s->share); s->last_cost = s->best_cost = cost; s->share = clamp(s->share + s->step); }
        else if (cost < s->best_cost) s->best_cost = cost;
        return s->share;
    }
    if (!(s->best_cost > 0.0) || cost < s->best_cost) { s->best_cost = cost; s->best_share = s->share; }
    if (s->last_cost > 0.0) {
        if (cost > s->last_cost * 1.01) { s->dir = -s->dir; s->step *= 0.5; }        // worse: turn round, smaller step
        else if (!(cost < s->last_cost * 0.99)) s->step *= 0.5;                       // flat: refine
    }
    s->last_cost = cost;
    if (s->step < min_step) { s->settled = 1; s->share = s->best_share; return s->share; }
    double next = clamp(s->share + (double)s->dir * s->step);
    if (next == s->share) { s->dir = -s->dir; next = clamp(s->share + (double)s->dir * s->step); }   // at a bound: look the other way
    s->share = next;
    return next;
}

extern "C" void uw_multi_destroy(uw_multi* m) {
    if (!m) return;
    for (size_t g = m->ctx.size(); g-- > 0;) if (m->ctx[g]) uw_gather_detach(m->ctx[g]);
    if (!m->ctx.empty() && m->ctx[0]) uw_gather_destroy(m->ctx[0]);
    for (auto* c : m->ctx) uw_destroy(c);
    delete m;
}

extern "C" uw_status uw_multi_create(const uw_config* cfg, const int32_t* devices, uint32_t n_devices, uw_multi** out) {
    if (!cfg || !out || !devices || n_devices == 0 || n_devices > UW_MAX_SEGMENTS)
        return mfail(nullptr, UW_ERR_INVALID, "uw_multi_create: bad argument (1 <= n_devices <= UW_MAX_SEGMENTS)");
    *out = nullptr;
    for (uint32_t a = 0; a < n_devices; ++a)
        for (uint32_t b = a + 1; b < n_devices; ++b)
            if (devices[a] == devices[b]) return mfail(nullptr, UW_ERR_INVALID, "uw_multi_create: a device is listed twice");
    uw_multi* m = new uw_multi();
    for (uint32_t g = 0; g < n_devices; ++g) {
        uw_config c = *cfg;
        c.device = devices[g];
        uw_ctx* ctx = nullptr;
        const uw_status st = uw_create(&c, &ctx);
        if (st != UW_OK) { g_multi_error = std::string("uw_multi_create: device ") + std::to_string(devices[g]) + ": " + uw_last_error(nullptr); uw_multi_destroy(m); return st; }
        m->ctx.push_back(ctx);
        if (!gather_supported(ctx, "uw_multi_create")) { g_multi_error = ctx->err; uw_multi_destroy(m); return UW_ERR_UNSUPPORTED; }
    }
    *out = m;
    return UW_OK;
}

extern "C" uw_status uw_multi_build(uw_multi* m, const int32_t* pos, uint32_t n, uint32_t flags, uw_gather_result* out) {
    if (!m) return UW_ERR_INVALID;
    if (!out || (!pos && n)) return mfail(m, UW_ERR_INVALID, "uw_multi_build: null argument");
    const uint32_t G = (uint32_t)m->ctx.size();
    uw_ctx* r = m->ctx[0];
    uint64_t vcap = 0, icap = 0;                          // 0 = the library's default estimate
    if (G >= 2) {
        // the rendering GPU's slab may grow to half the request (gather-aware partition): size the segments for that
        // from the start, unless 32-bit descriptor offsets would not cover G such segments
        const uint64_t per = ((uint64_t)n + 1) / 2, v = per * 192 + 4096, i = per * 640 + 16384;
        if (v * G <= 0xFFFFFFFFull && i * G <= 0xFFFFFFFFull) { vcap = v; icap = i; }
    }
    for (int attempt = 0; attempt < 4; ++attempt) {
        if (!m->have_arena || m->info.n_chunks < (n ? n : 1u) || vcap > m->info.seg_vcap || icap > m->info.seg_icap) {
            for (auto* c : m->ctx) { const uw_status st = uw_gather_detach(c); if (st != UW_OK) return mfail(m, st, c->err); }
            m->have_arena = false;
            uw_status st = uw_gather_create(r, G, n ? n : 1u, vcap, icap, &m->info);
            if (st != UW_OK) return mfail(m, st, r->err);
            for (uint32_t g = 0; g < G; ++g) {
                st = uw_gather_attach(m->ctx[g], &m->info, g);
                if (st != UW_OK) return mfail(m, st, m->ctx[g]->err);
            }
            m->have_arena = true;
        }
        // every device: pinned H2D of its slab + one fused launch, all asynchronous; device 0 last, so that the
        // peers are already computing while the render device's own work is being enqueued
        const auto t_begin = std::chrono::steady_clock::now();
        for (uint32_t k = 0; k < G; ++k) {
            const uint32_t g = (k + 1) % G;
            uint32_t first = 0, cnt = 0;
            uw_slab_bounds_weighted(n, G, g, 0, (uint32_t)(m->render_share * 1000.0 + 0.5), &first, &cnt);
            const uw_status st = uw_gather_build(m->ctx[g], pos + 3 * (size_t)first, cnt, first);
            if (st != UW_OK) return mfail(m, st, m->ctx[g]->err);
        }
        bool overflow = false;
        unsigned long long need_v = 0, need_i = 0;
        uw_status first_err = UW_OK; std::string first_msg;
        for (uint32_t g = 0; g < G; ++g) {               // validate every producer on the host (overflow, CUDA errors)
            const uw_status st = uw_sync(m->ctx[g]);
            if (st == UW_ERR_OOM && m->ctx[g]->B().last_gather) {
                overflow = true;
                need_v = std::max(need_v, m->ctx[g]->B().result.n_verts); need_i = std::max(need_i, m->ctx[g]->B().result.n_inds);
            } else if (st != UW_OK && first_err == UW_OK) { first_err = st; first_msg = m->ctx[g]->err; }
        }
        if (first_err != UW_OK) return mfail(m, first_err, first_msg);
        const uw_status wst = uw_gather_wait(r, flags, out);
        if (!overflow && wst == UW_OK) {
            if (G >= 2 && n >= 4096u * G) {               // choose the rendering GPU's share for the next request
                // cost of this request = its wall time per chunk (enqueue .. every mesh landed and the heads read), measured
                // at the share it ran with; the search hands back the share for the next request
                const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
                m->render_share = uw_share_search_next(&m->search, G, secs / (double)n);
            }
            return UW_OK;
        }
        if (!overflow) return mfail(m, wst, r->err);
        vcap = std::max<unsigned long long>(m->info.seg_vcap, need_v + need_v / 8 + 4096);
        icap = std::max<unsigned long long>(m->info.seg_icap, need_i + need_i / 8 + 16384);
        if (vcap == m->info.seg_vcap && icap == m->info.seg_icap) { vcap *= 2; icap *= 2; }
    }
    return mfail(m, UW_ERR_OOM, "uw_multi_build: segment overflow persisted");
}

// ---------------------------------------------------------------------------------------
// parity taps
// ---------------------------------------------------------------------------------------
extern "C" uw_status uw_debug_densities(uw_ctx* c, const int32_t* pos, uint32_t n, float* out) {
    if (!c) return UW_ERR_INVALID;
    if ((!pos || !out) && n) return fail(c, UW_ERR_INVALID, "uw_debug_densities: null argument");
    if (n == 0) return UW_OK;
    CU_TRY(c, cudaSetDevice(c->device));
    if (c->sets[0].busy || c->sets[1].busy) return fail(c, UW_ERR_NOT_READY, "uw_debug_densities: an async batch is in flight");
    uw_status st = ensure_chunks(c, n);
    if (st == UW_OK) st = ensure_dens(c);
    if (st == UW_OK) st = stage_positions(c, pos, n);
    if (st == UW_OK) CU_TRY(c, cudaMemsetAsync(c->d_guard, 0, sizeof(unsigned long long), c->stream));
    if (st == UW_OK) st = launch_noise(c, c->B().d_pos, n);
    if (st != UW_OK) return st;
    const DevCfg& d = c->dcfg;
    CU_TRY(c, cudaMemcpy2DAsync(out, (size_t)d.L3 * 4, c->B().d_dens, (size_t)d.dens_stride * 4, (size_t)d.L3 * 4, n,
                                cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaMemcpyAsync(c->h_guard, c->d_guard, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->guard_total = *c->h_guard;
    return UW_OK;
}

extern "C" uw_status uw_debug_cases(uw_ctx* c, const int32_t* pos, uint32_t n, uint8_t* out) {
    if (!c) return UW_ERR_INVALID;
    if (c->big_path) return fail(c, UW_ERR_UNSUPPORTED, "uw_debug_cases: not available for internal_size > 15");
    if ((!pos || !out) && n) return fail(c, UW_ERR_INVALID, "uw_debug_cases: null argument");
    if (n == 0) return UW_OK;
    CU_TRY(c, cudaSetDevice(c->device));
    if (c->sets[0].busy || c->sets[1].busy) return fail(c, UW_ERR_NOT_READY, "uw_debug_cases: an async batch is in flight");
    uw_status st = ensure_chunks(c, n);
    if (st == UW_OK) st = ensure_dens(c);
    if (st == UW_OK) st = stage_positions(c, pos, n);
    if (st == UW_OK) st = launch_noise(c, c->B().d_pos, n);
    if (st != UW_OK) return st;
    const DevCfg& d = c->dcfg;
    const size_t S3 = (size_t)d.S * d.S * d.S;
    if (n > c->B().cap_cases_chunks) {
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        CU_TRY(c, regrow(&c->B().d_cases, (size_t)n * S3));
        c->B().cap_cases_chunks = n;
    }
    k_classify_small<<<persistent_grid(c, n, c->classify_blocks_per_sm), 256, 0, c->stream>>>(d, c->d_mc, c->B().d_dens, n, c->B().d_counts, c->B().d_cases);
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaMemcpyAsync(out, c->B().d_cases, (size_t)n * S3, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return UW_OK;
}

extern "C" uw_status uw_iso_at(uw_ctx* c, const double* pts, uint32_t n, float* out) {
    if (!c) return UW_ERR_INVALID;
    if ((!pts || !out) && n) return fail(c, UW_ERR_INVALID, "uw_iso_at: null argument");
    if (n == 0) return UW_OK;
    CU_TRY(c, cudaSetDevice(c->device));
    char* base = nullptr;
    const size_t pts_bytes = ((size_t)n * 3 * sizeof(double) + 255) & ~(size_t)255;
    { uw_status st = scratch_get(c, pts_bytes + (size_t)n * sizeof(float), &base); if (st != UW_OK) return st; }
    double* d_pts = (double*)base; float* d_out = (float*)(base + pts_bytes);
    CU_TRY(c, cudaMemcpyAsync(d_pts, pts, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    unsigned blocks = (n + 255) / 256;
    if (blocks > (unsigned)c->num_sms * 8) blocks = c->num_sms * 8;
    k_iso_points<<<blocks, 256, 0, c->stream>>>(c->dcfg, c->d_perm, d_pts, n, d_out);
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaMemcpyAsync(out, d_out, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return UW_OK;
}

extern "C" uw_status uw_raycast_tris(uw_ctx* c, const float* origins, const float* dirs, uint32_t n_rays, int32_t wall_range, float* out_t) {
    if (!c) return UW_ERR_INVALID;
    if ((!origins || !dirs || !out_t) && n_rays) return fail(c, UW_ERR_INVALID, "uw_raycast_tris: null argument");
    if (!c->tris) return fail(c, UW_ERR_UNSUPPORTED, "uw_raycast_tris: the context was not created with UW_FLAG_TRIS");
    if (wall_range < 0 || wall_range >= c->cfg.chunk_size)
        return fail(c, UW_ERR_INVALID, "uw_raycast_tris: wall_range must be in [0, chunk_size)");
    if (n_rays == 0) return UW_OK;
    CU_TRY(c, cudaSetDevice(c->device));
    if (c->sets[0].busy || c->sets[1].busy) return fail(c, UW_ERR_NOT_READY, "uw_raycast_tris: an async batch is in flight");
    if (c->B().pending) { uw_status st = finish_build(c); if (st != UW_OK) return st; }
    const uint32_t n = c->B().last_n;
    if (n == 0) { for (uint32_t i = 0; i < n_rays; ++i) out_t[i] = -1.0f; return UW_OK; }
    if (c->ray_table_serial != c->build_serial || c->ray_table_set != c->cur) {           // first query after a build: index its chunks
        uint32_t cap = 1024;
        while (cap < 2u * n) cap *= 2;
        if (cap > c->ray_table_cap) {
            CU_TRY(c, cudaStreamSynchronize(c->stream));
            if (c->d_ray_table) { cudaFree(c->d_ray_table); c->d_ray_table = nullptr; c->ray_table_cap = 0; }
            CU_TRY(c, cudaMalloc((void**)&c->d_ray_table, (size_t)cap * sizeof(uint32_t)));
            c->ray_table_cap = cap;
        }
        c->ray_table_mask = cap - 1;
        CU_TRY(c, cudaMemsetAsync(c->d_ray_table, 0, (size_t)cap * sizeof(uint32_t), c->stream));
        unsigned blocks = (n + 255) / 256;
        if (blocks > (unsigned)c->num_sms * 8) blocks = c->num_sms * 8;
        k_chunk_table<<<blocks, 256, 0, c->stream>>>(c->B().d_descs, n, c->d_ray_table, c->ray_table_mask);
        CU_TRY(c, cudaGetLastError());
        c->ray_table_serial = c->build_serial; c->ray_table_set = c->cur;
    }
    char* base = nullptr;
    const size_t vb = ((size_t)n_rays * 3 * sizeof(float) + 255) & ~(size_t)255;
    { uw_status st = scratch_get(c, 2 * vb + (size_t)n_rays * sizeof(float), &base); if (st != UW_OK) return st; }
    float* d_o = (float*)base; float* d_d = (float*)(base + vb); float* d_t = (float*)(base + 2 * vb);
    CU_TRY(c, cudaMemcpyAsync(d_o, origins, (size_t)n_rays * 3 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(d_d, dirs, (size_t)n_rays * 3 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    unsigned long long blocks = ((unsigned long long)n_rays * 32 + 255) / 256;
    if (blocks > (unsigned long long)c->num_sms * 8) blocks = (unsigned long long)c->num_sms * 8;
    k_raycast<<<(unsigned)blocks, 256, 0, c->stream>>>(c->dcfg, c->B().d_descs, c->B().d_tris, c->B().d_tri_cell, c->d_ray_table, c->ray_table_mask,
                                                      d_o, d_d, n_rays, wall_range, d_t);
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaMemcpyAsync(out_t, d_t, (size_t)n_rays * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return UW_OK;
}

extern "C" uw_status uw_debug_vertex_colors(uw_ctx* c, const float* world_z, const uint32_t* level, uint32_t n, float* out_rgb) {
    if (!c) return UW_ERR_INVALID;
    if ((!world_z || !level || !out_rgb) && n) return fail(c, UW_ERR_INVALID, "uw_debug_vertex_colors: null argument");
    if (n == 0) return UW_OK;
    CU_TRY(c, cudaSetDevice(c->device));
    char* base = nullptr;
    const size_t a = ((size_t)n * 4 + 255) & ~(size_t)255;
    { uw_status st = scratch_get(c, 2 * a + (size_t)n * 12, &base); if (st != UW_OK) return st; }
    float* d_z = (float*)base; uint32_t* d_l = (uint32_t*)(base + a); float* d_o = (float*)(base + 2 * a);
    CU_TRY(c, cudaMemcpyAsync(d_z, world_z, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(d_l, level, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    unsigned blocks = (n + 255) / 256;
    if (blocks > (unsigned)c->num_sms * 8) blocks = c->num_sms * 8;
    k_vertex_colors<<<blocks, 256, 0, c->stream>>>(c->dcfg, c->d_mc, d_z, d_l, n, d_o);
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaMemcpyAsync(out_rgb, d_o, (size_t)n * 12, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return UW_OK;
}

// Measurement aid: sustained FFMA rate of the device in TFLOP/s (2 FLOP per FFMA), CUDA-event timed.
extern "C" uw_status uw_debug_ffma_peak(uw_ctx* c, double* tflops) {
    if (!c || !tflops) return UW_ERR_INVALID;
    CU_TRY(c, cudaSetDevice(c->device));
    float* d_out = nullptr;
    CU_TRY(c, cudaMalloc(&d_out, 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 8192, blocks = c->num_sms * 8;
    k_ffma_peak<<<blocks, 256, 0, c->stream>>>(d_out, 64, 1.0001f, 1e-4f);          // warm-up
    cudaEventRecord(e0, c->stream);
    k_ffma_peak<<<blocks, 256, 0, c->stream>>>(d_out, iters, 1.0001f, 1e-4f);
    cudaEventRecord(e1, c->stream);
    cudaError_t err = cudaStreamSynchronize(c->stream);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d_out);
    if (err != cudaSuccess) return fail(c, UW_ERR_CUDA, std::string("uw_debug_ffma_peak: ") + cudaGetErrorString(err));
    *tflops = 2.0 * 16.0 * (double)iters * 256.0 * (double)blocks / ((double)ms * 1e-3) / 1e12;
    return UW_OK;
}

#ifdef UW_PHASE_TIMING
// debug builds only (-DUW_PHASE_TIMING): accumulated per-phase cycle counts of thread 0 of every CTA
extern "C" int uw_debug_phase_cycles(unsigned long long out[16], int reset) {
    if (cudaMemcpyFromSymbol(out, g_phase, sizeof(unsigned long long) * 16) != cudaSuccess) return 1;
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_phase, z, sizeof z); }
    return 0;
}
extern "C" int uw_debug_cta_times(unsigned long long* out /*1024*4*/) {
    return cudaMemcpyFromSymbol(out, g_cta, sizeof(unsigned long long) * 4096) != cudaSuccess;
}
#endif
