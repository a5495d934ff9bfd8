/* uwcuda.h -- C ABI of the B200-native chunk builder (libuwcuda.so).
 *
 * Drop-in boundary for UnderwaterWorld's chunk-build hot path.  The reference has NO
 * FFI/plugin interface for this path -- the path is plain Rust methods on `Chunk`
 * (underwater_world/src/chunk.rs:88-349) called from `World::build_full_step` /
 * `World::build_step` (src/world.rs:113-145).  The entry points below are therefore what a
 * Rust `extern "C"` shim for that path binds (see INTEGRATION.md for the shim); each cites
 * the reference interface it replaces.
 *
 * No torch / C++ types cross this boundary: plain pointers, sizes and POD structs only.
 * Errors: integer status; message through uw_last_error().  Nothing throws or unwinds.
 * There is no CPU fallback: every build call needs a CUDA device (sm_100a).
 *
 * Threading: a uw_ctx is not thread-safe -- one per caller thread (the reference is
 * single-threaded: README.md:19, src/lib.rs:88-90).
 */
#ifndef UWCUDA_H
#define UWCUDA_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UW_ABI_VERSION 2

typedef enum uw_status {
    UW_OK              = 0,
    UW_ERR_INVALID     = 1,   /* bad argument / unsupported configuration            */
    UW_ERR_CUDA        = 2,   /* a CUDA runtime call failed (see uw_last_error)       */
    UW_ERR_NO_DEVICE   = 3,   /* no usable CUDA device -- there is no CPU fallback    */
    UW_ERR_OOM         = 4,   /* host or device allocation failed                     */
    UW_ERR_NOT_READY   = 5,   /* uw_batch_view before uw_batch_wait on an async batch */
    UW_ERR_UNSUPPORTED = 6
} uw_status;

/* uw_config.flags */
#define UW_FLAG_EXACT_F64   0x1u  /* evaluate EVERY density sample with the f64 reference-order path
                                     (verification mode; default = FP32 fast path + f64 guard band) */
#define UW_FLAG_INDEX32     0x2u  /* emit u32 indices instead of u16 (forced when internal_size > 22, where
                                     the reference's `ind as u16` (chunk.rs:243) could wrap)        */
#define UW_FLAG_KEEP_DENSITIES 0x4u /* keep per-chunk densities/cases of the last batch readable
                                     through uw_batch_densities / uw_batch_cases (debug taps)      */
#define UW_FLAG_STAGED      0x10u /* run the four stages as separate kernels with densities materialised in
                                     HBM (per-stage profiling / debugging) instead of the fused single-pass
                                     kernel.  Results are identical.                                  */
#define UW_FLAG_ORDERED     0x20u /* packed arenas in REQUEST order (decoupled look-back over per-chunk
                                     aggregates; deterministic layout, chunk i+1 follows chunk i).  Default:
                                     completion order (one atomic bump allocation per chunk, no inter-CTA
                                     dependency, cost-ordered hand-out: 52 us against 116 us for 2048 chunks).  Every chunk's OWN buffers are identical in
                                     both modes; only vert_offset / index_offset differ.              */
#define UW_FLAG_ANALYTIC_SKIP 0x40u /* chunks whose z layer provably holds no surface (every octave of the noise is
                                     clamped to [-1, 1], so iso = terrace(z) + p cannot reach iso_level there) are
                                     answered without evaluating the noise: blank-early above, solid below.  The
                                     reference encodes the same fact as world::MIN_Z / MAX_Z (world.rs:11-12,161).
                                     Outputs are identical; off by default so that benchmarks evaluate every sample. */
#define UW_FLAG_EXPORTABLE  0x80u /* the packed vertex / index arenas are CUDA VMM allocations that can be handed to
                                     another API or process as POSIX file descriptors: uw_export_arena_fd()
                                     (mesh hand-off to the renderer without the host round trip)  */
#define UW_FLAG_TRIS        0x8u  /* also emit the per-cell collision triangle lists (chunk.rs:167-174,
                                     245-250) -- SURVEY §8f-1                                       */

/* Compile-time constants of the reference made runtime.  uw_config_default() fills the
 * reference's values: src/chunk.rs:5-17, src/world.rs:11-12. */
typedef struct uw_config {
    int32_t  internal_size;  /* INTERNAL_SIZE   chunk.rs:6   = 12 (cells per axis, S)          */
    int32_t  chunk_size;     /* CHUNK_SIZE      chunk.rs:5   = 16 (world units per chunk)      */
    uint32_t octaves;        /* PERLIN_OCTAVES  chunk.rs:9   = 3  (1..4 supported)             */
    float    iso_level;      /* ISO_LEVEL       chunk.rs:10  = -0.1                            */
    float    max_height;     /* MAX_HEIGHT      chunk.rs:11  = 32                              */
    float    adj_z_mod;      /* ADJ_Z_MOD       chunk.rs:12  = 0.25                            */
    float    min_hue;        /* MIN_HUE         chunk.rs:14  = -150                            */
    float    max_hue;        /* MAX_HUE         chunk.rs:15  = 60                              */
    float    saturation;     /* SATURATION      chunk.rs:16  = 0.6                             */
    float    base_value;     /* BASE_VALUE      chunk.rs:17  = 0.4                             */
    float    min_z;          /* world::MIN_Z as f32, world.rs:12 = -2                          */
    float    max_z;          /* world::MAX_Z as f32, world.rs:11 =  2                          */
    uint32_t seed;           /* noise::Perlin::new(seed), state.rs:358-359 (reference: wall clock) */
    int32_t  device;         /* CUDA device ordinal; -1 = current device                       */
    uint32_t flags;          /* UW_FLAG_*                                                      */
    float    guard_eps;      /* guard band: samples with |iso - iso_level| < guard_eps are
                                re-evaluated in f64 reference order.  0 -> default 1e-5        */
    uint32_t reserved[4];
} uw_config;

/* == draw::VertColor, src/draw.rs:4-9: #[repr(C)] {pos:[f32;3], color:[f32;3]}, stride 24 */
typedef struct uw_vert {
    float pos[3];
    float color[3];
} uw_vert;

/* == util::Tri, src/util.rs:7-10 (cgmath Vector3<f32> x4), 48 bytes */
typedef struct uw_tri {
    float verts[3][3];
    float normal[3];
} uw_tri;

/* uw_chunk_desc.flags */
#define UW_CHUNK_BLANK_EARLY  0x1u  /* early_blank_check() was true, chunk.rs:131-133,276-280      */
#define UW_CHUNK_HAS_MESH     0x2u  /* num_inds > 0  <=>  Chunk::not_blank(), chunk.rs:291,344     */
#define UW_CHUNK_U16_OVERFLOW 0x4u  /* vert_count > 65536: the u16 view would wrap (chunk.rs:243)  */

/* One per requested chunk, in request order.  Offsets index the batch-wide packed arrays;
 * index VALUES are chunk-local (start at 0), exactly as the reference's per-chunk buffers. */
typedef struct uw_chunk_desc {
    int32_t  pos[3];        /* chunk position as passed to Chunk::new, chunk.rs:89             */
    uint32_t flags;
    uint32_t vert_offset;   /* first vertex of this chunk in verts[]                           */
    uint32_t vert_count;    /* == build.verts.len()                                            */
    uint32_t index_offset;  /* first index of this chunk in inds16[] / inds32[]                */
    uint32_t index_count;   /* == Chunk::num_inds(), chunk.rs:348                              */
} uw_chunk_desc;

/* Host-side view of a finished batch.  Pointers are into a pinned host arena owned by the
 * batch and stay valid until uw_batch_free().  The caller copies out, which mirrors the
 * reference's create_buffer_init copy (chunk.rs:292-304). */
typedef struct uw_batch_view {
    uint32_t             n_chunks;
    uint64_t             n_verts;
    uint64_t             n_inds;
    const uw_chunk_desc* descs;    /* [n_chunks]                                               */
    const uw_vert*       verts;    /* [n_verts]                                                */
    const uint16_t*      inds16;   /* [n_inds]  NULL when only u32 indices were emitted        */
    const uint32_t*      inds32;   /* [n_inds]  NULL unless UW_FLAG_INDEX32 / internal_size>22 */
    const uw_tri*        tris;     /* [n_inds/3]: triangle t = indices 3t..3t+2 (chunk c's start at
                                      index_offset/3); NULL unless UW_FLAG_TRIS                */
    const uint16_t*      tri_cell_start; /* [n_chunks][S^3+1]: first triangle (chunk-local) of every cell in
                                      scan order x,y,z -- the reference's per-cell Vec<Tri>
                                      (chunk.rs:167-174); NULL unless UW_FLAG_TRIS             */
} uw_batch_view;

/* Device-resident result (no host copies): raw device pointers for interop / benchmarking.
 * Valid until the next build on the same context. */
typedef struct uw_device_view {
    uint32_t n_chunks;
    uint64_t n_verts;        /* valid after uw_sync()                                          */
    uint64_t n_inds;
    const void* d_descs;     /* uw_chunk_desc[n_chunks]                                        */
    const void* d_verts;     /* uw_vert[n_verts]                                               */
    const void* d_inds16;    /* uint16_t[n_inds] or NULL                                       */
    const void* d_inds32;    /* uint32_t[n_inds] or NULL                                       */
    const void* d_densities; /* float[n_chunks][density_stride], idx = x*L*L + y*L + z (chunk.rs:351-353); NULL on the
                                default (fused) path, which never materialises them -- set UW_FLAG_KEEP_DENSITIES
                                (or use the staged / large-chunk paths)                                      */
    uint32_t    density_stride; /* floats per chunk (>= L^3, padded to 16 B)                   */
} uw_device_view;

/* Per-stage device times of the last build (CUDA events on the context's stream). */
typedef struct uw_stage_times {
    float noise_ms;      /* K1  density sampling                 (chunk.rs:105-129)            */
    float classify_ms;   /* K2  blank/solid vote + case counts   (chunk.rs:131-133,141-164)    */
    float scan_ms;       /* K3  chunk-level prefix sums          (order of chunk.rs:233-243)   */
    float emit_ms;       /* K4  vertex + index emission          (chunk.rs:178-243)            */
    float total_ms;
    uint32_t launches;   /* kernels launched by the last build                                 */
} uw_stage_times;

typedef struct uw_ctx   uw_ctx;
typedef struct uw_batch uw_batch;

/* ---- lifecycle ------------------------------------------------------------------------ */
uint32_t    uw_abi_version(void);
void        uw_config_default(uw_config* cfg);                      /* chunk.rs:5-17 defaults, seed 0 */
uw_status   uw_create(const uw_config* cfg, uw_ctx** out);          /* replaces noise::Perlin::new (state.rs:359) + consts */
void        uw_destroy(uw_ctx* ctx);                                /* free every uw_batch of the context first: a batch's views live in the context's pinned pool */
const char* uw_last_error(const uw_ctx* ctx);                       /* ctx may be NULL: last create error */

/* The permutation table of noise::Perlin::new(cfg.seed) (noise-0.8.2). */
uw_status   uw_perm_table(const uw_ctx* ctx, uint8_t out[256]);

/* ---- the hot path: Chunk::new + Chunk::build_full for n chunks (chunk.rs:89-103,266-313) -- */
/* chunk_pos_xyz: n x 3 int32, HOST memory.  Blocking: returns with the batch complete. */
uw_status   uw_build(uw_ctx* ctx, const int32_t* chunk_pos_xyz, uint32_t n, uw_batch** out);
/* Same, but returns once work is enqueued; uw_batch_wait() completes it.  Up to TWO batches may be in flight on
 * the default (fused, internal_size 10/12) path: submit batch k+1, then wait on batch k -- batch k's copies to the
 * host run on a second stream underneath batch k+1's kernel.  A third submit (or a second one on the staged /
 * large-chunk paths, or any device-resident / debug call while a batch is in flight) returns UW_ERR_NOT_READY. */
uw_status   uw_build_async(uw_ctx* ctx, const int32_t* chunk_pos_xyz, uint32_t n, uw_batch** out);
uw_status   uw_batch_wait(uw_batch* b);
uw_status   uw_batch_view_get(const uw_batch* b, uw_batch_view* out);
void        uw_batch_free(uw_batch* b);

/* Device-resident variant: positions already on the device, outputs stay on the device.
 * Enqueues on the context's stream and does NOT synchronise; call uw_sync() before reading
 * n_verts / n_inds through uw_device_view_get(). */
uw_status   uw_build_device(uw_ctx* ctx, const int32_t* d_chunk_pos_xyz, uint32_t n);
uw_status   uw_sync(uw_ctx* ctx);
uw_status   uw_device_view_get(uw_ctx* ctx, uw_device_view* out);

/* ---- parity taps (debug): stage outputs for n chunks into HOST buffers ---------------- */
/* densities: n * L^3 floats, idx = x*L*L + y*L + z  (build.isos, chunk.rs:119,351-353) */
uw_status   uw_debug_densities(uw_ctx* ctx, const int32_t* chunk_pos_xyz, uint32_t n, float* out);
/* cases: n * S^3 bytes, cell scan order x,y,z (triangulation_idx, chunk.rs:155-162) */
uw_status   uw_debug_cases(uw_ctx* ctx, const int32_t* chunk_pos_xyz, uint32_t n, uint8_t* out);
/* Run ONLY the extraction stages (K2-K4) on caller-supplied HOST densities (n * L^3 floats).
 * Lets tests feed oracle densities and demand bit-exact topology. */
uw_status   uw_build_from_densities(uw_ctx* ctx, const int32_t* chunk_pos_xyz, const float* densities,
                                    uint32_t n, uw_batch** out);
/* Batched density point queries: perlin_util::iso_at (perlin_util.rs:24-29) on n points
 * (x,y,z f64 triples, HOST) -> n floats (HOST).  SURVEY §8f-3. */
uw_status   uw_iso_at(uw_ctx* ctx, const double* points_xyz, uint32_t n, float* out);

/* Batched collision ray casts against the per-cell triangle lists (SURVEY §8f-1's consumer; replaces the per-boid loops of
 * boid.rs:175-240).  Context created with UW_FLAG_TRIS; the rays are tested against the chunks of the context's LAST
 * build (host or device-resident), whose triangle lists are still in HBM.  For ray i (origin, direction: n x 3 floats,
 * HOST) the candidate triangles are exactly the ones the reference gathers: every built chunk within +-wall_range
 * world units of the origin (boid.rs:177-208), and inside it the cells within +-wall_range cells of the origin's
 * cell (Chunk::tris_around, chunk.rs:315-342).  Each is tested with util::Tri::intersects (util.rs:22-59) with
 * range = wall_range; out_t[i] = the smallest Some(t), or -1 if every test returned None.  The reference's two uses
 * follow from it: "heading for a collision" = (0 <= t < wall_range), "direction is safe" = (t == -1). */
uw_status   uw_raycast_tris(uw_ctx* ctx, const float* origins_xyz, const float* dirs_xyz, uint32_t n_rays, int32_t wall_range,
                            float* out_t);

/* Vertex colour of chunk.rs:215-222 (util.rs:93-95 create_mix_ratio, :122-153 hsv_to_rgb, :106-112 to_srgb) for n
 * (world z, value level = corner_b index % 3) pairs (HOST) -> n RGB triples (HOST): the kernels' own colour code,
 * so the tests can sweep the whole hue range instead of the z values a mesh happens to hold. */
uw_status   uw_debug_vertex_colors(uw_ctx* ctx, const float* world_z, const uint32_t* level, uint32_t n, float* out_rgb);

/* ---- renderer hand-off without the host round trip (replaces the create_buffer_init copies of
 * chunk.rs:291-305 for a renderer that can import external memory) ------------------------------------------
 * Context created with UW_FLAG_EXPORTABLE, after uw_build_device() + uw_sync(): exports the arena that
 * uw_device_view_get() describes (which: 0 = vertices, 1 = indices) as a POSIX file descriptor
 * (cuMemExportToShareableHandle).  Importers: Vulkan VK_KHR_external_memory_fd (VkImportMemoryFdInfoKHR with
 * VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT, allocationSize = *bytes), OpenGL EXT_memory_object_fd, or another
 * CUDA context / process (cuMemImportFromShareableHandle).  The caller owns the fd.  *bytes is the size of the whole
 * allocation (>= the used part); offsets inside it are the descriptor's vert_offset / index_offset.  A later
 * build that has to GROW the arena replaces the allocation: export again when the capacity (bytes) changed. */
uw_status   uw_export_arena_fd(uw_ctx* ctx, int which, int* fd, uint64_t* bytes);

/* ---- multi-GPU: one region, G GPUs of one box, finished meshes on the rendering GPU ---------------------------
 * SURVEY §8e / BASELINE configs[2].  Chunks are independent (chunk.rs:89-129: a chunk's output is a function of
 * pos, seed and constants only), so a region is cut into contiguous slabs of the position list -- x-slabs for the
 * x-major window order of World::update_nearby (world.rs:164-170) -- one per GPU, with NO collective on the compute
 * path.  The only cross-GPU traffic is the optional gather of the finished meshes to the GPU that draws
 * (state.rs:500-508 consumes them there).  It is FUSED into the build: the rendering GPU owns one arena per buffer
 * kind, cut into one segment per producer; a producer GPU's fused kernel stores its vertices, indices and
 * descriptors straight into its segment through NVLink peer addresses while it computes (no staging copy, no
 * second pass), and its last CTA publishes a per-segment head {totals, epoch} after a system-wide fence.  The
 * consumer waits for the heads on its own stream (uw_gather_wait) -- one-sided, no rendezvous, no NCCL.
 *
 *   render process/GPU : uw_gather_create  -> uw_gather_info (plain bytes; hand it to the producers)
 *   every producer     : uw_gather_attach(info, segment)          (same process: peer access; other process: CUDA IPC)
 *                        uw_gather_build(pos, n, first_chunk)      async; outputs land in the render GPU's arenas
 *   render process/GPU : uw_gather_wait    -> uw_gather_result     the k-th wait returns when every segment's k-th
 *                                                                  build has landed
 * Descriptor i of the arena belongs to request chunk i (first_chunk + local index); its vert_offset / index_offset
 * are element offsets into the arena (segment s starts at s * seg_vcap / s * seg_icap), index VALUES stay
 * chunk-local as everywhere else.  Fused path only (internal_size 10 / 12, no UW_FLAG_TRIS / STAGED /
 * KEEP_DENSITIES); a segment that overflows its capacity fails the producer's uw_sync with UW_ERR_OOM.
 * uw_multi_* below drives all of this from ONE process (what a Rust `World` would call). */
#define UW_MAX_SEGMENTS 16
#define UW_GATHER_DESCS_TO_HOST 0x1u   /* uw_gather_wait also copies ALL descriptors (request order) to pinned host memory      */
#define UW_GATHER_DRAW_TO_HOST  0x2u   /* ... copies the DRAW LIST: the descriptors of the chunks that ended with a mesh only
                                          (Chunk::not_blank, what World::build_full_step keeps for rendering, world.rs:117-121;
                                          8 % of the chunks of a terrain region), per segment in completion order          */

typedef struct uw_gather_info {
    uint32_t abi_version, n_segments;
    int32_t  device;            /* render device ordinal (as seen by the owner process)                     */
    uint32_t index_bytes;       /* 2 or 4                                                                   */
    uint64_t owner_pid;
    uint64_t base, bytes;       /* the arena allocation: owner-process device address, size                 */
    uint64_t off_head, off_descs, off_verts, off_inds, off_draw;   /* byte offsets of the five parts inside it */
    uint64_t n_chunks;          /* descriptor capacity (chunks of the whole region)                         */
    uint64_t seg_vcap, seg_icap;/* capacity of ONE segment, in vertices / indices                           */
    uint8_t  ipc_handle[64];    /* cudaIpcMemHandle_t of the allocation                                     */
} uw_gather_info;

typedef struct uw_gather_segment {
    uint64_t first_chunk;       /* request index of the segment's first chunk                               */
    uint32_t n_chunks, n_mesh, n_blank, overflow;
    uint64_t n_verts, n_inds;   /* elements used in the segment (vertex allocations are padded to even counts) */
    uint64_t guard;             /* f64 guard-band re-evaluations                                            */
} uw_gather_segment;

typedef struct uw_gather_result {
    uint32_t n_segments, epoch;
    uint64_t n_chunks, n_verts, n_inds;                      /* sums over the segments                       */
    const void* d_descs;        /* uw_chunk_desc[n_chunks capacity], render GPU                             */
    const void* d_verts;        /* uw_vert[n_segments * seg_vcap]                                           */
    const void* d_inds;         /* index_bytes * [n_segments * seg_icap]                                    */
    uint64_t seg_vcap, seg_icap;
    const uw_chunk_desc* h_descs; /* pinned host copy of the descriptors (UW_GATHER_DESCS_TO_HOST), else NULL;
                                     valid until the next uw_gather_wait                                    */
    const uw_chunk_desc* h_draw;  /* pinned host copy of the draw list (UW_GATHER_DRAW_TO_HOST), else NULL: segment 0's
                                     n_mesh entries, then segment 1's, ...                                  */
    const void* d_draw;           /* the draw list on the render GPU: uw_chunk_desc[n_segments][n_chunks capacity]      */
    uint64_t n_draw;              /* sum of seg[].n_mesh                                                    */
    uw_gather_segment seg[UW_MAX_SEGMENTS];
} uw_gather_result;

/* seg_vcap / seg_icap = 0: sized for ceil(n_chunks / n_segments) chunks per segment like the library's own arenas. */
uw_status   uw_gather_create(uw_ctx* render_ctx, uint32_t n_segments, uint64_t n_chunks, uint64_t seg_vcap, uint64_t seg_icap,
                             uw_gather_info* out);
uw_status   uw_gather_destroy(uw_ctx* render_ctx);           /* producers detach first                        */
uw_status   uw_gather_attach(uw_ctx* ctx, const uw_gather_info* info, uint32_t segment);
uw_status   uw_gather_detach(uw_ctx* ctx);
/* Chunk::new + build_full for n chunks (HOST positions; request indices first_chunk .. first_chunk + n) into the
 * attached segment.  Returns once the pinned H2D copy and the kernel are enqueued; uw_sync() completes it locally.
 * If chunk_pos_xyz is itself page-locked memory the copy reads it directly: keep it unchanged until then. */
uw_status   uw_gather_build(uw_ctx* ctx, const int32_t* chunk_pos_xyz, uint32_t n, uint64_t first_chunk);
/* Same with positions already on the producer's device. */
uw_status   uw_gather_build_device(uw_ctx* ctx, const int32_t* d_chunk_pos_xyz, uint32_t n, uint64_t first_chunk);
uw_status   uw_gather_wait(uw_ctx* render_ctx, uint32_t flags, uw_gather_result* out);

/* Contiguous slab `part` of `parts` of a list of n chunks (remainder spread one per slab) -- the partition
 * uw_multi_build and bench.py use. */
void        uw_slab_bounds(uint32_t n, uint32_t parts, uint32_t part, uint32_t* first, uint32_t* count);
/* Gather-aware partition: the rendering GPU's own output does not cross NVLink, so when the gather is bound by the
 * rendering GPU's ingress (8 GPUs: 706 MB into one GPU) its slab should be larger.  Slab `render_part` holds
 * render_permille / 1000 of the n chunks (never less than an even share), the other slabs share the rest evenly; slabs
 * stay contiguous and in part order.  render_permille = 0: the even split of uw_slab_bounds.  uw_multi_build chooses the
 * share by a search over successive requests (uw_share_search_next); gather.RegionGather.tune measures a handful of
 * candidate shares during warm-up for one process per GPU. */
void        uw_slab_bounds_weighted(uint32_t n, uint32_t parts, uint32_t part, uint32_t render_part, uint32_t render_permille,
                                    uint32_t* first, uint32_t* count);

/* Choosing the rendering GPU's share over successive requests (what uw_multi_build does internally; exported so that
 * other drivers -- one process per GPU -- can run the same search and so that it can be tested without a GPU).
 * A damped hill climb on the measured cost of a whole request (any unit that is comparable between requests, e.g.
 * seconds per chunk of the slowest GPU): start at the even split, grow the share while the cost falls, turn round and
 * halve the step when it rises, settle on the cheapest share seen once the step is down to 1/64 of an even share, and
 * start over if the cost at that share later rises by more than 10 % (another workload).  Equalising the GPUs' kernel times instead stops short of the optimum, because the
 * rendering GPU's own kernel is slowed by the traffic arriving over NVLink (profiles/r02_gather_scaling.txt).
 * Zero-initialise the state; uw_share_search_next records the cost measured AT state->share (the first two costs are
 * discarded: cold requests) and returns the share to use for the next request (even split <= share <= 1/2; 0 when parts < 2). */
typedef struct uw_share_search {
    double   share;       /* share the next cost will be measured at; 0 = not started (even split)            */
    double   step;        /* current step                                                                     */
    double   last_cost;   /* cost at the previous share; 0 = none yet                                         */
    double   best_share;  /* cheapest share seen so far ...                                                   */
    double   best_cost;   /* ... and its cost                                                                 */
    int32_t  dir;         /* +1 / -1                                                                           */
    uint32_t moves;       /* requests seen                                                                     */
    uint32_t settled;     /* 1 = the step has shrunk to 1/64 of an even share: holding best_share             */
    uint32_t reserved;
} uw_share_search;
double      uw_share_search_next(uw_share_search* state, uint32_t parts, double cost);

/* One process, G GPUs: devices[0] renders.  uw_multi_build = World::build_full_step for a whole region
 * (world.rs:113-123) -- slabs, G fused launches (each GPU's H2D + kernel on its own stream), meshes gathered into
 * devices[0]'s arenas as they are produced, wait.  flags: UW_GATHER_DESCS_TO_HOST.  The result's pointers stay valid
 * until the next uw_multi_build / uw_multi_destroy. */
typedef struct uw_multi uw_multi;
uw_status   uw_multi_create(const uw_config* cfg, const int32_t* devices, uint32_t n_devices, uw_multi** out);
uw_status   uw_multi_build(uw_multi* m, const int32_t* chunk_pos_xyz, uint32_t n, uint32_t flags, uw_gather_result* out);
/* The rendering GPU's current share of a request in 1/1000 (0 = even split so far); adapts over successive builds. */
uint32_t    uw_multi_render_share(const uw_multi* m);
void        uw_multi_destroy(uw_multi* m);
const char* uw_multi_last_error(const uw_multi* m);           /* m may be NULL: last create error             */
/* Verification aid: blocking copy of `bytes` of device memory (any device of the process: unified addressing) to the
 * host -- how the tests read a gather arena back. */
uw_status   uw_debug_copy_to_host(const void* d_src, uint64_t bytes, void* h_dst);

/* ---- plumbing ------------------------------------------------------------------------- */
/* Run on an existing CUDA stream (cudaStream_t as void*), e.g. torch's current stream. */
uw_status   uw_set_stream(uw_ctx* ctx, void* cuda_stream);
uw_status   uw_get_stage_times(uw_ctx* ctx, uw_stage_times* out);
/* Enable/disable per-stage event timing (adds event records between stages). Default off. */
uw_status   uw_set_profiling(uw_ctx* ctx, int enabled);
/* Number of f64 guard-band re-evaluations in the last build (valid after sync). */
uw_status   uw_get_guard_count(uw_ctx* ctx, uint64_t* out);
/* Measurement aid: sustained FFMA rate of the device (TFLOP/s, 2 FLOP per FFMA) -- the "measured FP32
 * peak" beside the nominal 148 x 128 x 2 x 1.965 GHz = 74.4 TFLOP/s in the noise-stage roofline. */
uw_status   uw_debug_ffma_peak(uw_ctx* ctx, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* UWCUDA_H */
