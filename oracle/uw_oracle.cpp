// uw_oracle.cpp -- CPU restatement of UnderwaterWorld's chunk-build hot path.
//
// *** TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH. ***
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library.  The product (libuwcuda.so) never links or calls it.
//
// PARITY PIN STATUS: the reference has no tests or golden vectors, and
// cannot be compiled here (no Rust toolchain).  This restatement is pinned instead against
// OUTPUTS OF THE REFERENCE'S OWN SHIPPED BINARY (builds/web_build.zip, wasm), executed in
// the build container by oracle/wasm_forensics.py; the resulting vectors are committed in
// tests/golden/ (see oracle/gen_golden.py and DESIGN.md "Oracle pinning" for exactly which
// functions are pinned that way and which are not).
//
// Build: g++ -O3 -std=c++17 -ffp-contract=off -fno-fast-math -shared -fPIC  (see oracle/Makefile)
// -ffp-contract=off is REQUIRED: neither wasm nor default Rust codegen fuses mul+add.
//
// Everything here follows, by file:line,
//   /root/reference/underwater_world/src/chunk.rs, perlin_util.rs, marching_table.rs,
//   util.rs, draw.rs, world.rs, state.rs
// and, for the third-party crate noise-0.8.2 (+ rand-0.7.3 / rand_xorshift, un-vendored;
// Cargo.toml:20), the arithmetic recovered from the reference's shipped wasm
// (function 466 = noise::core::perlin::perlin_3d).

#include <stdint.h>
#include <stddef.h>
#include <math.h>
#include <string.h>
#include <array>
#include <atomic>
#include <chrono>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

#include "mc_tables_oracle.h"

extern "C" {

// Constants of src/chunk.rs:5-17 and src/world.rs:11-12, made runtime so that the same
// oracle serves S=12 (HEAD), S=10 (the shipped wasm) and S=64 (BASELINE config 4).
typedef struct uwo_config {
    int32_t  internal_size;   // INTERNAL_SIZE   chunk.rs:6   (12)
    int32_t  chunk_size;      // CHUNK_SIZE      chunk.rs:5   (16)
    uint32_t octaves;         // PERLIN_OCTAVES  chunk.rs:9   (3)
    float    iso_level;       // ISO_LEVEL       chunk.rs:10  (-0.1)
    float    max_height;      // MAX_HEIGHT      chunk.rs:11  (32)
    float    adj_z_mod;       // ADJ_Z_MOD       chunk.rs:12  (0.25)
    float    min_hue, max_hue, saturation, base_value;  // chunk.rs:14-17
    float    min_z, max_z;    // world.rs:11-12 as f32 (-2, 2)
} uwo_config;

typedef struct uwo_vert { float pos[3]; float color[3]; } uwo_vert;   // draw.rs:4-9 (24 B)
typedef struct uwo_tri  { float verts[3][3]; float normal[3]; } uwo_tri;  // util.rs:7-10 (48 B)

void uwo_config_default(uwo_config* c) {
    c->internal_size = 12; c->chunk_size = 16; c->octaves = 3;
    c->iso_level = -0.1f; c->max_height = 32.0f; c->adj_z_mod = 0.25f;
    c->min_hue = -150.0f; c->max_hue = 60.0f; c->saturation = 0.6f; c->base_value = 0.4f;
    c->min_z = -2.0f; c->max_z = 2.0f;
}

// ---------------------------------------------------------------------------------------
// noise::Perlin::new(seed) -> PermutationTable  (state.rs:359; SURVEY App. A.1,
// wasm@258850-259541).  XorShift128 seeded with bytes [1,0,0,0, seed LE x3]; identity
// table shuffled by rand-0.7.3 SliceRandom::shuffle: for i = 255..1 swap(i, gen_range(0,i+1))
// with the u32 widening-multiply rejection sampler.
// ---------------------------------------------------------------------------------------
void uwo_perm_table(uint32_t seed, uint8_t out[256]) {
    uint32_t x = 1u, y = seed, z = seed, w = seed;
    for (int i = 0; i < 256; ++i) out[i] = (uint8_t)i;
    for (uint32_t i = 255; i >= 1; --i) {
        const uint32_t range = i + 1u;
        const uint32_t zone = (range << __builtin_clz(range)) - 1u;
        uint32_t j;
        for (;;) {
            const uint32_t t = x ^ (x << 11);
            x = y; y = z; z = w;
            w = w ^ (w >> 19) ^ t ^ (t >> 8);
            const uint64_t m = (uint64_t)w * (uint64_t)range;
            if ((uint32_t)m <= zone) { j = (uint32_t)(m >> 32); break; }
        }
        const uint8_t tmp = out[i]; out[i] = out[j]; out[j] = tmp;
    }
}

// hash = perm[perm[perm[x&255] ^ (y&255)] ^ (z&255)]
static inline uint32_t hash3(const uint8_t* perm, int32_t ix, int32_t iy, int32_t iz) {
    return perm[perm[perm[ix & 255] ^ (iy & 255)] ^ (iz & 255)];
}

// 12-gradient dot decoded from the wasm br_table (h & 15).
static inline double grad3(uint32_t h, double x, double y, double z) {
    switch (h & 15u) {
        case 0: case 12: return x + y;
        case 1: case 13: return y - x;
        case 2:          return x - y;
        case 3:          return (-x) - y;
        case 4:          return x + z;
        case 5:          return z - x;
        case 6:          return x - z;
        case 7:          return (-x) - z;
        case 8:          return y + z;
        case 9: case 14: return z - y;
        case 10:         return y - z;
        default:         return (-y) - z;   // 11, 15
    }
}

static inline double fade5(double t) {
    double c = t < 0.0 ? 0.0 : t;
    c = c > 1.0 ? 1.0 : c;
    return (c * c * c) * (c * (c * 6.0 + (-15.0)) + 10.0);
}

// noise-0.8.2 core::perlin::perlin_3d (call site perlin_util.rs:13; SURVEY App. A.4,
// wasm func 466).  All f64, no fused operations.
double uwo_perlin3(const uint8_t* perm, double px, double py, double pz) {
    const double fx = floor(px), fy = floor(py), fz = floor(pz);
    const double dx = px - fx, dy = py - fy, dz = pz - fz;
    const int32_t ix = (int32_t)fx, iy = (int32_t)fy, iz = (int32_t)fz;
    const double dx1 = dx + (-1.0), dy1 = dy + (-1.0), dz1 = dz + (-1.0);

    const double g000 = grad3(hash3(perm, ix,     iy,     iz    ), dx,  dy,  dz );
    const double g100 = grad3(hash3(perm, ix + 1, iy,     iz    ), dx1, dy,  dz );
    const double g010 = grad3(hash3(perm, ix,     iy + 1, iz    ), dx,  dy1, dz );
    const double g110 = grad3(hash3(perm, ix + 1, iy + 1, iz    ), dx1, dy1, dz );
    const double g001 = grad3(hash3(perm, ix,     iy,     iz + 1), dx,  dy,  dz1);
    const double g101 = grad3(hash3(perm, ix + 1, iy,     iz + 1), dx1, dy,  dz1);
    const double g011 = grad3(hash3(perm, ix,     iy + 1, iz + 1), dx,  dy1, dz1);
    const double g111 = grad3(hash3(perm, ix + 1, iy + 1, iz + 1), dx1, dy1, dz1);

    const double a = fade5(dx), b = fade5(dy), c = fade5(dz);

    const double k0 = g000;
    const double k1 = g100 - g000;
    const double k2 = g010 - g000;
    const double k3 = g001 - g000;
    const double k4 = ((g000 + g110) - g100) - g010;
    const double k5 = ((g000 + g101) - g100) - g001;
    const double k6 = ((g000 + g011) - g010) - g001;
    const double k7 = ((((((g100 + g010) + g001) + g111) - g000) - g110) - g101) - g011;

    double r = ((((((k0 + k1 * a) + k2 * b) + k3 * c) + (k4 * a) * b) + (k5 * a) * c) + (k6 * b) * c)
               + ((k7 * a) * b) * c;
    r = r * 1.1547005383792515;   // 2/sqrt(3)
    return r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
}

// perlin_util.rs:6-22
double uwo_perlin3_octaves(const uint8_t* perm, double x, double y, double z, uint32_t octaves) {
    double total = 0.0, frequency = 1.0, amplitude = 1.0, max_value = 0.0;
    for (uint32_t o = 0; o < octaves; ++o) {
        total += uwo_perlin3(perm, x * frequency, y * frequency, z * frequency) * amplitude;
        max_value += amplitude;
        amplitude *= 0.5;
        frequency *= 2.0;
    }
    return total / max_value;
}

// perlin_util.rs:24-29.  Rust `%` on f32 is fmodf; `a + p - m` parses as (a + p) - m.
float uwo_iso_at(const uwo_config* c, const uint8_t* perm, double x, double y, double z) {
    const float p = (float)uwo_perlin3_octaves(perm, x, y, z, c->octaves);
    const float adj_z = ((float)z * (float)c->chunk_size) / c->max_height;
    return (adj_z + p) - fmodf(adj_z, c->adj_z_mod);
}

// chunk.rs:105-129 -- density lattice, idx = x*L*L + y*L + z (chunk.rs:351-353)
void uwo_densities(const uwo_config* c, const uint8_t* perm, const int32_t pos[3], float* isos) {
    const int L = c->internal_size + 1;
    const float size_scale = (float)c->chunk_size / (float)c->internal_size;   // chunk.rs:7 (f32)
    const int32_t off[3] = { pos[0] * c->chunk_size, pos[1] * c->chunk_size, pos[2] * c->chunk_size };  // chunk.rs:90-94
    const double cs = (double)c->chunk_size;
    size_t k = 0;
    for (int x = 0; x < L; ++x) {
        const double lx = (double)x * (double)size_scale;
        const double px = (lx + (double)off[0]) / cs;
        for (int y = 0; y < L; ++y) {
            const double ly = (double)y * (double)size_scale;
            const double py = (ly + (double)off[1]) / cs;
            for (int z = 0; z < L; ++z) {
                const double lz = (double)z * (double)size_scale;
                const double pz = (lz + (double)off[2]) / cs;
                isos[k++] = uwo_iso_at(c, perm, px, py, pz);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// colour helpers, util.rs:93-95, 106-112, 122-153 (all f32, unfused)
// ---------------------------------------------------------------------------------------
static inline float rem_euclid_f32(float a, float b) {   // Rust f32::rem_euclid
    const float r = fmodf(a, b);
    return r < 0.0f ? r + fabsf(b) : r;
}

void uwo_hsv_to_rgb(float hue_in, float saturation, float value, float out[3]) {
    const float hue = rem_euclid_f32(hue_in, 360.0f);
    const float c = value * saturation;
    const float h = hue / 60.0f;
    const float x = c * (1.0f - fabsf(fmodf(h, 2.0f) - 1.0f));
    const float m = value - c;
    float r, g, b;
    if      (0.0f <= h && h < 1.0f) { r = c;    g = x;    b = 0.0f; }
    else if (1.0f <= h && h < 2.0f) { r = x;    g = c;    b = 0.0f; }
    else if (2.0f <= h && h < 3.0f) { r = 0.0f; g = c;    b = x;    }
    else if (3.0f <= h && h < 4.0f) { r = 0.0f; g = x;    b = c;    }
    else if (4.0f <= h && h < 5.0f) { r = x;    g = 0.0f; b = c;    }
    else                            { r = c;    g = 0.0f; b = x;    }
    out[0] = (r + m) * 255.0f; out[1] = (g + m) * 255.0f; out[2] = (b + m) * 255.0f;
}

void uwo_to_srgb(const float in[3], float out[3]) {
    for (int i = 0; i < 3; ++i) out[i] = powf((in[i] / 255.0f + 0.055f) / 1.055f, 2.4f);
}

// chunk.rs:215-222
void uwo_vertex_color(const uwo_config* c, float world_z, uint32_t corner_b_idx, float out[3]) {
    const float world_z_ratio = world_z / (float)c->chunk_size;
    const float mix_ratio = (world_z_ratio - c->min_z) / (c->max_z - c->min_z);   // util.rs:93-95
    const float value_intensity = (float)(corner_b_idx % 3u) / 9.0f;
    const float hue = c->min_hue + (c->max_hue - c->min_hue) * mix_ratio;
    float rgb[3];
    uwo_hsv_to_rgb(hue, c->saturation, c->base_value + value_intensity, rgb);
    uwo_to_srgb(rgb, out);
}

// util.rs:12-21,61-64 (cgmath cross / magnitude / div, f32)
static inline void tri_new(const float v[3][3], uwo_tri* t) {
    memcpy(t->verts, v, sizeof(float) * 9);
    const float e1[3] = { v[1][0] - v[0][0], v[1][1] - v[0][1], v[1][2] - v[0][2] };
    const float e2[3] = { v[2][0] - v[0][0], v[2][1] - v[0][1], v[2][2] - v[0][2] };
    const float n[3] = { e1[1] * e2[2] - e1[2] * e2[1],
                         e1[2] * e2[0] - e1[0] * e2[2],
                         e1[0] * e2[1] - e1[1] * e2[0] };
    const float mag = sqrtf((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]);
    if (mag == 0.0f) { t->normal[0] = n[0]; t->normal[1] = n[1]; t->normal[2] = n[2]; }
    else { t->normal[0] = n[0] / mag; t->normal[1] = n[1] / mag; t->normal[2] = n[2] / mag; }
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// build_mesh, chunk.rs:135-264.  Corner order chunk.rs:144-153.
// ---------------------------------------------------------------------------------------
static const int CORNER_OFF[8][3] = {
    {0,0,0}, {1,0,0}, {1,0,1}, {0,0,1}, {0,1,0}, {1,1,0}, {1,1,1}, {0,1,1}
};

struct PairKey { size_t a[3]; size_t b[3]; };

struct CellKey { size_t x, y, z; bool operator==(const CellKey& o) const { return x == o.x && y == o.y && z == o.z; } };
struct CellKeyHash {   // stand-in for Rust's SipHash-1-3 (same role: hash 3 usize); only the cost profile matters
    size_t operator()(const CellKey& k) const {
        uint64_t h = 0x9E3779B97F4A7C15ull;
        for (uint64_t v : { (uint64_t)k.x, (uint64_t)k.y, (uint64_t)k.z }) {
            h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
            h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 31;
        }
        return (size_t)h;
    }
};

struct MeshOut {
    std::vector<uwo_vert> verts;
    std::vector<uint32_t> inds;     // un-truncated index (reference casts `ind as u16`, chunk.rs:243)
    std::vector<uint8_t>  cases;    // per cell, scan order
    std::vector<uwo_tri>  tris;     // flattened per-cell triangle lists (cell scan order)
    std::vector<uint32_t> tri_cell_start;  // S^3 + 1 offsets into tris
    std::vector<uint32_t> vpairs;   // per vertex: lattice index of corner_a, of corner_b (the reference's vert_pairs, chunk.rs:38)
};

// mode 0: FAITHFUL -- linear-search ordered-pair dedup (chunk.rs:233), colour per index
//         (chunk.rs:215-222), per-cell Tri lists inserted into a hash map (chunk.rs:167-174,245-250)
// mode 1: FAST     -- identical outputs, O(1) dedup through a directed-edge map; no Tri lists
static void build_mesh(const uwo_config* c, const int32_t pos[3], const float* isos, int mode,
                       bool want_tris, MeshOut& out) {
    const size_t S = (size_t)c->internal_size, L = S + 1;
    const float size_scale = (float)c->chunk_size / (float)c->internal_size;
    const int32_t off[3] = { pos[0] * c->chunk_size, pos[1] * c->chunk_size, pos[2] * c->chunk_size };
    const float iso_level = c->iso_level;

    std::vector<PairKey> vert_pairs;
    std::vector<int32_t> edge_map;   // mode 1: iso_idx_a * 6 + dir -> vertex index
    if (mode == 1) edge_map.assign(L * L * L * 6, -1);
    std::unordered_map<CellKey, std::vector<uwo_tri>, CellKeyHash> tris;

    out.cases.assign(S * S * S, 0);
    if (want_tris) out.tri_cell_start.assign(S * S * S + 1, 0);

    size_t cell = 0;
    for (size_t x = 0; x < S; ++x)
    for (size_t y = 0; y < S; ++y)
    for (size_t z = 0; z < S; ++z, ++cell) {
        size_t corners[8][3];
        for (int i = 0; i < 8; ++i) {
            corners[i][0] = x + CORNER_OFF[i][0]; corners[i][1] = y + CORNER_OFF[i][1]; corners[i][2] = z + CORNER_OFF[i][2];
        }
        unsigned tri_idx = 0;
        for (int i = 0; i < 8; ++i) {
            const float iso = isos[corners[i][0] * L * L + corners[i][1] * L + corners[i][2]];
            if (iso < iso_level) tri_idx |= 1u << i;     // strict <, chunk.rs:159
        }
        out.cases[cell] = (uint8_t)tri_idx;
        const uint64_t row = UWO_TRI_ROWS[tri_idx];

        std::vector<uwo_tri> pos_tris;
        if (mode == 0) pos_tris.reserve(16 / 3);
        float cur[3][3];

        for (int i = 0; i < 16; ++i) {
            const unsigned e = (unsigned)((row >> (4 * i)) & 0xF);
            if (e == 0xF) {
                if (mode == 0) tris[CellKey{x, y, z}] = std::move(pos_tris);   // chunk.rs:173-174
                break;
            }
            const unsigned ca = UWO_EDGE_CORNERS[e][0], cb = UWO_EDGE_CORNERS[e][1];
            const size_t* A = corners[ca];
            const size_t* B = corners[cb];
            const float sa[3] = { (float)A[0] * size_scale, (float)A[1] * size_scale, (float)A[2] * size_scale };
            const float sb[3] = { (float)B[0] * size_scale, (float)B[1] * size_scale, (float)B[2] * size_scale };
            const size_t ia = A[0] * L * L + A[1] * L + A[2];
            const size_t ib = B[0] * L * L + B[1] * L + B[2];
            const float iso_a = isos[ia], iso_b = isos[ib];
            const float t = (iso_level - iso_a) / (iso_b - iso_a);
            const float diff[3] = { sb[0] - sa[0], sb[1] - sa[1], sb[2] - sa[2] };
            const float mid[3] = { sa[0] + t * diff[0], sa[1] + t * diff[1], sa[2] + t * diff[2] };
            const float world_z = mid[2] + (float)off[2];

            uwo_vert v;
            v.pos[0] = mid[0] + (float)off[0]; v.pos[1] = mid[1] + (float)off[1]; v.pos[2] = world_z;

            size_t ind;
            if (mode == 0) {
                uwo_vertex_color(c, world_z, cb, v.color);   // computed for EVERY index, chunk.rs:215-222
                ind = vert_pairs.size();
                for (size_t k = 0; k < vert_pairs.size(); ++k) {   // chunk.rs:233
                    const PairKey& p = vert_pairs[k];
                    if (p.a[0] == A[0] && p.a[1] == A[1] && p.a[2] == A[2] &&
                        p.b[0] == B[0] && p.b[1] == B[1] && p.b[2] == B[2]) { ind = k; break; }
                }
                if (ind == vert_pairs.size()) {
                    vert_pairs.push_back(PairKey{{A[0], A[1], A[2]}, {B[0], B[1], B[2]}});
                    out.verts.push_back(v);
                    out.vpairs.push_back((uint32_t)ia); out.vpairs.push_back((uint32_t)ib);
                }
            } else {
                // direction code of the ordered pair: +x,-x,+y,-y(unused),+z,-z
                int dir;
                if (B[0] != A[0]) dir = B[0] > A[0] ? 0 : 1;
                else if (B[1] != A[1]) dir = B[1] > A[1] ? 2 : 3;
                else dir = B[2] > A[2] ? 4 : 5;
                int32_t& slot = edge_map[ia * 6 + dir];
                if (slot < 0) {
                    uwo_vertex_color(c, world_z, cb, v.color);
                    slot = (int32_t)out.verts.size();
                    out.verts.push_back(v);
                    out.vpairs.push_back((uint32_t)ia); out.vpairs.push_back((uint32_t)ib);
                }
                ind = (size_t)slot;
            }
            out.inds.push_back((uint32_t)ind);

            if (mode == 0 || want_tris) {
                memcpy(cur[i % 3], v.pos, sizeof(float) * 3);
                if (i % 3 == 2) {
                    uwo_tri tr; tri_new(cur, &tr);
                    if (mode == 0) pos_tris.push_back(tr);
                    if (want_tris) out.tris.push_back(tr);
                }
            }
        }
        if (want_tris) out.tri_cell_start[cell + 1] = (uint32_t)out.tris.size();
    }
}

// ---------------------------------------------------------------------------------------
// RULE mode: the deterministic parallel numbering of SURVEY.md App. B.4 (static edge
// ownership + prefix sums), executed serially.  Exists so that CPU tests can prove the rule
// the CUDA kernels implement reproduces the reference's first-seen numbering.
// Written independently of the product's derived tables (recomputes everything from rows).
// ---------------------------------------------------------------------------------------
static void owner_of(unsigned e, size_t cx, size_t cy, size_t cz, size_t& ox, size_t& oy, size_t& oz, unsigned& oe) {
    ox = cx; oy = cy; oz = cz; oe = e;
    switch (e) {
        case 4: case 5: case 6: case 7: case 10: return;
        case 0: case 1: case 2: case 3:
            if (cy > 0) { oy = cy - 1; oe = e + 4; } return;
        case 9:  if (cz > 0) { oz = cz - 1; oe = 10; } return;
        case 11: if (cx > 0) { ox = cx - 1; oe = 10; } return;
        case 8:
            if (cx > 0 && cz > 0) { ox = cx - 1; oz = cz - 1; oe = 10; }
            else if (cx > 0)      { ox = cx - 1; oe = 9; }
            else if (cz > 0)      { oz = cz - 1; oe = 11; }
            return;
    }
}

static void build_mesh_rule(const uwo_config* c, const int32_t pos[3], const float* isos, MeshOut& out) {
    const size_t S = (size_t)c->internal_size, L = S + 1, NC = S * S * S;
    const float size_scale = (float)c->chunk_size / (float)c->internal_size;
    const int32_t off[3] = { pos[0] * c->chunk_size, pos[1] * c->chunk_size, pos[2] * c->chunk_size };
    out.cases.assign(NC, 0);
    std::vector<uint32_t> vbase(NC + 1, 0), ibase(NC + 1, 0);
    auto cell_id = [&](size_t x, size_t y, size_t z) { return (x * S + y) * S + z; };
    // pass 1: classify + counts
    for (size_t x = 0; x < S; ++x) for (size_t y = 0; y < S; ++y) for (size_t z = 0; z < S; ++z) {
        unsigned cs = 0;
        for (int i = 0; i < 8; ++i) {
            const float iso = isos[(x + CORNER_OFF[i][0]) * L * L + (y + CORNER_OFF[i][1]) * L + (z + CORNER_OFF[i][2])];
            if (iso < c->iso_level) cs |= 1u << i;
        }
        const size_t id = cell_id(x, y, z);
        out.cases[id] = (uint8_t)cs;
        const uint64_t row = UWO_TRI_ROWS[cs];
        unsigned seen = 0, nown = 0, nidx = 0;
        for (int i = 0; i < 16; ++i) {
            const unsigned e = (unsigned)((row >> (4 * i)) & 0xF);
            if (e == 0xF) break;
            ++nidx;
            if (seen >> e & 1) continue;
            seen |= 1u << e;
            size_t ox, oy, oz; unsigned oe; owner_of(e, x, y, z, ox, oy, oz, oe);
            if (ox == x && oy == y && oz == z) ++nown;
        }
        vbase[id + 1] = nown; ibase[id + 1] = nidx;
    }
    for (size_t i = 0; i < NC; ++i) { vbase[i + 1] += vbase[i]; ibase[i + 1] += ibase[i]; }
    out.verts.resize(vbase[NC]);
    out.inds.resize(ibase[NC]);
    auto rank_in = [&](size_t x, size_t y, size_t z, unsigned edge) -> unsigned {
        const uint64_t row = UWO_TRI_ROWS[out.cases[cell_id(x, y, z)]];
        unsigned seen = 0, r = 0;
        for (int i = 0; i < 16; ++i) {
            const unsigned e = (unsigned)((row >> (4 * i)) & 0xF);
            if (e == 0xF) break;
            if (seen >> e & 1) continue;
            seen |= 1u << e;
            size_t ox, oy, oz; unsigned oe; owner_of(e, x, y, z, ox, oy, oz, oe);
            const bool owned = (ox == x && oy == y && oz == z);
            if (e == edge) return owned ? r : 0xFFFFFFFFu;
            if (owned) ++r;
        }
        return 0xFFFFFFFFu;
    };
    // pass 2: emit
    for (size_t x = 0; x < S; ++x) for (size_t y = 0; y < S; ++y) for (size_t z = 0; z < S; ++z) {
        const size_t id = cell_id(x, y, z);
        const uint64_t row = UWO_TRI_ROWS[out.cases[id]];
        unsigned seen = 0;
        for (int i = 0; i < 16; ++i) {
            const unsigned e = (unsigned)((row >> (4 * i)) & 0xF);
            if (e == 0xF) break;
            size_t ox, oy, oz; unsigned oe; owner_of(e, x, y, z, ox, oy, oz, oe);
            const unsigned r = rank_in(ox, oy, oz, oe);
            const uint32_t vi = vbase[cell_id(ox, oy, oz)] + r;
            out.inds[ibase[id] + i] = vi;
            if (!(seen >> e & 1) && ox == x && oy == y && oz == z) {
                const unsigned ca = UWO_EDGE_CORNERS[e][0], cb = UWO_EDGE_CORNERS[e][1];
                const size_t A[3] = { x + CORNER_OFF[ca][0], y + CORNER_OFF[ca][1], z + CORNER_OFF[ca][2] };
                const size_t B[3] = { x + CORNER_OFF[cb][0], y + CORNER_OFF[cb][1], z + CORNER_OFF[cb][2] };
                const float sa[3] = { (float)A[0] * size_scale, (float)A[1] * size_scale, (float)A[2] * size_scale };
                const float sb[3] = { (float)B[0] * size_scale, (float)B[1] * size_scale, (float)B[2] * size_scale };
                const float iso_a = isos[A[0] * L * L + A[1] * L + A[2]], iso_b = isos[B[0] * L * L + B[1] * L + B[2]];
                const float t = (c->iso_level - iso_a) / (iso_b - iso_a);
                const float mid[3] = { sa[0] + t * (sb[0] - sa[0]), sa[1] + t * (sb[1] - sa[1]), sa[2] + t * (sb[2] - sa[2]) };
                uwo_vert v;
                const float world_z = mid[2] + (float)off[2];
                v.pos[0] = mid[0] + (float)off[0]; v.pos[1] = mid[1] + (float)off[1]; v.pos[2] = world_z;
                uwo_vertex_color(c, world_z, cb, v.color);
                out.verts[vi] = v;
            }
            seen |= 1u << e;
        }
    }
}

extern "C" {

enum { UWO_FLAG_BLANK_EARLY = 1, UWO_FLAG_HAS_MESH = 2, UWO_FLAG_U16_OVERFLOW = 4 };

// Chunk::build_full (chunk.rs:266-313) from caller-supplied densities (isos_in != NULL) or from
// the density function.  Returns 0, or -1 if an output capacity was too small (counts still set).
// mode: 0 faithful, 1 fast, 2 rule.
int uwo_build_chunk(const uwo_config* c, const uint8_t* perm, const int32_t pos[3], int mode,
                    const float* isos_in, float* isos_out, uint8_t* cases_out,
                    uwo_vert* verts_out, uint32_t verts_cap, uint32_t* inds_out, uint32_t inds_cap,
                    uint32_t* n_verts, uint32_t* n_inds, uint32_t* flags,
                    uwo_tri* tris_out, uint32_t tris_cap, uint32_t* tri_cell_start_out) {
    const size_t S = (size_t)c->internal_size, L = S + 1;
    std::vector<float> isos;
    isos.reserve(L * L * L);   // chunk.rs:57
    if (isos_in) isos.assign(isos_in, isos_in + L * L * L);
    else { isos.resize(L * L * L); uwo_densities(c, perm, pos, isos.data()); }
    if (isos_out) memcpy(isos_out, isos.data(), sizeof(float) * L * L * L);

    *n_verts = 0; *n_inds = 0; *flags = 0;
    bool blank = true;   // chunk.rs:131-133: all(iso > ISO_LEVEL)
    for (float v : isos) if (!(v > c->iso_level)) { blank = false; break; }
    if (blank) {
        *flags = UWO_FLAG_BLANK_EARLY;
        if (cases_out) memset(cases_out, 0, S * S * S);
        if (tri_cell_start_out) memset(tri_cell_start_out, 0, sizeof(uint32_t) * (S * S * S + 1));
        return 0;
    }
    MeshOut out;
    if (mode == 2) build_mesh_rule(c, pos, isos.data(), out);
    else build_mesh(c, pos, isos.data(), mode, tris_out != nullptr, out);
    if (cases_out) memcpy(cases_out, out.cases.data(), S * S * S);
    *n_verts = (uint32_t)out.verts.size();
    *n_inds = (uint32_t)out.inds.size();
    if (*n_inds > 0) *flags |= UWO_FLAG_HAS_MESH;          // chunk.rs:291
    if (*n_verts > 65536u) *flags |= UWO_FLAG_U16_OVERFLOW; // `ind as u16` would wrap, chunk.rs:243
    int rc = 0;
    if (verts_out) { if (*n_verts <= verts_cap) memcpy(verts_out, out.verts.data(), sizeof(uwo_vert) * *n_verts); else rc = -1; }
    if (inds_out)  { if (*n_inds  <= inds_cap)  memcpy(inds_out,  out.inds.data(),  sizeof(uint32_t) * *n_inds);  else rc = -1; }
    if (tris_out && mode != 2) {
        if (out.tris.size() <= tris_cap) memcpy(tris_out, out.tris.data(), sizeof(uwo_tri) * out.tris.size()); else rc = -1;
        if (tri_cell_start_out) memcpy(tri_cell_start_out, out.tri_cell_start.data(), sizeof(uint32_t) * (S * S * S + 1));
    }
    return rc;
}

// The ordered corner pair behind every vertex (Build::vert_pairs, chunk.rs:38,233-240) as lattice indices
// x*L*L + y*L + z, two per vertex, in vertex order -- lets the tests bound each vertex's position error by the
// conditioning of ITS edge, (iso_b - iso_a).  Returns the vertex count, or -1 if cap (in vertices) is too small.
int uwo_vertex_pairs(const uwo_config* c, const uint8_t* perm, const int32_t pos[3], const float* isos_in,
                     uint32_t* pairs_out, uint32_t cap) {
    const size_t S = (size_t)c->internal_size, L = S + 1;
    std::vector<float> isos(L * L * L);
    if (isos_in) isos.assign(isos_in, isos_in + L * L * L);
    else uwo_densities(c, perm, pos, isos.data());
    bool blank = true;
    for (float v : isos) if (!(v > c->iso_level)) { blank = false; break; }
    if (blank) return 0;
    MeshOut out;
    build_mesh(c, pos, isos.data(), 1, false, out);
    const size_t nv = out.verts.size();
    if (nv > cap) return -1;
    memcpy(pairs_out, out.vpairs.data(), sizeof(uint32_t) * 2 * nv);
    return (int)nv;
}

// ---------------------------------------------------------------------------------------
// Collision ray casts (SURVEY 8f-1's consumer): util::Tri::intersects (util.rs:22-59) over the triangles
// Chunk::tris_around (chunk.rs:315-342) returns for the chunks boid.rs:175-208 visits.  cgmath (un-vendored,
// cgmath 0.18 per the reference's Cargo.toml) is restated from its published source: dot = (x*x' + y*y') + z*z',
// cross = (y z' - z y', z x' - x z', x y' - y x'), vector * scalar and +/- element-wise; f32, unfused.
// PARITY UNPINNED for this function: the shipped wasm's copy of boid.rs was not executed (see wasm_forensics.py).
// ---------------------------------------------------------------------------------------
static inline float dot3(const float a[3], const float b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
static inline void cross3(const float a[3], const float b[3], float o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}

// util.rs:22-59.  Returns t, or -1 for None.
float uwo_tri_intersects(const uwo_tri* tr, const float pos[3], const float dir[3], float range) {
    const float EPSILON = 1e-5f;                                        // util.rs:3
    const float dot_normal_dir = dot3(tr->normal, dir);
    if (fabsf(dot_normal_dir) < EPSILON) return -1.0f;
    const float d0[3] = { tr->verts[0][0] - pos[0], tr->verts[0][1] - pos[1], tr->verts[0][2] - pos[2] };
    const float t = dot3(tr->normal, d0) / dot_normal_dir;
    if (t < 0.0f || t > range) return -1.0f;
    const float ip[3] = { pos[0] + dir[0] * t, pos[1] + dir[1] * t, pos[2] + dir[2] * t };
    const float* v0 = tr->verts[0]; const float* v1 = tr->verts[1]; const float* v2 = tr->verts[2];
    const float e0[3] = { v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2] };
    const float e1[3] = { v2[0] - v1[0], v2[1] - v1[1], v2[2] - v1[2] };
    const float e2[3] = { v0[0] - v2[0], v0[1] - v2[1], v0[2] - v2[2] };
    const float p0[3] = { ip[0] - v0[0], ip[1] - v0[1], ip[2] - v0[2] };
    const float p1[3] = { ip[0] - v1[0], ip[1] - v1[1], ip[2] - v1[2] };
    const float p2[3] = { ip[0] - v2[0], ip[1] - v2[1], ip[2] - v2[2] };
    float n0[3], n1[3], n2[3];
    cross3(e0, p0, n0); cross3(e1, p1, n1); cross3(e2, p2, n2);
    const float q0 = dot3(n0, tr->normal), q1 = dot3(n1, tr->normal), q2 = dot3(n2, tr->normal);
    return (q0 >= 0.0f && q1 >= 0.0f && q2 >= 0.0f) ? t : -1.0f;
}

// For every ray: the triangles boid.rs:175-208 gathers around its origin (chunks within +-wall_range world units,
// cells within +-wall_range cells: Chunk::tris_around) out of the given chunks, each tested with Tri::intersects
// (range = wall_range as f32, boid.rs:221); out_t = the smallest Some(t), or -1 if every test returned None.
void uwo_raycast(const uwo_config* c, const uint8_t* perm, const int32_t* chunk_pos, uint32_t n_chunks,
                 const float* origins, const float* dirs, uint32_t n_rays, int32_t wall_range, float* out_t) {
    const size_t S = (size_t)c->internal_size, L = S + 1, NC = S * S * S;
    struct Built { int32_t pos[3]; std::vector<uwo_tri> tris; std::vector<uint32_t> start; };
    std::vector<Built> built(n_chunks);
    for (uint32_t i = 0; i < n_chunks; ++i) {
        Built& b = built[i];
        memcpy(b.pos, chunk_pos + 3 * (size_t)i, sizeof b.pos);
        b.start.assign(NC + 1, 0);
        std::vector<float> isos(L * L * L);
        uwo_densities(c, perm, b.pos, isos.data());
        bool blank = true;
        for (float v : isos) if (!(v > c->iso_level)) { blank = false; break; }
        if (blank) continue;
        MeshOut out;
        build_mesh(c, b.pos, isos.data(), 1, true, out);
        b.tris = out.tris; b.start = out.tri_cell_start;
    }
    const float cs = (float)c->chunk_size, wr = (float)wall_range;
    for (uint32_t r = 0; r < n_rays; ++r) {
        const float* p = origins + 3 * (size_t)r;
        const float* d = dirs + 3 * (size_t)r;
        int32_t ws[3], we[3];
        for (int k = 0; k < 3; ++k) { ws[k] = (int32_t)floorf((p[k] - wr) / cs); we[k] = (int32_t)floorf((p[k] + wr) / cs); }   // boid.rs:177-183
        float best = -1.0f;
        for (int32_t a = ws[0]; a <= we[0]; ++a) for (int32_t b = ws[1]; b <= we[1]; ++b) for (int32_t cc = ws[2]; cc <= we[2]; ++cc) {
            const Built* bc = nullptr;
            for (const Built& q : built) if (q.pos[0] == a && q.pos[1] == b && q.pos[2] == cc) { bc = &q; break; }   // world.get_chunk
            if (!bc) continue;
            const int32_t cp[3] = { a, b, cc };
            int32_t lo[3], hi[3];
            for (int k = 0; k < 3; ++k) {
                const float local = p[k] - (float)cp[k] * cs;                 // boid.rs:186,190,201
                const float pct = local / cs;
                const int32_t mid = (int32_t)floorf(pct * (float)c->internal_size);   // chunk.rs:316-318
                lo[k] = std::max(mid - wall_range, 0); hi[k] = std::min(mid + wall_range, (int32_t)c->internal_size);
            }
            for (int32_t x = lo[0]; x <= hi[0]; ++x) for (int32_t y = lo[1]; y <= hi[1]; ++y) for (int32_t z = lo[2]; z <= hi[2]; ++z) {
                if (x >= (int32_t)S || y >= (int32_t)S || z >= (int32_t)S) continue;   // no such key in the map
                const size_t cell = ((size_t)x * S + y) * S + z;
                for (uint32_t j = bc->start[cell]; j < bc->start[cell + 1]; ++j) {
                    const float t = uwo_tri_intersects(&bc->tris[j], p, d, wr);
                    if (t >= 0.0f && (best < 0.0f || t < best)) best = t;
                }
            }
        }
        out_t[r] = best;
    }
}

// CPU baseline driver: builds n chunks exactly as World::build_full_step does one by one
// (world.rs:113-123), on `nthreads` host threads over independent chunks.  Outputs are
// dropped like Build::finish() (chunk.rs:66-77); totals + a checksum are returned so the
// work cannot be optimised away.  Returns elapsed seconds (steady_clock).
double uwo_build_batch_timed(const uwo_config* c, const uint8_t* perm, const int32_t* pos_xyz, uint32_t n,
                             int mode, int nthreads, uint64_t* tot_verts, uint64_t* tot_inds,
                             uint64_t* n_blank, uint64_t* n_mesh, double* checksum) {
    if (nthreads < 1) nthreads = 1;
    std::atomic<uint32_t> next(0);
    std::vector<uint64_t> tv(nthreads, 0), ti(nthreads, 0), nb(nthreads, 0), nm(nthreads, 0);
    std::vector<double> cs(nthreads, 0.0);
    auto worker = [&](int tid) {
        const size_t S = (size_t)c->internal_size, L = S + 1;
        for (;;) {
            const uint32_t i = next.fetch_add(1);
            if (i >= n) break;
            std::vector<float> isos;
            isos.reserve(L * L * L);
            isos.resize(L * L * L);
            uwo_densities(c, perm, pos_xyz + 3 * (size_t)i, isos.data());
            bool blank = true;
            for (float v : isos) if (!(v > c->iso_level)) { blank = false; break; }
            if (blank) { nb[tid]++; continue; }
            MeshOut out;
            if (mode == 2) build_mesh_rule(c, pos_xyz + 3 * (size_t)i, isos.data(), out);
            else build_mesh(c, pos_xyz + 3 * (size_t)i, isos.data(), mode, false, out);
            tv[tid] += out.verts.size(); ti[tid] += out.inds.size();
            if (!out.inds.empty()) nm[tid]++;
            for (const auto& v : out.verts) cs[tid] += (double)v.pos[0] + (double)v.color[1];
        }
    };
    const auto t0 = std::chrono::steady_clock::now();
    if (nthreads == 1) worker(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) th.emplace_back(worker, t);
        for (auto& t : th) t.join();
    }
    const auto t1 = std::chrono::steady_clock::now();
    *tot_verts = *tot_inds = *n_blank = *n_mesh = 0; *checksum = 0.0;
    for (int t = 0; t < nthreads; ++t) { *tot_verts += tv[t]; *tot_inds += ti[t]; *n_blank += nb[t]; *n_mesh += nm[t]; *checksum += cs[t]; }
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
