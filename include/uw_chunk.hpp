// uw_chunk.hpp -- header-only C++ host mirror of the reference's Rust chunk API on top of the C ABI.
//
// The reference is compiled code (Rust); its toolchain is not in this image, so the host side above
// include/uwcuda.h is mirrored here in C++ with the same names, argument meaning and error behaviour
// (the reference `unwrap()`s / panics; this throws uw::Error):
//
//   noise::Perlin::new(seed)             src/state.rs:359   -> uw::Perlin(seed)
//   Chunk::new(pos)                      src/chunk.rs:89    -> uw::Chunk::create(pos)
//   Chunk::build_full(&perlin,&device)   src/chunk.rs:266   -> chunk.build_full(builder)
//   Chunk::build_partial(..) -> bool     src/chunk.rs:270   -> chunk.build_partial(builder)
//   Chunk::not_blank()                   src/chunk.rs:344   -> chunk.not_blank()
//   Chunk::verts_buffer_slice()          src/chunk.rs:346   -> chunk.verts_buffer_slice()
//   Chunk::inds_buffer_slice()           src/chunk.rs:347   -> chunk.inds_buffer_slice()
//   Chunk::num_inds()                    src/chunk.rs:348   -> chunk.num_inds()
//   World::build_full_step (one chunk)   src/world.rs:113   -> uw::build_chunks(builder, positions) (a batch),
//                                                            uw::build_chunks_pipelined (a stream of batches)
//
// Link with -luwcuda.  No CPU fallback: constructing a ChunkBuilder without a CUDA device throws.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "uwcuda.h"

namespace uw {

constexpr int CHUNK_SIZE = 16;        // src/chunk.rs:5
constexpr int INTERNAL_SIZE = 12;     // src/chunk.rs:6
constexpr unsigned PERLIN_OCTAVES = 3;  // src/chunk.rs:9
constexpr float ISO_LEVEL = -0.1f;    // src/chunk.rs:10

struct Error : std::runtime_error {
    uw_status status;
    Error(uw_status s, const std::string& msg) : std::runtime_error(msg), status(s) {}
};

class Perlin {   // only the seed matters; the permutation table is rebuilt inside the library
public:
    static constexpr uint32_t DEFAULT_SEED = 0;
    explicit Perlin(uint32_t seed = DEFAULT_SEED) : seed_(seed) {}
    uint32_t seed() const { return seed_; }   // noise::Seedable::seed
private:
    uint32_t seed_;
};

using VertColor = uw_vert;   // src/draw.rs:4-9

class ChunkBuilder {
public:
    explicit ChunkBuilder(const Perlin& perlin = Perlin(), int device = -1, uint32_t flags = 0) {
        uw_config cfg;
        uw_config_default(&cfg);
        cfg.seed = perlin.seed();
        cfg.device = device;
        cfg.flags = flags;
        const uw_status st = uw_create(&cfg, &ctx_);
        if (st != UW_OK) throw Error(st, std::string("uw_create: ") + uw_last_error(nullptr));
    }
    explicit ChunkBuilder(const uw_config& cfg) {
        const uw_status st = uw_create(&cfg, &ctx_);
        if (st != UW_OK) throw Error(st, std::string("uw_create: ") + uw_last_error(nullptr));
    }
    ~ChunkBuilder() { uw_destroy(ctx_); }
    ChunkBuilder(const ChunkBuilder&) = delete;
    ChunkBuilder& operator=(const ChunkBuilder&) = delete;

    uw_ctx* ctx() const { return ctx_; }
    void check(uw_status st, const char* what) const {
        if (st != UW_OK) throw Error(st, std::string(what) + ": " + uw_last_error(ctx_));
    }
    std::array<uint8_t, 256> perm_table() const {
        std::array<uint8_t, 256> t{};
        check(uw_perm_table(ctx_, t.data()), "uw_perm_table");
        return t;
    }
    // perlin_util::iso_at on n points (src/perlin_util.rs:24-29)
    std::vector<float> iso_at(const std::vector<std::array<double, 3>>& pts) const {
        std::vector<float> out(pts.size());
        check(uw_iso_at(ctx_, pts.empty() ? nullptr : pts[0].data(), (uint32_t)pts.size(), out.data()), "uw_iso_at");
        return out;
    }

private:
    uw_ctx* ctx_ = nullptr;
};

class Chunk {
public:
    static Chunk create(std::array<int32_t, 3> pos) { return Chunk(pos); }   // Chunk::new
    explicit Chunk(std::array<int32_t, 3> pos)
        : pos_(pos), chunk_offset_{pos[0] * CHUNK_SIZE, pos[1] * CHUNK_SIZE, pos[2] * CHUNK_SIZE} {}   // chunk.rs:90-94

    void build_full(const ChunkBuilder& b) {
        uw_batch* batch = nullptr;
        b.check(uw_build(b.ctx(), pos_.data(), 1, &batch), "uw_build");
        adopt(b, batch, 0);
        uw_batch_free(batch);
    }
    // The reference slices a build over frames to keep ONE CPU thread responsive (chunk.rs:19-20);
    // the GPU build has no partial state: one step, returns true.
    bool build_partial(const ChunkBuilder& b) {
        if (!done_) build_full(b);
        return true;
    }
    // the two create_buffer_init copies of chunk.rs:291-305, out of the batch's pinned arena
    void adopt(const ChunkBuilder& b, const uw_batch* batch, uint32_t i) {
        uw_batch_view v;
        b.check(uw_batch_view_get(batch, &v), "uw_batch_view_get");
        const uw_chunk_desc& d = v.descs[i];
        flags_ = d.flags;
        num_inds_ = d.index_count;
        verts_.clear();
        inds_.clear();
        inds32_.clear();
        index32_ = v.inds32 != nullptr;
        if (d.flags & UW_CHUNK_HAS_MESH) {
            verts_.assign(v.verts + d.vert_offset, v.verts + d.vert_offset + d.vert_count);
            // a builder created with UW_FLAG_INDEX32 (or internal_size > 22, where `ind as u16` could wrap,
            // chunk.rs:243) returns u32 indices only: inds16 is NULL then
            if (index32_) inds32_.assign(v.inds32 + d.index_offset, v.inds32 + d.index_offset + d.index_count);
            else          inds_.assign(v.inds16 + d.index_offset, v.inds16 + d.index_offset + d.index_count);
        }
        done_ = true;
    }

    bool not_blank() const { return !verts_.empty(); }                       // chunk.rs:344
    const std::vector<VertColor>& verts_buffer_slice() const {               // chunk.rs:346 (unwrap -> panic)
        if (!not_blank()) throw std::logic_error("called verts_buffer_slice() on a blank chunk");
        return verts_;
    }
    const std::vector<uint16_t>& inds_buffer_slice() const {                 // chunk.rs:347
        if (!not_blank()) throw std::logic_error("called inds_buffer_slice() on a blank chunk");
        if (index32_) throw std::logic_error("this builder emits u32 indices: use inds32_buffer_slice()");
        return inds_;
    }
    // u32 variant (wgpu::IndexFormat::Uint32 instead of state.rs:506's Uint16) for UW_FLAG_INDEX32 builders
    const std::vector<uint32_t>& inds32_buffer_slice() const {
        if (!not_blank()) throw std::logic_error("called inds32_buffer_slice() on a blank chunk");
        if (!index32_) throw std::logic_error("this builder emits u16 indices: use inds_buffer_slice()");
        return inds32_;
    }
    bool index32() const { return index32_; }
    size_t num_inds() const { return num_inds_; }                            // chunk.rs:348
    bool blank_early() const { return flags_ & UW_CHUNK_BLANK_EARLY; }
    const std::array<int32_t, 3>& pos() const { return pos_; }
    const std::array<int32_t, 3>& chunk_offset() const { return chunk_offset_; }

private:
    std::array<int32_t, 3> pos_, chunk_offset_;
    std::vector<VertColor> verts_;
    std::vector<uint16_t> inds_;
    std::vector<uint32_t> inds32_;
    bool index32_ = false;
    size_t num_inds_ = 0;
    uint32_t flags_ = 0;
    bool done_ = false;
};

// Batched World::build_full_step (src/world.rs:113-123): one call, many chunks.
inline std::vector<Chunk> build_chunks(const ChunkBuilder& b, const std::vector<std::array<int32_t, 3>>& positions) {
    std::vector<Chunk> out;
    out.reserve(positions.size());
    uw_batch* batch = nullptr;
    b.check(uw_build(b.ctx(), positions.empty() ? nullptr : positions[0].data(), (uint32_t)positions.size(), &batch), "uw_build");
    for (uint32_t i = 0; i < positions.size(); ++i) {
        out.emplace_back(positions[i]);
        out.back().adopt(b, batch, i);
    }
    uw_batch_free(batch);
    return out;
}

// Streaming variant for a loader that hands over many batches: batch k+1 is submitted (uw_build_async) before
// batch k is collected (uw_batch_wait), so k's copy to the host overlaps k+1's kernel.  `sink(k, chunks)` is called
// in batch order.
template <class Sink>
inline void build_chunks_pipelined(const ChunkBuilder& b, const std::vector<std::vector<std::array<int32_t, 3>>>& batches, Sink sink) {
    auto submit = [&](const std::vector<std::array<int32_t, 3>>& pos) {
        uw_batch* h = nullptr;
        b.check(uw_build_async(b.ctx(), pos.empty() ? nullptr : pos[0].data(), (uint32_t)pos.size(), &h), "uw_build_async");
        return h;
    };
    auto collect = [&](size_t k, uw_batch* h) {
        struct Guard { uw_batch* h; ~Guard() { uw_batch_free(h); } } guard{h};
        b.check(uw_batch_wait(h), "uw_batch_wait");
        std::vector<Chunk> out;
        out.reserve(batches[k].size());
        for (uint32_t i = 0; i < batches[k].size(); ++i) {
            out.emplace_back(batches[k][i]);
            out.back().adopt(b, h, i);
        }
        sink(k, std::move(out));
    };
    if (batches.empty()) return;
    uw_batch* prev = submit(batches[0]);
    for (size_t k = 1; k < batches.size(); ++k) {
        uw_batch* next = nullptr;
        try { next = submit(batches[k]); } catch (...) { uw_batch_free(prev); throw; }
        try { collect(k - 1, prev); } catch (...) { uw_batch_free(next); throw; }
        prev = next;
    }
    collect(batches.size() - 1, prev);
}

}  // namespace uw
