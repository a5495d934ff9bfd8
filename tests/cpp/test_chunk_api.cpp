// Drives the C++ host mirror (include/uw_chunk.hpp) the way the reference's World drives Chunk
// (src/world.rs:113-123): Chunk::new(pos) -> build_full -> not_blank / num_inds / buffer slices.
// Prints one line per chunk; tests/test_cpp_mirror.py compares the lines with the oracle.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "uw_chunk.hpp"

static uint64_t fnv(const void* p, size_t n, uint64_t h = 1469598103934665603ull) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

int main(int argc, char** argv) {
    const uint32_t seed = argc > 1 ? (uint32_t)strtoul(argv[1], nullptr, 10) : 0;
    try {
        uw::ChunkBuilder builder{uw::Perlin(seed)};
        std::vector<std::array<int32_t, 3>> positions = {{0, 0, -1}, {0, 0, 0}, {3, -2, 2}, {-5, 7, -4}, {1, 1, -2}};
        // one at a time, like the reference ...
        for (const auto& p : positions) {
            uw::Chunk c = uw::Chunk::create(p);
            c.build_full(builder);
            if (!c.build_partial(builder)) return 3;
            if (c.not_blank())
                printf("chunk %d %d %d not_blank=1 num_inds=%zu verts=%zu inds_hash=%016llx pos0=%a\n", p[0], p[1], p[2], c.num_inds(),
                       c.verts_buffer_slice().size(), (unsigned long long)fnv(c.inds_buffer_slice().data(), c.num_inds() * 2),
                       (double)c.verts_buffer_slice()[0].pos[0]);
            else
                printf("chunk %d %d %d not_blank=0 num_inds=%zu blank_early=%d\n", p[0], p[1], p[2], c.num_inds(), (int)c.blank_early());
        }
        // ... and as one batch: must agree chunk by chunk
        auto batch = uw::build_chunks(builder, positions);
        for (size_t i = 0; i < positions.size(); ++i) {
            uw::Chunk c(positions[i]);
            c.build_full(builder);
            if (c.num_inds() != batch[i].num_inds() || c.not_blank() != batch[i].not_blank()) { printf("BATCH MISMATCH %zu\n", i); return 4; }
            if (c.not_blank() && memcmp(c.inds_buffer_slice().data(), batch[i].inds_buffer_slice().data(), c.num_inds() * 2)) return 5;
        }
        // ... and as a stream of batches with two in flight
        {
            std::vector<std::vector<std::array<int32_t, 3>>> stream;
            for (int k = 0; k < 4; ++k) stream.push_back({positions[k], positions[(k + 1) % positions.size()], {k, -k, -1}});
            size_t seen = 0, bad = 0;
            uw::build_chunks_pipelined(builder, stream, [&](size_t k, std::vector<uw::Chunk> chunks) {
                if (k != seen++) ++bad;
                for (size_t i = 0; i < chunks.size(); ++i) {
                    uw::Chunk c(stream[k][i]);
                    c.build_full(builder);
                    if (c.num_inds() != chunks[i].num_inds()) ++bad;
                    else if (c.not_blank() && memcmp(c.inds_buffer_slice().data(), chunks[i].inds_buffer_slice().data(), c.num_inds() * 2)) ++bad;
                }
            });
            printf("pipelined batches=%zu bad=%zu\n", seen, bad);
            if (seen != stream.size() || bad) return 6;
        }
        // a UW_FLAG_INDEX32 builder returns u32 indices only (inds16 == NULL): same values, wider type
        {
            uw_config cfg;
            uw_config_default(&cfg);
            cfg.seed = seed; cfg.flags |= UW_FLAG_INDEX32;
            uw::ChunkBuilder b32{cfg};
            size_t bad32 = 0, meshes32 = 0;
            auto batch32 = uw::build_chunks(b32, positions);
            for (size_t i = 0; i < positions.size(); ++i) {
                if (batch32[i].num_inds() != batch[i].num_inds()) { ++bad32; continue; }
                if (!batch32[i].not_blank()) continue;
                ++meshes32;
                if (!batch32[i].index32()) ++bad32;
                const auto& a = batch32[i].inds32_buffer_slice();
                const auto& r = batch[i].inds_buffer_slice();
                for (size_t k = 0; k < a.size(); ++k) bad32 += a[k] != (uint32_t)r[k];
                bool threw16 = false;
                try { batch32[i].inds_buffer_slice(); } catch (const std::logic_error&) { threw16 = true; }
                bad32 += !threw16;
            }
            printf("index32 meshes=%zu bad=%zu\n", meshes32, bad32);
            if (bad32 || !meshes32) return 7;
        }
        // the reference panics when slicing a blank chunk's buffers (chunk.rs:346): here it throws
        uw::Chunk blank({0, 0, 5});
        blank.build_full(builder);
        bool threw = false;
        try { blank.verts_buffer_slice(); } catch (const std::logic_error&) { threw = true; }
        printf("blank_slice_throws=%d\n", (int)threw);
        printf("OK\n");
        return 0;
    } catch (const uw::Error& e) {
        printf("uw::Error status=%d %s\n", (int)e.status, e.what());
        return e.status == UW_ERR_NO_DEVICE ? 42 : 1;
    }
}
