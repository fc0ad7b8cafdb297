// tier2.h -- host-side codestream front door of libj2kgpu.so: main header, tile-part index and tier-2 (packet headers)
// of an ISO/IEC 15444-1 / -15 codestream -> the flat job tables of include/j2kgpu.h.
//
// In the reference this is the work of internal/codestream (ReadHeader parser.go:44-124, ReadTilePartHeader :894-982) and
// of a tier-2 that does not exist yet: decodeTile is a placeholder (decoder.go:375-380) and internal/tcd/t2.go is a toy
// (unary "tag tree", 3-bit lengths).  north_star keeps parsing in Go; with no Go toolchain in this
// image the host side above the C ABI is written here in C++ (the role of `buildGPUJob` in INTEGRATION.md), so that a real
// codestream can be handed to the library as bytes.  Pure host code: no CUDA in this translation unit.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/j2kgpu.h"

struct j2kgpu_parsed {
    j2k_image_t image{};
    std::vector<j2k_tilecomp_t> tilecomps;
    std::vector<j2k_cblk_t> cblks;
    const uint8_t *blob = nullptr;      // the caller's codestream when every block's bytes are contiguous in it (zero copy) ...
    uint64_t blob_len = 0;
    std::vector<uint8_t> owned;         // ... else the codestream followed by the concatenated multi-layer blocks
    uint32_t layers = 0, tiles = 0, tile_parts = 0, packets = 0;
    uint32_t progression = 0;
    uint32_t plt_packets = 0;           // packets whose length a PLT marker announced (checked against the parsed length)
    uint32_t tlm_tile_parts = 0;        // tile-parts listed by TLM markers (checked against Psot)
    std::string err;
};

// The same in steps, for callers that spread the tiles of several codestreams over one set of threads: begin reads the main
// header and the tile-part index, tile(t) runs tier-2 of one tile (thread-safe for distinct tiles), finish merges the tiles
// into the tables and frees the frame (free: without merging).
struct j2k_t2_frame;
int j2k_tier2_begin(const uint8_t *cs, uint64_t len, uint32_t reduce, j2k_t2_frame **f, std::string &err);
uint32_t j2k_tier2_tiles(const j2k_t2_frame *f);
void j2k_tier2_tile(j2k_t2_frame *f, uint32_t t);
int j2k_tier2_finish(j2k_t2_frame *f, j2kgpu_parsed &out);
void j2k_tier2_free(j2k_t2_frame *f);

// threads = tiles parsed concurrently (0 = hardware concurrency).  Returns J2KGPU_OK or a J2KGPU_E_* code with out.err set.
int j2k_tier2_parse(const uint8_t *cs, uint64_t len, uint32_t reduce, uint32_t threads, j2kgpu_parsed &out);
