// mk_stream3.cuh — configuration and launch interface of k_stream3 (mk_stream3.cu).
#pragma once
#include "mk_common.cuh"

#ifndef S3_TILE
#define S3_TILE 12288      // tile-proper bytes (multiple of 2048)
#endif
#ifndef S3_NS
#define S3_NS 7            // ring stages
#endif
#ifndef S3_NF
#define S3_NF 12           // front warps (one tile each, round robin, ahead of the ring)
#endif
#ifndef S3_NP
#define S3_NP 18           // probe warps
#endif
#ifndef S3_AHEAD
#define S3_AHEAD (S3_NF + 8)   // tiles (= rounds of the grid) the front warps may run ahead of the loader
#endif
#ifndef S3_FPF
#define S3_FPF 1            // front warps prefetch the tile they take this many iterations from now into L2
#endif
#ifndef S3_QN
#define S3_QN 2048         // item queue slots
#endif
#ifndef S3_MAX_WBITS
#define S3_MAX_WBITS 15    // filter words = 2^15 (128 KB)
#endif
#define S3_THREADS (32 * (2 + S3_NF + S3_NP))   // loader, dispatcher, front, probe
#define FLAG_ARENA_FULL 16u   // the item arena was too small: the host sizes it from the cursor and runs again

struct S3Args {
    const uint8_t *text;
    u64 nbytes;
    u64 line_base;          // '\n' bytes in front of the text (its low two bits give the record phase)
    u32 tile_bytes;         // multiple of 64, <= S3_TILE
    u32 n_tiles;            // tiles [tile_begin, n_tiles) are processed by this launch
    u32 tile_begin;
    uint8_t *desc;          // one byte per tile (status << 2 | value mod 4), zeroed before the first launch
    u64 *ttab;              // one entry per tile (ready | item count | arena offset), zeroed before the first launch
    u32 *arena;             // items of all tiles (front warps -> dispatcher)
    u64 arena_cap;          // items
    u64 *arena_cursor;      // zeroed before the first launch
    const u32 *bitmap;      // two-plane core filter (mk_s3_filter_add)
    u32 bitmap_bytes;
    u64 *cand_pos;
    u64 *cand_count;
    u64 cand_cap;
    u32 *flags;
    u64 *total_newlines;    // += newlines of every tile
    u64 *wd;
    u64 *stats;             // development aid (-DS3_STATS)
    u32 stat_cta;
    int TL;                 // k-mer length
    int prew;               // 16-byte vectors in front of a block's first position (SHIFTED geometry)
    int shift_d;            // 2 * (16 prew - pre) bits
};

__device__ __forceinline__ void emit_hit3(const S3Args &A, u64 pos)
{
    u64 idx = atomicAdd((unsigned long long *)A.cand_count, 1ull);
    if (idx < A.cand_cap) A.cand_pos[idx] = pos;
}

int mk_s3_word_bits(int mw);
size_t mk_s3_smem_bytes(u32 bitmap_bytes);
void mk_s3_filter_add(std::vector<u32> &bitmap, int mw, u64 q);
size_t mk_s3_arena_items(size_t nbytes, u32 tile_bytes);
int mk_s3_launch(mk_ctx *ctx, const S3Args &a, bool raw, u32 grid);
