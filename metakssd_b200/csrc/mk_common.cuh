// mk_common.cuh — shared declarations of libmkssd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <vector>
#include <limits.h>
#include "mkssd_b200.h"

typedef unsigned long long u64;
typedef unsigned int u32;

// two-hash Bloom filter over the pass set (inner windows of 22+ bits): 2^MK_BLOOM_WBITS 32-bit words.
// 15 = 128 KB (fewer first-level false positives, 6-stage ring), 14 = 64 KB (10-stage ring)
#ifndef MK_BLOOM_WBITS
#define MK_BLOOM_WBITS 15
#endif
#define MK_HALO 32            // bytes of left context staged in front of every text tile
#define MK_MAX_TILE 24576     // tile-proper bytes (multiple of 64); three stages fit beside the bitmap
#ifndef MK_STREAM_THREADS
#define MK_STREAM_THREADS 512
#endif
#define MK_HITCAP 1024        // first-level bitmap hits queued per tile

// ---- sketch parameters mirrored on the device (values of iseq2comem.c:54-86) -------------------
struct KParams {
    int k, subk, drlevel, outctx, TL, crvs_shift, dim_end;
    u64 tupmask, domask, undomask, lowmask;
    int code_shift;        // 2*TL - 4*outctx
    u32 hashsize;
    // bloom / probe geometry
    int mw;                // 4*subk bits of inner substring
    int pre;               // k + subk - 1 bases before the k-mer end where the inner window starts
    int spare;             // 1 when mw < 22 (window aligned at bit 2)
    int prew;              // 16-base words staged in front of a 32-position block (1 or 2)
    int shift_s;           // pre-shift (bits) aligning position j's window at bit 2j (+2*spare)
    u32 ptab_mask;
};

struct Scratch {
    void *p = nullptr;
    size_t bytes = 0;
};

enum {
    SB_TILE_DESC = 0, SB_CAND_CODE, SB_CAND_POS, SB_COUNTERS, SB_ACC_KEYS, SB_ACC_CNT, SB_ACC_POS,
    SB_IT_CODE, SB_IT_CNT, SB_IT_POS, SB_SORT_K0, SB_SORT_V0, SB_SORT_K1, SB_SORT_V1, SB_HIST,
    SB_R_CODE, SB_R_CNT, SB_R_FILE, SB_R_PROBE, SB_SLOT_KEYS, SB_SLOT_VALS, SB_OUT_CODE, SB_OUT_CNT,
    SB_OUT_KEY, SB_SEG_COUNTS, SB_TEXT, SB_SCAN_TMP, SB_FA_STATE, SB_FA_CNT, SB_FA_DENSE, SB_FA_OFF,
    SB_CQ_KEYS, SB_CQ_IDX, SB_C_REF, SB_C_IDX, SB_C_QRY, SB_C_QCNT, SB_C_HITVAL, SB_C_POS, SB_C_STORE_S,
    SB_C_STORE_C, SB_C_STATS, SB_C_NH, SB_SYN_CDF, SB_SYN_SPC, SB_MISC, SB_FILE_OFF, SB_RUN_CODE,
    SB_RUN_POS, SB_RUN_CNT, SB_TILE_DESC8, SB_S3_ARENA, SB_X_CUT, SB_X_SHDR, SB_X_SCODE, SB_X_SPOS, SB_X_SCNT,
    SB_X_RCODE, SB_X_RPOS, SB_X_RCNT, SB_X_QCODE, SB_X_QCNT, SB_X_QN, SB_X_HITS, SB_NUM
};

struct ResidentComponent {         // one MarkerDB component kept on the device (mk_markerdb_load)
    u32 *d_ref = nullptr;
    u64 *d_index = nullptr;
    u64 r = 0;
    int n_species = 0;
};

struct mk_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    cudaEvent_t copy_ev[2] = {nullptr, nullptr};
    mk_info info;
    KParams kp;
    u32 *d_bitmap = nullptr;
    u32 bitmap_words = 0;
    u32 *d_bitmap3 = nullptr;       // two-plane core filter of k_stream3 (mk_stream3.cu)
    u32 bitmap3_words = 0;
    u64 *d_ptab = nullptr;
    bool no_tables = false;         // context created without a permutation (composite only)
    // per-call options of the secondary sketchers (reset by the entry points that set them)
    int verify_quality = INT_MIN;   // fastq2co(): a base counts iff (signed char)quality byte >= Q (INT_MIN: no check)
    u32 line_limit = 4095;          // fgets(buf, LEN): a line of LEN - 1 bytes or more is split by the reference
    u32 emit_lo = 0, emit_hi = 0xFFFFFFFFu;   // codes are written iff emit_lo <= occurrences <= emit_hi
    bool borrow_output = false;     // sketches returned as views into the pinned staging block (mk_ctx_set_borrowed_output)
    bool fasta_dedup = false;       // `dist -u`: uniq_fasta2co(), codes occurring once per genome
    void *d_trace = nullptr;        // development aid: phase timestamps of CTA 0 (mk_debug_set_trace)
    Scratch sb[SB_NUM];
    void *h_pinned = nullptr;       // small pinned staging block
    size_t h_pinned_bytes = 0;
    mk_profile prof;
    char err[512];
    int pos_bits = 64;              // significant bits of candidate positions (set per call)
    // composite state
    int comp_species = 0;
    u64 comp_nhits = 0;
    std::vector<int32_t> comp_lists_flat;
    std::vector<const int32_t *> comp_lists_ptr;
    std::vector<ResidentComponent> mdb;
    // multi-GPU (mk_comm.cu): NCCL communicator, this rank, and per component the MarkerDB slice size of every rank
    void *comm = nullptr;
    int rank = 0, world = 1;
    std::vector<std::vector<u64>> mdb_shard_sizes;
    // host text handed from mk_fastq_koc_host to the stream driver (uploaded there, chunk by chunk)
    const uint8_t *h_src = nullptr;       // pipelined upload (pinned or pageable)
    const uint8_t *h_src_all = nullptr;   // same pointer; plain upload if the pipelined path is not taken
    std::vector<cudaEvent_t> chunk_ev;
    u64 h_maxpos = 0;                     // read-back slot of mk_runs_finalize_device
    cudaEvent_t evx[2] = {nullptr, nullptr};   // around the first exchange of the sharded step
    bool classic_only = false;            // k_stream_ws's shared-memory base assumption failed on this driver: unit-pulling kernel
    u64 last_block_need = 0;              // largest block the last sharded step saw on this rank (sent to it, or merged by it)
    u64 h_xflag = 0;                      // read-back slot of the sharded step's collective overflow flag
    u64 last_newlines = 0;                // line_base + newlines of the shard mk_fastq_partial_device saw last
    // device copy of the last single-file -A sketch (codes / counts in on-disk order, component bounds)
    const u32 *last_out_code = nullptr;
    const uint16_t *last_out_cnt = nullptr;
    std::vector<u64> last_seg;
};

// Development aid (env MK_TIMING=1): host wall clock between phase marks, with a stream sync at
// every mark, printed to stderr.
#include <chrono>
struct MkPhaseClock {
    bool on;
    cudaStream_t st;
    std::chrono::steady_clock::time_point t;
    MkPhaseClock(cudaStream_t s) : on(getenv("MK_TIMING") != nullptr), st(s) { if (on) { cudaStreamSynchronize(st); t = std::chrono::steady_clock::now(); } }
    void mark(const char *what)
    {
        if (!on) return;
        cudaStreamSynchronize(st);
        auto n = std::chrono::steady_clock::now();
        const char *rk = getenv("RANK");
        fprintf(stderr, "[mk timing%s%s] %-22s %8.3f ms\n", rk ? " r" : "", rk ? rk : "", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            snprintf(ctx->err, sizeof(ctx->err), "%s:%d: %s -> %s", __FILE__, __LINE__, #call,     \
                     cudaGetErrorString(e_));                                                      \
            return MK_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

#define CKR(expr)                                                                                  \
    do {                                                                                           \
        int r_ = (expr);                                                                           \
        if (r_ != MK_OK) return r_;                                                                \
    } while (0)

// grow-only pinned host staging block (device -> host result copies run at full PCIe speed from it)
static inline int mk_pinned(mk_ctx *ctx, size_t bytes, void **out)
{
    if (ctx->h_pinned_bytes < bytes) {
        if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
        ctx->h_pinned = nullptr;
        ctx->h_pinned_bytes = 0;
        size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaMallocHost(&ctx->h_pinned, want);
        if (e != cudaSuccess) {
            snprintf(ctx->err, sizeof(ctx->err), "cudaMallocHost(%zu): %s", want, cudaGetErrorString(e));
            return MK_ERR_NOMEM;
        }
        ctx->h_pinned_bytes = want;
    }
    *out = ctx->h_pinned;
    return MK_OK;
}

// grow-only device scratch
template <class T>
static inline int mk_scratch(mk_ctx *ctx, int id, size_t n, T **out)
{
    size_t need = n * sizeof(T);
    if (need < 256) need = 256;
    Scratch &s = ctx->sb[id];
    if (s.bytes < need) {
        if (s.p) cudaFree(s.p);
        s.p = nullptr;
        s.bytes = 0;
        size_t want = need + need / 4;
        cudaError_t e = cudaMalloc(&s.p, want);
        if (e != cudaSuccess) {
            e = cudaMalloc(&s.p, need);
            want = need;
        }
        if (e != cudaSuccess) {
            snprintf(ctx->err, sizeof(ctx->err), "cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e));
            cudaGetLastError();
            return MK_ERR_NOMEM;
        }
        s.bytes = want;
    }
    *out = (T *)s.p;
    return MK_OK;
}

#define LAUNCH_COUNT(ctx) ((ctx)->prof.kernel_launches++)

// ---- device helpers ---------------------------------------------------------------------------
// 2-bit code of an ASCII nucleotide (A/a=0 C/c=1 G/g=2 T/t=3), valid for those 8 bytes only.
__device__ __forceinline__ u32 mk_code2(u32 c) { return ((c >> 1) ^ (c >> 2)) & 3u; }
// exactly the bytes global_basic.c:62-69 maps to 0..3
__device__ __forceinline__ bool mk_is_acgt(u32 c)
{
    return ((c & 0xC0u) == 0x40u) && ((0x0010008Au >> (c & 31u)) & 1u);
}

__device__ __forceinline__ u32 mk_lane() { return threadIdx.x & 31u; }

// internal entry points shared between translation units
int mk_radix_sort_pairs(mk_ctx *ctx, u64 **keys, u64 **vals, u64 *keys_alt, u64 *vals_alt, u64 n, int begin_bit,
                        int end_bit);
int mk_exclusive_scan_u32(mk_ctx *ctx, const u32 *d_in, u32 *d_out, u64 n, u64 *total_host);
int mk_finalize_candidates(mk_ctx *ctx, const u64 *d_cand_code, const u64 *d_cand_pos, u64 n_cand, long long keep_below,
                           const u64 *d_file_off, int n_files, bool with_counts, mk_sketch *out);
int mk_reduce_candidates(mk_ctx *ctx, const u64 *d_cand_code, const u64 *d_cand_pos, u64 n_cand, long long keep_below,
                         const u64 *d_file_off, int n_files, int code_bits, u64 **d_it_key, u32 **d_it_cnt,
                         u64 **d_it_pos, u64 *n_items);
int mk_order_and_emit(mk_ctx *ctx, u64 *d_it_key, u32 *d_it_cnt, u64 *d_it_pos, u64 n_items, int n_files,
                      bool with_counts, bool drop_zero_code, mk_sketch *out);
int mk_stream_fastq(mk_ctx *ctx, const uint8_t *d_text, size_t nbytes, u64 pos_base, u64 line_base, bool raw_mode,
                    u64 **d_cand_code, u64 **d_cand_pos, u64 *n_cand, u64 *n_newlines);
int mk_runs_finalize_blocks(mk_ctx *ctx, const u64 *d_code, const u64 *d_firstpos, const u32 *d_count, int W, u64 cap,
                            const u64 *h_counts, mk_sketch *out);
int mk_composite_reserve(mk_ctx *ctx, u64 extra);
int mk_composite_component_dev(mk_ctx *ctx, int component, const u32 *d_qry, const uint16_t *d_qcnt, u64 q);
int mk_tail_cut(mk_ctx *ctx, const uint8_t *d_text, size_t nbytes, long long *keep_below);
int mk_tail_cut_fq2co(mk_ctx *ctx, const uint8_t *d_text, size_t nbytes, u64 n_newlines, long long *keep_below);
int mk_fasta_compact(mk_ctx *ctx, const uint8_t *d_text, size_t nbytes, const u64 *h_offsets, int n_files,
                     uint8_t **d_dense, u64 *dense_bytes, u64 **d_dense_off);
