// mk_fasta.cu — FASTA text -> dense base stream, the front half of the FASTA sketcher.
//
// Restates the tokenizer of fasta2co() (/root/reference/iseq2comem.c:245-281) as a parallel
// compaction: ACGT (either case) are kept (upper-cased); '\n' and '\r' vanish (k-mers span line
// breaks); '>' starts a header that is skipped through its '\n' and leaves ONE reset marker;
// every other byte leaves a reset marker ('N').  "Inside a header" at byte i  <=>  the latest of
// {'>', '\n', file start} before i is a '>', which makes the header state a last-non-zero scan.
// The dense stream is then handed to the same streaming kernel as FASTQ sequence lines (RAW mode):
// reset markers are non-ACGT bytes, so no k-mer can contain one.
#include "mk_common.cuh"

#define FA_THREADS 256
#define FA_SPAN 64
#define FA_TILE (FA_THREADS * FA_SPAN)

// first index j in [0, n] with off[j] >= g
__device__ __forceinline__ int lower_bound_u64(const u64 *__restrict__ off, int n, u64 g)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (off[mid] < g) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// last special byte kind in [g0, g1): 0 none, 1 newline / file start, 2 '>'
__device__ __forceinline__ u32 span_summary(const uint8_t *__restrict__ text, u64 g0, u64 g1,
                                            const u64 *__restrict__ file_off, int n_files)
{
    u32 last = 0;
    int nb = lower_bound_u64(file_off, n_files, g0);
    u64 next_start = nb < n_files ? file_off[nb] : ~0ull;
    for (u64 g = g0; g < g1; g++) {
        if (g == next_start) {
            last = 1;
            nb++;
            while (nb < n_files && file_off[nb] == g) nb++;
            next_start = nb < n_files ? file_off[nb] : ~0ull;
        }
        uint8_t c = text[g];
        if (c == '\n') last = 1;
        else if (c == '>') last = 2;
    }
    return last;
}

// "last non-zero" inclusive scan across the block; returns the value entering this thread
__device__ __forceinline__ u32 block_last_nonzero_excl(u32 v, u32 *ws /*[9]*/, u32 *tile_last)
{
    u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    u32 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (u32)o && incl == 0) incl = t;
    }
    if (lane == 31) ws[wid] = incl;
    __syncthreads();
    u32 prev_lane = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) prev_lane = 0;
    u32 carry = 0;
    for (u32 w = 0; w < wid; w++) if (ws[w]) carry = ws[w];
    u32 all = 0;
    for (u32 w = 0; w < FA_THREADS / 32; w++) if (ws[w]) all = ws[w];
    *tile_last = all;
    return prev_lane ? prev_lane : carry;
}

__global__ void __launch_bounds__(FA_THREADS)
k_fa_summary(const uint8_t *__restrict__ text, u64 n, const u64 *__restrict__ file_off, int n_files,
             u32 *__restrict__ tile_state)
{
    __shared__ u32 ws[9];
    u64 g0 = (u64)blockIdx.x * FA_TILE + (u64)threadIdx.x * FA_SPAN;
    u64 g1 = g0 + FA_SPAN < n ? g0 + FA_SPAN : n;
    u32 v = g0 < n ? span_summary(text, g0, g1, file_off, n_files) : 0;
    u32 last;
    block_last_nonzero_excl(v, ws, &last);
    if (threadIdx.x == 0) tile_state[blockIdx.x] = last;
}

// single block: tile_in[t] = last non-zero of tile_state[0..t-1] (1 = not in a header by default)
__global__ void __launch_bounds__(1024) k_fa_propagate(const u32 *__restrict__ tile_state, u32 m, u32 *__restrict__ tile_in)
{
    __shared__ u32 ws[33];
    __shared__ u32 carry_s;
    if (threadIdx.x == 0) carry_s = 1;
    __syncthreads();
    u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (u32 base = 0; base < m; base += 1024) {
        u32 idx = base + threadIdx.x;
        u32 x = idx < m ? tile_state[idx] : 0;
        u32 incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (u32)o && incl == 0) incl = t;
        }
        if (lane == 31) ws[wid] = incl;
        __syncthreads();
        u32 prev_lane = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) prev_lane = 0;
        u32 c = carry_s;
        for (u32 w = 0; w < wid; w++) if (ws[w]) c = ws[w];
        if (idx < m) tile_in[idx] = prev_lane ? prev_lane : c;
        u32 all = carry_s;
        for (u32 w = 0; w < 32; w++) if (ws[w]) all = ws[w];
        __syncthreads();
        if (threadIdx.x == 0) carry_s = all;
        __syncthreads();
    }
}

template <bool WRITE>
__global__ void __launch_bounds__(FA_THREADS)
k_fa_compact(const uint8_t *__restrict__ text, u64 n, const u64 *__restrict__ file_off, int n_files,
             const u32 *__restrict__ tile_in, u32 *__restrict__ tile_cnt, const u64 *__restrict__ tile_off,
             uint8_t *__restrict__ dense, u64 *__restrict__ dense_off)
{
    __shared__ u32 ws[9];
    __shared__ u32 ws2[9];
    u64 g0 = (u64)blockIdx.x * FA_TILE + (u64)threadIdx.x * FA_SPAN;
    u64 g1 = g0 + FA_SPAN < n ? g0 + FA_SPAN : n;
    u32 v = g0 < n ? span_summary(text, g0, g1, file_off, n_files) : 0;
    u32 last;
    u32 in_state = block_last_nonzero_excl(v, ws, &last);
    if (in_state == 0) in_state = tile_in[blockIdx.x];
    bool in_hdr = in_state == 2;

    // pass A: count kept bytes of my span
    u32 kept = 0;
    {
        bool h = in_hdr;
        int nb = g0 < n ? lower_bound_u64(file_off, n_files, g0) : n_files;
        u64 next_start = nb < n_files ? file_off[nb] : ~0ull;
        for (u64 g = g0; g < g1 && g0 < n; g++) {
            if (g == next_start) {
                h = false;
                nb++;
                while (nb < n_files && file_off[nb] == g) nb++;
                next_start = nb < n_files ? file_off[nb] : ~0ull;
            }
            uint8_t c = text[g];
            if (h) { if (c == '\n') h = false; continue; }
            if (c == '\n' || c == '\r') continue;
            if (c == '>') h = true;
            kept++;
        }
    }
    // block exclusive scan of kept
    u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    u32 incl = kept;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (u32)o) incl += t;
    }
    if (lane == 31) ws2[wid] = incl;
    __syncthreads();
    u32 woff = 0, total = 0;
    for (u32 w = 0; w < FA_THREADS / 32; w++) {
        if (w < wid) woff += ws2[w];
        total += ws2[w];
    }
    u32 excl = woff + incl - kept;
    if (!WRITE) {
        if (threadIdx.x == 0) tile_cnt[blockIdx.x] = total;
        return;
    }
    // pass B: write
    u64 o = tile_off[blockIdx.x] + excl;
    bool h = in_hdr;
    if (g0 < n) {
        int nb = lower_bound_u64(file_off, n_files, g0);
        u64 next_start = nb < n_files ? file_off[nb] : ~0ull;
        for (u64 g = g0; g < g1; g++) {
            if (g == next_start) {
                h = false;
                while (nb < n_files && file_off[nb] == g) { dense_off[nb] = o; nb++; }
                next_start = nb < n_files ? file_off[nb] : ~0ull;
            }
            uint8_t c = text[g];
            if (h) { if (c == '\n') h = false; continue; }
            if (c == '\n' || c == '\r') continue;
            uint8_t outc;
            if (mk_is_acgt(c)) outc = c & 0xDF;
            else { outc = 'N'; if (c == '>') h = true; }
            dense[o++] = outc;
        }
    }
}

__global__ void k_widen_u32_u64(const u32 *in, u64 *out, u64 n)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

// Returns the dense stream (device, padded), its length, and the dense offset of every file
// (device array of n_files + 1 entries).
int mk_fasta_compact(mk_ctx *ctx, const uint8_t *d_text, size_t nbytes, const u64 *h_offsets, int n_files,
                     uint8_t **d_dense, u64 *dense_bytes, u64 **d_dense_off)
{
    u64 n_tiles = (nbytes + FA_TILE - 1) / FA_TILE;
    if (n_tiles == 0) n_tiles = 1;
    if (n_tiles > 0x7FFFFFFFull) return MK_ERR_UNSUPPORTED;
    u64 *d_foff, *d_doff, *d_toff;
    u32 *state, *cnt;
    uint8_t *dense;
    CKR(mk_scratch(ctx, SB_FILE_OFF, (size_t)n_files + 1, &d_foff));
    CKR(mk_scratch(ctx, SB_FA_OFF, (size_t)n_files + 1 + n_tiles, &d_doff));
    d_toff = d_doff + n_files + 1;
    CKR(mk_scratch(ctx, SB_FA_STATE, (size_t)2 * n_tiles, &state));
    CKR(mk_scratch(ctx, SB_FA_CNT, (size_t)2 * n_tiles, &cnt));
    CKR(mk_scratch(ctx, SB_FA_DENSE, nbytes + 256, &dense));
    CK(cudaMemcpyAsync(d_foff, h_offsets, (size_t)(n_files + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    ctx->prof.h2d_bytes += (u64)(n_files + 1) * 8;
    u32 *state_in = state + n_tiles, *cnt_scan = cnt + n_tiles;
    k_fa_summary<<<(unsigned)n_tiles, FA_THREADS, 0, ctx->stream>>>(d_text, nbytes, d_foff, n_files, state);
    LAUNCH_COUNT(ctx);
    k_fa_propagate<<<1, 1024, 0, ctx->stream>>>(state, (u32)n_tiles, state_in);
    LAUNCH_COUNT(ctx);
    k_fa_compact<false><<<(unsigned)n_tiles, FA_THREADS, 0, ctx->stream>>>(d_text, nbytes, d_foff, n_files, state_in, cnt,
                                                                         nullptr, nullptr, nullptr);
    LAUNCH_COUNT(ctx);
    u64 total = 0;
    CKR(mk_exclusive_scan_u32(ctx, cnt, cnt_scan, n_tiles, &total));
    if (total >= 0xFFFFFFFFull) {
        snprintf(ctx->err, sizeof(ctx->err), "FASTA batch larger than 4 Gi bases per call; split the batch");
        return MK_ERR_UNSUPPORTED;
    }
    k_widen_u32_u64<<<(unsigned)((n_tiles + 255) / 256), 256, 0, ctx->stream>>>(cnt_scan, d_toff, n_tiles);
    LAUNCH_COUNT(ctx);
    // files that start at or beyond the end of the text (empty trailing files) get `total`
    std::vector<u64> init(n_files + 1, total);
    CK(cudaMemcpyAsync(d_doff, init.data(), (size_t)(n_files + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_fa_compact<true><<<(unsigned)n_tiles, FA_THREADS, 0, ctx->stream>>>(d_text, nbytes, d_foff, n_files, state_in, cnt,
                                                                        d_toff, dense, d_doff);
    LAUNCH_COUNT(ctx);
    CK(cudaMemsetAsync(dense + total, 0, 64, ctx->stream));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    *d_dense = dense;
    *dense_bytes = total;
    *d_dense_off = d_doff;
    return MK_OK;
}
