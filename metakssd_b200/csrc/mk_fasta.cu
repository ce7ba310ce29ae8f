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

#define FA_THREADS 256                       // 8 warps per block, one tile per warp
#define FA_STEP 512                          // bytes per warp step: 16 per lane, one 128-bit load
#define FA_TILE (64 * FA_STEP)               // 32 KB per warp tile
#define FA_TILES_PER_BLOCK (FA_THREADS / 32)

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

// ---- per-lane classification of 16 bytes, SIMD in registers -----------------------------------------
// 0x80 in every byte of w equal to c (c < 0x80), exact
__device__ __forceinline__ u32 eq_bytes(u32 w, u32 c4)
{
    u32 t = ((w ^ c4) & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~(t | w) & 0x80808080u;
}
// 0x80 in every byte that is one of ACGT in either case, exact
__device__ __forceinline__ u32 acgt_bytes(u32 w)
{
    u32 ta = ((w ^ 0x41414141u) & 0x5F5F5F5Fu) + 0x7F7F7F7Fu;   // bit 7 clear <=> byte is 'A' or 'a' (given bit 7 of w clear)
    u32 tc = ((w ^ 0x43434343u) & 0x5F5F5F5Fu) + 0x7F7F7F7Fu;
    u32 tg = ((w ^ 0x47474747u) & 0x5F5F5F5Fu) + 0x7F7F7F7Fu;
    u32 tt = ((w ^ 0x54545454u) & 0x5F5F5F5Fu) + 0x7F7F7F7Fu;
    return ~((ta & tc & tg & tt) | w) & 0x80808080u;
}
// the four 0x80 flags of each of four words -> 16-bit mask (bit i = byte i)
__device__ __forceinline__ u32 flags16(u32 f0, u32 f1, u32 f2, u32 f3)
{
    u32 m = 0;
    m = __funnelshift_r(m, __umulhi(f0, 0x02040810u), 4);
    m = __funnelshift_r(m, __umulhi(f1, 0x02040810u), 4);
    m = __funnelshift_r(m, __umulhi(f2, 0x02040810u), 4);
    m = __funnelshift_r(m, __umulhi(f3, 0x02040810u), 4);
    return m >> 16;
}

struct FaLane {
    u32 w[4];        // the 16 bytes
    u32 valid;       // bytes inside the text
    u32 nl, crnl, gt, acgt;   // 16-bit masks
};

__device__ __forceinline__ void fa_load(const uint8_t *__restrict__ text, u64 g, u64 n, FaLane &L)
{
    L.w[0] = L.w[1] = L.w[2] = L.w[3] = 0;
    if (g + 16 <= n) {
        uint4 v = *reinterpret_cast<const uint4 *>(text + g);
        L.w[0] = v.x; L.w[1] = v.y; L.w[2] = v.z; L.w[3] = v.w;
        L.valid = 0xFFFFu;
    } else if (g < n) {               // last, partial vector of the text: byte loads, nothing read past the end
        const u32 m = (u32)(n - g);
        for (u32 i = 0; i < m; i++) L.w[i >> 2] |= (u32)text[g + i] << (8 * (i & 3));
        L.valid = (1u << m) - 1u;
    } else {
        L.valid = 0;
    }
    u32 fn[4], fr[4], fg[4], fa[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        fn[i] = eq_bytes(L.w[i], 0x0A0A0A0Au);
        fr[i] = eq_bytes(L.w[i], 0x0D0D0D0Du);
        fg[i] = eq_bytes(L.w[i], 0x3E3E3E3Eu);
        fa[i] = acgt_bytes(L.w[i]);
    }
    L.nl = flags16(fn[0], fn[1], fn[2], fn[3]) & L.valid;
    L.crnl = (flags16(fr[0], fr[1], fr[2], fr[3]) & L.valid) | L.nl;
    L.gt = flags16(fg[0], fg[1], fg[2], fg[3]) & L.valid;
    L.acgt = flags16(fa[0], fa[1], fa[2], fa[3]) & L.valid;
}

// Header state of the 512 bytes of a warp step as a carry chain: adding the '>' bits (starts) to the
// mask of non-terminator bytes makes a carry run from a '>' through the header up to and including its
// '\n' (or the last byte of a file).  Bit i of the result = byte i of this lane is inside a header.
// `end` = terminator bits ('\n', or the byte before a file start), `gt` = '>' bits that may start a
// header; wcin / wcout = state entering / leaving the warp step.
__device__ __forceinline__ u32 fa_header_bits(u32 end, u32 gt, u32 wcin, u32 &wcout)
{
    const u32 lane = threadIdx.x & 31;
    const u32 M = ~end & 0xFFFFu, S = gt;
    const u32 g = (M + S) >> 16;                    // carry out of my 16 bytes without / with carry in
    const u32 p = ((M + S + 1u) >> 16) & ~g;
    const u32 G = __ballot_sync(0xffffffffu, g), P = __ballot_sync(0xffffffffu, p);
    const u32 A = G | P;
    const u64 W = (u64)A + (u64)G + (u64)wcin;      // carry look-ahead across the 32 lanes
    const u32 C = (u32)W ^ A ^ G;                    // bit l = carry into lane l
    wcout = (u32)(W >> 32);
    const u32 cin = (C >> lane) & 1u;
    return ((M + S + cin) ^ M ^ S) & 0xFFFFu;
}

// file starts inside (s0, s0 + FA_STEP]: the byte before each one ends any header and starts none
__device__ __forceinline__ void fa_file_ends(const u64 *__restrict__ file_off, int n_files, int nb, u64 s0, u64 g,
                                             u32 &end, u32 &gt)
{
    for (int j = nb; j < n_files && file_off[j] <= s0 + FA_STEP; j++) {
        const u64 b = file_off[j];
        if (b > s0 && b - 1 >= g && b - 1 < g + 16) {
            end |= 1u << (u32)(b - 1 - g);
            gt &= ~(1u << (u32)(b - 1 - g));
        }
    }
}

// pass 1, one warp per tile, assuming the tile is entered outside a header:
//   tile_state = kind of the last event (0 none, 1 terminator / file start, 2 '>'),
//   cnt0 = bytes kept, pre = those of them that lie before the first terminator (dropped instead if the
//   tile is in fact entered inside a header)
__global__ void __launch_bounds__(FA_THREADS)
k_fa_summary(const uint8_t *__restrict__ text, u64 n, const u64 *__restrict__ file_off, int n_files, u64 n_tiles,
             u32 *__restrict__ tile_state, u32 *__restrict__ cnt0, u32 *__restrict__ pre)
{
    const u32 lane = threadIdx.x & 31;
    const u64 tile = (u64)blockIdx.x * FA_TILES_PER_BLOCK + (threadIdx.x >> 5);
    if (tile >= n_tiles) return;
    const u64 t0 = tile * FA_TILE;
    int nb = lower_bound_u64(file_off, n_files, t0);
    u32 kind = (nb < n_files && file_off[nb] == t0) ? 1u : 0u;    // a file starting exactly here
    bool seen_end = kind != 0;
    u32 wc = 0, kept_all = 0, kept_pre = 0;
    for (u64 s0 = t0; s0 < t0 + FA_TILE && s0 < n; s0 += FA_STEP) {
        const u64 g = s0 + 16ull * lane;
        // Fast path (nearly every step of a genome): outside a header, the first terminator of the tile already
        // seen, the whole step inside the text and inside one file, and no '>' in it.  Then nothing but '\n' and
        // '\r' is dropped and the only possible event is a terminator: three byte compares per word and one
        // population count, no masks, no carry chain.
        if (wc == 0 && seen_end && s0 + FA_STEP <= n && !(nb < n_files && file_off[nb] <= s0 + FA_STEP)) {
            const uint4 v = *reinterpret_cast<const uint4 *>(text + g);
            const u32 w[4] = {v.x, v.y, v.z, v.w};
            u32 fn = 0, fd = 0, fg = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const u32 a = eq_bytes(w[i], 0x0A0A0A0Au);
                fn |= a;
                fd += __popc(a | eq_bytes(w[i], 0x0D0D0D0Du));
                fg |= eq_bytes(w[i], 0x3E3E3E3Eu);
            }
            if (!__any_sync(0xffffffffu, fg != 0)) {
                kept_all += 16u - fd;
                if (__any_sync(0xffffffffu, fn != 0)) kind = 1u;
                continue;
            }
        }
        FaLane L;
        fa_load(text, g, n, L);
        u32 end = L.nl, gt = L.gt;
        if (nb < n_files && file_off[nb] <= s0 + FA_STEP) {
            fa_file_ends(file_off, n_files, nb, s0, g, end, gt);
            while (nb < n_files && file_off[nb] < s0 + FA_STEP) nb++;
        }
        u32 wcout;
        const u32 hdr = fa_header_bits(end, gt, wc, wcout);
        wc = wcout;
        const u32 keep = ~hdr & ~L.crnl & L.valid;
        kept_all += __popc(keep);
        const u32 E = __ballot_sync(0xffffffffu, end != 0);
        if (!seen_end) {
            if (E == 0) kept_pre += __popc(keep);
            else {
                const u32 fl = __ffs(E) - 1;
                if (lane < fl) kept_pre += __popc(keep);
                else if (lane == fl) kept_pre += __popc(keep & ((2u << (__ffs(end) - 1)) - 1u));
                seen_end = true;
            }
        }
        const u32 ev = end | gt;
        const u32 V = __ballot_sync(0xffffffffu, ev != 0);
        if (V) {
            const u32 top = 31 - __clz(V);
            const u32 k = (gt >> (31 - __clz(ev | 1u))) & 1u ? 2u : 1u;   // (only lane `top` matters)
            kind = __shfl_sync(0xffffffffu, k, top);
        }
    }
    kept_all = __reduce_add_sync(0xffffffffu, kept_all);
    kept_pre = __reduce_add_sync(0xffffffffu, kept_pre);
    if (lane == 0) { tile_state[tile] = kind; cnt0[tile] = kept_all; pre[tile] = kept_pre; }
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

// bytes each tile really keeps, given the state it is entered in
__global__ void k_fa_counts(const u32 *__restrict__ tile_in, const u32 *__restrict__ cnt0, const u32 *__restrict__ pre,
                            u64 n_tiles, u32 *__restrict__ cnt)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_tiles) cnt[i] = cnt0[i] - (tile_in[i] == 2 ? pre[i] : 0u);
}

// pass 2: the same walk with the true entry state; kept bytes go to the dense stream (ACGT upper-cased,
// everything else 'N'), file starts record their dense offset.  Every warp stages its output in a 1 KB
// shared-memory ring and writes it out in aligned 16-byte vectors (byte stores only for the first and
// last few bytes of the tile).
__global__ void __launch_bounds__(FA_THREADS)
k_fa_write(const uint8_t *__restrict__ text, u64 n, const u64 *__restrict__ file_off, int n_files, u64 n_tiles,
           const u32 *__restrict__ tile_in, const u64 *__restrict__ tile_off, uint8_t *__restrict__ dense,
           u64 *__restrict__ dense_off)
{
    __shared__ __align__(16) uint8_t ring_all[FA_TILES_PER_BLOCK][1024];
    const u32 lane = threadIdx.x & 31;
    uint8_t *ring = ring_all[threadIdx.x >> 5];
    const u64 tile = (u64)blockIdx.x * FA_TILES_PER_BLOCK + (threadIdx.x >> 5);
    if (tile >= n_tiles) return;
    const u64 t0 = tile * FA_TILE;
    int nb = lower_bound_u64(file_off, n_files, t0);
    const bool starts_file = nb < n_files && file_off[nb] == t0;
    u32 wc = (!starts_file && tile_in[tile] == 2) ? 1u : 0u;
    const u64 out0 = tile_off[tile];
    const u64 base = out0 & ~15ull;              // ring byte i (absolute count) <-> dense[base + i]
    const u32 skew = (u32)(out0 - base);
    u32 fill = skew, flushed = 0;                // absolute byte counts since `base`
    auto flush = [&](bool final) {
        __syncwarp();
        const u32 v_end = fill >> 4;             // full vectors available
        for (u32 v = (flushed >> 4) + lane; v < v_end; v += 32) {
            const uint4 q = *reinterpret_cast<const uint4 *>(ring + ((v << 4) & 1023u));
            if (v == 0 && skew) {                // the first vector also holds bytes of the tile before
                const uint8_t *b = reinterpret_cast<const uint8_t *>(&q);
                for (u32 i = skew; i < 16; i++) dense[base + i] = b[i];
            } else {
                *reinterpret_cast<uint4 *>(dense + base + ((u64)v << 4)) = q;
            }
        }
        flushed = v_end << 4;
        if (final) {                              // the last, partial vector
            const u32 i = flushed + lane;
            if (lane < 16 && i < fill && i >= skew) dense[base + i] = ring[i & 1023u];
        }
        __syncwarp();
    };
    for (u64 s0 = t0; s0 < t0 + FA_TILE && s0 < n; s0 += FA_STEP) {
        const u64 g = s0 + 16ull * lane;
        FaLane L;
        fa_load(text, g, n, L);
        u32 end = L.nl, gt = L.gt;
        const bool files_here = nb < n_files && file_off[nb] <= s0 + FA_STEP;
        if (files_here) fa_file_ends(file_off, n_files, nb, s0, g, end, gt);
        u32 hdr = 0;
        if (wc != 0 || __any_sync(0xffffffffu, gt != 0)) {     // (else: no header in or entering this step, no carry chain)
            u32 wcout;
            hdr = fa_header_bits(end, gt, wc, wcout);
            wc = wcout;
        }
        const u32 keep = ~hdr & ~L.crnl & L.valid;
        const u32 cnt = __popc(keep);
        u32 incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 x = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (u32)o) incl += x;
        }
        const u32 total = __shfl_sync(0xffffffffu, incl, 31);
        u32 o = fill + incl - cnt;               // absolute byte index of my first kept byte
        if (files_here) {                       // dense offset of every file starting inside this step
            for (; nb < n_files && file_off[nb] < s0 + FA_STEP; nb++) {
                const u64 b = file_off[nb];
                if (b >= g && b < g + 16) dense_off[nb] = base + o + __popc(keep & ((1u << (u32)(b - g)) - 1u));
            }
        }
        if (keep) {
            u32 ob[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const u32 m = (acgt_bytes(L.w[i]) >> 7) * 0xFFu;           // 0xFF in every ACGT byte
                ob[i] = ((L.w[i] & 0xDFDFDFDFu) & m) | (0x4E4E4E4Eu & ~m);
            }
#pragma unroll
            for (int i = 0; i < 16; i++)
                if (keep & (1u << i)) ring[(o++) & 1023u] = (uint8_t)(ob[i >> 2] >> (8 * (i & 3)));
        }
        fill += total;
        flush(false);
    }
    flush(true);
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
    if ((uintptr_t)d_text & 15) {
        snprintf(ctx->err, sizeof(ctx->err), "device text pointer must be 16-byte aligned");
        return MK_ERR_ARG;
    }
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
    CKR(mk_scratch(ctx, SB_FA_CNT, (size_t)4 * n_tiles, &cnt));
    CKR(mk_scratch(ctx, SB_FA_DENSE, nbytes + 256, &dense));
    CK(cudaMemcpyAsync(d_foff, h_offsets, (size_t)(n_files + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    ctx->prof.h2d_bytes += (u64)(n_files + 1) * 8;
    u32 *state_in = state + n_tiles, *cnt_scan = cnt + n_tiles, *cnt0 = cnt + 2 * n_tiles, *pre = cnt + 3 * n_tiles;
    const unsigned nblk = (unsigned)((n_tiles + FA_TILES_PER_BLOCK - 1) / FA_TILES_PER_BLOCK);
    k_fa_summary<<<nblk, FA_THREADS, 0, ctx->stream>>>(d_text, nbytes, d_foff, n_files, n_tiles, state, cnt0, pre);
    LAUNCH_COUNT(ctx);
    k_fa_propagate<<<1, 1024, 0, ctx->stream>>>(state, (u32)n_tiles, state_in);
    LAUNCH_COUNT(ctx);
    k_fa_counts<<<(unsigned)((n_tiles + 255) / 256), 256, 0, ctx->stream>>>(state_in, cnt0, pre, n_tiles, cnt);
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
    k_fa_write<<<nblk, FA_THREADS, 0, ctx->stream>>>(d_text, nbytes, d_foff, n_files, n_tiles, state_in, d_toff, dense, d_doff);
    LAUNCH_COUNT(ctx);
    CK(cudaMemsetAsync(dense + total, 0, 64, ctx->stream));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    *d_dense = dense;
    *dense_bytes = total;
    *d_dense_off = d_doff;
    return MK_OK;
}
