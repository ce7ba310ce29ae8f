// mk_stream.cu — the dominant kernel of the path: stream sequence text from HBM, find the
// k-mers whose inner substring is in the .shuf pass set, and emit (code, position) candidates.
//
// Replaces the per-base loop of mt_shortreads2koc() (/root/reference/iseq2comem.c:672-720) and,
// in RAW mode (a pre-compacted base stream), of fasta2co() (iseq2comem.c:245-293).
//
// Design (one persistent CTA per SM, 512 threads):
//   * text tiles are staged HBM -> shared memory with 1-D TMA bulk copies (cp.async.bulk +
//     mbarrier), double buffered, tiles claimed in file order with an atomic ticket;
//   * FASTQ record structure is resolved exactly like four fgets() calls per record: every tile
//     counts its '\n' bytes and the line number of the tile start is obtained with a decoupled
//     look-back over per-tile descriptors (single pass over HBM, no pre-scan);
//   * only bytes of sequence lines (line % 4 == 1) are turned into work: aligned 32-byte blocks
//     that intersect a sequence line are compacted into an item list;
//   * per item, 48 bases are packed to 2 bits and the 32 inner windows are tested against a
//     2^20-bit shared-memory bitmap of  S ∪ revcomp(S)  (S = pass set of the .shuf permutation);
//     character validity is NOT checked here — bitmap hits (~1 %) are queued and verified exactly
//     from the text (22 valid ACGT bytes inside one line, canonical strand, exact table), so the
//     fast path is: 2 funnel shifts, 1 AND, 1 LDS, 2 funnel shifts per position.
#include "mk_common.cuh"

struct StreamArgs {
    const uint8_t *text;
    u64 nbytes;
    u64 pos_base;
    u64 line_base;
    u32 tile_bytes;
    u32 n_tiles;
    u64 *tile_desc;
    u32 *tile_counter;
    const u32 *bitmap;
    u32 bitmap_bytes;
    const u64 *ptab;
    u64 *cand_code;
    u64 *cand_pos;
    u64 *cand_count;
    u64 cand_cap;
    u32 *flags;
    u64 *total_newlines;
    KParams kp;
};

#define TBUF_STRIDE (MK_HALO + MK_MAX_TILE + 96) // keeps both buffers 128-byte aligned
#define FLAG_LONG_LINE 2u

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ u32 mbar_try_wait(u64 *bar, u32 parity)
{
    u32 ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, u32 bytes, u64 *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ u64 ld_volatile_u64(const u64 *p)
{
    u64 v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(u64 *p, u64 v)
{
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ u64 warp_sum_u64(u64 v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 512-thread exclusive scan (two barriers); ws holds 16 warp slots + total
__device__ __forceinline__ u32 block_excl_scan_512(u32 v, u32 *ws /*[17]*/, u32 *total)
{
    u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    u32 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (u32)o) incl += t;
    }
    if (lane == 31) ws[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        u32 s = lane < 16 ? ws[lane] : 0;
        u32 si = s;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= (u32)o) si += t;
        }
        if (lane < 16) ws[lane] = si - s;
        if (lane == 15) ws[16] = si;
    }
    __syncthreads();
    *total = ws[16];
    return ws[wid] + incl - v;
}

// Decoupled look-back over tile descriptors (status in bits 63:62: 1 = aggregate, 2 = inclusive
// prefix).  Called by all 32 lanes of warp 0; returns the exclusive prefix of `agg`.
__device__ __forceinline__ u64 tile_lookback(u64 *desc, u32 tile, u64 agg)
{
    const u32 lane = threadIdx.x & 31;
    const u64 VMASK = (1ull << 62) - 1;
    if (tile == 0) {
        if (lane == 0) st_volatile_u64(&desc[0], (2ull << 62) | agg);
        return 0;
    }
    if (lane == 0) st_volatile_u64(&desc[tile], (1ull << 62) | agg);
    u64 excl = 0;
    long long look = (long long)tile - 1;
    for (;;) {
        long long idx = look - (long long)lane;
        u64 d;
        do {
            d = idx >= 0 ? ld_volatile_u64(&desc[idx]) : (2ull << 62);
        } while (__any_sync(0xffffffffu, (d >> 62) == 0));
        u32 m2 = __ballot_sync(0xffffffffu, (d >> 62) == 2);
        u64 val = d & VMASK;
        if (m2) {
            u32 first = __ffs(m2) - 1;
            excl += warp_sum_u64(lane <= first ? val : 0);
            break;
        }
        excl += warp_sum_u64(val);
        look -= 32;
    }
    if (lane == 0) st_volatile_u64(&desc[tile], (2ull << 62) | ((excl + agg) & VMASK));
    return excl;
}

// ---- exact verification of one bitmap hit ------------------------------------------------------
// t points at the LAST base of the candidate k-mer inside the shared-memory text.  Returns true and
// the sketch code when the TL bytes ending at t are all ACGT and the canonical k-mer's inner
// substring is in the pass set (iseq2comem.c:682-699).
__device__ __forceinline__ bool verify_kmer(const uint8_t *t, const KParams &kp, const u64 *__restrict__ ptab, u64 *code)
{
    u64 fwd = 0, rc = 0;
    const int TL = kp.TL;
    for (int i = TL - 1; i >= 0; i--) {
        u32 c = t[-i];
        if (!mk_is_acgt(c)) return false;
        u64 b = mk_code2(c);
        fwd = (fwd << 2) | b;
        rc = (rc >> 2) | ((b ^ 3ull) << kp.crvs_shift);
    }
    fwd &= kp.tupmask;
    u64 u = fwd < rc ? fwd : rc;
    u32 dim = (u32)((u & kp.domask) >> (2 * kp.outctx));
    u32 h = (dim * 0x9E3779B1u) >> 11;
    u32 pf = 0;
    for (;;) {
        h &= kp.ptab_mask;
        u64 e = ptab[h];
        u32 key = (u32)e;
        if (key == 0) return false;
        if (key == dim + 1) {
            pf = (u32)(e >> 32);
            break;
        }
        h++;
    }
    *code = (((u & kp.undomask) + ((u & kp.lowmask) << kp.code_shift)) >> (4 * kp.drlevel)) + (u64)pf;
    return true;
}

// 16 ASCII bases -> 32 bits, base i at bits [2i, 2i+2) (garbage for non-ACGT bytes, by design)
__device__ __forceinline__ u32 pack16(uint4 v)
{
    u32 x0 = (((v.x >> 1) ^ (v.x >> 2)) & 0x03030303u) * 0x01041040u;
    u32 x1 = (((v.y >> 1) ^ (v.y >> 2)) & 0x03030303u) * 0x01041040u;
    u32 x2 = (((v.z >> 1) ^ (v.z >> 2)) & 0x03030303u) * 0x01041040u;
    u32 x3 = (((v.w >> 1) ^ (v.w >> 2)) & 0x03030303u) * 0x01041040u;
    return (x0 >> 24) | ((x1 >> 16) & 0x0000FF00u) | ((x2 >> 8) & 0x00FF0000u) | (x3 & 0xFF000000u);
}

// Probe the 32 k-mer end positions of one aligned 32-byte block against the bitmap.
// blk = shared-memory address of the block's first byte.  Bit j of the result = position j hit.
template <int ROTOFF, u32 WORDMASK, int PREW>
__device__ __forceinline__ u32 probe_block(const uint8_t *blk, const u32 *bm, int shift_s)
{
    u32 W[PREW + 3];
    const uint4 *q = reinterpret_cast<const uint4 *>(blk - 16 * PREW);
#pragma unroll
    for (int i = 0; i < PREW + 2; i++) W[i] = pack16(q[i]);
    W[PREW + 2] = 0;
    u32 A[4];
    A[0] = __funnelshift_r(W[0], W[1], shift_s);
    A[1] = __funnelshift_r(W[1], W[2], shift_s);
    A[2] = __funnelshift_r(W[2], W[3], shift_s);
    A[3] = 0;
    u32 hits = 0;
#pragma unroll
    for (int j = 0; j < 32; j++) {
        const int o = 2 * j, o2 = 2 * j + ROTOFF;
        u32 v = __funnelshift_r(A[o >> 5], A[(o >> 5) + 1], o & 31);
        u32 r = __funnelshift_r(A[o2 >> 5], A[(o2 >> 5) + 1], o2 & 31);
        u32 word = *reinterpret_cast<const u32 *>(reinterpret_cast<const char *>(bm) + (v & WORDMASK));
        u32 rot = __funnelshift_l(word, word, r);
        hits = __funnelshift_l(rot, hits, 1);
    }
    return __brev(hits);
}

// ---- the kernel --------------------------------------------------------------------------------
template <int ROTOFF, u32 WORDMASK, int PREW, bool RAW>
__global__ void __launch_bounds__(MK_STREAM_THREADS, 1) k_stream(const __grid_constant__ StreamArgs A)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const u32 tid = threadIdx.x;
    const u32 lane = tid & 31, wid = tid >> 5;
    const u32 bm_bytes = (A.bitmap_bytes + 127u) & ~127u;
    u32 *bm = reinterpret_cast<u32 *>(smem);
    uint8_t *tbuf = smem + bm_bytes;
    uint8_t *p = tbuf + 2 * TBUF_STRIDE;
    uint16_t *starts = reinterpret_cast<uint16_t *>(p); p += MK_MAXL * 2;
    uint16_t *ends = reinterpret_cast<uint16_t *>(p);   p += MK_MAXL * 2;
    u32 *posmask = reinterpret_cast<u32 *>(p);          p += (MK_MAX_TILE / 32) * 4;
    uint16_t *items = reinterpret_cast<uint16_t *>(p);  p += (MK_MAX_TILE / 32) * 2;
    uint16_t *hitq = reinterpret_cast<uint16_t *>(p);   p += MK_HITCAP * 2;
    u32 *ws = reinterpret_cast<u32 *>(p);               p += 20 * 4;
    u32 *ws2 = reinterpret_cast<u32 *>(p);              p += 20 * 4;
    u64 *bar = reinterpret_cast<u64 *>(p);              p += 16;
    u64 *s_P = reinterpret_cast<u64 *>(p);              p += 8;
    u32 *s_tile = reinterpret_cast<u32 *>(p);           p += 8;
    u32 *s_nhits = reinterpret_cast<u32 *>(p);          p += 8;

    // bitmap -> shared memory
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(A.bitmap);
        uint4 *dst = reinterpret_cast<uint4 *>(bm);
        for (u32 i = tid; i < A.bitmap_bytes / 16; i += MK_STREAM_THREADS) dst[i] = src[i];
    }
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const u32 TB = A.tile_bytes;
    auto issue_load = [&](int stage, u32 t) {
        u64 T = (u64)t * TB;
        u64 rem = A.nbytes - T;
        u32 tb = rem < TB ? (u32)rem : TB;
        uint8_t *dst = tbuf + stage * TBUF_STRIDE;
        const uint8_t *src = A.text + T;
        u32 bytes = tb;
        if (t > 0) { src -= MK_HALO; bytes += MK_HALO; } else { dst += MK_HALO; }
        bytes = (bytes + 15u) & ~15u;
        fence_proxy_async(); // order earlier generic-proxy writes to this buffer before the bulk copy
        mbar_expect_tx(&bar[stage], bytes);
        tma_load_1d(dst, src, bytes, &bar[stage]);
    };

    if (tid == 0) {
        u32 t = atomicAdd(A.tile_counter, 1u);
        s_tile[0] = t;
        if (t < A.n_tiles) issue_load(0, t);
    }
    __syncthreads();
    u32 cur = s_tile[0];
    int stage = 0;
    u32 parity0 = 0, parity1 = 0;

    while (cur < A.n_tiles) {
        if (tid == 0) {
            u32 t = atomicAdd(A.tile_counter, 1u);
            s_tile[stage ^ 1] = t;
            if (t < A.n_tiles) issue_load(stage ^ 1, t);
            *s_nhits = 0;
        }
        if (stage == 0) { mbar_wait(&bar[0], parity0); parity0 ^= 1; }
        else            { mbar_wait(&bar[1], parity1); parity1 ^= 1; }

        uint8_t *tx = tbuf + stage * TBUF_STRIDE; // tx[0..HALO) = left context, tile bytes from tx+HALO
        const u64 T = (u64)cur * TB;
        const u64 rem = A.nbytes - T;
        const u32 tb = rem < TB ? (u32)rem : TB;
        if (cur == 0 && tid < MK_HALO / 4) reinterpret_cast<u32 *>(tx)[tid] = 0; // no text before the file
        if (tb < TB) { // last tile: blank everything past the end of the text
            for (u32 i = MK_HALO + tb + tid; i < MK_HALO + TB; i += MK_STREAM_THREADS) tx[i] = 0;
            fence_proxy_async();
        }
        // zero posmask (2 words per thread)
        posmask[2 * tid] = 0;
        posmask[2 * tid + 1] = 0;
        __syncthreads();

        u32 n_items = 0;
        if (RAW) {
            // every byte is a base-stream byte: all blocks below tb are items
            u32 nblk = (tb + 31) >> 5;
            for (u32 b = tid; b < nblk; b += MK_STREAM_THREADS) {
                items[b] = (uint16_t)b;
                u32 left = tb - 32 * b;
                posmask[b] = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
            }
            n_items = nblk;
            __syncthreads();
        } else {
            // ---- A1: newline mask of my 64 bytes ----
            u64 nl = 0;
            const u32 off = tid * 64;
            if (off < TB) {
                const uint4 *q = reinterpret_cast<const uint4 *>(tx + MK_HALO + off);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    uint4 v = q[c];
                    u32 w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        u32 y = w4[j] ^ 0x0A0A0A0Au;
                        u32 t7 = (y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
                        u32 f = ~(t7 | y) & 0x80808080u;            // 0x80 where the byte is '\n'
                        u32 nib = __umulhi(f, 0x02040810u) & 0xFu;   // gather bits 7,15,23,31
                        nl |= (u64)nib << (16 * c + 4 * j);
                    }
                }
            }
            const u32 cnt = __popcll(nl);
            u32 total;
            const u32 excl = block_excl_scan_512(cnt, ws, &total);
            if (wid == 0) {
                u64 e = tile_lookback(A.tile_desc, cur, (u64)total);
                if (lane == 0) {
                    *s_P = A.line_base + e;
                    if (cur == A.n_tiles - 1) *A.total_newlines = A.line_base + e + total;
                    if (total == 0 && tb == TB && TB >= 4096) atomicOr(A.flags, FLAG_LONG_LINE);
                }
            }
            __syncthreads();
            const u64 P = *s_P;                     // '\n' bytes before the tile == line index of its first byte
            const u64 q0 = (P + 2) >> 2;            // record number of the first sequence line touching the tile
            const u32 in_seq = ((P & 3) == 1);
            const u32 n_seq = (u32)(((P + total + 3) >> 2) - ((P + 3) >> 2)) + in_seq;

            // ---- markers -> posmask, in windows of MK_MAXL sequence lines ----
            for (u32 w0 = 0; w0 < n_seq; w0 += MK_MAXL) {
                const u32 wn = (n_seq - w0) < MK_MAXL ? (n_seq - w0) : MK_MAXL;
                for (u32 i = tid; i < wn; i += MK_STREAM_THREADS) {
                    starts[i] = (uint16_t)((w0 == 0 && i == 0 && in_seq) ? 0 : tb);
                    ends[i] = (uint16_t)tb;
                }
                __syncthreads();
                if (cnt) {
                    u64 m = nl;
                    u64 line = P + excl;            // index of the line my first newline terminates
                    while (m) {
                        u32 bit = __ffsll((long long)m) - 1;
                        m &= m - 1;
                        u32 pos = off + bit;
                        u32 ph = (u32)line & 3u;
                        if (ph <= 1) {
                            u64 ord = (line >> 2) - q0;
                            if (ord >= w0 && ord < (u64)w0 + wn) {
                                if (ph == 0) starts[ord - w0] = (uint16_t)(pos + 1 > tb ? tb : pos + 1);
                                else         ends[ord - w0] = (uint16_t)pos;
                            }
                        }
                        line++;
                    }
                }
                __syncthreads();
                for (u32 i = tid; i < wn; i += MK_STREAM_THREADS) {
                    u32 s = starts[i], e = ends[i];
                    if (e > s) {
                        u32 b0 = s >> 5, b1 = (e - 1) >> 5;
                        for (u32 b = b0; b <= b1; b++) {
                            u32 msk = 0xffffffffu;
                            if (b == b0) msk &= 0xffffffffu << (s & 31);
                            if (b == b1) msk &= 0xffffffffu >> (31 - ((e - 1) & 31));
                            atomicOr(&posmask[b], msk);
                        }
                    }
                }
                __syncthreads();
            }
            // ---- compact blocks that hold sequence bytes into the item list ----
            const u32 a0 = posmask[2 * tid] != 0, a1 = posmask[2 * tid + 1] != 0;
            u32 tot2;
            u32 ioff = block_excl_scan_512(a0 + a1, ws2, &tot2);
            if (a0) items[ioff++] = (uint16_t)(2 * tid);
            if (a1) items[ioff] = (uint16_t)(2 * tid + 1);
            n_items = tot2;
            __syncthreads();
        }

        // ---- B: probe ----
        for (u32 it = tid; it < n_items; it += MK_STREAM_THREADS) {
            const u32 b = items[it];
            u32 hits = probe_block<ROTOFF, WORDMASK, PREW>(tx + MK_HALO + 32 * b, bm, A.kp.shift_s);
            hits &= posmask[b];
            while (hits) {
                u32 j = __ffs(hits) - 1;
                hits &= hits - 1;
                u32 slot = atomicAdd(s_nhits, 1u);
                u32 pos = 32 * b + j;
                if (slot < MK_HITCAP) {
                    hitq[slot] = (uint16_t)pos;
                } else { // queue full: verify in place
                    u64 code;
                    if (verify_kmer(tx + MK_HALO + pos, A.kp, A.ptab, &code)) {
                        u64 idx = atomicAdd((unsigned long long *)A.cand_count, 1ull);
                        if (idx < A.cand_cap) {
                            A.cand_code[idx] = code;
                            A.cand_pos[idx] = A.pos_base + T + pos;
                        }
                    }
                }
            }
        }
        __syncthreads();
        // ---- C: exact verification of queued hits ----
        {
            u32 nh = *s_nhits;
            if (nh > MK_HITCAP) nh = MK_HITCAP;
            for (u32 h = tid; h < nh; h += MK_STREAM_THREADS) {
                u32 pos = hitq[h];
                u64 code;
                if (verify_kmer(tx + MK_HALO + pos, A.kp, A.ptab, &code)) {
                    u64 idx = atomicAdd((unsigned long long *)A.cand_count, 1ull);
                    if (idx < A.cand_cap) {
                        A.cand_code[idx] = code;
                        A.cand_pos[idx] = A.pos_base + T + pos;
                    }
                }
            }
        }
        __syncthreads();
        cur = s_tile[stage ^ 1];
        stage ^= 1;
    }
}

// ---- tail rule ---------------------------------------------------------------------------------
// A record counts only if all four fgets() calls succeed (iseq2comem.c:673).  Equivalent rule on
// the byte stream: a sequence line is kept iff some '\n' at offset <= n-2 follows its own
// terminator.  With e2 = last '\n' at offset <= n-2 and e1 = last '\n' before e2, every k-mer
// position > e1 belongs to a dropped line.  One warp scans backwards; writes e1 (or -1).
__global__ void k_tail_cut(const uint8_t *__restrict__ text, u64 n, long long *out)
{
    const u32 lane = threadIdx.x;
    long long found[2] = {-1, -1};
    int nfound = 0;
    long long hi = (long long)n - 2; // highest offset considered
    while (hi >= 0 && nfound < 2) {
        long long idx = hi - (long long)lane;
        bool is_nl = idx >= 0 && text[idx] == '\n';
        u32 m = __ballot_sync(0xffffffffu, is_nl);
        while (m && nfound < 2) {
            u32 l = __ffs(m) - 1; // smallest lane = highest offset
            found[nfound++] = hi - (long long)l;
            m &= m - 1;
        }
        hi -= 32;
    }
    if (lane == 0) *out = (nfound == 2) ? found[1] : -1;
}

int mk_tail_cut(mk_ctx *ctx, const uint8_t *d_text, size_t nbytes, long long *keep_below)
{
    long long *d_out;
    CKR(mk_scratch(ctx, SB_MISC, 64, &d_out));
    k_tail_cut<<<1, 32, 0, ctx->stream>>>(d_text, (u64)nbytes, d_out);
    LAUNCH_COUNT(ctx);
    CK(cudaMemcpyAsync(keep_below, d_out, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->prof.d2h_bytes += 8;
    return MK_OK;
}

// ---- newline count (shard line_base for multi-GPU) ---------------------------------------------
__global__ void __launch_bounds__(256) k_count_newlines(const uint8_t *__restrict__ text, u64 n, u64 *out)
{
    u64 cnt = 0;
    u64 nvec = n / 16;
    const uint4 *v = reinterpret_cast<const uint4 *>(text);
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (u64)gridDim.x * blockDim.x) {
        uint4 q = v[i];
        u32 w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            u32 y = w4[j] ^ 0x0A0A0A0Au;
            u32 t7 = (y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
            cnt += __popc(~(t7 | y) & 0x80808080u);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (u64 i = nvec * 16; i < n; i++) cnt += text[i] == '\n';
    cnt = warp_sum_u64(cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd((unsigned long long *)out, cnt);
}

extern "C" int mk_count_newlines_device(mk_ctx *ctx, const void *d_text, size_t nbytes, uint64_t *count)
{
    if (!ctx || !count || ((uintptr_t)d_text & 15)) return MK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    u64 *d_out;
    CKR(mk_scratch(ctx, SB_MISC, 64, &d_out));
    CK(cudaMemsetAsync(d_out, 0, 8, ctx->stream));
    k_count_newlines<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((const uint8_t *)d_text, (u64)nbytes, d_out);
    LAUNCH_COUNT(ctx);
    CK(cudaMemcpyAsync(count, d_out, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MK_OK;
}

// ---- host driver -------------------------------------------------------------------------------
typedef void (*stream_kernel_t)(const StreamArgs);

template <int ROTOFF, u32 WORDMASK>
static stream_kernel_t pick2(int prew, bool raw)
{
    if (prew == 1) return raw ? k_stream<ROTOFF, WORDMASK, 1, true> : k_stream<ROTOFF, WORDMASK, 1, false>;
    return raw ? k_stream<ROTOFF, WORDMASK, 2, true> : k_stream<ROTOFF, WORDMASK, 2, false>;
}
static stream_kernel_t pick_kernel(const KParams &kp, bool raw)
{
    if (kp.mw >= 20) return pick2<17, 0x1FFFCu>(kp.prew, raw);
    if (kp.mw == 16) return pick2<13, 0x1FFCu>(kp.prew, raw);
    if (kp.mw == 12) return pick2<9, 0x1FCu>(kp.prew, raw);
    return nullptr;
}

static size_t stream_smem_bytes(u32 bitmap_bytes)
{
    size_t s = (bitmap_bytes + 127u) & ~127u;
    s += 2 * TBUF_STRIDE;
    s += MK_MAXL * 2 * 2;
    s += (MK_MAX_TILE / 32) * 4 + (MK_MAX_TILE / 32) * 2;
    s += MK_HITCAP * 2;
    s += 20 * 4 * 2 + 16 + 8 + 8 + 8;
    return s + 64;
}

int mk_stream_fastq(mk_ctx *ctx, const uint8_t *d_text, size_t nbytes, u64 pos_base, u64 line_base, bool raw_mode,
                    u64 **d_cand_code, u64 **d_cand_pos, u64 *n_cand, u64 *n_newlines)
{
    *n_cand = 0;
    if (n_newlines) *n_newlines = line_base;
    if (nbytes == 0) return MK_OK;
    if ((uintptr_t)d_text & 15) {
        snprintf(ctx->err, sizeof(ctx->err), "device text pointer must be 16-byte aligned");
        return MK_ERR_ARG;
    }
    const KParams &kp = ctx->kp;
    stream_kernel_t kern = pick_kernel(kp, raw_mode);
    if (!kern) {
        snprintf(ctx->err, sizeof(ctx->err), "unsupported inner substring width subk=%d", kp.subk);
        return MK_ERR_UNSUPPORTED;
    }
    u32 tile_bytes = raw_mode ? 16384u : 32768u;
    if (const char *e = getenv(raw_mode ? "MK_RAW_TILE_BYTES" : "MK_TILE_BYTES")) {
        u32 v = (u32)atoi(e);
        if (v >= 64 && v <= MK_MAX_TILE) tile_bytes = v & ~63u;
    }
    u64 n_tiles64 = (nbytes + tile_bytes - 1) / tile_bytes;
    if (n_tiles64 > 0x7FFFFFFFull) return MK_ERR_UNSUPPORTED;
    u32 n_tiles = (u32)n_tiles64;

    u64 *desc, *counters;
    CKR(mk_scratch(ctx, SB_TILE_DESC, (size_t)n_tiles, &desc));
    CKR(mk_scratch(ctx, SB_COUNTERS, 8, &counters));
    double rate = (double)kp.dim_end / (double)(1ull << (4 * kp.subk));
    if (rate > 1.0) rate = 1.0;
    u64 cap = (u64)((double)nbytes * rate * 1.25) + 65536;
    size_t smem = stream_smem_bytes(ctx->bitmap_words * 4);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    for (int attempt = 0; attempt < 2; attempt++) {
        u64 *cc, *cp;
        CKR(mk_scratch(ctx, SB_CAND_CODE, (size_t)cap, &cc));
        CKR(mk_scratch(ctx, SB_CAND_POS, (size_t)cap, &cp));
        CK(cudaMemsetAsync(desc, 0, (size_t)n_tiles * 8, ctx->stream));
        CK(cudaMemsetAsync(counters, 0, 64, ctx->stream));
        StreamArgs a;
        a.text = d_text; a.nbytes = nbytes; a.pos_base = pos_base; a.line_base = line_base;
        a.tile_bytes = tile_bytes; a.n_tiles = n_tiles; a.tile_desc = desc;
        a.cand_count = counters + 0; a.total_newlines = counters + 1;
        a.tile_counter = (u32 *)(counters + 2); a.flags = (u32 *)(counters + 2) + 1;
        a.bitmap = ctx->d_bitmap; a.bitmap_bytes = ctx->bitmap_words * 4; a.ptab = ctx->d_ptab;
        a.cand_code = cc; a.cand_pos = cp; a.cand_cap = cap; a.kp = kp;
        u32 grid = n_tiles < (u32)ctx->sm_count ? n_tiles : (u32)ctx->sm_count;
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
        kern<<<grid, MK_STREAM_THREADS, smem, ctx->stream>>>(a);
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        LAUNCH_COUNT(ctx);
        CK(cudaGetLastError());
        u64 h[4];
        CK(cudaMemcpyAsync(h, counters, 32, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->prof.d2h_bytes += 32;
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        ctx->prof.stream_kernel_ms += ms;
        ctx->prof.stream_kernel_launches++;
        ctx->prof.stream_kernel_bytes += nbytes;
        u32 flags = (u32)(h[2] >> 32);
        if (!raw_mode && (flags & FLAG_LONG_LINE)) {
            snprintf(ctx->err, sizeof(ctx->err), "FASTQ line longer than 4095 bytes");
            return MK_ERR_LONG_LINE;
        }
        if (h[0] > cap) { // candidate buffer too small: size it exactly and run again
            cap = h[0] + 1024;
            continue;
        }
        *n_cand = h[0];
        if (n_newlines) *n_newlines = raw_mode ? line_base : h[1];
        *d_cand_code = cc;
        *d_cand_pos = cp;
        return MK_OK;
    }
    snprintf(ctx->err, sizeof(ctx->err), "candidate buffer overflow persisted");
    return MK_ERR_NOMEM;
}
