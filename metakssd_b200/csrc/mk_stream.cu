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
#include "mk_stream_dev.cuh"
#include "mk_stream3.cuh"


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
        u32 spins = 0;
        do {
            d = idx >= 0 ? ld_volatile_u64(&desc[idx]) : (2ull << 62);
            if (++spins > WD_LIMIT) return ~0ull;      // gave up (the caller raises the watchdog flag)
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


// Filter word at byte offset `off`.  FIXED = true: the filter is the first object of the kernel's dynamic shared
// memory and the kernel has no static shared memory, so it starts at CTA-shared address MK_DYN_SMEM_BASE (the
// 1 KB below it is reserved by the system on sm_100) and the load is ONE `LDS R, [R + 0x400]`; through the generic
// pointer the compiler adds the (shared::cluster) window base in a separate instruction per lookup.  The
// kernel checks the assumption once (FLAG_SMEM_BASE).
#define MK_DYN_SMEM_BASE 1024
template <bool FIXED>
__device__ __forceinline__ u32 filter_word(const u32 *bm, u32 off)
{
    if constexpr (FIXED) {
        u32 w;
        asm("ld.shared.u32 %0, [%1+1024];" : "=r"(w) : "r"(off));
        return w;
    } else {
        return *reinterpret_cast<const u32 *>(reinterpret_cast<const char *>(bm) + off);
    }
}

template <int ROTOFF, u32 WORDMASK, int J, bool FIXED>
__device__ __forceinline__ void probe_one(const u32 (&A)[4], const u32 *bm, u32 &hits)
{
    constexpr int NEEDV = ROTOFF;     // bits 0..1 of v are masked off, bits 2..ROTOFF-1 index the word
    u32 v = take_bits<2 * J, NEEDV>(A);
    u32 r = take_bits<2 * J + ROTOFF, 5>(A);
    u32 word = filter_word<FIXED>(bm, v & WORDMASK);
    u32 rot = __funnelshift_l(word, word, r);
    hits = __funnelshift_l(rot, hits, 1);
    if constexpr (J < 31) probe_one<ROTOFF, WORDMASK, J + 1, FIXED>(A, bm, hits);
}

// Packed bases of one block: A[0..2], position j's inner window starting at bit 2j (+2*spare).
template <int PREW>
__device__ __forceinline__ void load_block(const uint8_t *blk, int shift_s, u32 (&A)[4])
{
    u32 W[PREW + 3];
    const uint4 *q = reinterpret_cast<const uint4 *>(blk - 16 * PREW);
#ifdef MK_ROTATED_BLOCK_LOADS
    // lanes alternate the order of their 16-byte pieces so that neighbouring blocks never hit the same
    // banks in the same wavefront (measured: no gain over the plain order, kept for reference)
    const bool rot = (threadIdx.x >> 2) & 1;
    if (PREW == 1) {
        u32 w0 = pack16(q[rot ? 1 : 0]), w1 = pack16(q[rot ? 2 : 1]), w2 = pack16(q[rot ? 0 : 2]);
        W[0] = rot ? w2 : w0; W[1] = rot ? w0 : w1; W[2] = rot ? w1 : w2;
    } else {
        u32 w0 = pack16(q[rot ? 2 : 0]), w1 = pack16(q[rot ? 3 : 1]), w2 = pack16(q[rot ? 0 : 2]), w3 = pack16(q[rot ? 1 : 3]);
        W[0] = rot ? w2 : w0; W[1] = rot ? w3 : w1; W[2] = rot ? w0 : w2; W[PREW + 1] = rot ? w1 : w3;
    }
#else
#pragma unroll
    for (int i = 0; i < PREW + 2; i++) W[i] = pack16(q[i]);
#endif
    W[PREW + 2] = 0;
    A[0] = __funnelshift_r(W[0], W[1], shift_s);
    A[1] = __funnelshift_r(W[1], W[2], shift_s);
    A[2] = __funnelshift_r(W[2], W[3], shift_s);
    A[3] = 0;
}

// Probe the 32 k-mer end positions of one aligned 32-byte block against the bitmap.
// blk = shared-memory address of the block's first byte.  Bit j of the result = position j hit.
// A[0..2] receive the packed bases, position j's inner window starting at bit 2j.
template <int ROTOFF, u32 WORDMASK, int PREW, bool FIXED = false>
__device__ __forceinline__ u32 probe_block(const uint8_t *blk, const u32 *bm, int shift_s, u32 (&A)[4])
{
    load_block<PREW>(blk, shift_s, A);
    u32 hits = 0;
    probe_one<ROTOFF, WORDMASK, 0, FIXED>(A, bm, hits);
    return __brev(hits);
}

// Second hash of the two-hash Bloom filter (inner windows of 22 bits or more): word from window bits
// 6..5+MK_BLOOM_WBITS, bit from window bits 19..23 xor-ed onto the lowest five bits (the first hash uses
// bits 2..6+MK_BLOOM_WBITS).  Must mirror set_bit() in
// mk_api.cu.
__device__ __forceinline__ bool second_hash_hit(const u32 (&A)[4], u32 j, const u32 *bm)
{
    u32 lo = j < 16 ? A[0] : A[1];
    u32 hi = j < 16 ? A[1] : A[2];
    u32 v = __funnelshift_r(lo, hi, (2 * j) & 31);
    u32 word = *reinterpret_cast<const u32 *>(reinterpret_cast<const char *>(bm) + ((v >> 4) & (((1u << MK_BLOOM_WBITS) - 1u) << 2)));
    // bit index: window bits 19..23 folded onto bits 0..4 (bits 0 and 1 are the ones the first hash does
    // not see); stored in natural order (bit b <-> 1 << b), unlike the first hash's rotate trick
    return (word >> (((v >> 19) ^ v) & 31u)) & 1u;        // (v carries more than the 24 window bits: only bits < 24 may be used)
}


#define MAXBLK (MK_MAX_TILE / 32)
#define MAXCHUNK (MK_MAX_TILE / 2048)   // scan chunks: one warp pass = 32 lanes x 64 bytes
#define NWARPS (MK_STREAM_THREADS / 32)
#define WHITS 32                         // verified-later hits buffered per warp

// inclusive prefix XOR over the 32 bits of x
__device__ __forceinline__ u32 prefix_xor(u32 x)
{
    x ^= x << 1; x ^= x << 2; x ^= x << 4; x ^= x << 8; x ^= x << 16;
    return x;
}

// ---- the kernel --------------------------------------------------------------------------------
// Per CTA three tiles (cur, next, next-next) are resident in shared memory.  One iteration has a
// single block-wide barrier; after it every warp pulls work units until none are left:
//
//   P  probe units of `cur`  : 32 items (aligned 32-byte blocks holding sequence bytes) per pull.
//                              Hits that pass both Bloom hashes are queued; the LAST warp to leave
//                              the probe loop verifies the queue exactly (22 valid ACGT bytes,
//                              canonical strand, exact pass table) and then recycles the stage of
//                              `cur` (new ticket, TMA issued).
//   M  mask units of `next`  : per 32-byte block, newline mask + line number -> mask of sequence
//                              bytes (branch-free prefix-parity arithmetic); blocks with any
//                              sequence byte are appended to next's item list.  Needs the line
//                              number of `next`, which warp 0 resolves first (decoupled look-back;
//                              every predecessor published its count at least one iteration ago).
//   S  scan units of `next-next`: 2 KB chunks scanned for '\n' (mask per block, count per chunk);
//                              the warp completing the last chunk publishes the tile's count.
struct StreamSmem {
    u32 nlm[3][MAXBLK];          // newline mask of each 32-byte block
    uint16_t exw[3][MAXBLK];     // '\n' bytes before the block inside its chunk
    u32 ctot[3][MAXCHUNK];       // '\n' bytes per chunk
    u32 cpre[3][MAXCHUNK];       // '\n' bytes before each chunk inside the tile
    u32 posmask[2][MAXBLK];      // double buffered: tile being probed / tile being masked
    uint16_t items[2][MAXBLK];
    u64 bar[3];
    u64 P[3];                    // '\n' bytes before each staged tile (line number of its first byte)
    u32 tile[3];                 // ticket held by each stage
    u32 tot[3];                  // '\n' bytes in each staged tile
    u32 scur[3];                 // scan-chunk cursor per stage
    u32 sdone[3];                // scanned chunks per stage
    u32 n_items[2];              // items per buffer
    u32 icur[2], pdone[2], mcur;    // item cursor / completed probe units per item buffer
    volatile u32 resolved;       // iteration stamp: line number of `next` is available
};

template <int ROTOFF, u32 WORDMASK, int PREW, bool RAW>
__global__ void __launch_bounds__(MK_STREAM_THREADS, 1) k_stream(const __grid_constant__ StreamArgs A)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const u32 tid = threadIdx.x;
    const u32 lane = tid & 31, wid = tid >> 5;
    const u32 bm_bytes = (A.bitmap_bytes + 127u) & ~127u;
    u32 *bm = reinterpret_cast<u32 *>(smem);
    uint8_t *tbuf = smem + bm_bytes;
    StreamSmem &S = *reinterpret_cast<StreamSmem *>(tbuf + 3 * TBUF_STRIDE);

    {   // bitmap -> shared memory
        const uint4 *src = reinterpret_cast<const uint4 *>(A.bitmap);
        uint4 *dst = reinterpret_cast<uint4 *>(bm);
        for (u32 i = tid; i < A.bitmap_bytes / 16; i += MK_STREAM_THREADS) dst[i] = src[i];
    }
    if (tid == 0) {
        for (int s = 0; s < 3; s++) { mbar_init(&S.bar[s], 1); S.scur[s] = 0; S.sdone[s] = 0; S.tot[s] = 0; S.P[s] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        S.n_items[0] = 0; S.n_items[1] = 0; S.icur[0] = S.icur[1] = 0; S.pdone[0] = S.pdone[1] = 0; S.mcur = 0; S.resolved = 0;
    }
    __syncthreads();

    const u32 TB = A.tile_bytes;
    const u32 NBLK = TB / 32;
    const u32 NCHUNK = (TB + 2047u) / 2048u;
    const u32 NMUNIT = (NBLK + 63u) / 64u;
    auto tile_len = [&](u32 t) -> u32 {
        u64 rem = A.nbytes - (u64)t * TB;
        return rem < TB ? (u32)rem : TB;
    };
    auto claim_and_load = [&](int stage) {   // one thread
        u32 t = atomicAdd(A.tile_counter, 1u);
        S.tile[stage] = t;
        S.scur[stage] = 0;
        S.sdone[stage] = 0;
        if (t >= A.n_tiles) return;
        u32 tb = tile_len(t);
        uint8_t *dst = tbuf + stage * TBUF_STRIDE;
        const uint8_t *src = A.text + (u64)t * TB;
        u32 bytes = tb;
        if (t > 0) { src -= MK_HALO; bytes += MK_HALO; } else { dst += MK_HALO; }
        bytes = (bytes + 15u) & ~15u;
        fence_proxy_async();
        mbar_expect_tx(&S.bar[stage], bytes);
        tma_load_1d(dst, src, bytes, &S.bar[stage]);
    };
    auto watchdog = [&](u32 site, u64 a, u64 b, u64 c, u64 d) {
        if (atomicOr(A.flags, FLAG_WATCHDOG) & FLAG_WATCHDOG) return;   // first report wins
        A.wd[0] = site; A.wd[1] = blockIdx.x; A.wd[2] = wid; A.wd[3] = a; A.wd[4] = b; A.wd[5] = c; A.wd[6] = d;
        A.wd[7] = ((u64)S.tile[0] << 42) | ((u64)S.tile[1] << 21) | (u64)S.tile[2];
    };
    // S units: scan the tile in `stage` (chunks pulled dynamically by whole warps).
    auto scan_units = [&](int stage, u32 use_parity) {
        const u32 t = S.tile[stage];
        if (t >= A.n_tiles) return;                        // uniform per CTA
        uint8_t *tx = tbuf + stage * TBUF_STRIDE;
        const u32 tb = tile_len(t);
        bool waited = false;
        for (;;) {
            u32 c = 0;
            if (lane == 0) c = atomicAdd(&S.scur[stage], 1u);
            c = __shfl_sync(0xffffffffu, c, 0);
            if (c >= NCHUNK) break;
            if (!waited) {
                if (!mbar_wait(&S.bar[stage], use_parity) && lane == 0) watchdog(1, stage, use_parity, t, c);
                waited = true;
            }
            const u32 off = c * 2048u + lane * 64u;        // my 64 bytes (two blocks)
            // blank what lies outside the text so that stale bytes can never look like sequence
            if (t == 0 && c == 0 && lane < MK_HALO / 4) reinterpret_cast<u32 *>(tx)[lane] = 0;
            if (off + 64u > tb && off < TB) {
                u32 from = off > tb ? off : tb;
                for (u32 i = from; i < off + 64u; i++) tx[MK_HALO + i] = 0;
            }
            u32 m0 = 0, m1 = 0;
            if (!RAW && off < TB) {
                // lane l owns bytes [64 l, 64 l + 64) of the chunk: with a 64-byte lane stride the four
                // 128-bit loads would be 4-way bank conflicted; lane l therefore starts at piece (l/2) % 4
                // and the 16-bit piece masks are placed by piece index afterwards.
                const uint4 *q = reinterpret_cast<const uint4 *>(tx + MK_HALO + off);
                const u32 rot = (lane >> 1) & 3u;
#pragma unroll
                for (int k4 = 0; k4 < 4; k4++) {
                    const u32 pc = (k4 + rot) & 3u;
                    uint4 v = q[pc];
                    u32 w[4] = {v.x, v.y, v.z, v.w};
                    u32 m = 0;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        // exact "byte == 0x0A": bit 7 of t7 is set iff the low 7 bits of (byte ^ 0x0A) are non-zero
                        u32 t7 = ((w[j] ^ 0x0A0A0A0Au) & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
                        u32 f = ~(t7 | w[j]) & 0x80808080u;
                        u32 nib = __umulhi(f, 0x02040810u);   // bits 7,15,23,31 -> bits 0..3 (junk above)
                        m = __funnelshift_r(m, nib, 4);
                    }
                    m >>= 16;                                  // 16-bit newline mask of this piece
                    m <<= (pc & 1u) * 16u;
                    if (pc & 2u) m1 |= m; else m0 |= m;
                }
            }
            if (!RAW) {
                const u32 c0 = __popc(m0), cnt = c0 + __popc(m1);
                u32 incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    u32 v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= (u32)o) incl += v;
                }
                const u32 b0 = c * 64u + 2u * lane;
                if (b0 < NBLK) {
                    S.nlm[stage][b0] = m0;
                    S.exw[stage][b0] = (uint16_t)(incl - cnt);
                }
                if (b0 + 1 < NBLK) {
                    S.nlm[stage][b0 + 1] = m1;
                    S.exw[stage][b0 + 1] = (uint16_t)(incl - cnt + c0);
                }
                if (lane == 31) S.ctot[stage][c] = incl;
            }
            fence_proxy_async();
            __threadfence_block();
            __syncwarp();
            u32 fin = 0;
            if (lane == 0) fin = atomicAdd(&S.sdone[stage], 1u);
            fin = __shfl_sync(0xffffffffu, fin, 0);
            if (fin == NCHUNK - 1 && !RAW) {               // this warp completed the tile's scan
                __threadfence_block();
                u32 v = lane < NCHUNK ? S.ctot[stage][lane] : 0;
                u32 incl = v;
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) {
                    u32 x = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= (u32)o) incl += x;
                }
                if (lane < NCHUNK) S.cpre[stage][lane] = incl - v;
                u32 total = __shfl_sync(0xffffffffu, incl, 15);   // NCHUNK <= 12 lanes carry values
                if (lane == 0) {
                    S.tot[stage] = total;
                    st_volatile_u64(&A.tile_desc[t], (1ull << 62) | (u64)total);
                    if (total == 0 && tb == TB && TB >= 4096) atomicOr(A.flags, FLAG_LONG_LINE);
                }
            }
        }
    };
    // Line number of tile t (warp 0, all lanes).  This CTA resolved tile prev_t < t earlier, so
    //   newlines before t = newlines through prev_t + sum of the counts of tiles prev_t+1 .. t-1,
    // and every one of those counts was published when its tile was scanned, i.e. at least one
    // iteration before this call.  All loads are independent: one L2 round trip, no chaining on
    // other tiles' prefixes.
    long long prev_t = -1;
    u64 prev_incl = 0;
    auto resolve_tile = [&](u32 t, u32 total) -> u64 {
        const u64 VMASK = (1ull << 62) - 1;
        u64 sum = 0;
        for (long long i0 = prev_t + 1; i0 < (long long)t; i0 += 128) {
            u64 d[4];
#pragma unroll
            for (int k4 = 0; k4 < 4; k4++) {
                long long idx = i0 + 32 * k4 + (long long)lane;
                d[k4] = idx < (long long)t ? ld_volatile_u64(&A.tile_desc[idx]) : (1ull << 62);
            }
#pragma unroll
            for (int k4 = 0; k4 < 4; k4++) {
                long long idx = i0 + 32 * k4 + (long long)lane;
                for (u32 n = 0; (d[k4] >> 62) == 0; n++) {
                    if (n > WD_LIMIT) { watchdog(2, t, (u64)idx, (u64)prev_t, 0); break; }
                    __nanosleep(100);
                    d[k4] = ld_volatile_u64(&A.tile_desc[idx]);
                }
                sum += d[k4] & VMASK;
            }
        }
        u64 excl = prev_incl + warp_sum_u64(sum);
        prev_t = (long long)t;
        prev_incl = excl + total;
        return excl;
    };
    // M units: sequence-byte masks + item list of the tile in `stage`, into buffer `buf`.
    // A unit is 64 consecutive blocks (two per lane), pulled dynamically.
    auto mask_units = [&](int stage, int buf, bool all_warps) {
        const u32 t = S.tile[stage];
        if (t >= A.n_tiles) return;
        const u32 tb = tile_len(t);
        const u32 P = RAW ? 0u : (u32)S.P[stage];
        (void)all_warps;
        for (;;) {
            u32 u = 0;
            if (lane == 0) u = atomicAdd(&S.mcur, 1u);
            u = __shfl_sync(0xffffffffu, u, 0);
            if (u >= NMUNIT) break;
            u32 pm[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const u32 b = u * 64u + 32u * h + lane;
                pm[h] = 0;
                if (b < NBLK) {
                    if (RAW) {
                        u32 lo = 32 * b;
                        pm[h] = lo >= tb ? 0u : (tb - lo >= 32 ? 0xffffffffu : ((1u << (tb - lo)) - 1u));
                    } else {
                        // line(i) = P + '\n' before the block + '\n' among bytes < i of the block; a byte is
                        // a sequence byte iff line(i) % 4 == 1 and it is not the '\n' itself.  The two low
                        // bits of the running count are prefix parities of the newline mask.
                        const u32 nl = S.nlm[stage][b];
                        const u32 s0 = (P + S.cpre[stage][b >> 6] + S.exw[stage][b]) & 3u;
                        const u32 tgt = (1u - s0) & 3u;          // count % 4 that puts a byte on phase 1
                        u32 p0 = 0, p1 = 0;
                        if (nl) {
                            p0 = prefix_xor(nl << 1);            // parity of '\n' strictly before each byte
                            p1 = prefix_xor((nl & p0) << 1);     // carries out of bit 0
                        }
                        pm[h] = (p0 ^ ((tgt & 1u) ? 0u : ~0u)) & (p1 ^ ((tgt & 2u) ? 0u : ~0u)) & ~nl;
                    }
                    S.posmask[buf][b] = pm[h];
                }
            }
            const u32 m0 = __ballot_sync(0xffffffffu, pm[0] != 0);
            const u32 m1 = __ballot_sync(0xffffffffu, pm[1] != 0);
            const u32 n0 = __popc(m0);
            u32 base = 0;
            if (lane == 0 && (m0 | m1)) base = atomicAdd(&S.n_items[buf], n0 + (u32)__popc(m1));
            base = __shfl_sync(0xffffffffu, base, 0);
            const u32 lt = (1u << lane) - 1u;
            if (pm[0]) S.items[buf][base + __popc(m0 & lt)] = (uint16_t)(u * 64u + lane);
            if (pm[1]) S.items[buf][base + n0 + __popc(m1 & lt)] = (uint16_t)(u * 64u + 32u + lane);
        }
    };

    // ---- prologue: three tickets; scan the first two; resolve + mask the first ----------------
    if (tid == 0)
        for (int s = 0; s < 3; s++) claim_and_load(s);
    __syncthreads();
    u32 uses0 = 0, uses1 = 0, uses2 = 0;       // completed uses of each stage's mbarrier (parity)
    scan_units(0, 0); uses0 = 1;
    scan_units(1, 0); uses1 = 1;
    __syncthreads();
    if (!RAW && wid == 0 && S.tile[0] < A.n_tiles) {
        u64 e = resolve_tile(S.tile[0], S.tot[0]);
        if (lane == 0) {
            S.P[0] = A.line_base + e;
            if (S.tile[0] == A.n_tiles - 1) *A.total_newlines = A.line_base + e + S.tot[0];
        }
    }
    __syncthreads();
    mask_units(0, 0, true);

    int stage = 0, buf = 0;
    u32 iter = 0;
    for (;;) {
        __syncthreads();                                   // the only block-wide barrier per tile
        const u32 cur = S.tile[stage];
        if (cur >= A.n_tiles) break;
        iter++;
        const int st1 = stage == 2 ? 0 : stage + 1;       // next tile
        const int st2 = st1 == 2 ? 0 : st1 + 1;           // tile after next
        uint8_t *tx = tbuf + stage * TBUF_STRIDE;
        const u64 T = (u64)cur * TB;
        const u32 n_items = S.n_items[buf];
        if (tid == 0) {                                    // next iteration's cursors; nobody touches them now
            S.mcur = 0;                                    // (M units are not pulled before `resolved` is stamped)
            S.icur[buf ^ 1] = 0;
            S.pdone[buf ^ 1] = 0;
        }
        const bool tr = A.trace && blockIdx.x == 0 && iter <= 64 && lane == 0;
        u64 *trp = A.trace + ((u64)(iter - 1) * NWARPS + wid) * 8;
        if (tr) { trp[0] = clock64(); trp[6] = n_items; trp[7] = cur; }

        // warp 0: line number of the next tile, then stamp it
        if (wid == 0) {
            const u32 nt = S.tile[st1];
            if (!RAW && nt < A.n_tiles) {
                u64 e = resolve_tile(nt, S.tot[st1]);
                if (lane == 0) {
                    S.P[st1] = A.line_base + e;
                    if (nt == A.n_tiles - 1) *A.total_newlines = A.line_base + e + S.tot[st1];
                }
            }
            __syncwarp();
            if (lane == 0) { __threadfence_block(); S.resolved = iter; }
        }
        if (tr) trp[1] = clock64();

        // ---- P units: pulled by every warp; warps 0 and 1 (warp 0 is busy with the look-back) take
        //      scan and mask units first and only then whatever probe units are left -----------------
        const u32 nP = (n_items + 31u) >> 5;
        auto probe_units = [&]() {
            for (;;) {
                u32 base = 0;
                if (lane == 0) base = atomicAdd(&S.icur[buf], 32u);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= n_items) break;
                const u32 it = base + lane;
                if (it < n_items) {
                    const u32 b = S.items[buf][it];
                    u32 Aw[4];
                    u32 hits = probe_block<ROTOFF, WORDMASK, PREW>(tx + MK_HALO + 32 * b, bm, A.kp.shift_s, Aw);
                    hits &= S.posmask[buf][b];
                    while (hits) {
                        u32 j = __ffs(hits) - 1;
                        hits &= hits - 1;
                        if (A.two_hash && !second_hash_hit(Aw, j, bm)) continue;
                        emit_hit(A, T + 32 * b + j);
                    }
                }
                __syncwarp();
                if (lane == 0) {                           // the warp completing the last unit recycles the stage
                    __threadfence_block();
                    if (atomicAdd(&S.pdone[buf], 1u) == nP - 1) {
                        S.n_items[buf] = 0;
                        claim_and_load(stage);
                    }
                }
            }
        };
        if (nP == 0 && tid == 0) { S.n_items[buf] = 0; claim_and_load(stage); }
        if (wid >= 2) probe_units();
        if (tr) trp[2] = clock64();
        // ---- S units of the tile two ahead -----------------------------------------------------
        {
            u32 par = st2 == 0 ? (uses0 & 1u) : (st2 == 1 ? (uses1 & 1u) : (uses2 & 1u));
            scan_units(st2, par);
            if (st2 == 0) uses0++; else if (st2 == 1) uses1++; else uses2++;
        }
        if (tr) trp[3] = clock64();
        // ---- M units of the next tile (after its line number is known) -------------------------
        for (u32 n = 0; S.resolved != iter; n++) {
            if (n > WD_LIMIT) { if (lane == 0) watchdog(3, iter, S.resolved, cur, S.tile[st1]); break; }
            __nanosleep(40);
        }
        if (tr) trp[4] = clock64();
        __threadfence_block();
        mask_units(st1, buf ^ 1, false);
        if (tr) trp[5] = clock64();
        if (wid < 2) probe_units();
        stage = st1;
        buf ^= 1;
    }
}

// ================================================================================================
// k_stream_ws — warp-specialised, barrier-free variant of the same pipeline (the default).
//
// The unit-pulling kernel above spends a large share of its issue slots on shared-memory atomics,
// fetch loops and one block barrier per tile.  Here every warp has a fixed role and a plain loop;
// tiles flow through a ring of NS shared-memory stages and the roles hand a stage over with
// mbarriers (arrive / try_wait.parity), never with __syncthreads.  A multi-warp role waits through
// its leader warp (the others park on a named barrier).  Tickets are groups of WS_G consecutive tiles.
//
//   loader      (1 lane)    waits until the probe group released the stage, takes the next tile of
//                           its group, starts the TMA                              -> full[s]
//   resolver    (1 warp)    claims the CTA's next group ticket and sums the newline counts published
//                           for the tickets since its previous group             -> mailbox gready[]
//   count-ahead (WS_NCW)    a team that counts the newlines of whole groups straight from global
//                           memory, WS_WINDOW x gridDim groups ahead of the load front, and publishes
//                           one packed descriptor per group (read by every CTA's resolver)
//   front       (2 teams)   every warp of the team scans one 2 KB chunk (newline mask + in-tile prefix of
//                           newline counts per 32-byte block), the team meets on a named barrier, then
//                           every warp masks one 64-block unit (sequence-byte mask per block, item list
//                           of the tile)                                             -> ready[s]
//                           (-DWS_SPLIT_FRONT: the earlier separate scan and mask teams with a `scanned`
//                           mbarrier between them; 1 % slower)
//   probe       (2 groups)  probe a static, per-tile rotated share of the items   -> done[s]
//
// Teams / groups take alternating tiles (k % 2).  The k-th tile of a CTA always uses stage k % NS, so
// every role derives stage and mbarrier parity from its own loop counter.
// ================================================================================================
#ifndef WS_NS
#define WS_NS 6        // ring stages
#endif
#ifndef WS_FIXED_FILTER_BASE
#define WS_FIXED_FILTER_BASE true   // filter lookups as LDS [R + 0x400] (see filter_word())
#endif
#ifndef WS_TILE
#define WS_TILE 12288  // tile-proper bytes (multiple of 2048)
#endif
#define WS_BLK (WS_TILE / 32)
#define WS_CHUNK (WS_TILE / 2048)
#ifndef WS_NSW
#define WS_NSW 6       // scan warps
#endif
#ifndef WS_NMW
#define WS_NMW 6       // mask warps
#endif
#define WS_NFW (WS_NSW + WS_NMW)
#ifndef WS_NPG
#define WS_NPG 2       // probe groups: group g probes the tiles with k % WS_NPG == g
#endif
#ifndef WS_NPW
#define WS_NPW 7       // probe warps per group
#endif
#ifndef WS_NCW
#define WS_NCW 4        // count-ahead warps
#endif
#ifndef WS_WINDOW
#define WS_WINDOW 2     // the count pass may run this many x gridDim tiles ahead of the load front
#endif
#ifndef WS_G
#define WS_G 4          // tiles per ticket (consecutive tiles of one CTA)
#endif
#define WS_ROLE0 (2 + WS_NCW)   // first front-end warp (0 = loader, 1 = resolver, then the count warps)
#ifndef WS_SPLIT_FRONT
#define WS_FUSED_FRONT 1
#endif
#ifdef WS_FUSED_FRONT
// one front team per probe group does scan AND mask of its tile (a named barrier in between): every warp scans one
// 2 KB chunk and masks one 64-block unit, so the two stages take half the time each and one mbarrier hand-over
// is gone
#define WS_SCT (WS_NFW / WS_NPG)
#define WS_MKT (WS_NFW / WS_NPG)
#else
#define WS_SCT (WS_NSW / WS_NPG)   // warps per scan team
#define WS_MKT (WS_NMW / WS_NPG)   // warps per mask team
#endif
#define WS_THREADS (32 * (WS_ROLE0 + WS_NFW + WS_NPG * WS_NPW))
static_assert(WS_NSW % WS_NPG == 0 && WS_NMW % WS_NPG == 0, "teams");
static_assert(WS_G == 4 && WS_TILE < 16384, "descriptor packing");
#ifndef WS_PF_DIST
#define WS_PF_DIST 3    // L2 prefetch distance in units of gridDim tiles
#endif
#define WS_TBUF (MK_HALO + WS_TILE + 96)

struct WsStage {
    u32 nlm[WS_BLK];             // newline mask of each block; the mask warps overwrite it in place with
                                 // the block's sequence-byte mask (same index, same thread)
    uint16_t exw[WS_BLK];
    uint16_t items[WS_BLK];
    u32 ctot[WS_CHUNK];
    u32 cpre[WS_CHUNK];
};
template <int NS>
struct WsSmem {
    WsStage st[NS];
    u64 full[NS], scanned[NS], ready[NS], done[NS];
    u64 gready[2], gfree[2];     // group mailbox resolver -> loader
    u64 gq_P[2][4];              // newlines before each tile of the group
    u32 gq_g[2];
    u32 cticket, csum[2];        // count team: current group ticket, its packed per-tile newline counts so far
    u32 abort;                   // a leader's wait gave up: the whole role leaves
    u64 Pn[2 * NS];           // newlines before the k-th tile of this CTA (index k % (2 NS))
    u32 tile[NS], n_items[NS], scnt[NS];
};

#ifndef WS_CNT_BATCH
#define WS_CNT_BATCH 12 // 16-byte loads in flight per lane of a count warp
#endif
__device__ __forceinline__ uint4 ld_nc_u128(const void *p)
{
    uint4 v;
#ifndef WS_COUNT_NO_HINT
    // evict_last: the line must still be in L2 when the tile's TMA load arrives a few tile periods later
    // (without the hint the streaming load is the first to go and DRAM traffic doubles)
    asm volatile("{\n\t.reg .b64 pol;\n\tcreatepolicy.fractional.L2::evict_last.b64 pol, 1.0;\n\t"
                 "ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], pol;\n\t}"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
#else
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
#endif
    return v;
}
__device__ __forceinline__ void ld_volatile_v2(const u64 *p, u32 &lo, u32 &hi)
{
    asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(lo), "=r"(hi) : "l"(p) : "memory");
}

// cold path, kept out of line so that the roles' loops stay compact in the instruction cache
__device__ __noinline__ void ws_watchdog_report(u32 *flags, u64 *wd, u32 site, u32 wid, u64 a, u64 b, u64 c, u64 d, u64 tiles)
{
    if (atomicOr(flags, FLAG_WATCHDOG) & FLAG_WATCHDOG) return;   // first report wins
    wd[0] = site; wd[1] = blockIdx.x; wd[2] = wid; wd[3] = a; wd[4] = b; wd[5] = c; wd[6] = d; wd[7] = tiles;
}

template <int ROTOFF, u32 WORDMASK, int PREW, bool RAW, int NS>
__global__ void __launch_bounds__(WS_THREADS, 1) k_stream_ws(const __grid_constant__ StreamArgs A)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const u32 tid = threadIdx.x;
    const u32 lane = tid & 31, wid = tid >> 5;
    // byte-compare constants kept in registers, so that (w ^ c) & m is ONE three-input LOP3 (with immediates
    // the compiler needs two instructions per word)
    // (derived from a launch parameter in a way no compiler pass can fold: tile_bytes < 2^31)
    const u32 c0a = 0x0A0A0A0Au + (A.tile_bytes >> 31), c7f = 0x7F7F7F7Fu + (A.tile_bytes >> 31);
    const u32 bm_bytes = (A.bitmap_bytes + 127u) & ~127u;
    u32 *bm = reinterpret_cast<u32 *>(smem);
    uint8_t *tbuf = smem + bm_bytes;
    static_assert(NS % WS_NPG == 0, "teams");
    WsSmem<NS> &S = *reinterpret_cast<WsSmem<NS> *>(tbuf + NS * WS_TBUF);

    {
        const uint4 *src = reinterpret_cast<const uint4 *>(A.bitmap);
        uint4 *dst = reinterpret_cast<uint4 *>(bm);
        for (u32 i = tid; i < A.bitmap_bytes / 16; i += WS_THREADS) dst[i] = src[i];
    }
    if (tid == 0) {
        for (int s = 0; s < NS; s++) {
            mbar_init(&S.full[s], 1);
            mbar_init(&S.scanned[s], WS_SCT);
            mbar_init(&S.ready[s], WS_MKT);
            mbar_init(&S.done[s], WS_NPW);
            S.n_items[s] = 0; S.tile[s] = 0xFFFFFFFFu; S.scnt[s] = 0;
        }
        for (int q = 0; q < 2; q++) { mbar_init(&S.gready[q], 1); mbar_init(&S.gfree[q], 1); }
        S.cticket = 0; S.csum[0] = 0; S.csum[1] = 0; S.abort = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (WS_FIXED_FILTER_BASE && tid == 0 && ((u32)__cvta_generic_to_shared(bm) & 0xFFFFFFu) != MK_DYN_SMEM_BASE)
        atomicOr(A.flags, FLAG_SMEM_BASE);      // (filter_word<true> would read the wrong words: the host refuses the result)
    __syncthreads();   // the only block-wide barrier of the kernel

    const u32 TB = A.tile_bytes;
    const u32 NBLK = TB / 32;
    const u32 NCHUNK = (TB + 2047u) / 2048u;
    const u32 NMUNIT = (NBLK + 63u) / 64u;
    const u32 gb = A.tile_begin / WS_G;                                  // first group of this launch
    const u32 n_groups = (A.n_tiles - A.tile_begin + WS_G - 1) / WS_G;   // groups of this launch (tickets are relative)
    auto tile_len = [&](u32 t) -> u32 {
        u64 rem = A.nbytes - (u64)t * TB;
        return rem < TB ? (u32)rem : TB;
    };
    auto watchdog = [&](u32 site, u64 a, u64 b, u64 c, u64 d) {
        ws_watchdog_report(A.flags, A.wd, site, wid, a, b, c, d,
                           ((u64)S.tile[0] << 42) | ((u64)S.tile[1] << 21) | (u64)S.tile[2]);
    };
    auto stamp = [&](u32 k, u32 slot) {      // development aid (build with -DMK_TRACE): clock64() stamps of CTA 0
#ifdef MK_TRACE
        if (A.trace && blockIdx.x == 0 && k < 48 && lane == 0) A.trace[((u64)k * 32 + wid) * 4 + slot] = clock64();
#else
        (void)k; (void)slot;
#endif
    };
    // wait with a watchdog; false = gave up (the role then leaves the kernel)
    auto wait_on = [&](u64 *bar, u32 parity, u32 site, u32 k) -> bool {
        if (mbar_wait(bar, parity)) return true;
        if (lane == 0) watchdog(site, k, parity, wid, 0);
        return false;
    };

    // A multi-warp role waits through its leader warp: only the leader polls the mbarrier, the other
    // warps block on a named barrier, which costs no issue slots while they wait.
    auto role_wait = [&](u64 *bar, u64 *bar2, u32 parity, u32 site, u32 k, u32 barid, u32 nwarps, bool leader) -> bool {
        if (leader) {
            bool ok = mbar_wait(bar, parity);
            if (ok && bar2) ok = mbar_wait(bar2, parity);
            if (!ok && lane == 0) { watchdog(site, k, parity, wid, 0); S.abort = 1; }
        }
        named_bar_sync(barid, 32u * nwarps);
        return *(volatile u32 *)&S.abort == 0;
    };

    if (wid == 0) {
        // ======================= loader =========================================================
        // One lane.  Tickets are groups of WS_G consecutive tiles; the resolver warp hands over the next
        // group and the line number of its first tile through a two-slot mailbox, so a freed stage is
        // refilled without any global round trip.
        if (lane == 0) {
            u32 ended = 0, j = 0, g = 0, i = 0, cnt = 0;
            u64 P[4] = {0, 0, 0, 0};
            const u64 LB = A.line_base + (A.line_base_ptr ? *A.line_base_ptr : 0ull);
            bool finished = false;
            for (u32 k = 0;; k++) {
                const u32 s = k % NS, v = k / NS;
                if (v > 0 && !wait_on(&S.done[s], (v - 1) & 1u, 10, k)) break;
                stamp(k, 0);
                if (!finished && i == cnt) {            // next group of this CTA
                    const u32 slot = j & 1u;
                    if (!wait_on(&S.gready[slot], (j >> 1) & 1u, 11, k)) break;
                    g = S.gq_g[slot];
#pragma unroll
                    for (int q = 0; q < 4; q++) P[q] = S.gq_P[slot][q];
                    mbar_arrive(&S.gfree[slot]);
                    j++;
                    if (g >= n_groups) finished = true;
                    else {
                        i = 0;
                        g += gb;
                        const u32 left = A.n_tiles - g * WS_G;
                        cnt = left < (u32)WS_G ? left : (u32)WS_G;
                    }
                }
                const u32 t = finished ? 0xFFFFFFFFu : g * WS_G + i;
                S.n_items[s] = 0;
                S.tile[s] = t;
                if (t >= A.n_tiles) {       // end marker: one per probe group, in consecutive ring slots
                    mbar_arrive(&S.full[s]);
                    if (++ended == WS_NPG) break;
                    continue;
                }
                S.Pn[k % (2 * NS)] = LB + P[i & 3u];
                i++;
                stamp(k, 1);
                const u32 tb = tile_len(t);
                uint8_t *dst = tbuf + s * WS_TBUF;
                const uint8_t *src = A.text + (u64)t * TB;
                u32 bytes = tb;
                if (t > 0) { src -= MK_HALO; bytes += MK_HALO; } else { dst += MK_HALO; }
                bytes = (bytes + 15u) & ~15u;
                fence_proxy_async();
                mbar_expect_tx(&S.full[s], bytes);
                tma_load_1d_last_use(dst, src, bytes, &S.full[s]);
                if (RAW) {                  // (FASTQ mode: the count-ahead pass has pulled the tile into L2)
                    const u64 pf = (u64)t + (u64)WS_PF_DIST * gridDim.x * WS_G;
                    if (pf < A.n_tiles) {
                        const u64 o = pf * TB;
                        const u64 rem = A.nbytes - o;
                        prefetch_l2(A.text + o, (u32)((rem < TB ? rem : TB) + 15u) & ~15u);
                    }
                }
            }
        }
    } else if (wid == 1) {
        // ======================= resolver =======================================================
        // Claims this CTA's groups one to two ahead of the loader and computes the number of newlines
        // before each: newlines through the CTA's previous group + the group counts the count-ahead
        // pass published for the tickets in between (one look-back per WS_G tiles, off the ring).
        u64 run_incl = 0;
        long long prev_g = -1;
        for (u32 j = 0;; j++) {
            const u32 slot = j & 1u;
            if (j >= 2 && !wait_on(&S.gfree[slot], ((j >> 1) - 1u) & 1u, 20, j)) break;
            u32 g = 0;
            if (lane == 0) g = atomicAdd(A.tile_counter, 1u);
            g = __shfl_sync(0xffffffffu, g, 0);
            u64 P = 0;
            u32 c4[4] = {0, 0, 0, 0};
            if (!RAW && g < n_groups) {
                u32 sum = 0, own_lo = 0, own_hi = 0;
                bool good = true;
                for (long long i0 = prev_g + 1; i0 <= (long long)g; i0 += 256) {
                    u32 lo[8], hi[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const long long idx = i0 + 32 * q + (long long)lane;
                        lo[q] = 0; hi[q] = 1u << 30;
                        if (idx <= (long long)g) ld_volatile_v2(&A.tile_desc[gb + idx], lo[q], hi[q]);
                    }
                    // (descriptors are published well ahead: the retry path is cold and kept out of the unrolled code)
                    u32 ready = hi[0] & hi[1] & hi[2] & hi[3] & hi[4] & hi[5] & hi[6] & hi[7];
                    for (u32 n = 0; __any_sync(0xffffffffu, (ready >> 30) == 0); n++) {
                        if (n > WD_LIMIT) { if (lane == 0) watchdog(22, g, (u64)i0, (u64)prev_g, j); good = false; break; }
                        __nanosleep(200);
#pragma unroll 1
                        for (int q = 0; q < 8; q++) {
                            const long long idx = i0 + 32 * q + (long long)lane;
                            if (idx <= (long long)g && (hi[q] >> 30) == 0) ld_volatile_v2(&A.tile_desc[gb + idx], lo[q], hi[q]);
                        }
                        ready = hi[0] & hi[1] & hi[2] & hi[3] & hi[4] & hi[5] & hi[6] & hi[7];
                    }
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const long long idx = i0 + 32 * q + (long long)lane;
                        if (idx == (long long)g) { own_lo = lo[q]; own_hi = hi[q]; }
                        else sum += (lo[q] & 0xFFFFu) + (lo[q] >> 16) + (hi[q] & 0xFFFFu) + ((hi[q] >> 16) & 0x3FFFu);
                    }
                }
                sum = __reduce_add_sync(0xffffffffu, sum);
                own_lo = __reduce_add_sync(0xffffffffu, own_lo);
                own_hi = __reduce_add_sync(0xffffffffu, own_hi);
                if (!__all_sync(0xffffffffu, good)) break;
                P = run_incl + sum;
                c4[0] = own_lo & 0xFFFFu; c4[1] = own_lo >> 16; c4[2] = own_hi & 0xFFFFu; c4[3] = (own_hi >> 16) & 0x3FFFu;
                run_incl = P + c4[0] + c4[1] + c4[2] + c4[3];
                prev_g = (long long)g;
                if (lane == 0 && g == n_groups - 1)
                    *A.total_newlines = A.line_base + (A.line_base_ptr ? *A.line_base_ptr : 0ull) + run_incl;
            }
            if (lane == 0) {
                S.gq_g[slot] = g;
                u64 p = P;
#pragma unroll
                for (int i = 0; i < 4; i++) { S.gq_P[slot][i] = p; p += c4[i]; }
                __threadfence_block();
                mbar_arrive(&S.gready[slot]);
            }
            if (g >= n_groups) break;
        }
    } else if (wid < WS_ROLE0) {
        // ======================= count-ahead pass ===============================================
        // Record structure needs the number of '\n' before every tile.  Counting is cheap and has no
        // dependencies, so it runs ahead of the pipeline on its own tickets, straight from global
        // memory (which also pulls the tile into L2 for the TMA load that follows), and publishes one
        // descriptor per tile.
        if (!RAW) {
            const u32 window = (u32)WS_WINDOW * gridDim.x;
            // The WS_NCW count warps of the CTA work as a team on one group ticket (its tiles interleaved
            // among them), so a group's count is complete one tile-time after it was claimed.
            const u32 cw = wid - 2;
            for (;;) {
                if (cw == 0 && lane == 0) {
                    u32 c = atomicAdd(A.count_counter, 1u);
                    if (c >= window && c < n_groups) {   // flow control against the load front
                        for (u32 n = 0; c >= ld_volatile_u32(A.tile_counter) + window; n++) {
                            if (n > WD_LIMIT) { watchdog(50, c, 0, 0, 0); c = 0xFFFFFFFFu; break; }
                            __nanosleep(256);
                        }
                    }
                    S.cticket = c;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * WS_NCW) : "memory");
                const u32 c = S.cticket;
                if (c >= n_groups) break;
                for (u32 i = cw; i < (u32)WS_G; i += WS_NCW) {
                    const u32 t = (gb + c) * WS_G + i;
                    if (t >= A.n_tiles) break;
                    const u32 tb = tile_len(t);
                    const uint8_t *base = A.text + (u64)t * TB;
                    {   // the tile this warp is likely to count a group from now: into L2 already
                        const u64 pf = (u64)t + (u64)gridDim.x * WS_G;
                        if (lane == 0 && pf < A.n_tiles) {
                            const u64 o = pf * TB, rem = A.nbytes - o;
                            prefetch_l2(A.text + o, (u32)((rem < TB ? rem : TB) + 15u) & ~15u);
                        }
                    }
                    u32 acc_lo = 0, acc_hi = 0;         // per-byte-lane counters, folded to 16 bit per batch
                    if (tb == TB && TB % (512u * WS_CNT_BATCH) == 0) {
                        // full tile: WS_CNT_BATCH independent 16-byte loads per lane in flight, 4 instructions
                        // per word (the last one, f >> 7 accumulated, on the FMA pipe)
                        for (u32 v0 = 0; v0 < TB / 16u; v0 += 32u * WS_CNT_BATCH) {
                            uint4 q[WS_CNT_BATCH];
#pragma unroll
                            for (int jj = 0; jj < WS_CNT_BATCH; jj++) q[jj] = ld_nc_u128(base + 16u * (v0 + 32u * jj + lane));
                            u32 acc = 0;
#pragma unroll
                            for (int jj = 0; jj < WS_CNT_BATCH; jj++) {
                                u32 w[4] = {q[jj].x, q[jj].y, q[jj].z, q[jj].w};
#pragma unroll
                                for (int x = 0; x < 4; x++) {
                                    u32 t7 = ((w[x] ^ c0a) & c7f) + c7f;
                                    u32 f = ~(t7 | w[x]) & 0x80808080u;     // 0x80 where the byte is a newline
                                    acc = __umulhi(f, 1u << 25) + acc;      // += f >> 7 (<= 4 * WS_CNT_BATCH per byte lane)
                                }
                            }
                            acc_lo += acc & 0x00FF00FFu;
                            acc_hi += (acc >> 8) & 0x00FF00FFu;
                        }
                    } else {
                        const u32 nvec = (tb + 15u) >> 4;
#pragma unroll 1
                        for (u32 v = lane; v < nvec; v += 32u) {
                            uint4 q = ld_nc_u128(base + 16u * v);
                            u32 w[4] = {q.x, q.y, q.z, q.w};
                            u32 acc = 0;
#pragma unroll
                            for (int x = 0; x < 4; x++) {
                                u32 t7 = ((w[x] ^ 0x0A0A0A0Au) & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
                                u32 f = ~(t7 | w[x]) & 0x80808080u;
                                const u32 o = 16u * v + 4u * x;
                                if (o + 4u > tb) f &= o >= tb ? 0u : ((1u << (8u * (tb - o))) - 1u);   // end of the text
                                acc += f >> 7;
                            }
                            acc_lo += acc & 0x00FF00FFu;
                            acc_hi += (acc >> 8) & 0x00FF00FFu;
                        }
                    }
                    u32 total = acc_lo + acc_hi;
                    total = (total & 0xFFFFu) + (total >> 16);
                    total = __reduce_add_sync(0xffffffffu, total);
                    // descriptor = four 14-bit per-tile counts at bit 16 i (a tile has < 2^14 bytes)
                    if (lane == 0 && total) atomicAdd(&S.csum[i >> 1], total << (16u * (i & 1u)));
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * WS_NCW) : "memory");
                if (cw == 0 && lane == 0) {
                    st_volatile_u64(&A.tile_desc[gb + c], (1ull << 62) | ((u64)S.csum[1] << 32) | (u64)S.csum[0]);
                    S.csum[0] = 0; S.csum[1] = 0;   // (the team adds to them again only after the next barrier)
                }
            }
        }
    } else if (wid < WS_ROLE0 + WS_NFW) {
        // ======================= front end: scan warps and mask warps ==============================
        // Each role works as WS_NPG teams on alternating tiles (team = k % WS_NPG), so that a warp has
        // WS_NPG tile periods for its share of a tile.
        const u32 fw = wid - WS_ROLE0;
        const bool is_scan = fw < WS_NSW;
#ifdef WS_FUSED_FRONT
        const u32 team = fw / WS_SCT, member = fw % WS_SCT;
#else
        const u32 team = (is_scan ? fw : fw - WS_NSW) / (is_scan ? WS_SCT : WS_MKT);
        const u32 member = (is_scan ? fw : fw - WS_NSW) % (is_scan ? WS_SCT : WS_MKT);
#endif
        auto scan = [&](u32 s, u32 t, u32 k) {
            WsStage &G = S.st[s];
            uint8_t *tx = tbuf + s * WS_TBUF;
            const u32 tb = tile_len(t);
            for (u32 c = member; c < NCHUNK; c += WS_SCT) {
                const u32 off = c * 2048u + lane * 64u;        // my 64 bytes (two blocks)
                if (t == 0 && c == 0 && lane < MK_HALO / 4) reinterpret_cast<u32 *>(tx)[lane] = 0;
                if (off + 64u > tb && off < TB) {               // blank what lies outside the text (last tile only)
                    u32 from = off > tb ? off : tb;
#pragma unroll 1
                    for (u32 i = from; i < off + 64u; i++) tx[MK_HALO + i] = 0;
                }
                if (RAW) continue;
                u32 m0 = 0, m1 = 0;
                if (off < TB) {
                    const uint4 *q = reinterpret_cast<const uint4 *>(tx + MK_HALO + off);
                    // The four 16-byte pieces are read in a per-lane rotated order (conflict-free; plain
                    // order costs 1 %); piece k4 of the loop is piece (k4 + rot) & 3 of the 64 bytes, so the
                    // 64-bit mask is the loop-order concatenation rotated left by 16 rot bits.
                    const u32 rot = (lane >> 1) & 3u;
                    u32 r[4];
#pragma unroll
                    for (int k4 = 0; k4 < 4; k4++) {
                        const u32 pc = (k4 + rot) & 3u;
                        uint4 v = q[pc];
                        u32 w[4] = {v.x, v.y, v.z, v.w};
                        u32 m = 0;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            u32 t7 = ((w[j] ^ c0a) & c7f) + c7f;
                            u32 f = ~(t7 | w[j]) & 0x80808080u;            // 0x80 where the byte is '\n' (exact)
                            m = __funnelshift_r(m, __umulhi(f, 0x02040810u), 4);
                        }
                        r[k4] = m;                                        // the piece's 16 bits sit in the top half
                    }
                    u32 lo = __byte_perm(r[0], r[1], 0x7632), hi = __byte_perm(r[2], r[3], 0x7632);
                    if (rot & 2u) { const u32 t = lo; lo = hi; hi = t; }
                    const u32 sh = (rot & 1u) * 16u;
                    m0 = __funnelshift_l(hi, lo, sh);
                    m1 = __funnelshift_l(lo, hi, sh);
                }
                const u32 c0 = __popc(m0), cnt = c0 + __popc(m1);
                u32 incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    u32 x = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= (u32)o) incl += x;
                }
                const u32 b0 = c * 64u + 2u * lane;
                if (b0 < NBLK) { G.nlm[b0] = m0; G.exw[b0] = (uint16_t)(incl - cnt); }
                if (b0 + 1 < NBLK) { G.nlm[b0 + 1] = m1; G.exw[b0 + 1] = (uint16_t)(incl - cnt + c0); }
                if (lane == 31) G.ctot[c] = incl;
            }
            fence_proxy_async();
            __syncwarp();
#ifdef WS_FUSED_FRONT
            named_bar_sync(4 + WS_NPG + team, 32u * WS_SCT);     // the whole tile is scanned
            if (member == 0 && !RAW) {      // a line of 4095+ bytes (fgets splits it, MK_ERR_LONG_LINE) covers a whole chunk
                const u32 v = lane < NCHUNK ? G.ctot[lane] : 1u;
                const u32 cend = (lane + 1u) * 2048u < TB ? (lane + 1u) * 2048u : TB;
                if (__any_sync(0xffffffffu, lane < NCHUNK && v == 0 && cend <= tb) && lane == 0)
                    atomicOr(A.flags, FLAG_MAYBE_LONG);
            }
            return;
#endif
            // The warp of the team that finishes the tile's scan last turns the per-chunk counts into the
            // in-tile prefix the mask stage needs (the counts other CTAs look back on come from the
            // count-ahead pass, not from here).
            u32 old = 0;
            if (lane == 0) { __threadfence_block(); old = atomicAdd(&S.scnt[s], 1u); }
            old = __shfl_sync(0xffffffffu, old, 0);
            if (old == WS_SCT - 1 && !RAW) {
                __threadfence_block();
                u32 v = lane < NCHUNK ? G.ctot[lane] : 0;
                u32 incl = v;
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) {
                    u32 x = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= (u32)o) incl += x;
                }
                if (lane < NCHUNK) G.cpre[lane] = incl - v;
                // a line of 4095+ bytes (fgets splits it, MK_ERR_LONG_LINE) covers a whole chunk
                const u32 cend = (lane + 1u) * 2048u < TB ? (lane + 1u) * 2048u : TB;
                if (__any_sync(0xffffffffu, lane < NCHUNK && v == 0 && cend <= tb) && lane == 0)
                    atomicOr(A.flags, FLAG_MAYBE_LONG);
            }
            if (old == WS_SCT - 1 && lane == 0) S.scnt[s] = 0;
            __syncwarp();
            if (lane == 0) { __threadfence_block(); mbar_arrive(&S.scanned[s]); }
        };
        auto mask = [&](u32 s, u32 t, u32 k) {
            WsStage &G = S.st[s];
            const u32 tb = tile_len(t);
            const u32 P = RAW ? 0u : (u32)S.Pn[k];      // k = tile sequence number mod 2 NS
#ifdef WS_FUSED_FRONT
            u32 cpre_incl = 0, cpre_v = 0;              // in-tile prefix of the chunks' newline counts, per warp
            if (!RAW) {
                cpre_v = lane < NCHUNK ? G.ctot[lane] : 0;
                cpre_incl = cpre_v;
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) {
                    u32 x = __shfl_up_sync(0xffffffffu, cpre_incl, o);
                    if (lane >= (u32)o) cpre_incl += x;
                }
            }
#endif
            for (u32 u = member; u < NMUNIT; u += WS_MKT) {
#ifdef WS_FUSED_FRONT
                const u32 cpre_u = __shfl_sync(0xffffffffu, cpre_incl - cpre_v, u & 31u);     // (unit u = chunk u)
#endif
                u32 pm[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const u32 b = u * 64u + 32u * h + lane;
                    pm[h] = 0;
                    if (b < NBLK) {
                        if (RAW) {
                            u32 lo = 32 * b;
                            pm[h] = lo >= tb ? 0u : (tb - lo >= 32 ? 0xffffffffu : ((1u << (tb - lo)) - 1u));
                        } else {
                            const u32 nl = G.nlm[b];
#ifdef WS_FUSED_FRONT
                            const u32 s0 = (P + cpre_u + G.exw[b]) & 3u;
#else
                            const u32 s0 = (P + G.cpre[b >> 6] + G.exw[b]) & 3u;
#endif
                            const u32 tgt = (1u - s0) & 3u;
                            // (no "if (nl)": some lane of the warp always has a newline, the branch only costs)
                            const u32 p0 = prefix_xor(nl << 1);
                            const u32 p1 = prefix_xor((nl & p0) << 1);
                            pm[h] = (p0 ^ ((tgt & 1u) ? 0u : ~0u)) & (p1 ^ ((tgt & 2u) ? 0u : ~0u)) & ~nl;
                        }
                        G.nlm[b] = pm[h];
                    }
                }
                const u32 ma = __ballot_sync(0xffffffffu, pm[0] != 0);
                const u32 mb = __ballot_sync(0xffffffffu, pm[1] != 0);
                const u32 na = __popc(ma);
                u32 base = 0;
                if (lane == 0 && (ma | mb)) base = atomicAdd(&S.n_items[s], na + (u32)__popc(mb));
                base = __shfl_sync(0xffffffffu, base, 0);
                const u32 lt = (1u << lane) - 1u;
                if (pm[0]) G.items[base + __popc(ma & lt)] = (uint16_t)(u * 64u + lane);
                if (pm[1]) G.items[base + na + __popc(mb & lt)] = (uint16_t)(u * 64u + 32u + lane);
            }
            __syncwarp();
            if (lane == 0) { __threadfence_block(); mbar_arrive(&S.ready[s]); }
        };
#ifdef WS_FUSED_FRONT
        (void)is_scan;
        for (u32 k = team, s = team % NS, par = (team / NS) & 1u, k2 = team % (2 * NS);; k += WS_NPG) {
            if (k != team) {
                s += WS_NPG; if (s >= NS) { s -= NS; par ^= 1u; }
                k2 += WS_NPG; if (k2 >= 2 * NS) k2 -= 2 * NS;
            }
            if (!role_wait(&S.full[s], nullptr, par, 30, k, 4 + team, WS_SCT, member == 0)) break;
            const u32 t = S.tile[s];
            if (t >= A.n_tiles) {                       // end marker: wake this team's probe group, then leave
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.ready[s]);
                break;
            }
            stamp(k, 0);
            scan(s, t, k2);
            stamp(k, 1);
            mask(s, t, k2);
            stamp(k, 3);
        }
#else
        if (is_scan) {
            // (stage, parity and the 2 NS-periodic index follow k incrementally: no divisions in the loops)
            for (u32 k = team, s = team % NS, par = (team / NS) & 1u, k2 = team % (2 * NS);; k += WS_NPG) {
                if (k != team) {
                    s += WS_NPG; if (s >= NS) { s -= NS; par ^= 1u; }
                    k2 += WS_NPG; if (k2 >= 2 * NS) k2 -= 2 * NS;
                }
                if (!role_wait(&S.full[s], nullptr, par, 30, k, 4 + team, WS_SCT, member == 0)) break;
                const u32 t = S.tile[s];
                if (t >= A.n_tiles) break;                  // (one end marker per team)
                stamp(k, 0);
                scan(s, t, k2);
                stamp(k, 1);
            }
        } else {
            for (u32 k = team, s = team % NS, par = (team / NS) & 1u, k2 = team % (2 * NS);; k += WS_NPG) {
                if (k != team) {
                    s += WS_NPG; if (s >= NS) { s -= NS; par ^= 1u; }
                    k2 += WS_NPG; if (k2 >= 2 * NS) k2 -= 2 * NS;
                }
                // (an end marker's `scanned` never completes: look at the tile id first)
                if (!role_wait(&S.full[s], nullptr, par, 32, k, 4 + WS_NPG + team, WS_MKT, member == 0)) break;
                const u32 t = S.tile[s];
                if (t >= A.n_tiles) {                       // end marker: wake this team's probe group, then leave
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&S.ready[s]);
                    break;
                }
                if (!role_wait(&S.scanned[s], nullptr, par, 33, k, 4 + WS_NPG + team, WS_MKT, member == 0)) break;
                stamp(k, 2);
                mask(s, t, k2);
                stamp(k, 3);
            }
        }
#endif
    } else {
        // ======================= probe ==========================================================
        const u32 pg = (wid - WS_ROLE0 - WS_NFW) / WS_NPW, pw = (wid - WS_ROLE0 - WS_NFW) % WS_NPW;
        for (u32 k = pg, s = pg % NS, par = (pg / NS) & 1u, rot = 0;; k += WS_NPG) {
            if (k != pg) {
                s += WS_NPG; if (s >= NS) { s -= NS; par ^= 1u; }
                if (++rot == WS_NPW) rot = 0;          // rot = (k / WS_NPG) % WS_NPW
            }
            stamp(k, 0);
            if (!role_wait(&S.ready[s], nullptr, par, 40, k, 2 + pg, WS_NPW, pw == 0)) break;
            stamp(k, 1);
            const u32 t = S.tile[s];
            if (t >= A.n_tiles) break;
            WsStage &G = S.st[s];
            const uint8_t *tx = tbuf + s * WS_TBUF;
            const u64 T = (u64)t * TB;
            const u32 n = S.n_items[s];
            const bool two_hash = A.two_hash != 0;
            for (u32 r = pw >= rot ? pw - rot : pw + WS_NPW - rot; r * 32u < n; r += WS_NPW) {
                const u32 it = r * 32u + lane;
                if (it < n) {
                    const u32 b = G.items[it];
                    u32 Aw[4];
                    u32 hits = probe_block<ROTOFF, WORDMASK, PREW, WS_FIXED_FILTER_BASE>(tx + MK_HALO + 32 * b, bm, A.kp.shift_s, Aw);
                    hits &= G.nlm[b];                       // sequence-byte mask by now
                    while (hits) {
                        u32 j = __ffs(hits) - 1;
                        hits &= hits - 1;
                        if (two_hash && !second_hash_hit(Aw, j, bm)) continue;
                        emit_hit(A, T + 32 * b + j);
                    }
                }
            }
            stamp(k, 2);
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.done[s]);
        }
    }
}

// ---- exact verification of the hit list -------------------------------------------------------
// One thread per position that passed the shared-memory filter (true members of the pass set
// plus ~0.025 % Bloom false positives): the TL bytes ending there must all be ACGT — hence lie
// inside one line — and the canonical k-mer's inner substring must be in the pass set
// (iseq2comem.c:682-699).  Writes the sketch code (or EMPTY for a rejected hit) and the global
// position.  Text comes from global memory: ~10 hits per 24 KB tile, mostly L2 hits.
__global__ void __launch_bounds__(256)
k_verify(const uint8_t *__restrict__ text, u64 n, u64 pos_base, KParams kp, const u64 *__restrict__ ptab,
         u64 *__restrict__ cand_code, u64 *__restrict__ cand_pos, u64 nbytes, int quality, u32 line_limit)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 pos = cand_pos[i];
    u64 code = ~0ull;
    if (pos + 1 >= (u64)kp.TL) {
        u64 c;
        if (verify_kmer(text + pos, kp, ptab, &c)) code = c;
    }
    if (code != ~0ull && quality != INT_MIN) {
        // fastq2co() (iseq2comem.c:366-368): every base of the k-mer needs (signed char)quality >= Q, the quality byte
        // being the one in the same column of the record's fourth line.  Line start: back to the previous '\n';
        // quality line: after the second '\n' that follows the k-mer.
        u64 a = pos;                                   // first byte of the sequence line
        for (u32 s = 0; a > 0 && text[a - 1] != '\n' && s <= line_limit; s++) a--;
        u64 e = pos + 1;                               // end of the sequence line
        for (u32 s = 0; e < nbytes && text[e] != '\n' && s <= line_limit; s++) e++;
        u64 p2 = e + 1;                                // end of the '+' line
        for (u32 s = 0; p2 < nbytes && text[p2] != '\n' && s <= line_limit; s++) p2++;
        const u64 qa = p2 + 1;
        const u64 col0 = pos + 1 - (u64)kp.TL - a;
        bool ok = e < nbytes && p2 < nbytes && qa + col0 + (u64)kp.TL <= nbytes;
        for (int j = 0; ok && j < kp.TL; j++) ok = (int)(signed char)text[qa + col0 + (u64)j] >= quality;
        if (!ok) code = ~0ull;
    }
    cand_code[i] = code;
    cand_pos[i] = pos_base + pos;
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

// fastq2co() (iseq2comem.c:323-419) uses a record after the first only if its fourth line ended with a newline:
// with N newlines in the text, sequence lines after newline number 4 floor(N / 4) are dropped (N < 4: only the
// first record exists, it is always used).  One warp finds the last four newlines.
__global__ void k_last_newlines(const uint8_t *__restrict__ text, u64 n, long long *out /*[4]*/)
{
    const u32 lane = threadIdx.x;
    long long found[4] = {-1, -1, -1, -1};
    int nfound = 0;
    long long hi = (long long)n - 1;
    while (hi >= 0 && nfound < 4) {
        long long idx = hi - (long long)lane;
        bool is_nl = idx >= 0 && text[idx] == '\n';
        u32 m = __ballot_sync(0xffffffffu, is_nl);
        while (m && nfound < 4) {
            u32 l = __ffs(m) - 1;                 // smallest lane = highest offset
            found[nfound++] = hi - (long long)l;
            m &= m - 1;
        }
        hi -= 32;
    }
    if (lane == 0)
        for (int i = 0; i < 4; i++) out[i] = found[i];
}

int mk_tail_cut_fq2co(mk_ctx *ctx, const uint8_t *d_text, size_t nbytes, u64 n_newlines, long long *keep_below)
{
    *keep_below = LLONG_MAX;
    if (n_newlines < 4 || nbytes == 0) return MK_OK;
    long long *d_out, h[4];
    CKR(mk_scratch(ctx, SB_MISC, 64, &d_out));
    k_last_newlines<<<1, 32, 0, ctx->stream>>>(d_text, (u64)nbytes, d_out);
    LAUNCH_COUNT(ctx);
    CK(cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->prof.d2h_bytes += sizeof(h);
    const long long cut = h[n_newlines % 4];         // newline number 4 floor(N / 4), counted from the end
    *keep_below = cut >= 0 ? cut + 1 : LLONG_MAX;
    return MK_OK;
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

// ---- exact line-length check (only when the stream kernel saw a newline-free chunk) ------------
// fgets(buf, 4096) splits a line once 4095 bytes came without a newline; what the reference does with
// the pieces depends on stale buffer contents, so such input is refused.  A run of >= 4095 newline-free
// bytes covers an aligned 2 KB chunk entirely: one warp per such chunk measures the run around it.
__global__ void __launch_bounds__(256) k_long_line_check(const uint8_t *__restrict__ text, u64 n, u32 *__restrict__ flag, u64 limit)
{
    const u32 lane = threadIdx.x & 31;
    const u64 nchunks = n / 2048;
    const u64 warp0 = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 c = warp0; c < nchunks; c += nwarps) {
        const uint4 *q = reinterpret_cast<const uint4 *>(text + c * 2048 + (u64)lane * 64);
        u32 any = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint4 v = q[j];
            u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                u32 t7 = ((w[i] ^ 0x0A0A0A0Au) & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
                any |= ~(t7 | w[i]) & 0x80808080u;
            }
        }
        if (__any_sync(0xffffffffu, any != 0)) continue;
        const u64 lo = c * 2048, hi = lo + 2048;
        u64 back = 0, fwd = 0;                       // newline-free bytes right before lo / right after hi
        while (back + 2048 < limit && back < lo) {
            const u64 d = back + lane;               // looks at byte lo - 1 - d
            const bool nl = d < lo && text[lo - 1 - d] == '\n';
            const u32 m = __ballot_sync(0xffffffffu, nl);
            if (m) { back += __ffs(m) - 1; break; }
            back = back + 32 < lo ? back + 32 : lo;
        }
        while (back + 2048 + fwd < limit && hi + fwd < n) {
            const u64 p = hi + fwd + lane;
            const bool nl = p < n && text[p] == '\n';
            const u32 m = __ballot_sync(0xffffffffu, nl);
            if (m) { fwd += __ffs(m) - 1; break; }
            fwd = hi + fwd + 32 < n ? fwd + 32 : n - hi;
        }
        if (back + 2048 + fwd >= limit && lane == 0) atomicOr(flag, 1u);
    }
}

// ---- host driver -------------------------------------------------------------------------------
typedef void (*stream_kernel_t)(const StreamArgs);

template <int ROTOFF, u32 WORDMASK, int NS>
static stream_kernel_t pick2_ws(int prew, bool raw)
{
    if (prew == 1) return raw ? k_stream_ws<ROTOFF, WORDMASK, 1, true, NS> : k_stream_ws<ROTOFF, WORDMASK, 1, false, NS>;
    return raw ? k_stream_ws<ROTOFF, WORDMASK, 2, true, NS> : k_stream_ws<ROTOFF, WORDMASK, 2, false, NS>;
}
static int ws_stages(const KParams &kp) { return (kp.mw == 20 || (kp.mw >= 22 && MK_BLOOM_WBITS == 15)) ? 6 : 10; }
static stream_kernel_t pick_kernel_ws(const KParams &kp, bool raw)
{
    if (kp.mw >= 22) return pick2_ws<MK_BLOOM_WBITS + 2, ((1u << MK_BLOOM_WBITS) - 1u) << 2, MK_BLOOM_WBITS == 15 ? 6 : 10>(kp.prew, raw);   // two-hash filter
    if (kp.mw == 20) return pick2_ws<17, 0x1FFFCu, 6>(kp.prew, raw);    // 128 KB exact bitmap, 6 stages
    if (kp.mw == 16) return pick2_ws<13, 0x1FFCu, 10>(kp.prew, raw);
    if (kp.mw == 12) return pick2_ws<9, 0x1FCu, 10>(kp.prew, raw);
    return nullptr;
}
static size_t stream_smem_bytes_ws(u32 bitmap_bytes, int ns)
{
    return ((bitmap_bytes + 127u) & ~127u) + (size_t)ns * WS_TBUF + (ns == 6 ? sizeof(WsSmem<6>) : sizeof(WsSmem<10>)) + 64;
}

template <int ROTOFF, u32 WORDMASK>
static stream_kernel_t pick2(int prew, bool raw)
{
    if (prew == 1) return raw ? k_stream<ROTOFF, WORDMASK, 1, true> : k_stream<ROTOFF, WORDMASK, 1, false>;
    return raw ? k_stream<ROTOFF, WORDMASK, 2, true> : k_stream<ROTOFF, WORDMASK, 2, false>;
}
static stream_kernel_t pick_kernel(const KParams &kp, bool raw)
{
    if (kp.mw >= 22) return pick2<MK_BLOOM_WBITS + 2, ((1u << MK_BLOOM_WBITS) - 1u) << 2>(kp.prew, raw);
    if (kp.mw == 20) return pick2<17, 0x1FFFCu>(kp.prew, raw);
    if (kp.mw == 16) return pick2<13, 0x1FFCu>(kp.prew, raw);
    if (kp.mw == 12) return pick2<9, 0x1FCu>(kp.prew, raw);
    return nullptr;
}

static size_t stream_smem_bytes(u32 bitmap_bytes)
{
    size_t s = (bitmap_bytes + 127u) & ~127u;
    s += 3 * TBUF_STRIDE;
    s += sizeof(StreamSmem);
    return s + 64;
}

int mk_stream_fastq(mk_ctx *ctx, const uint8_t *d_text, size_t nbytes, u64 pos_base, u64 line_base, bool raw_mode,
                    u64 **d_cand_code, u64 **d_cand_pos, u64 *n_cand, u64 *n_newlines)
{
    *n_cand = 0;
    if (n_newlines) *n_newlines = line_base;
    if (ctx->no_tables) {
        snprintf(ctx->err, sizeof(ctx->err), "this context was created without a .shuf permutation (composite only)");
        return MK_ERR_ARG;
    }
    if (nbytes == 0) return MK_OK;
    if ((uintptr_t)d_text & 15) {
        snprintf(ctx->err, sizeof(ctx->err), "device text pointer must be 16-byte aligned");
        return MK_ERR_ARG;
    }
    const KParams &kp = ctx->kp;
    // MK_STREAM_IMPL selects the kernel: default k_stream_ws (warp-specialised ring); "v3" = k_stream3 (front
    // end ahead of the ring, mk_stream3.cu) and "classic" = the unit-pulling kernel, both kept as cross-checks
    const char *impl = getenv("MK_STREAM_IMPL");
    const bool v3 = !ctx->classic_only && impl && !strcmp(impl, "v3");
    const bool ws = !ctx->classic_only && (v3 || !(impl && !strcmp(impl, "classic")));      // (v3 shares the chunked-upload path with ws)
    stream_kernel_t kern = v3 ? nullptr : (ws ? pick_kernel_ws(kp, raw_mode) : pick_kernel(kp, raw_mode));
    if (!kern && !v3) {
        snprintf(ctx->err, sizeof(ctx->err), "unsupported inner substring width subk=%d", kp.subk);
        return MK_ERR_UNSUPPORTED;
    }
    // tile-proper bytes: a multiple of 64 (two 32-byte probe blocks per scanning thread)
    const u32 max_tile = v3 ? (u32)S3_TILE : (ws ? (u32)WS_TILE : (u32)MK_MAX_TILE);
    u32 tile_bytes = raw_mode ? (max_tile < 16384u ? max_tile : 16384u) : max_tile;
    if (const char *e = getenv(raw_mode ? "MK_RAW_TILE_BYTES" : "MK_TILE_BYTES")) {
        u32 v = (u32)atoi(e);
        v &= ~63u;
        if (v < 64u) v = 64u;
        if (v > max_tile) v = max_tile;
        tile_bytes = v;
    }
    u64 n_tiles64 = (nbytes + tile_bytes - 1) / tile_bytes;
    if (n_tiles64 > 0x7FFFFFFFull) return MK_ERR_UNSUPPORTED;
    u32 n_tiles = (u32)n_tiles64;

    u64 *desc, *counters;
    CKR(mk_scratch(ctx, SB_TILE_DESC, (size_t)n_tiles, &desc));
    CKR(mk_scratch(ctx, SB_COUNTERS, 16, &counters));
    // hit list capacity: members of S ∪ revcomp(S) (2 x pass rate) plus filter false positives
    double rate = 2.0 * (double)kp.dim_end / (double)(1ull << (4 * kp.subk)) + 0.003;
    if (rate > 1.0) rate = 1.0;
    u64 cap = (u64)((double)nbytes * rate * 0.75) + 65536;
    size_t smem = 0;
    if (!v3) {
        smem = ws ? stream_smem_bytes_ws(ctx->bitmap_words * 4, ws_stages(kp)) : stream_smem_bytes(ctx->bitmap_words * 4);
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    uint8_t *desc8 = nullptr;
    u64 *ttab = nullptr;
    size_t arena_items = v3 ? mk_s3_arena_items(nbytes, tile_bytes) : 0;
    if (v3) {
        CKR(mk_scratch(ctx, SB_TILE_DESC8, (size_t)n_tiles + 64, &desc8));
        CKR(mk_scratch(ctx, SB_TILE_DESC, (size_t)n_tiles, &ttab));
    }
    const u32 group = v3 ? 1u : (u32)WS_G;      // launches of the chunked path start on a ticket-group boundary

    // Host source pending (mk_fastq_koc_host): upload and sketch in chunks, the copy of chunk i+1 under
    // the kernel of chunk i.  Each chunk is one launch over its tile range; the line count carries over
    // through device memory, candidates accumulate in the same list.
    const uint8_t *h_src = ctx->h_src;
    ctx->h_src = nullptr;
    if (!ws) h_src = nullptr;
    auto upload_all = [&]() -> int {     // (fallback: plain chunked upload on the compute stream)
        const size_t CH = (size_t)256 << 20;
        for (size_t o = 0; o < nbytes; o += CH) {
            size_t m = nbytes - o < CH ? nbytes - o : CH;
            CK(cudaMemcpyAsync(const_cast<uint8_t *>(d_text) + o, ctx->h_src_all + o, m, cudaMemcpyHostToDevice, ctx->stream));
        }
        return MK_OK;
    };
    if (ctx->h_src_all && !h_src) { CKR(upload_all()); }
    ctx->h_src_all = nullptr;

    for (int attempt = 0; attempt < 3; attempt++) {
        u64 *cc, *cp;
        CKR(mk_scratch(ctx, SB_CAND_CODE, (size_t)cap, &cc));
        CKR(mk_scratch(ctx, SB_CAND_POS, (size_t)cap, &cp));
        u32 *arena = nullptr;
        if (v3) {
            CKR(mk_scratch(ctx, SB_S3_ARENA, arena_items, &arena));
            CK(cudaMemsetAsync(desc8, 0, (size_t)n_tiles + 64, ctx->stream));
            CK(cudaMemsetAsync(ttab, 0, (size_t)n_tiles * 8, ctx->stream));
        } else CK(cudaMemsetAsync(desc, 0, (size_t)n_tiles * 8, ctx->stream));
        CK(cudaMemsetAsync(counters, 0, 128, ctx->stream));
        S3Args a3;
        a3.text = d_text; a3.nbytes = nbytes; a3.line_base = line_base; a3.tile_bytes = tile_bytes;
        a3.n_tiles = n_tiles; a3.tile_begin = 0; a3.desc = desc8;
        a3.flags = (u32 *)(counters + 2) + 1;
        a3.ttab = ttab; a3.arena = arena; a3.arena_cap = arena_items; a3.arena_cursor = counters + 7;
        a3.bitmap = ctx->d_bitmap3; a3.bitmap_bytes = ctx->bitmap3_words * 4;
        a3.cand_count = counters + 0; a3.total_newlines = counters + 1; a3.wd = counters + 8;
        a3.TL = kp.TL;
        a3.stats = nullptr; a3.stat_cta = 0;
#ifdef S3_STATS
        u64 *d_stats = nullptr;
        if (getenv("MK_S3_STATS")) {
            CKR(mk_scratch(ctx, SB_MISC, 128, &d_stats));
            CK(cudaMemsetAsync(d_stats, 0, 128 * 8, ctx->stream));
            a3.stats = d_stats;
            a3.stat_cta = (u32)atoi(getenv("MK_S3_STATS"));
        }
#endif
        a3.prew = (kp.pre + 15) / 16; a3.shift_d = 2 * (16 * a3.prew - kp.pre);
        StreamArgs a;
        a.text = d_text; a.nbytes = nbytes; a.pos_base = pos_base; a.line_base = line_base;
        a.tile_bytes = tile_bytes; a.n_tiles = n_tiles; a.tile_begin = 0; a.line_base_ptr = nullptr; a.tile_desc = desc;
        a.cand_count = counters + 0; a.total_newlines = counters + 1;
        a.tile_counter = (u32 *)(counters + 2); a.flags = (u32 *)(counters + 2) + 1;
        a.count_counter = (u32 *)(counters + 3);
        a.bitmap = ctx->d_bitmap; a.bitmap_bytes = ctx->bitmap_words * 4; a.ptab = ctx->d_ptab; a.two_hash = ctx->kp.mw >= 22 ? 1u : 0u;
        a.cand_code = cc; a.cand_pos = cp; a.cand_cap = cap; a.kp = kp; a.trace = (u64 *)ctx->d_trace; a.wd = counters + 8;
        a3.cand_pos = cp; a3.cand_cap = cap;
        const u32 threads = ws ? WS_THREADS : MK_STREAM_THREADS;
        u64 *total_slot = counters + 1;
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
        if (h_src && attempt == 0) {
            size_t chunk_bytes = (size_t)192 << 20;
            if (const char *e = getenv("MK_CHUNK_BYTES")) chunk_bytes = (size_t)atoll(e);   // (tests: force many chunks)
            u32 tiles_per_chunk = (u32)(chunk_bytes / tile_bytes) / group * group;
            if (tiles_per_chunk < group) tiles_per_chunk = group;
            const u32 nchunk = (n_tiles + tiles_per_chunk - 1) / tiles_per_chunk;
            while (ctx->chunk_ev.size() < nchunk) {
                cudaEvent_t e;
                CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                ctx->chunk_ev.push_back(e);
            }
            // the copy stream starts after whatever the compute stream still has queued on the buffer
            CK(cudaEventRecord(ctx->copy_ev[0], ctx->stream));
            CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ev[0], 0));
            for (u32 c = 0; c < nchunk; c++) {
                const u32 t0 = c * tiles_per_chunk, t1 = (c + 1 == nchunk) ? n_tiles : (c + 1) * tiles_per_chunk;
                const size_t o = (size_t)t0 * tile_bytes;
                const size_t m = (c + 1 == nchunk) ? nbytes - o : (size_t)(t1 - t0) * tile_bytes;
                CK(cudaMemcpyAsync(const_cast<uint8_t *>(d_text) + o, h_src + o, m, cudaMemcpyHostToDevice, ctx->copy_stream));
                CK(cudaEventRecord(ctx->chunk_ev[c], ctx->copy_stream));
                CK(cudaStreamWaitEvent(ctx->stream, ctx->chunk_ev[c], 0));
                if (c) {
                    CK(cudaMemsetAsync(a.tile_counter, 0, 4, ctx->stream));
                    CK(cudaMemsetAsync(a.count_counter, 0, 4, ctx->stream));
                }
                a.tile_begin = t0;
                a.n_tiles = t1;
                a.line_base_ptr = c ? total_slot : nullptr;                 // total through the chunk before
                total_slot = counters + ((c & 1u) ? 4 : 1);
                a.total_newlines = total_slot;
                const u32 nt = t1 - t0;
                const u32 grid = nt < (u32)ctx->sm_count ? nt : (u32)ctx->sm_count;
                if (v3) {       // (line phase and newline total carry over through the tile descriptors / the counter)
                    a3.tile_begin = t0; a3.n_tiles = t1;
                    CKR(mk_s3_launch(ctx, a3, raw_mode, grid));
                } else {
                    kern<<<grid, threads, smem, ctx->stream>>>(a);
                    LAUNCH_COUNT(ctx);
                    CK(cudaGetLastError());
                }
            }
            if (v3) total_slot = counters + 1;
            ctx->prof.h2d_bytes += nbytes;
        } else {
            u32 grid = n_tiles < (u32)ctx->sm_count ? n_tiles : (u32)ctx->sm_count;
            if (v3) CKR(mk_s3_launch(ctx, a3, raw_mode, grid));
            else {
                kern<<<grid, threads, smem, ctx->stream>>>(a);
                LAUNCH_COUNT(ctx);
                CK(cudaGetLastError());
            }
        }
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        u64 h[16];
        CK(cudaMemcpyAsync(h, counters, 128, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->prof.d2h_bytes += 128;
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        ctx->prof.stream_kernel_ms += ms;
        ctx->prof.stream_kernel_launches++;
        ctx->prof.stream_kernel_bytes += nbytes;
#ifdef S3_STATS
        if (a3.stats) {
            u64 st[128];
            CK(cudaMemcpy(st, a3.stats, sizeof(st), cudaMemcpyDeviceToHost));
            fprintf(stderr, "[s3 hist, 8K-cycle bins] scan:");
            for (int i = 0; i < 16; i++) fprintf(stderr, " %llu", st[32 + i]);
            fprintf(stderr, " | lookback:");
            for (int i = 0; i < 16; i++) fprintf(stderr, " %llu", st[48 + i]);
            fprintf(stderr, "\n");
            fprintf(stderr, "[s3 stats, CTA %u, kilo-cycles] loader: free-wait %llu (%llu waits) | dispatcher (%llu items): table-wait %llu full-wait %llu room-wait %llu | front (%llu tiles, %llu items): ahead-wait %llu scan %llu lookback %llu (%llu retries, %llu windows) emit %llu publish %llu loop-total %llu | probe (%llu claims, %llu items): claim-wait %llu work %llu\n",
                    a3.stat_cta, st[0] / 1000, st[1], st[27], st[24] / 1000, st[25] / 1000, st[26] / 1000, st[9], st[13], st[8] / 1000, st[10] / 1000,
                    st[11] / 1000, st[20], st[21], st[12] / 1000, st[15] / 1000, st[14] / 1000, st[17], st[18], st[16] / 1000, st[19] / 1000);
        }
#endif
        u32 flags = (u32)(h[2] >> 32);
        if (flags & FLAG_WATCHDOG) {
            snprintf(ctx->err, sizeof(ctx->err),
                     "k_stream watchdog: site %llu block %llu warp %llu a=%llu b=%llu c=%llu d=%llu tiles=%llu/%llu/%llu n_tiles=%u",
                     (unsigned long long)h[8], (unsigned long long)h[9], (unsigned long long)h[10],
                     (unsigned long long)h[11], (unsigned long long)h[12], (unsigned long long)h[13],
                     (unsigned long long)h[14], (unsigned long long)(h[15] >> 42),
                     (unsigned long long)((h[15] >> 21) & 0x1FFFFF), (unsigned long long)(h[15] & 0x1FFFFF), n_tiles);
            return MK_ERR_CUDA;
        }
        if (ws && !v3 && !ctx->classic_only && getenv("MK_DEBUG_FORCE_SMEM_BASE_FLAG")) flags |= FLAG_SMEM_BASE;   // (tests: take the fallback)
        if (flags & FLAG_SMEM_BASE) {
            // This driver places dynamic shared memory elsewhere: the lookups of this launch read the wrong words.
            // The unit-pulling kernel addresses the filter through the generic pointer: use it from now on (slower,
            // same results; -DWS_FIXED_FILTER_BASE=false rebuilds k_stream_ws without the assumption).
            if (ctx->classic_only) {
                snprintf(ctx->err, sizeof(ctx->err), "k_stream: shared-memory base check failed twice");
                return MK_ERR_UNSUPPORTED;
            }
            fprintf(stderr, "mkssd_b200: dynamic shared memory does not start at CTA-shared address %d on this driver; "
                            "using the unit-pulling stream kernel (rebuild with -DWS_FIXED_FILTER_BASE=false for k_stream_ws)\n",
                    MK_DYN_SMEM_BASE);
            ctx->classic_only = true;
            return mk_stream_fastq(ctx, d_text, nbytes, pos_base, line_base, raw_mode, d_cand_code, d_cand_pos, n_cand, n_newlines);
        }
        if (v3 && (flags & FLAG_ARENA_FULL)) {       // unusually many short sequence lines: the cursor says what is needed
            arena_items = (size_t)h[7] + 1024;
            continue;
        }
        bool long_line = !raw_mode && (flags & FLAG_LONG_LINE);
        if (!raw_mode && !long_line && (flags & FLAG_MAYBE_LONG)) {     // rare: measure the line exactly
            u32 *d_flag = (u32 *)(counters + 6), h_flag = 0;
            CK(cudaMemsetAsync(d_flag, 0, 4, ctx->stream));
            k_long_line_check<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d_text, (u64)nbytes, d_flag, (u64)ctx->line_limit);
            LAUNCH_COUNT(ctx);
            CK(cudaMemcpyAsync(&h_flag, d_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            long_line = h_flag != 0;
        }
        if (long_line) {
            snprintf(ctx->err, sizeof(ctx->err), "FASTQ line of %u bytes or more (fgets(…, %u) would split it)", ctx->line_limit, ctx->line_limit + 1);
            return MK_ERR_LONG_LINE;
        }
        if (h[0] > cap) { // candidate buffer too small: size it exactly and run again
            cap = h[0] + 1024;
            continue;
        }
        if (h[0]) {
            k_verify<<<(unsigned)((h[0] + 255) / 256), 256, 0, ctx->stream>>>(d_text, h[0], pos_base, kp, ctx->d_ptab, cc, cp, (u64)nbytes,
                                                                             ctx->verify_quality, ctx->line_limit);
            LAUNCH_COUNT(ctx);
            CK(cudaGetLastError());
        }
        *n_cand = h[0];
        if (n_newlines) *n_newlines = raw_mode ? line_base : (v3 ? line_base + h[1] : h[total_slot - counters]);
        *d_cand_code = cc;
        *d_cand_pos = cp;
        return MK_OK;
    }
    snprintf(ctx->err, sizeof(ctx->err), "candidate buffer overflow persisted");
    return MK_ERR_NOMEM;
}
