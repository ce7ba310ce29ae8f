// mk_stream3.cu — k_stream3, the line-centric stream kernel (default since round 2).
//
// Same job as k_stream_ws (mk_stream.cu): stream the FASTQ text once from HBM, find the k-mers whose
// inner substring is in the .shuf pass set, append their end positions to the hit list that k_verify
// turns into (code, position) candidates.  Replaces the per-base loop of mt_shortreads2koc()
// (/root/reference/iseq2comem.c:672-720) and, in RAW mode, of fasta2co() (iseq2comem.c:245-293).
//
// What changed against k_stream_ws, and why (profiles/r1_k_stream_ncu.md: 1.02 warp instructions per
// base, half of them spent on record structure, every byte compared with '\n' twice, and a shared-
// memory ring that holds a tile through scan + mask + probe):
//
//   * Record structure is resolved AHEAD of the ring, from registers.  Front warps read whole tiles
//     straight from global memory (coalesced 128-bit loads, L2 evict_last so that the ring's TMA load
//     finds the tile in L2), find the newlines ONCE (64-bit mask per lane after a shuffle transpose),
//     publish the tile's newline count and get the line phase of the tile start from a decoupled
//     look-back over one-BYTE tile descriptors (status + count mod 4: fgets()x4 record structure only
//     needs the line number mod 4; one 8-byte load per lane covers 256 predecessors).  They run up to
//     S3_AHEAD rounds in front of the loader, so none of their latencies (global loads, look-back)
//     keeps a tile in shared memory: a ring stage is held for load + probe only.
//   * Line-centric items.  From the masks the front warp derives the sequence lines directly (the
//     newline that ends a line with index = 1 mod 4, and the newline before it) and emits items of 48
//     k-mer end positions aligned to 16 text bytes, clipped to the line: no per-block prefix-parity
//     masks, ~98 % of the probed positions are real k-mer positions (71 % with 32-byte blocks).  Items
//     go to a global arena (4 bytes per 48 positions, ~4 % of the text) indexed by a per-tile table.
//   * Cross-tile item queue.  A dispatcher warp moves the items of every staged tile into one shared-
//     memory queue that the probe warps pull in chunks of 32, so probe lanes stay full whatever a
//     tile's item count is; a per-stage counter of outstanding items recycles the stage.
//   * Two positions per filter lookup.  Windows at positions i and i+1 share an 11-base core; the
//     filter holds two bit planes over hashed cores ("some member ends with this core" / "some member
//     starts with it"), read by ONE shared-memory load: 3 instructions and half a load per position
//     instead of 6 and one.  Positions passing their plane are re-tested with the other core; what
//     passes both goes to k_verify (exact), so the result is unchanged.
//   * Tiles are dealt round robin (tile = block + k * grid): no ticket atomics, and every tile a
//     look-back waits for is being worked on by a front warp that never waits on the ring beyond the
//     S3_AHEAD window.
#include "mk_common.cuh"
#include "mk_stream_dev.cuh"
#include "mk_stream3.cuh"

#define S3_CH (S3_TILE / 2048)
#define S3_TBUF (MK_HALO + S3_TILE + 96)
#define S3_NONE (-64)              // "no newline before": prev + TL <= 0 for every supported k
#define S3_UNKNOWN 0xFFFFFFFFu
#define S3_QSLACK (32 * S3_NP + 64)
#ifndef S3_PROBE_SLEEP
#define S3_PROBE_SLEEP 64          // ns between two looks of an idle probe warp at the queue
#endif
#ifndef S3_PATIENCE
#define S3_PATIENCE 16             // looks a probe warp grants the dispatcher before taking a partial chunk
#endif
static_assert((S3_QN & (S3_QN - 1)) == 0 && S3_QN >= S3_QSLACK + S3_TILE / 16 + 64, "queue size");
static_assert(S3_TILE % 2048 == 0 && S3_TILE / 16 < 1024, "tile size");
static_assert(S3_NS <= 16, "stage field of an item");
static_assert(2 + S3_NF + S3_NP <= 32, "warps");

struct S3Smem {
    u64 full[S3_NS], freeb[S3_NS];
    u32 tile[S3_NS];
    u32 rem[S3_NS];              // items of the staged tile not yet probed
    u32 fcur[S3_NF];             // front warps: items written for the tile in hand
    u32 q_res, q_head, q_final, loader_k, abort;
    u32 Q[S3_QN];
};

__device__ __forceinline__ u32 lanemask_lt()
{
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ u32 lds_volatile(const u32 *p)
{
    u32 v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void sts_volatile(u32 *p, u32 v)
{
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void st_volatile_u8(uint8_t *p, u32 v)
{
    asm volatile("st.volatile.global.u8 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// streaming 128-bit load of the front pass: not kept in L1 (every byte is read once per warp), kept in
// L2 (evict_last) for the TMA load of the same tile a few rounds later
__device__ __forceinline__ uint4 ldg_front(const void *p)
{
    uint4 v;
#ifdef S3_PLAIN_LDG
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
#endif
    asm volatile("{\n\t.reg .b64 pol;\n\tcreatepolicy.fractional.L2::evict_last.b64 pol, 1.0;\n\t"
                 "ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], pol;\n\t}"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// window extraction: 32 bits of the 128-bit packed stream W starting at bit O; only bits [0, NEED) of
// the result are used, which lets single-word cases run on the FMA pipe (shr_fma)
template <int O, int NEED>
__device__ __forceinline__ u32 s3_take(const u32 (&W)[4])
{
    constexpr int w = O >> 5, sh = O & 31;
    static_assert(O + NEED <= 128, "window outside the packed block");
    if (sh + NEED <= 32) return shr_fma<sh>(W[w]);
    return __funnelshift_r(W[w], W[w + 1], sh);
}

// One lookup = positions I (plane A: the core is the window's last bases) and I+1 (plane B: the core
// is the window's first bases).  Core of position I: bits [2I+2, 2I+2+cb).  Word index = core bits
// [5, 5+WB), bit = core bits [0,5) (plane A) / that + 1 mod 32 (plane B).
template <u32 WORDMASK, int I, int END>
__device__ __forceinline__ void s3_probe_pairs(const u32 (&W)[4], const u32 *bm, u32 &acc)
{
    constexpr int C = 2 * I + 2;
    u32 a = s3_take<C + 3, 17>(W);
    u32 r = s3_take<C, 5>(W);
    u32 word = *reinterpret_cast<const u32 *>(reinterpret_cast<const char *>(bm) + (a & WORDMASK));
    u32 rot = __funnelshift_r(word, word, r);
    acc = __funnelshift_r(acc, rot, 2);
    if constexpr (I + 2 < END) s3_probe_pairs<WORDMASK, I + 2, END>(W, bm, acc);
}

// both planes of one position P (run time): window = 32 bits from bit 2P of W
template <u32 WORDMASK>
__device__ __forceinline__ bool s3_second_level(const u32 (&W)[4], u32 p, const u32 *bm)
{
    const u32 ws = p >> 4;
    const u32 lo = ws == 0 ? W[0] : (ws == 1 ? W[1] : W[2]);
    const u32 hi = ws == 0 ? W[1] : (ws == 1 ? W[2] : W[3]);
    const u32 v = __funnelshift_r(lo, hi, (2u * p) & 31u);
    const u32 wB = *reinterpret_cast<const u32 *>(reinterpret_cast<const char *>(bm) + ((v >> 3) & WORDMASK));
    const u32 wA = *reinterpret_cast<const u32 *>(reinterpret_cast<const char *>(bm) + ((v >> 5) & WORDMASK));
    return ((wB >> ((v + 1u) & 31u)) & (wA >> ((v >> 2) & 31u)) & 1u) != 0;
}

// 16 bytes -> 16-bit newline mask (exact "byte == 0x0A"), in the TOP half of the result
__device__ __forceinline__ u32 s3_newline_mask16(uint4 v, u32 c0a, u32 c7f)
{
    const u32 w[4] = {v.x, v.y, v.z, v.w};
    u32 m = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        u32 t7 = ((w[j] ^ c0a) & c7f) + c7f;
        u32 f = ~(t7 | w[j]) & 0x80808080u;            // 0x80 where the byte is '\n'
        m = __funnelshift_r(m, __umulhi(f, 0x02040810u), 4);
    }
    return m;
}

// Development aid (build with -DS3_STATS, run with MK_S3_STATS=1): cycles CTA 0 spends per role and site.
#ifdef S3_STATS
#define S3_T0() const long long st_t0_ = clock64()
#define S3_ACC(slot) do { if (blockIdx.x == A.stat_cta && lane == 0 && A.stats) atomicAdd((unsigned long long *)&A.stats[slot], (unsigned long long)(clock64() - st_t0_)); } while (0)
#define S3_CNT(slot, v) do { if (blockIdx.x == A.stat_cta && lane == 0 && A.stats) atomicAdd((unsigned long long *)&A.stats[slot], (unsigned long long)(v)); } while (0)
#define S3_HIST(base) do { long long dt_ = (clock64() - st_t0_) >> 13; S3_CNT((base) + (dt_ > 15 ? 15 : dt_), 1); } while (0)
#else
#define S3_HIST(base) do {} while (0)
#define S3_T0() do {} while (0)
#define S3_ACC(slot) do {} while (0)
#define S3_CNT(slot, v) do {} while (0)
#endif

__device__ __noinline__ void s3_watchdog_report(u32 *flags, u64 *wd, u32 site, u32 wid, u64 a, u64 b, u64 c, u64 d)
{
    if (atomicOr(flags, FLAG_WATCHDOG) & FLAG_WATCHDOG) return;   // first report wins
    wd[0] = site; wd[1] = blockIdx.x; wd[2] = wid; wd[3] = a; wd[4] = b; wd[5] = c; wd[6] = d; wd[7] = 0;
}

// per-tile table entry: bit 63 = ready, bits 40..55 = item count, bits 0..39 = arena offset (items)
#define S3_TT_READY (1ull << 63)

// WORDMASK = ((1 << word bits) - 1) << 2;  SHIFTED: the window of position p starts D = 16 PREW - pre
// bases into the block (D != 0 for geometries other than k + subk = 17)
template <u32 WORDMASK, bool SHIFTED, bool RAW>
__global__ void __launch_bounds__(S3_THREADS, 1) k_stream3(const __grid_constant__ S3Args A)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const u32 tid = threadIdx.x;
    const u32 lane = tid & 31, wid = tid >> 5;
    const u32 bm_bytes = (A.bitmap_bytes + 127u) & ~127u;
    u32 *bm = reinterpret_cast<u32 *>(smem);
    uint8_t *tbuf = smem + bm_bytes;
    S3Smem &S = *reinterpret_cast<S3Smem *>(tbuf + S3_NS * S3_TBUF);

    {
        const uint4 *src = reinterpret_cast<const uint4 *>(A.bitmap);
        uint4 *dst = reinterpret_cast<uint4 *>(bm);
        for (u32 i = tid; i < A.bitmap_bytes / 16; i += S3_THREADS) dst[i] = src[i];
    }
    if (tid == 0) {
        for (int s = 0; s < S3_NS; s++) {
            mbar_init(&S.full[s], 1);
            mbar_init(&S.freeb[s], 1);
            S.tile[s] = 0xFFFFFFFFu; S.rem[s] = 0;
        }
        S.q_res = 0; S.q_head = 0; S.q_final = S3_UNKNOWN; S.loader_k = 0; S.abort = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();   // the only block-wide barrier of the kernel

    const u32 TB = A.tile_bytes;
    auto tile_len = [&](u32 t) -> u32 {
        u64 rem = A.nbytes - (u64)t * TB;
        return rem < TB ? (u32)rem : TB;
    };
    auto tile_of = [&](u32 k) -> u64 { return (u64)A.tile_begin + blockIdx.x + (u64)k * gridDim.x; };
    auto watchdog = [&](u32 site, u64 a, u64 b, u64 c, u64 d) {
        s3_watchdog_report(A.flags, A.wd, site, wid, a, b, c, d);
        sts_volatile(&S.abort, 1u);
    };
    auto aborted = [&]() -> bool { return lds_volatile(&S.abort) != 0; };

    if (wid == 0) {
        // ======================= loader (one lane) ================================================
        if (lane == 0) {
            for (u32 k = 0;; k++) {
                const u32 s = k % S3_NS;
                const u64 t64 = tile_of(k);
                if (t64 >= A.n_tiles) break;
                const u32 t = (u32)t64;
                if (k >= S3_NS) {
                    S3_T0();
                    bool ok = mbar_wait(&S.freeb[s], ((k / S3_NS) - 1u) & 1u);
                    S3_ACC(0); S3_CNT(1, 1);
                    if (!ok) { watchdog(10, k, s, S.tile[s], S.rem[s]); break; }
                    if (aborted()) break;
                }
                const u32 tb = tile_len(t);
                uint8_t *dst = tbuf + s * S3_TBUF;
                const uint8_t *src = A.text + (u64)t * TB;
                u32 bytes = tb;
                if (t > 0) { src -= MK_HALO; bytes += MK_HALO; } else { dst += MK_HALO; }
                bytes = (bytes + 15u) & ~15u;
                fence_proxy_async();
                mbar_expect_tx(&S.full[s], bytes);
                tma_load_1d_last_use(dst, src, bytes, &S.full[s]);
                sts_volatile(&S.loader_k, k + 1u);         // (front warps may work S3_AHEAD rounds past this)
            }
        }
    } else if (wid == 1) {
        // ======================= dispatcher =======================================================
        // Moves the items of every staged tile from the arena into the shared-memory queue.  The table
        // entry and the first 256 items of the next tile travel while this tile's TMA load completes.
        u64 ent_next = 0;
        {
            const u64 t0 = tile_of(0);
            if (t0 < A.n_tiles && lane == 0) ent_next = ld_volatile_u64(&A.ttab[t0]);
        }
        u32 qres = 0;
        for (u32 k = 0;; k++) {
            const u32 s = k % S3_NS, par = (k / S3_NS) & 1u;
            const u64 t64 = tile_of(k);
            if (t64 >= A.n_tiles) break;
            const u32 t = (u32)t64;
            // table entry of this tile (normally published rounds ago)
            u64 ent = ent_next;
            bool bad = false;
            if (lane == 0) {
                S3_T0();
                for (u32 n = 0; !(ent & S3_TT_READY); n++) {
                    if (n > WD_LIMIT || aborted()) { bad = true; break; }
                    if (n) __nanosleep(100);
                    ent = ld_volatile_u64(&A.ttab[t]);
                }
                S3_ACC(24);
                const u64 tn = tile_of(k + 1);
                if (tn < A.n_tiles) ent_next = ld_volatile_u64(&A.ttab[tn]);
            }
            if (__any_sync(0xffffffffu, bad)) { if (lane == 0 && !aborted()) watchdog(20, t, k, 0, 0); break; }
            ent = __shfl_sync(0xffffffffu, ent, 0);
            const u32 n = (u32)(ent >> 40) & 0xFFFFu;
            const u32 *items = A.arena + (ent & ((1ull << 40) - 1ull));
            u32 pre[8];
#pragma unroll
            for (int i = 0; i < 8; i++) pre[i] = (u32)i * 32u + lane < n ? __ldcg(items + i * 32 + lane) : 0u;
            {
                S3_T0();
                if (!mbar_wait(&S.full[s], par)) { if (lane == 0) watchdog(21, t, k, s, par); break; }
                S3_ACC(25);
            }
            if (lane == 0) { S.tile[s] = t; S.rem[s] = n; }
            if (n == 0) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.freeb[s]);
                continue;
            }
            {   // room: slots that a probe warp may still be reading are never overwritten
                S3_T0();
                bool room = true;
                for (u32 w = 0; (int)(qres + n - lds_volatile(&S.q_head)) > (int)(S3_QN - S3_QSLACK); w++) {
                    if (w > WD_LIMIT || aborted()) { room = false; break; }
                    __nanosleep(100);
                }
                S3_ACC(26);
                if (!room) { if (lane == 0 && !aborted()) watchdog(22, t, k, qres, n); break; }
            }
            const u32 tag = s << 26;
#pragma unroll
            for (int i = 0; i < 8; i++)
                if ((u32)i * 32u + lane < n) S.Q[(qres + i * 32 + lane) & (S3_QN - 1)] = tag | pre[i];
            for (u32 i = 256u + lane; i < n; i += 32u) S.Q[(qres + i) & (S3_QN - 1)] = tag | __ldcg(items + i);
            qres += n;
            __threadfence_block();
            __syncwarp();
            if (lane == 0) sts_volatile(&S.q_res, qres);
            S3_CNT(27, n);
        }
        __syncwarp();
        if (lane == 0) { __threadfence_block(); sts_volatile(&S.q_final, qres); }
    } else if (wid < 2 + S3_NF) {
        // ======================= front: newlines -> line phase -> items (ahead of the ring) ========
        const u32 fw = wid - 2;
        // byte-compare constants kept in registers ((w ^ c) & m is then ONE three-input LOP3)
        const u32 c0a = 0x0A0A0A0Au + (A.tile_bytes >> 31), c7f = 0x7F7F7F7Fu + (A.tile_bytes >> 31);
        const u32 lt = lanemask_lt();
        const int TL = A.TL;
        u32 *items = nullptr;            // arena slice of the tile in hand
        u32 cap = 0;

        // items of one line fragment: k-mer end positions [xlo, xhi] (tile relative), 48 per item,
        // first block aligned down to 16 bytes
        auto item_word = [&](int xlo, int xhi, u32 i) -> u32 {
            const int bx = (xlo & ~15) + 48 * (int)i;
            const u32 lo = i == 0 ? (u32)(xlo & 15) : 0u;
            const int rest = xhi - bx;
            const u32 hi = rest > 47 ? 47u : (u32)rest;
            return ((u32)(bx >> 4) << 16) | (lo << 6) | hi;
        };
        auto emit_coop = [&](int xlo, int xhi) {      // all lanes, warp-uniform arguments
            if (xhi < xlo) return;
            const u32 n = (u32)(xhi - (xlo & ~15)) / 48u + 1u;
            u32 base = 0;
            if (lane == 0) base = atomicAdd(&S.fcur[fw], n);
            base = __shfl_sync(0xffffffffu, base, 0);
            for (u32 i = lane; i < n; i += 32u)
                if (base + i < cap) items[base + i] = item_word(xlo, xhi, i);
        };

        for (u32 j = fw;; j += S3_NF) {
            const u64 t64 = tile_of(j);
            if (t64 >= A.n_tiles) break;
            const u32 t = (u32)t64;
#ifdef S3_STATS
            const long long st_loop_ = clock64();
#endif
            {   // stay within S3_AHEAD rounds of the loader (the text must still be in L2 when the ring loads it)
                S3_T0();
                bool bad = false;
                for (u32 n = 0; j >= lds_volatile(&S.loader_k) + S3_AHEAD; n++) {
                    if (n > WD_LIMIT || aborted()) { bad = true; break; }
                    __nanosleep(200);
                }
                S3_ACC(8); S3_CNT(9, 1);
                if (bad) { if (lane == 0 && !aborted()) watchdog(30, j, lds_volatile(&S.loader_k), 0, 0); break; }
            }
            const u32 tb = tile_len(t);
            const u32 tb16 = (tb + 15u) & ~15u;
            const uint8_t *gx = A.text + (u64)t * TB;
            {   // the tile this warp takes S3_FPF iterations from now: towards L2 already
                const u64 tn = tile_of(j + S3_FPF * S3_NF);
                if (lane == 0 && tn < A.n_tiles) {
                    const u64 o = tn * TB, rem = A.nbytes - o;
                    prefetch_l2(A.text + o, (u32)((rem < TB ? rem : TB) + 15u) & ~15u);
                }
            }
            S3_T0();
            if (lane == 0) S.fcur[fw] = 0;
            __syncwarp();
            u32 total = 0, P = 0;
            int carry_last = t == 0 ? -1 : S3_NONE;
            u32 M0[S3_CH], M1[S3_CH], ST[S3_CH];
            if (!RAW) {
                // ---- phase 1: newline masks of the whole tile, kept in registers -----------------
#pragma unroll
                for (int c = 0; c < S3_CH; c++) {
                    u32 m0 = 0, m1 = 0;
                    const u32 off = (u32)c * 2048u + lane * 64u;
                    if ((u32)c * 2048u < TB) {
                        // coalesced: piece q of the lane = bytes [512 q + 16 lane, +16) of the chunk
                        u32 pm[4];
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const u32 o = (u32)c * 2048u + 512u * q + 16u * lane;
                            uint4 v = make_uint4(0, 0, 0, 0);
                            if (o < tb16 && o < TB) v = ldg_front(gx + o);
                            pm[q] = s3_newline_mask16(v, c0a, c7f);
                        }
                        // transpose: lane L owns bytes [64 L, 64 L + 64) = pieces (L / 8, 4 (L % 8) + i), i = 0..3
                        const u32 pa = __byte_perm(pm[0], pm[1], 0x7632), pb = __byte_perm(pm[2], pm[3], 0x7632);
                        const u32 src = 4u * (lane & 7u);
                        u32 h[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const u32 xa = __shfl_sync(0xffffffffu, pa, src + i), xb = __shfl_sync(0xffffffffu, pb, src + i);
                            const u32 x = (lane & 16u) ? xb : xa;
                            h[i] = (lane & 8u) ? (x >> 16) : (x & 0xFFFFu);
                        }
                        m0 = h[0] | (h[1] << 16);
                        m1 = h[2] | (h[3] << 16);
                        if (tb < TB) {                       // last tile of the text: bytes past its end do not exist
                            const int v = (int)tb - (int)off;
                            if (v < 64) {
                                m0 &= v <= 0 ? 0u : (v >= 32 ? 0xFFFFFFFFu : ((1u << v) - 1u));
                                m1 &= v <= 32 ? 0u : ((1u << (v - 32)) - 1u);
                            }
                        }
                    }
                    const u32 cnt = __popc(m0) + __popc(m1);
                    const u32 B = __ballot_sync(0xffffffffu, cnt != 0);
                    const u32 b0 = __ballot_sync(0xffffffffu, cnt & 1u), b1 = __ballot_sync(0xffffffffu, cnt & 2u);
                    const u32 pre = (total + __popc(b0 & lt) + 2u * __popc(b1 & lt)) & 3u;
                    total += __reduce_add_sync(0xffffffffu, cnt);
                    const int lastpos = (int)off + (m1 ? 63 - __clz(m1) : 31 - __clz(m0));
                    const u32 below = B & lt;
                    const int pl = __shfl_sync(0xffffffffu, lastpos, (31 - __clz(below)) & 31);
                    const int prevnl = below ? pl : carry_last;
                    const int nl = __shfl_sync(0xffffffffu, lastpos, (31 - __clz(B)) & 31);
                    if (B) carry_last = nl;
                    else {   // a newline-free piece inside the text: the host measures the line exactly (k_long_line_check)
                        const u32 cend = ((u32)c + 1u) * 2048u < TB ? ((u32)c + 1u) * 2048u : TB;
                        if ((u32)c * 2048u < TB && cend <= tb && lane == 0) atomicOr(A.flags, FLAG_MAYBE_LONG);
                    }
                    M0[c] = m0; M1[c] = m1; ST[c] = (u32)(prevnl + 64) | (pre << 16);
                }
            }
            S3_ACC(10); S3_HIST(32);
            // arena slice: an upper bound of the tile's items (fragments <= newlines / 4 + 2, 48 positions per item)
            cap = (tb / 48u + 2u * (total / 4u + 2u) + 4u) & ~3u;
            u64 aoff = 0;
            if (lane == 0) aoff = atomicAdd((unsigned long long *)A.arena_cursor, (unsigned long long)cap);
            if (!RAW) {
                // ---- phase 2: publish the count, look back for the line phase --------------------
                S3_T0();
                if (lane == 0) {
                    st_volatile_u8(A.desc + t, 4u | (total & 3u));
                    if (total) atomicAdd((unsigned long long *)A.total_newlines, (unsigned long long)total);
                }
                const long long base = (long long)(t & ~7u);
                u32 acc = 0;
                bool good = true;
                for (u32 w = 0;; w++) {
                    const long long t0 = base - 8ll * (long long)(lane + 32u * w);
                    u64 d = 0;
                    u32 F = 0;
                    bool pending = true;                 // this lane's eight descriptors are not all published yet
                    for (u32 n = 0;; n++) {
                        // (only lanes with unpublished descriptors look again: the descriptors of the ~500 tiles
                        // in flight sit in two or three L2 lines that every front warp of the chip reads)
                        if (pending) {
                            d = t0 >= 0 ? ld_volatile_u64(reinterpret_cast<const u64 *>(A.desc + t0)) : 0x0808080808080808ull;
                            if (w == 0 && lane == 0) {   // descriptors of tiles >= t: neutral (aggregate, 0)
                                const u32 nv = t & 7u;
                                const u64 keepv = nv ? ((1ull << (8u * nv)) - 1ull) : 0ull;
                                d = (d & keepv) | (0x0404040404040404ull & ~keepv);
                            }
                        }
                        const u64 incl = d & 0x0808080808080808ull;
                        const u64 keep = incl ? (~0ull << (8u * (7u - ((u32)__clzll(incl) >> 3)))) : ~0ull;
                        const u64 ready = (d | (d >> 1)) & 0x0404040404040404ull;
                        const bool notready = ((~ready) & 0x0404040404040404ull & keep) != 0;
                        pending = notready;
                        F = __ballot_sync(0xffffffffu, incl != 0);
                        const u32 NR = __ballot_sync(0xffffffffu, notready);
                        const u32 rel = F ? ((2u << (__ffs(F) - 1)) - 1u) : 0xFFFFFFFFu;   // lanes up to the nearest inclusive one
                        if ((NR & rel) == 0) {
                            const u64 vals = d & 0x0303030303030303ull & keep;
                            u32 sum = __popcll(vals & 0x0101010101010101ull) + 2u * __popcll(vals & 0x0202020202020202ull);
                            if (!((rel >> lane) & 1u)) sum = 0;
                            acc += __reduce_add_sync(0xffffffffu, sum);
                            break;
                        }
                        if (n > WD_LIMIT || aborted()) { good = false; break; }
                        S3_CNT(20, 1);
                        __nanosleep(n < 4 ? (100u << n) : 1600u);
                    }
                    S3_CNT(21, 1);
                    if (!good || F) break;
                }
                if (!good) { if (lane == 0 && !aborted()) watchdog(31, t, j, 0, 0); break; }
                if (lane == 0) st_volatile_u8(A.desc + t, 8u | ((acc + total) & 3u));
                P = ((u32)A.line_base + acc) & 3u;
                S3_ACC(11); S3_HIST(48);
            }
            aoff = __shfl_sync(0xffffffffu, aoff, 0);
            const bool fits = aoff + cap <= A.arena_cap;
            if (!fits) { cap = 0; if (lane == 0) atomicOr(A.flags, FLAG_ARENA_FULL); }
            items = A.arena + (fits ? aoff : 0);
            {
                S3_T0();
                if (RAW) {
                    emit_coop(t == 0 ? TL - 1 : 0, (int)tb - 1);
                } else {
                    // ---- phase 3: sequence lines -> items ---------------------------------------------
#pragma unroll
                    for (int c = 0; c < S3_CH; c++) {
                        if ((u32)c * 2048u >= TB) break;
                        const u32 off = (u32)c * 2048u + lane * 64u;
                        const u32 m0 = M0[c], m1 = M1[c];
                        const u32 pre = ST[c] >> 16;
                        const int prevnl = (int)(ST[c] & 0xFFFFu) - 64;
                        const u32 cnt = __popc(m0) + __popc(m1);
                        if (!__any_sync(0xffffffffu, cnt > 4u)) {
                            // at most one of a lane's (<= 4) newlines ends a line with index = 1 mod 4
                            const u32 js = (1u - P - pre) & 3u;
                            const u64 m = ((u64)m1 << 32) | m0;
                            u64 mm = m;
                            if (js > 0) mm &= mm - 1;
                            if (js > 1) mm &= mm - 1;
                            if (js > 2) mm &= mm - 1;
                            int prev = prevnl;
                            if (js > 0 && mm) prev = (int)off + 63 - __clzll((long long)(m & ((mm & (0 - mm)) - 1)));
                            const int e = (int)off + __ffsll((long long)mm) - 1;
                            const int xlo = prev + TL < 0 ? 0 : prev + TL, xhi = e - 1;
                            const u32 n = (mm && xhi >= xlo) ? (u32)(xhi - (xlo & ~15)) / 48u + 1u : 0u;
                            const u32 big = __ballot_sync(0xffffffffu, n > 6u);
                            if (n && n <= 6u) {
                                const u32 base = atomicAdd(&S.fcur[fw], n);
                                for (u32 i = 0; i < n; i++)
                                    if (base + i < cap) items[base + i] = item_word(xlo, xhi, i);
                            }
                            for (u32 bg = big; bg; bg &= bg - 1) {        // long lines: all lanes write
                                const int l = __ffs(bg) - 1;
                                emit_coop(__shfl_sync(0xffffffffu, xlo, l), __shfl_sync(0xffffffffu, xhi, l));
                            }
                        } else {
                            // rare: five or more newlines inside 64 bytes — walk the chunk's newlines in order
                            u32 rank = __shfl_sync(0xffffffffu, pre, 0);    // newlines before the chunk, mod 4
                            int prev = __shfl_sync(0xffffffffu, prevnl, 0);
                            for (u32 Bm = __ballot_sync(0xffffffffu, cnt != 0); Bm; Bm &= Bm - 1) {
                                const int l = __ffs(Bm) - 1;
                                const u32 lm0 = __shfl_sync(0xffffffffu, m0, l), lm1 = __shfl_sync(0xffffffffu, m1, l);
                                const u32 loff = (u32)c * 2048u + (u32)l * 64u;
                                for (u64 mm = ((u64)lm1 << 32) | lm0; mm; mm &= mm - 1) {
                                    const int e = (int)loff + __ffsll((long long)mm) - 1;
                                    if (((P + rank) & 3u) == 1u) emit_coop(prev + TL < 0 ? 0 : prev + TL, e - 1);
                                    prev = e;
                                    rank++;
                                }
                            }
                        }
                    }
                    if (((P + total) & 3u) == 1u)               // the line that runs past the end of the tile
                        emit_coop(carry_last + TL < 0 ? 0 : carry_last + TL, (int)tb - 1);
                }
                S3_ACC(12);
            }
            // publish the tile's table entry (after its items)
            {
                S3_T0();
                __threadfence();
                __syncwarp();
                if (lane == 0) {
                    u32 n = lds_volatile(&S.fcur[fw]);
                    if (n > cap) n = cap;                        // (only when the arena is full: the host runs again)
                    S3_CNT(13, n);
                    st_volatile_u64(&A.ttab[t], S3_TT_READY | ((u64)n << 40) | (fits ? aoff : 0ull));
                }
                S3_ACC(15);
            }
#ifdef S3_STATS
            S3_CNT(14, clock64() - st_loop_);
#endif
        }
    } else {
        // ======================= probe ==========================================================
        // A warp claims up to 32 queue slots the dispatcher has published (compare-and-swap on the head,
        // never beyond q_res), so a warp never sits on items while waiting for others that may need the
        // stage they pin.  A partial claim is made only after a short patience (keeps lanes full).
        for (;;) {
            u32 h = 0, cnt = 0;
            bool bad = false;
            S3_T0();
            if (lane == 0) {
                for (u32 waited = 0;;) {
                    const u32 head = lds_volatile(&S.q_head);
                    const u32 fin = lds_volatile(&S.q_final);
                    const int avail = (int)(lds_volatile(&S.q_res) - head);
                    if (avail >= 32 || (avail > 0 && (fin != S3_UNKNOWN || waited >= S3_PATIENCE))) {
                        const u32 c = avail < 32 ? (u32)avail : 32u;
                        if (atomicCAS(&S.q_head, head, head + c) == head) { h = head; cnt = c; break; }
                        continue;
                    }
                    if (fin != S3_UNKNOWN && head >= fin) break;          // the queue is complete and empty
                    if (++waited > WD_LIMIT || aborted()) { bad = true; break; }
                    __nanosleep(S3_PROBE_SLEEP);
                }
                if (bad && !aborted()) watchdog(40, lds_volatile(&S.q_head), lds_volatile(&S.q_final), lds_volatile(&S.q_res), 0);
            }
            h = __shfl_sync(0xffffffffu, h, 0);
            cnt = __shfl_sync(0xffffffffu, cnt, 0);
            S3_ACC(16); S3_CNT(17, 1); S3_CNT(18, cnt);
            if (cnt == 0) break;
#ifdef S3_STATS
            const long long st_t1_ = clock64();
#endif
            const bool mine = lane < cnt;
            u32 stage = 31u;
            if (mine) {
                const u32 it = S.Q[(h + lane) & (S3_QN - 1)];
                stage = (it >> 26) & 15u;
                const u32 q = (it >> 16) & 1023u, lo = (it >> 6) & 63u, hi = it & 63u;
                const uint8_t *blk = tbuf + stage * S3_TBUF + MK_HALO + 16u * q;
                u32 W[4];
                if (!SHIFTED) {
                    const uint4 *v = reinterpret_cast<const uint4 *>(blk - 16);
#pragma unroll
                    for (int i = 0; i < 4; i++) W[i] = pack16(v[i]);
                } else {
                    const uint4 *v = reinterpret_cast<const uint4 *>(blk - 16 * A.prew);
                    u32 X[5];
#pragma unroll
                    for (int i = 0; i < 5; i++) X[i] = pack16(v[i]);
#pragma unroll
                    for (int i = 0; i < 4; i++) W[i] = __funnelshift_r(X[i], X[i + 1], A.shift_d);
                }
                u32 h0 = 0, h1 = 0;
                s3_probe_pairs<WORDMASK, 0, 32>(W, bm, h0);
                s3_probe_pairs<WORDMASK, 32, 48>(W, bm, h1);
                h1 >>= 16;
                // positions [lo, hi]
                const u64 range = ((2ull << hi) - 1ull) & ~((1ull << lo) - 1ull);
                h0 &= (u32)range;
                h1 &= (u32)(range >> 32);
                const u64 T = (u64)S.tile[stage] * TB + 16u * q;
                while (h0 | h1) {
                    u32 p;
                    if (h0) { p = __ffs(h0) - 1; h0 &= h0 - 1; } else { p = 32 + __ffs(h1) - 1; h1 &= h1 - 1; }
                    if (s3_second_level<WORDMASK>(W, p, bm)) emit_hit3(A, T + p);
                }
            }
            // hand the items back: the lane group of each stage subtracts its count; zero recycles the stage
            const u32 grp = __match_any_sync(0xffffffffu, stage);
            if (mine && lane == (u32)(__ffs(grp) - 1)) {
                const u32 c = __popc(grp);
                if (atomicSub(&S.rem[stage], c) == c) mbar_arrive(&S.freeb[stage]);
            }
#ifdef S3_STATS
            S3_CNT(19, clock64() - st_t1_);
#endif
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------
typedef void (*s3_kernel_t)(const S3Args);

template <u32 WORDMASK>
static s3_kernel_t s3_pick2(bool shifted, bool raw)
{
    if (shifted) return raw ? k_stream3<WORDMASK, true, true> : k_stream3<WORDMASK, true, false>;
    return raw ? k_stream3<WORDMASK, false, true> : k_stream3<WORDMASK, false, false>;
}

int mk_s3_word_bits(int mw)
{
    int wb = mw - 2 - 5;                 // core bits above the five that select the bit
    if (wb > S3_MAX_WBITS) wb = S3_MAX_WBITS;
    return wb;
}

static s3_kernel_t s3_pick(int mw, bool shifted, bool raw)
{
    switch (mk_s3_word_bits(mw)) {
    case 15: return s3_pick2<((1u << 15) - 1u) << 2>(shifted, raw);
    case 14: return s3_pick2<((1u << 14) - 1u) << 2>(shifted, raw);
    case 13: return s3_pick2<((1u << 13) - 1u) << 2>(shifted, raw);
    case 9: return s3_pick2<((1u << 9) - 1u) << 2>(shifted, raw);
    case 5: return s3_pick2<((1u << 5) - 1u) << 2>(shifted, raw);
    default: return nullptr;
    }
}

size_t mk_s3_smem_bytes(u32 bitmap_bytes)
{
    return ((bitmap_bytes + 127u) & ~127u) + (size_t)S3_NS * S3_TBUF + sizeof(S3Smem) + 64;
}

// items the front warps may need for a text of nbytes: 48 positions per item plus one per sequence-line
// fragment; sized for ordinary reads (>= ~60 bytes per line), the kernel reports what it really needed
size_t mk_s3_arena_items(size_t nbytes, u32 tile_bytes)
{
    // every tile reserves its upper bound (bytes / 48 + 2 per possible sequence-line fragment); one newline
    // per 32 text bytes is assumed here, the kernel reports what it really needed if that is exceeded
    const size_t n_tiles = (nbytes + tile_bytes - 1) / tile_bytes;
    return n_tiles * (tile_bytes / 48 + 12) + nbytes / 64 + 65536;
}

// Two-plane core filter of the pass set (host side of s3_probe_pairs / s3_second_level).
// q = inner window as it appears in the text, first base in the lowest bits, mw bits.
void mk_s3_filter_add(std::vector<u32> &bitmap, int mw, u64 q)
{
    const int cb = mw - 2;
    const u32 wm = (1u << mk_s3_word_bits(mw)) - 1u;
    const u32 coreA = (u32)(q >> 2);                          // the window's last cb/2 bases
    bitmap[(coreA >> 5) & wm] |= 1u << (coreA & 31u);
    const u32 coreB = (u32)(q & ((1ull << cb) - 1ull));       // its first cb/2 bases
    bitmap[(coreB >> 5) & wm] |= 1u << (((coreB & 31u) + 1u) & 31u);
}

int mk_s3_launch(mk_ctx *ctx, const S3Args &a, bool raw, u32 grid)
{
    const KParams &kp = ctx->kp;
    s3_kernel_t kern = s3_pick(kp.mw, a.shift_d != 0, raw);
    if (!kern) {
        snprintf(ctx->err, sizeof(ctx->err), "unsupported inner substring width subk=%d", kp.subk);
        return MK_ERR_UNSUPPORTED;
    }
    const size_t smem = mk_s3_smem_bytes(a.bitmap_bytes);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, S3_THREADS, smem, ctx->stream>>>(a);
    LAUNCH_COUNT(ctx);
    CK(cudaGetLastError());
    return MK_OK;
}
