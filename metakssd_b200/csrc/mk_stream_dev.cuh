// mk_stream_dev.cuh — device-side helpers shared by the stream kernels (mk_stream.cu, mk_stream3.cu).
#pragma once
#include "mk_common.cuh"

struct StreamArgs {
    const uint8_t *text;
    u64 nbytes;
    u64 pos_base;
    u64 line_base;
    u32 tile_bytes;
    u32 n_tiles;            // tiles [tile_begin, n_tiles) are processed by this launch
    u32 tile_begin;         // multiple of the ticket group size (k_stream_ws; 0 for a whole-text launch)
    const u64 *line_base_ptr;   // optional: added to line_base (total of the launch before, chunked host path)
    u64 *tile_desc;
    u32 *tile_counter;
    u32 *count_counter;     // tickets of the count-ahead pass (k_stream_ws)
    const u32 *bitmap;
    u32 bitmap_bytes;
    const u64 *ptab;
    u32 two_hash;           // 1: the bitmap is a two-hash Bloom filter (inner window of 22+ bits)
    u64 *cand_code;
    u64 *cand_pos;
    u64 *cand_count;
    u64 cand_cap;
    u32 *flags;
    u64 *total_newlines;
    u64 *trace;             // optional per-warp phase timestamps of CTA 0 (development aid)
    u64 *wd;                // watchdog diagnostics: [site, block, warp, a, b, c, d, e]
    KParams kp;
};

#define TBUF_STRIDE (MK_HALO + MK_MAX_TILE + 96) // keeps the stage buffers 128-byte aligned
#define FLAG_LONG_LINE 2u
#define FLAG_MAYBE_LONG 8u   // some 2 KB chunk holds no newline: the host runs the exact line-length check
#define FLAG_SMEM_BASE 32u   // the dynamic shared memory of k_stream_ws does not start where filter_word<true> assumes
#define FLAG_WATCHDOG 4u     // a wait inside k_stream gave up (diagnostics in StreamArgs::wd)
#define WD_LIMIT (1u << 21)

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
#ifndef MK_WAIT_HINT
#define MK_WAIT_HINT 2000       // try_wait suspend-time hint (ns); small enough that the poll-count watchdog still fires within seconds
#endif
__device__ __forceinline__ u32 mbar_try_wait(u64 *bar, u32 parity)
{
    u32 ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"((u32)MK_WAIT_HINT)
        : "memory");
    return ok;
}
// false = gave up (the caller raises the watchdog flag); every failed try_wait suspends the warp for
// a short hardware-defined time, so the loop is kept to the bare minimum of instructions
__device__ __forceinline__ bool mbar_wait(u64 *bar, u32 parity)
{
    for (u32 n = 0; !mbar_try_wait(bar, parity); n++)
        if (n > WD_LIMIT) return false;
    return true;
}
__device__ __forceinline__ void named_bar_sync(u32 id, u32 nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
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
// same, with an L2 evict_first hint: the text is not needed again once it sits in shared memory
__device__ __forceinline__ void tma_load_1d_last_use(void *dst, const void *src, u32 bytes, u64 *bar)
{
    asm volatile("{\n\t.reg .b64 pol;\n\tcreatepolicy.fractional.L2::evict_first.b64 pol, 1.0;\n\t"
                 "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], pol;\n\t}" ::"r"(
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

// The probe is bound by the integer ALU pipe (shifts, logic); the FMA pipe, which also executes
// IMAD / IMAD.HI, is mostly idle.  Right shifts by a compile-time amount are therefore written as
// "high half of a multiply by 2^(32-s)" so that ptxas places them on the FMA pipe.
#ifndef MK_ALU_SHIFTS
template <int SH>
__device__ __forceinline__ u32 shr_fma(u32 x)
{
    if (SH == 0) return x;
    return __umulhi(x, 1u << ((32 - SH) & 31));
}
#else
template <int SH>
__device__ __forceinline__ u32 shr_fma(u32 x) { return x >> SH; }
#endif

// 16 ASCII bases -> 32 bits, base i at bits [2i, 2i+2) (garbage for non-ACGT bytes, by design)
__device__ __forceinline__ u32 pack16(uint4 v)
{
    // code = ((c >> 1) ^ (c >> 2)) & 3 = ((c ^ (c >> 1)) >> 1) & 3: one shift, one LOP3; the remaining
    // ">> 1" is folded into the gathering multiplier (0x01041040 >> 1)
    u32 x0 = ((v.x ^ shr_fma<1>(v.x)) & 0x06060606u) * 0x00820820u;
    u32 x1 = ((v.y ^ shr_fma<1>(v.y)) & 0x06060606u) * 0x00820820u;
    u32 x2 = ((v.z ^ shr_fma<1>(v.z)) & 0x06060606u) * 0x00820820u;
    u32 x3 = ((v.w ^ shr_fma<1>(v.w)) & 0x06060606u) * 0x00820820u;
    return __byte_perm(__byte_perm(x0, x1, 0x0073), __byte_perm(x2, x3, 0x7300), 0x7610);
}

// window extraction for position J: 32 bits of the packed bases starting at bit offset O; only
// bits [0, NEED) of the result are used, which lets single-word cases run on the FMA pipe.
template <int O, int NEED>
__device__ __forceinline__ u32 take_bits(const u32 (&A)[4])
{
    constexpr int W = O >> 5, SH = O & 31;
    if (SH + NEED <= 32) return shr_fma<SH>(A[W]);
    return __funnelshift_r(A[W], A[W + 1], SH);
}

// A position that passed the shared-memory filter: appended to the global hit list; k_verify turns
// it into a (code, position) candidate or drops it.
__device__ __forceinline__ void emit_hit(const StreamArgs &A, u64 pos)
{
    u64 idx = atomicAdd((unsigned long long *)A.cand_count, 1ull);
    if (idx < A.cand_cap) A.cand_pos[idx] = pos;
}

__device__ __forceinline__ u32 ld_volatile_u32(const u32 *p)
{
    u32 v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mbar_arrive(u64 *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p, u32 bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
