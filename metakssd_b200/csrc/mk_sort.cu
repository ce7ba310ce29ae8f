// mk_sort.cu — device-wide primitives written for this library: exclusive scan (u32) and a
// stable LSD radix sort of (u64 key, u64 value) pairs with warp-match ranking.
//
// They serve the "count accumulation / first-occurrence ranking / slot-order" tail of the
// sketching path, which handles ~2^-12 (L3) to 2^-8 (L2) of the input volume; the design goal is
// few launches and coalesced traffic, not peak sort throughput.
#include "mk_common.cuh"

// ------------------------------------------------------------------------------------------------
// exclusive scan
// ------------------------------------------------------------------------------------------------
#define SCAN_THREADS 256
#define SCAN_IPT 16
#define SCAN_CHUNK (SCAN_THREADS * SCAN_IPT)

__device__ __forceinline__ u32 block_excl_scan_256(u32 v, u32 *ws /*[9]*/, u32 *total)
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
        u32 s = lane < 8 ? ws[lane] : 0;
        u32 si = s;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= (u32)o) si += t;
        }
        if (lane < 8) ws[lane] = si - s;
        if (lane == 7) ws[8] = si;
    }
    __syncthreads();
    u32 r = ws[wid] + incl - v;
    *total = ws[8];
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const u32 *__restrict__ in, u64 n, u32 *__restrict__ bsum)
{
    __shared__ u32 ws[9];
    u64 base = (u64)blockIdx.x * SCAN_CHUNK;
    u32 s = 0;
    for (int i = 0; i < SCAN_IPT; i++) {
        u64 idx = base + (u64)i * SCAN_THREADS + threadIdx.x;
        if (idx < n) s += in[idx];
    }
    u32 total;
    block_excl_scan_256(s, ws, &total);
    if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

// single block: in-place exclusive scan of m values, writes grand total to *total.
// Every thread owns one contiguous run of 4 * R4 values (read as 16-byte vectors, v must be 16-byte
// aligned): one pass to sum the run, one block-wide scan of the 1024 sums, one pass to write the
// prefixes — two memory round trips whatever m is (the old sweep-by-sweep version cost one per 8 K values).
__global__ void __launch_bounds__(1024) k_scan_single(u32 *__restrict__ v, u32 m, u64 *__restrict__ total)
{
    __shared__ u32 ws[33];
    const u32 lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const u32 R4 = (m + 4095u) / 4096u;                  // vectors per thread
    const u32 base = threadIdx.x * 4u * R4;
    u32 sum = 0;
    for (u32 j = 0; j < R4; j++) {
        const u32 i = base + 4u * j;
        if (i + 3u < m) {
            const uint4 q = *reinterpret_cast<const uint4 *>(v + i);
            sum += q.x + q.y + q.z + q.w;
        } else {
            for (u32 e = i; e < m && e < i + 4u; e++) sum += v[e];
        }
    }
    u32 incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (u32)o) incl += t;
    }
    if (lane == 31) ws[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        u32 s = ws[lane];
        u32 si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= (u32)o) si += t;
        }
        ws[lane] = si - s;
        if (lane == 31) ws[32] = si;
    }
    __syncthreads();
    u32 off = ws[wid] + incl - sum;
    for (u32 j = 0; j < R4; j++) {
        const u32 i = base + 4u * j;
        if (i + 3u < m) {
            const uint4 q = *reinterpret_cast<const uint4 *>(v + i);
            uint4 o4;
            o4.x = off; o4.y = off + q.x; o4.z = o4.y + q.y; o4.w = o4.z + q.z;
            off = o4.w + q.w;
            *reinterpret_cast<uint4 *>(v + i) = o4;
        } else {
            for (u32 e = i; e < m && e < i + 4u; e++) { const u32 x = v[e]; v[e] = off; off += x; }
        }
    }
    if (threadIdx.x == 0 && total) *total = ws[32];
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_down(const u32 *__restrict__ in, u32 *__restrict__ out, u64 n, const u32 *__restrict__ bsum)
{
    __shared__ u32 ws[9];
    u64 base = (u64)blockIdx.x * SCAN_CHUNK + (u64)threadIdx.x * SCAN_IPT;
    u32 x[SCAN_IPT];
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_IPT; i++) {
        u64 idx = base + i;
        x[i] = idx < n ? in[idx] : 0;
        s += x[i];
    }
    u32 total;
    u32 off = block_excl_scan_256(s, ws, &total) + bsum[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_IPT; i++) {
        u64 idx = base + i;
        if (idx < n) out[idx] = off;
        off += x[i];
    }
}

int mk_exclusive_scan_u32(mk_ctx *ctx, const u32 *d_in, u32 *d_out, u64 n, u64 *total_host)
{
    if (n == 0) {
        if (total_host) *total_host = 0;
        return MK_OK;
    }
    u64 nb = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
    u32 *bsum;
    CKR(mk_scratch(ctx, SB_SCAN_TMP, nb + 8, &bsum));
    u64 *d_total = (u64 *)(bsum + ((nb + 1) & ~(u64)1) + 2); // 8-byte aligned slot after the sums
    k_scan_reduce<<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(d_in, n, bsum);
    LAUNCH_COUNT(ctx);
    k_scan_single<<<1, 1024, 0, ctx->stream>>>(bsum, (u32)nb, d_total);
    LAUNCH_COUNT(ctx);
    k_scan_down<<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(d_in, d_out, n, bsum);
    LAUNCH_COUNT(ctx);
    CK(cudaGetLastError());
    if (total_host) {
        CK(cudaMemcpyAsync(total_host, d_total, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->prof.d2h_bytes += 8;
    }
    return MK_OK;
}

// ------------------------------------------------------------------------------------------------
// radix sort: 8-bit digits, 3 kernels per pass (histogram, digit-major scan, stable scatter)
// ------------------------------------------------------------------------------------------------
#define RS_THREADS 256
#define RS_WARPS 8
#define RS_ROUNDS 8
#define RS_TILE (RS_THREADS * RS_ROUNDS) // 2048 keys per tile

__global__ void __launch_bounds__(RS_THREADS)
k_radix_hist(const u64 *__restrict__ keys, u64 n, int shift, u32 *__restrict__ hist, u32 nblocks, u64 chunk)
{
    __shared__ u32 h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    u64 lo = (u64)blockIdx.x * chunk;
    u64 hi = lo + chunk < n ? lo + chunk : n;
    for (u64 i = lo + threadIdx.x; i < hi; i += RS_THREADS) {
        u32 d = (u32)(keys[i] >> shift) & 255u;
        atomicAdd(&h[d], 1u);
    }
    __syncthreads();
    hist[(u64)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS)
k_radix_scatter(const u64 *__restrict__ keys, const u64 *__restrict__ vals, u64 *__restrict__ keys_out,
                u64 *__restrict__ vals_out, u64 n, int shift, const u32 *__restrict__ hist, u32 nblocks, u64 chunk)
{
    __shared__ u32 cnt[RS_WARPS][256];
    __shared__ u32 gbase[256];
    __shared__ u32 goff[256];
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    gbase[threadIdx.x] = hist[(u64)threadIdx.x * nblocks + blockIdx.x];
    u64 lo = (u64)blockIdx.x * chunk;
    u64 hi = lo + chunk < n ? lo + chunk : n;
    for (u64 tile = lo; tile < hi; tile += RS_TILE) {
        for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&cnt[0][0])[i] = 0;
        __syncthreads();
        u64 k[RS_ROUNDS], v[RS_ROUNDS];
        u32 rank[RS_ROUNDS];
#pragma unroll
        for (int r = 0; r < RS_ROUNDS; r++) {
            u64 i = tile + (u64)w * (32 * RS_ROUNDS) + (u64)r * 32 + lane;
            bool valid = i < hi;
            k[r] = valid ? keys[i] : 0;
            v[r] = valid ? vals[i] : 0;
            u32 d = valid ? ((u32)(k[r] >> shift) & 255u) : 256u;
            u32 peers = __match_any_sync(0xffffffffu, d);
            u32 leader = __ffs(peers) - 1;
            u32 before = __popc(peers & ((1u << lane) - 1u));
            u32 old = 0;
            if (lane == leader && valid) {
                old = cnt[w][d];
                cnt[w][d] = old + __popc(peers);
            }
            old = __shfl_sync(0xffffffffu, old, leader);
            rank[r] = old + before;
            __syncwarp();
        }
        __syncthreads();
        {
            u32 d = threadIdx.x, run = 0;
#pragma unroll
            for (int ww = 0; ww < RS_WARPS; ww++) {
                u32 c = cnt[ww][d];
                cnt[ww][d] = run;
                run += c;
            }
            u32 gb = gbase[d];
            goff[d] = gb;
            gbase[d] = gb + run;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RS_ROUNDS; r++) {
            u64 i = tile + (u64)w * (32 * RS_ROUNDS) + (u64)r * 32 + lane;
            if (i < hi) {
                u32 d = (u32)(k[r] >> shift) & 255u;
                u32 pos = goff[d] + cnt[w][d] + rank[r];
                keys_out[pos] = k[r];
                vals_out[pos] = v[r];
            }
        }
        __syncthreads();
    }
}

// Sorts (keys, vals) by key bits [begin_bit, end_bit). On return *keys / *vals point at the
// buffers holding the sorted data (either the originals or the alternates).
int mk_radix_sort_pairs(mk_ctx *ctx, u64 **keys, u64 **vals, u64 *keys_alt, u64 *vals_alt, u64 n, int begin_bit,
                        int end_bit)
{
    if (n <= 1 || end_bit <= begin_bit) return MK_OK;
    if (n >= 0xFFFFFFFFull) {
        snprintf(ctx->err, sizeof(ctx->err), "radix sort: n=%llu exceeds 32-bit positions", (unsigned long long)n);
        return MK_ERR_UNSUPPORTED;
    }
    u64 tiles = (n + RS_TILE - 1) / RS_TILE;
    u32 maxb = (u32)ctx->sm_count * 4;
    u32 nblocks = (u32)(tiles < maxb ? tiles : maxb);
    u64 chunk = ((tiles + nblocks - 1) / nblocks) * RS_TILE;
    nblocks = (u32)((n + chunk - 1) / chunk);
    u32 *hist;
    CKR(mk_scratch(ctx, SB_HIST, (size_t)256 * nblocks + 8, &hist));
    u64 *kin = *keys, *vin = *vals, *kout = keys_alt, *vout = vals_alt;
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        k_radix_hist<<<nblocks, RS_THREADS, 0, ctx->stream>>>(kin, n, shift, hist, nblocks, chunk);
        LAUNCH_COUNT(ctx);
        k_scan_single<<<1, 1024, 0, ctx->stream>>>(hist, 256 * nblocks, nullptr);
        LAUNCH_COUNT(ctx);
        k_radix_scatter<<<nblocks, RS_THREADS, 0, ctx->stream>>>(kin, vin, kout, vout, n, shift, hist, nblocks, chunk);
        LAUNCH_COUNT(ctx);
        u64 *t;
        t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    CK(cudaGetLastError());
    *keys = kin;
    *vals = vin;
    return MK_OK;
}
