// mk_comm.cu — the multi-GPU step inside the library: NCCL exchange of runs by code range, rank-local
// composite against a MarkerDB sharded on the same boundaries (SURVEY.md §8(e), §8(b)·1).
//
// The reference is a single process (no counterpart); what is reproduced is its RESULT on the whole file:
//   * reads are sharded by contiguous record ranges; every rank reduces its shard to runs
//     (code, first global position, count) sorted by code — mk_fastq_partial_device;
//   * the code space [0, 2^code_bits) is cut into `world` ranges that hold the same share of the codes
//     (quantiles of the canonical-k-mer law, range_edge()); ONE grouped ncclSend/ncclRecv step moves every
//     run to the owner of its range.  Blocks have a fixed capacity (`max_runs` per pair) and carry their
//     count in a header, so no size is exchanged beforehand and nothing is read back to the host before
//     the data moves; unused slots hold an empty marker that the merge skips.  A block that does not fit
//     fails the step on every rank together (one 8-byte all-reduce), and says how large it had to be;
//   * the owner merges (counts add up and saturate at 65535, first position = minimum), probes ITS slice
//     of the MarkerDB (mk_markerdb_load_sharded keeps the codes of its range) and sends rank 0 the
//     (species, count) hits — a block whose size rank 0 knows from the load — and the merged runs;
//   * rank 0 packs the gathered blocks (distinct codes, ascending ranges: no second accumulate), reproduces
//     the reference's hash-slot order from them (the one step that needs all codes of a component in one
//     table) and the per-species statistics from the gathered hits.
//
// NCCL is loaded with dlopen("libnccl.so.2") when a communicator is first asked for: the library itself
// has no link-time dependency on it (it loads on a CPU box, and inside a torch process it picks up the
// NCCL torch has already loaded).
#include "mk_common.cuh"
#include <dlfcn.h>
#include <nccl.h>

#define EMPTY64 0xFFFFFFFFFFFFFFFFull

namespace {
struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;

bool load_nccl(char *err, size_t errlen)
{
    if (g_nccl.ok) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.h) break;
    }
    if (!g_nccl.h) {
        snprintf(err, errlen, "dlopen(libnccl.so.2): %s", dlerror());
        return false;
    }
#define SYM(field, name)                                                                     \
    *(void **)(&g_nccl.field) = dlsym(g_nccl.h, name);                                       \
    if (!g_nccl.field) { snprintf(err, errlen, "NCCL symbol %s is missing", name); return false; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv")
    SYM(GetErrorString, "ncclGetErrorString") SYM(AllReduce, "ncclAllReduce")
#undef SYM
    g_nccl.ok = true;
    return true;
}
}   // namespace

#define NK(call)                                                                                       \
    do {                                                                                               \
        ncclResult_t r_ = (call);                                                                      \
        if (r_ != ncclSuccess) {                                                                       \
            snprintf(ctx->err, sizeof(ctx->err), "%s:%d: %s -> %s", __FILE__, __LINE__, #call,         \
                     g_nccl.GetErrorString(r_));                                                       \
            return MK_ERR_CUDA;                                                                        \
        }                                                                                              \
    } while (0)

extern "C" int mk_comm_unique_id(void *id, size_t bytes)
{
    char err[256];
    if (!id || bytes < sizeof(ncclUniqueId)) return MK_ERR_ARG;
    if (!load_nccl(err, sizeof err)) return MK_ERR_UNSUPPORTED;
    ncclUniqueId u;
    if (g_nccl.GetUniqueId(&u) != ncclSuccess) return MK_ERR_CUDA;
    memcpy(id, &u, sizeof u);
    return MK_OK;
}

extern "C" int mk_comm_init(mk_ctx *ctx, const void *id, int rank, int world)
{
    if (!ctx || !id || world < 1 || rank < 0 || rank >= world) return MK_ERR_ARG;
    if (!load_nccl(ctx->err, sizeof ctx->err)) return MK_ERR_UNSUPPORTED;
    CK(cudaSetDevice(ctx->device));
    if (ctx->comm) { g_nccl.CommDestroy((ncclComm_t)ctx->comm); ctx->comm = nullptr; }
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclComm_t c;
    NK(g_nccl.CommInitRank(&c, world, u, rank));
    ctx->comm = c;
    ctx->rank = rank;
    ctx->world = world;
    return MK_OK;
}

extern "C" int mk_comm_destroy(mk_ctx *ctx)
{
    if (!ctx) return MK_ERR_ARG;
    if (ctx->comm && g_nccl.ok) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        g_nccl.CommDestroy((ncclComm_t)ctx->comm);
    }
    ctx->comm = nullptr;
    ctx->world = 1;
    ctx->rank = 0;
    return MK_OK;
}

// First code of rank p's range.  A code leads with the high bases of the canonical k-mer (the smaller of a k-mer and
// its reverse complement), so over the code space [0, 1) the codes of unbiased sequence follow the density 2 (1 - x)
// of the minimum of two uniform values; the ranges are the quantiles of that law,
//     edge_p = 2^code_bits * (1 - sqrt(1 - p / world)),
// which gives every owner the same share of the runs and of the MarkerDB (equal-width ranges gave the first of two
// owners two thirds).  Integer arithmetic only, so every rank and the MarkerDB loader agree on the cut.
static u64 isqrt_u128(unsigned __int128 x)
{
    if (x == 0) return 0;
    u64 r = (u64)sqrtl((long double)x);
    while ((unsigned __int128)r * r > x) r--;
    while ((unsigned __int128)(r + 1) * (r + 1) <= x) r++;
    return r;
}

static inline u64 range_edge(int p, int world, int code_bits)
{
    if (p <= 0) return 0;
    if (p >= world) return ~0ull;
    const unsigned __int128 x = (((unsigned __int128)(unsigned)(world - p)) << (2 * code_bits)) / (unsigned)world;
    u64 r = isqrt_u128(x);
    if ((unsigned __int128)r * r < x) r++;                 // ceiling
    return ((u64)1 << code_bits) - r;
}

struct RangeEdges { u64 e[65]; };

// ---- MarkerDB slice of this rank -------------------------------------------------------------------
extern "C" int mk_markerdb_load_sharded(mk_ctx *ctx, int component, const uint32_t *ref_codes, const uint64_t *ref_index,
                                        int n_species)
{
    if (!ctx || !ref_index || n_species <= 0 || component < 0 || component >= ctx->info.component_num) return MK_ERR_ARG;
    const int W = ctx->world, cb = ctx->info.code_bits, ccb = ctx->info.comp_code_bits;
    if ((size_t)component >= ctx->mdb_shard_sizes.size()) ctx->mdb_shard_sizes.resize((size_t)component + 1);
    std::vector<u64> &sizes = ctx->mdb_shard_sizes[(size_t)component];
    sizes.assign((size_t)W, 0);
    std::vector<u64> edges((size_t)W + 1);
    for (int p = 0; p <= W; p++) edges[(size_t)p] = p == W ? ~0ull : range_edge(p, W, cb);
    const u64 lo = edges[(size_t)ctx->rank], hi = edges[(size_t)ctx->rank + 1];
    std::vector<u32> codes;
    std::vector<u64> index((size_t)n_species + 1, 0);
    for (int s = 0; s < n_species; s++) {
        for (u64 i = ref_index[s]; i < ref_index[s + 1]; i++) {
            const u64 full = ((u64)ref_codes[i] << ccb) | (u64)component;      // code % component_num == component
            int p = (int)(((unsigned __int128)full * (unsigned)W) >> cb);        // a first guess of the owner, then corrected
            if (p >= W) p = W - 1;
            while (p > 0 && full < edges[(size_t)p]) p--;
            while (p + 1 < W && full >= edges[(size_t)p + 1]) p++;
            sizes[(size_t)p]++;
            if (full >= lo && (full < hi || ctx->rank == W - 1)) codes.push_back(ref_codes[i]);
        }
        index[(size_t)s + 1] = codes.size();
    }
    return mk_markerdb_load(ctx, component, codes.data(), (const uint64_t *)index.data(), n_species);
}

// ---- exchange ------------------------------------------------------------------------------------------
// cut[p] = first run whose code belongs to rank p (runs are sorted by code); cut[world] = n
__global__ void k_range_cuts(const u64 *__restrict__ code, u64 n, int world, RangeEdges E, u64 *__restrict__ cut)
{
    int p = threadIdx.x;
    if (p > world) return;
    if (p == world) { cut[p] = n; return; }
    const u64 edge = E.e[p];
    u64 a = 0, b = n;
    while (a < b) { u64 m = (a + b) >> 1; if (code[m] < edge) a = m + 1; else b = m; }
    cut[p] = a;
}

// send layout: hdr[world][2] = (count, overflow), then code / pos / cnt as [world][cap] arrays; slots past a
// block's count hold the empty marker
__global__ void __launch_bounds__(256)
k_pack_blocks(const u64 *__restrict__ code, const u64 *__restrict__ pos, const u32 *__restrict__ cnt, const u64 *__restrict__ cut,
              int world, u64 cap, u64 *__restrict__ hdr, u64 *__restrict__ o_code, u64 *__restrict__ o_pos, u32 *__restrict__ o_cnt)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (u64)world * cap) return;
    const int p = (int)(i / cap);
    const u64 j = i - (u64)p * cap;
    const u64 lo = cut[p], m = cut[p + 1] - lo;
    if (j == 0) { hdr[2 * p] = m < cap ? m : cap; hdr[2 * p + 1] = m > cap ? m : 0ull; }      // (overflow: the size it needed)
    if (j < m) { o_code[i] = code[lo + j]; o_pos[i] = pos[lo + j]; o_cnt[i] = cnt[lo + j]; }
    else { o_code[i] = EMPTY64; o_pos[i] = 0; o_cnt[i] = 0; }
}

// one grouped send/recv step: every rank sends block p of (hdr, code, pos, cnt) to rank p.  `to_root`: only rank 0
// receives (block 0 of every rank), everybody else just sends.
static int exchange_blocks(mk_ctx *ctx, bool to_root, u64 cap, const u64 *s_hdr, const u64 *s_code, const u64 *s_pos,
                           const u32 *s_cnt, u64 *r_hdr, u64 *r_code, u64 *r_pos, u32 *r_cnt)
{
    const int W = ctx->world, me = ctx->rank;
    ncclComm_t comm = (ncclComm_t)ctx->comm;
    NK(g_nccl.GroupStart());
    for (int p = 0; p < W; p++) {
        if (to_root && p != 0) continue;
        NK(g_nccl.Send(s_hdr + 2 * p, 2, ncclUint64, p, comm, ctx->stream));
        NK(g_nccl.Send(s_code + (u64)p * cap, cap, ncclUint64, p, comm, ctx->stream));
        NK(g_nccl.Send(s_pos + (u64)p * cap, cap, ncclUint64, p, comm, ctx->stream));
        NK(g_nccl.Send(s_cnt + (u64)p * cap, cap, ncclUint32, p, comm, ctx->stream));
    }
    if (!to_root || me == 0) {
        for (int p = 0; p < W; p++) {
            NK(g_nccl.Recv(r_hdr + 2 * p, 2, ncclUint64, p, comm, ctx->stream));
            NK(g_nccl.Recv(r_code + (u64)p * cap, cap, ncclUint64, p, comm, ctx->stream));
            NK(g_nccl.Recv(r_pos + (u64)p * cap, cap, ncclUint64, p, comm, ctx->stream));
            NK(g_nccl.Recv(r_cnt + (u64)p * cap, cap, ncclUint32, p, comm, ctx->stream));
        }
    }
    NK(g_nccl.GroupEnd());
    return MK_OK;
}

// merged runs -> query arrays of one component (file code = code >> comp_code_bits, count clamped to 16 bits)
__global__ void __launch_bounds__(256)
k_runs_to_query(const u64 *__restrict__ code, const u32 *__restrict__ cnt, u64 n, int component, u32 comp_mask, int ccb,
                u32 *__restrict__ q_code, uint16_t *__restrict__ q_cnt, u64 *__restrict__ q_n)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    bool take = false;
    u64 c = 0;
    if (i < n) { c = code[i]; take = c != EMPTY64 && (u32)(c & comp_mask) == (u32)component; }
    const u32 m = __ballot_sync(0xffffffffu, take);
    if (!m) return;
    const u32 lane = threadIdx.x & 31;
    u64 base = 0;
    if (lane == 0) base = atomicAdd((unsigned long long *)q_n, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (take) {
        const u64 o = base + __popc(m & ((1u << lane) - 1u));
        q_code[o] = (u32)(c >> ccb);
        const u32 k = cnt[i];
        q_cnt[o] = (uint16_t)(k > 65535u ? 65535u : k);
    }
}

// hits of this rank -> one block for rank 0: [count][ (species << 32 | count) x cap ]
__global__ void __launch_bounds__(256)
k_pack_hits(const u32 *__restrict__ store_s, const u32 *__restrict__ store_c, u64 n, u64 cap, u64 *__restrict__ blk)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) blk[0] = n;
    if (i < cap) blk[1 + i] = i < n ? (((u64)store_s[i] << 32) | store_c[i]) : EMPTY64;
}
__global__ void __launch_bounds__(256)
k_unpack_hits(const u64 *__restrict__ blk, u64 cap, u32 *__restrict__ store_s, u32 *__restrict__ store_c, u64 base)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap || i >= blk[0]) return;
    const u64 v = blk[1 + i];
    store_s[base + i] = (u32)(v >> 32);
    store_c[base + i] = (u32)v;
}

// The sharded step after the rank-local partial sketch.  runs = this rank's runs sorted by code (device).
static int sharded_tail(mk_ctx *ctx, const mk_runs &runs, u64 max_runs, mk_sketch *out, mk_species_stat *stats)
{
    const int W = ctx->world, me = ctx->rank;
    const u64 cap = max_runs;
    if (out) memset(out, 0, sizeof(*out));
    MkPhaseClock pc(ctx->stream);
    CK(cudaEventRecord(ctx->ev2, ctx->stream));
    // ---- 1. runs to their owners --------------------------------------------------------------------
    u64 *cut, *s_hdr, *s_code, *s_pos, *r_hdr, *r_code, *r_pos;
    u32 *s_cnt, *r_cnt;
    CKR(mk_scratch(ctx, SB_X_CUT, (size_t)W + 8, &cut));
    CKR(mk_scratch(ctx, SB_X_SHDR, (size_t)4 * W + 8, &s_hdr));
    r_hdr = s_hdr + 2 * W;
    CKR(mk_scratch(ctx, SB_X_SCODE, (size_t)W * cap, &s_code));
    CKR(mk_scratch(ctx, SB_X_SPOS, (size_t)W * cap, &s_pos));
    CKR(mk_scratch(ctx, SB_X_SCNT, (size_t)W * cap, &s_cnt));
    CKR(mk_scratch(ctx, SB_X_RCODE, (size_t)W * cap, &r_code));
    CKR(mk_scratch(ctx, SB_X_RPOS, (size_t)W * cap, &r_pos));
    CKR(mk_scratch(ctx, SB_X_RCNT, (size_t)W * cap, &r_cnt));
    RangeEdges E;
    for (int p = 0; p <= W; p++) E.e[p] = range_edge(p, W, ctx->info.code_bits);
    k_range_cuts<<<1, W + 1, 0, ctx->stream>>>((const u64 *)runs.d_code, runs.n, W, E, cut);
    LAUNCH_COUNT(ctx);
    const u64 tot = (u64)W * cap;
    k_pack_blocks<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>((const u64 *)runs.d_code, (const u64 *)runs.d_firstpos,
                                                                        runs.d_count, cut, W, cap, s_hdr, s_code, s_pos, s_cnt);
    LAUNCH_COUNT(ctx);
    CK(cudaGetLastError());
    pc.mark("x: cuts + pack");
    if (!ctx->evx[0]) { CK(cudaEventCreate(&ctx->evx[0])); CK(cudaEventCreate(&ctx->evx[1])); }
    CK(cudaEventRecord(ctx->evx[0], ctx->stream));
    CKR(exchange_blocks(ctx, false, cap, s_hdr, s_code, s_pos, s_cnt, r_hdr, r_code, r_pos, r_cnt));
    CK(cudaEventRecord(ctx->evx[1], ctx->stream));
    pc.mark("x: runs exchange");
    // ---- 2. merge on the owner (empty slots are skipped) -------------------------------------------------
    mk_runs merged;
    CKR(mk_runs_merge_device(ctx, (const uint64_t *)r_code, (const uint64_t *)r_pos, r_cnt, tot, &merged));
    u64 h_hdr[2 * 64];
    CK(cudaMemcpyAsync(h_hdr, r_hdr, sizeof(u64) * 2 * (size_t)W, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->prof.d2h_bytes += sizeof(u64) * 2 * (u64)W;
    {
        float ms = 0;       // (the stream has been synchronised: both events are complete)
        if (cudaEventElapsedTime(&ms, ctx->evx[0], ctx->evx[1]) == cudaSuccess) ctx->prof.exchange_wait_ms += ms;
    }
    bool overflow = merged.n > cap;
    ctx->last_block_need = merged.n;
    for (int p = 0; p < W; p++) {
        overflow = overflow || h_hdr[2 * p + 1] != 0;
        const u64 need = h_hdr[2 * p + 1] ? h_hdr[2 * p + 1] : h_hdr[2 * p];
        if (need > ctx->last_block_need) ctx->last_block_need = need;
    }
    pc.mark("x: owner merge");
    // ---- 3. rank-local composite against this rank's MarkerDB slice ---------------------------------------
    const bool with_composite = !ctx->mdb.empty() && !ctx->mdb_shard_sizes.empty();
    u64 *hit_blk = nullptr;
    u64 hit_cap = 0;
    if (with_composite) {
        const int S = ctx->mdb[0].n_species;
        CKR(mk_composite_begin(ctx, S));
        u32 *q_code;
        uint16_t *q_cnt;
        u64 *q_n;
        CKR(mk_scratch(ctx, SB_X_QCODE, (size_t)merged.n + 1, &q_code));
        CKR(mk_scratch(ctx, SB_X_QCNT, (size_t)merged.n + 1, &q_cnt));
        CKR(mk_scratch(ctx, SB_X_QN, 8, &q_n));
        for (int c = 0; c < (int)ctx->mdb.size(); c++) {
            u64 q = merged.n;             // one component: every merged run is a query code
            CK(cudaMemsetAsync(q_n, 0, 8, ctx->stream));
            if (merged.n) {
                k_runs_to_query<<<(unsigned)((merged.n + 255) / 256), 256, 0, ctx->stream>>>(
                    (const u64 *)merged.d_code, merged.d_count, merged.n, c, (u32)ctx->info.component_num - 1u,
                    ctx->info.comp_code_bits, q_code, q_cnt, q_n);
                LAUNCH_COUNT(ctx);
            }
            if (ctx->info.component_num > 1) {
                CK(cudaMemcpyAsync(&q, q_n, 8, cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaStreamSynchronize(ctx->stream));
            }
            CKR(mk_composite_component_dev(ctx, c, q_code, q_cnt, q));
        }
        // hits -> rank 0, in a block sized by the slice (rank 0 knows every slice size from the load)
        for (int p = 0; p < W; p++) {
            u64 sz = 0;
            for (const auto &v : ctx->mdb_shard_sizes) sz += v[(size_t)p];
            if (p == me || me == 0) hit_cap = sz > hit_cap ? sz : hit_cap;
        }
        u64 my_cap = 0;
        for (const auto &v : ctx->mdb_shard_sizes) my_cap += v[(size_t)me];
        CKR(mk_scratch(ctx, SB_X_HITS, (size_t)(my_cap + 1) + (me == 0 ? (size_t)W * (hit_cap + 1) : 0), &hit_blk));
        if (me != 0) {
            k_pack_hits<<<(unsigned)((my_cap + 256) / 256), 256, 0, ctx->stream>>>((const u32 *)ctx->sb[SB_C_STORE_S].p,
                                                                                 (const u32 *)ctx->sb[SB_C_STORE_C].p,
                                                                                 ctx->comp_nhits, my_cap, hit_blk);
            LAUNCH_COUNT(ctx);
        }
    }
    pc.mark("x: slice composite");
    // ---- 4. merged runs and hits to rank 0 -----------------------------------------------------------------
    {
        // block 0 of the send arrays <- this rank's merged runs (padded)
        u64 *one_cut;
        CKR(mk_scratch(ctx, SB_X_CUT, (size_t)W + 8, &one_cut));
        u64 h_cut[2] = {0, merged.n};             // (more than cap: the block's header carries the overflow flag)
        CK(cudaMemcpyAsync(one_cut, h_cut, 16, cudaMemcpyHostToDevice, ctx->stream));
        k_pack_blocks<<<(unsigned)((cap + 255) / 256), 256, 0, ctx->stream>>>((const u64 *)merged.d_code, (const u64 *)merged.d_firstpos,
                                                                            merged.d_count, one_cut, 1, cap, s_hdr, s_code, s_pos, s_cnt);
        LAUNCH_COUNT(ctx);
        CKR(exchange_blocks(ctx, true, cap, s_hdr, s_code, s_pos, s_cnt, r_hdr, r_code, r_pos, r_cnt));
        // a block that did not fit anywhere fails the step on EVERY rank (one more reduction of a single word)
        ctx->h_xflag = overflow ? 1 : 0;
        CK(cudaMemcpyAsync(one_cut + 4, &ctx->h_xflag, 8, cudaMemcpyHostToDevice, ctx->stream));
        NK(g_nccl.AllReduce(one_cut + 4, one_cut + 4, 1, ncclUint64, ncclMax, (ncclComm_t)ctx->comm, ctx->stream));
        if (with_composite) {
            ncclComm_t comm = (ncclComm_t)ctx->comm;
            NK(g_nccl.GroupStart());
            if (me != 0) {
                u64 my_cap = 0;
                for (const auto &v : ctx->mdb_shard_sizes) my_cap += v[(size_t)me];
                NK(g_nccl.Send(hit_blk, my_cap + 1, ncclUint64, 0, comm, ctx->stream));
            } else {
                u64 my_cap = 0;
                for (const auto &v : ctx->mdb_shard_sizes) my_cap += v[0];
                for (int p = 1; p < W; p++) {
                    u64 sz = 0;
                    for (const auto &v : ctx->mdb_shard_sizes) sz += v[(size_t)p];
                    NK(g_nccl.Recv(hit_blk + (my_cap + 1) + (u64)p * (hit_cap + 1), sz + 1, ncclUint64, p, comm, ctx->stream));
                }
            }
            NK(g_nccl.GroupEnd());
        }
    }
    {
        u64 *flag;
        CKR(mk_scratch(ctx, SB_X_CUT, (size_t)W + 8, &flag));
        CK(cudaMemcpyAsync(&ctx->h_xflag, flag + 4, 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    cudaEventRecord(ctx->ev3, ctx->stream);
    pc.mark("x: gather to rank 0");
    if (me != 0) {
        CK(cudaStreamSynchronize(ctx->stream));
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3) == cudaSuccess) ctx->prof.exchange_ms += ms;
        if (overflow || ctx->h_xflag) {
            snprintf(ctx->err, sizeof(ctx->err), "sharded step: more than max_runs = %llu runs for one code range", (unsigned long long)cap);
            return MK_ERR_NOMEM;
        }
        return MK_OK;
    }
    // ---- 5. rank 0: hits of all slices -> statistics; merged ranges -> slot order ----------------------------
    CK(cudaMemcpyAsync(h_hdr, r_hdr, sizeof(u64) * 2 * (size_t)W, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3) == cudaSuccess) ctx->prof.exchange_ms += ms;
    }
    for (int p = 0; p < W; p++) overflow = overflow || h_hdr[2 * p + 1] != 0;
    if (overflow || ctx->h_xflag) {
        snprintf(ctx->err, sizeof(ctx->err), "sharded step: more than max_runs = %llu runs for one code range", (unsigned long long)cap);
        return MK_ERR_NOMEM;
    }
    if (with_composite) {
        u64 my_cap = 0;
        for (const auto &v : ctx->mdb_shard_sizes) my_cap += v[0];
        u64 extra = 0;
        for (int p = 1; p < W; p++)
            for (const auto &v : ctx->mdb_shard_sizes) extra += v[(size_t)p];
        CKR(mk_composite_reserve(ctx, extra));
        // counts of the received blocks (device) -> host, to place them one after another
        std::vector<u64> hn((size_t)W, 0);
        for (int p = 1; p < W; p++)
            CK(cudaMemcpyAsync(&hn[(size_t)p], hit_blk + (my_cap + 1) + (u64)p * (hit_cap + 1), 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (int p = 1; p < W; p++) {
            u64 sz = 0;
            for (const auto &v : ctx->mdb_shard_sizes) sz += v[(size_t)p];
            if (hn[(size_t)p]) {
                k_unpack_hits<<<(unsigned)((sz + 255) / 256), 256, 0, ctx->stream>>>(hit_blk + (my_cap + 1) + (u64)p * (hit_cap + 1), sz,
                                                                                   (u32 *)ctx->sb[SB_C_STORE_S].p,
                                                                                   (u32 *)ctx->sb[SB_C_STORE_C].p, ctx->comp_nhits);
                LAUNCH_COUNT(ctx);
                ctx->comp_nhits += hn[(size_t)p];
            }
        }
        if (stats) CKR(mk_composite_stats(ctx, stats));
    }
    pc.mark("x: hits -> statistics");
    if (out) {       // the gathered blocks hold disjoint ascending code ranges, every code once
        u64 h_n[64];
        for (int p = 0; p < W; p++) h_n[p] = h_hdr[2 * p];
        return mk_runs_finalize_blocks(ctx, r_code, r_pos, r_cnt, W, cap, h_n, out);
    }
    return MK_OK;
}

extern "C" int mk_comm_last_block_need(mk_ctx *ctx, uint64_t *need)
{
    if (!ctx || !need) return MK_ERR_ARG;
    *need = ctx->last_block_need;
    return MK_OK;
}

extern "C" int mk_fastq_koc_sharded_device(mk_ctx *ctx, const void *d_text, size_t nbytes, uint64_t pos_base, uint64_t line_base,
                                           int is_last, uint64_t max_runs, mk_sketch *out, mk_species_stat *stats)
{
    if (!ctx || !ctx->comm || ctx->world > 64 || max_runs == 0) return MK_ERR_ARG;
    mk_runs runs;
    CKR(mk_fastq_partial_device(ctx, d_text, nbytes, pos_base, line_base, is_last, &runs));
    return sharded_tail(ctx, runs, max_runs, ctx->rank == 0 ? out : nullptr, ctx->rank == 0 ? stats : nullptr);
}

extern "C" int mk_fastq_koc_sharded_host(mk_ctx *ctx, const void *h_text, size_t nbytes, uint64_t pos_base, uint64_t line_base,
                                         int is_last, uint64_t max_runs, mk_sketch *out, mk_species_stat *stats)
{
    if (!ctx || !ctx->comm || ctx->world > 64 || max_runs == 0) return MK_ERR_ARG;
    mk_runs runs;
    CKR(mk_fastq_partial_host(ctx, h_text, nbytes, pos_base, line_base, is_last, &runs));
    return sharded_tail(ctx, runs, max_runs, ctx->rank == 0 ? out : nullptr, ctx->rank == 0 ? stats : nullptr);
}
