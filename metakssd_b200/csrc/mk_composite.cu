// mk_composite.cu — MarkerDB intersection of `metakssd composite`.
//
// Replaces the dictionary build + probe loop of get_species_abundance()
// (/root/reference/command_composite.c:535-566) and the per-species order statistics of
// command_composite.c:598-613.  Only membership matters for the result (the reference's own
// 32-bit wrap-around probe arithmetic is an implementation detail of its dictionary), so the query
// codes go into a device hash and every MarkerDB code is probed by one thread; hits are compacted
// in MarkerDB order (component-major) into a per-context store.  Statistics: one radix sort of
// (species << 16 | count) and one thread per species.
#include "mk_common.cuh"
#include <algorithm>

#define EMPTY64 0xFFFFFFFFFFFFFFFFull

__device__ __forceinline__ u32 mixc(u32 x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

__global__ void __launch_bounds__(256)
k_cq_insert(const u32 *__restrict__ qry, u64 q, u64 *__restrict__ keys, u32 *__restrict__ idxmin, u32 mask)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q) return;
    u64 key = qry[i];
    u32 h = mixc((u32)key) & mask;
    for (;;) {
        u64 old = atomicCAS((unsigned long long *)&keys[h], EMPTY64, key);
        if (old == EMPTY64 || old == key) break;
        h = (h + 1) & mask;
    }
    atomicMin(&idxmin[h], (u32)i); // a duplicated query code resolves to its first occurrence (:537-545)
}

__global__ void __launch_bounds__(256)
k_cq_probe(const u32 *__restrict__ ref, u64 r, const u64 *__restrict__ keys, const u32 *__restrict__ idxmin, u32 mask,
           const uint16_t *__restrict__ qcnt, u32 *__restrict__ flag, u32 *__restrict__ val)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r) return;
    u64 key = ref[i];
    u32 h = mixc((u32)key) & mask;
    u32 f = 0, v = 0;
    for (;;) {
        u64 k = keys[h];
        if (k == EMPTY64) break;
        if (k == key) { f = 1; v = qcnt[idxmin[h]]; break; }
        h = (h + 1) & mask;
    }
    flag[i] = f;
    val[i] = v;
}

__global__ void __launch_bounds__(256)
k_cq_gather(const u32 *__restrict__ flag, const u32 *__restrict__ val, const u32 *__restrict__ pos, u64 r,
            const u64 *__restrict__ ref_index, int n_species, u32 *__restrict__ store_s, u32 *__restrict__ store_c,
            u64 store_base)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r || !flag[i]) return;
    int lo = 0, hi = n_species - 1; // species s with ref_index[s] <= i < ref_index[s+1]
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (ref_index[mid] <= i) lo = mid; else hi = mid - 1;
    }
    u64 o = store_base + pos[i];
    store_s[o] = (u32)lo;
    store_c[o] = val[i];
}

extern "C" int mk_composite_begin(mk_ctx *ctx, int n_species)
{
    if (!ctx || n_species <= 0) return MK_ERR_ARG;
    ctx->comp_species = n_species;
    ctx->comp_nhits = 0;
    ctx->comp_lists_flat.clear();
    ctx->comp_lists_ptr.clear();
    return MK_OK;
}

// room for `extra` more hits in the per-context store, keeping its content
int mk_composite_reserve(mk_ctx *ctx, u64 extra)
{
    u64 need = ctx->comp_nhits + extra;
    Scratch &ss = ctx->sb[SB_C_STORE_S], &sc = ctx->sb[SB_C_STORE_C];
    if (ss.bytes < need * 4 || sc.bytes < need * 4) {
        u32 *ns, *nc;
        size_t bytes = (size_t)(need + need / 2) * 4 + 256;
        CK(cudaMalloc(&ns, bytes));
        CK(cudaMalloc(&nc, bytes));
        if (ctx->comp_nhits) {
            CK(cudaMemcpyAsync(ns, ss.p, ctx->comp_nhits * 4, cudaMemcpyDeviceToDevice, ctx->stream));
            CK(cudaMemcpyAsync(nc, sc.p, ctx->comp_nhits * 4, cudaMemcpyDeviceToDevice, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        }
        if (ss.p) cudaFree(ss.p);
        if (sc.p) cudaFree(sc.p);
        ss.p = ns; ss.bytes = bytes;
        sc.p = nc; sc.bytes = bytes;
    }
    return MK_OK;
}

// One MarkerDB component against the query codes [qry_lo, qry_hi).  The MarkerDB side is either the
// host arrays of the call (uploaded now) or a component made resident by mk_markerdb_load().
static int composite_component(mk_ctx *ctx, const uint32_t *ref_codes, const uint64_t *ref_index, const u32 *res_ref,
                               const u64 *res_index, u64 r, int n_species, const uint32_t *qry_codes,
                               const uint16_t *qry_counts, uint64_t qry_lo, uint64_t qry_hi,
                               const u32 *dev_qry = nullptr, const uint16_t *dev_qcnt = nullptr, bool slice = false)
{
    CK(cudaSetDevice(ctx->device));
    const u64 q = qry_hi - qry_lo;
    // The reference sizes its dictionary with nextPrime((int)(q / 0.6)) (command_composite.c:535).  q == 0 gives
    // a table of 0 slots: both loops run zero times, the component contributes no hits and the run goes on.
    // q == 1 gives 1 slot and HASH() then takes K % (hash_sz - 1) = K % 0: the reference dies with SIGFPE.
    if (q == 0 || r == 0) return MK_OK;
    if (q == 1 && !slice) {     // (a code-range slice of a sharded query with one code is not the reference's one-code query)
        snprintf(ctx->err, sizeof(ctx->err), "composite: query component with exactly one code (the reference divides by zero)");
        return MK_ERR_EMPTY_QUERY;
    }
    if (r >= 0xFFFFFFFFull || q >= 0x7FFFFFFFull) return MK_ERR_UNSUPPORTED;
    cudaEvent_t e0 = ctx->ev2, e1 = ctx->ev3;
    CK(cudaEventRecord(e0, ctx->stream));
    u32 *d_ref, *d_qry, *d_flag, *d_val, *d_pos, *d_idx, *store_s, *store_c;
    u64 *d_index, *d_keys;
    uint16_t *d_qcnt;
    if (res_ref) {
        d_ref = const_cast<u32 *>(res_ref);
        d_index = const_cast<u64 *>(res_index);
    } else {
        CKR(mk_scratch(ctx, SB_C_REF, (size_t)r, &d_ref));
        CKR(mk_scratch(ctx, SB_C_IDX, (size_t)n_species + 1, &d_index));
    }
    CKR(mk_scratch(ctx, SB_C_QRY, (size_t)q, &d_qry));
    CKR(mk_scratch(ctx, SB_C_QCNT, (size_t)q, &d_qcnt));
    CKR(mk_scratch(ctx, SB_C_HITVAL, (size_t)2 * r, &d_flag));
    d_val = d_flag + r;
    CKR(mk_scratch(ctx, SB_C_POS, (size_t)r, &d_pos));
    u64 cap = 1024;
    while (cap < 2 * q) cap <<= 1;
    CKR(mk_scratch(ctx, SB_CQ_KEYS, (size_t)cap, &d_keys));
    CKR(mk_scratch(ctx, SB_CQ_IDX, (size_t)cap, &d_idx));
    // grow the hit store, keeping what earlier components appended
    CKR(mk_composite_reserve(ctx, r));
    store_s = (u32 *)ctx->sb[SB_C_STORE_S].p;
    store_c = (u32 *)ctx->sb[SB_C_STORE_C].p;
    if (!res_ref) {
        CK(cudaMemcpyAsync(d_ref, ref_codes, (size_t)r * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d_index, ref_index, (size_t)(n_species + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        ctx->prof.h2d_bytes += r * 4 + (u64)(n_species + 1) * 8;
    }
    if (dev_qry) {                      // the query is the sketch this context has just produced: already on the device
        d_qry = const_cast<u32 *>(dev_qry) + qry_lo;
        d_qcnt = const_cast<uint16_t *>(dev_qcnt) + qry_lo;
    } else {
        CK(cudaMemcpyAsync(d_qry, qry_codes + qry_lo, (size_t)q * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d_qcnt, qry_counts + qry_lo, (size_t)q * 2, cudaMemcpyHostToDevice, ctx->stream));
        ctx->prof.h2d_bytes += q * 6;
    }
    CK(cudaMemsetAsync(d_keys, 0xFF, (size_t)cap * 8, ctx->stream));
    CK(cudaMemsetAsync(d_idx, 0xFF, (size_t)cap * 4, ctx->stream));
    k_cq_insert<<<(unsigned)((q + 255) / 256), 256, 0, ctx->stream>>>(d_qry, q, d_keys, d_idx, (u32)(cap - 1));
    LAUNCH_COUNT(ctx);
    k_cq_probe<<<(unsigned)((r + 255) / 256), 256, 0, ctx->stream>>>(d_ref, r, d_keys, d_idx, (u32)(cap - 1), d_qcnt,
                                                                    d_flag, d_val);
    LAUNCH_COUNT(ctx);
    u64 nh = 0;
    CKR(mk_exclusive_scan_u32(ctx, d_flag, d_pos, r, &nh));
    k_cq_gather<<<(unsigned)((r + 255) / 256), 256, 0, ctx->stream>>>(d_flag, d_val, d_pos, r, d_index, n_species,
                                                                     store_s, store_c, ctx->comp_nhits);
    LAUNCH_COUNT(ctx);
    CK(cudaGetLastError());
    CK(cudaEventRecord(e1, ctx->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ctx->prof.composite_ms += ms;
    ctx->comp_nhits += nh;
    return MK_OK;
}

extern "C" int mk_composite_component(mk_ctx *ctx, const uint32_t *ref_codes, const uint64_t *ref_index, int n_species,
                                      const uint32_t *qry_codes, const uint16_t *qry_counts, uint64_t qry_lo,
                                      uint64_t qry_hi)
{
    if (!ctx || !ref_index || n_species != ctx->comp_species || qry_hi < qry_lo) return MK_ERR_ARG;
    return composite_component(ctx, ref_codes, ref_index, nullptr, nullptr, ref_index[n_species], n_species, qry_codes,
                               qry_counts, qry_lo, qry_hi);
}

// ---- resident MarkerDB: load once, intersect many samples ------------------------------------------
extern "C" int mk_markerdb_load(mk_ctx *ctx, int component, const uint32_t *ref_codes, const uint64_t *ref_index,
                                int n_species)
{
    if (!ctx || component < 0 || component >= 65536 || !ref_index || n_species <= 0) return MK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if ((size_t)component >= ctx->mdb.size()) ctx->mdb.resize((size_t)component + 1);
    ResidentComponent &m = ctx->mdb[(size_t)component];
    if (m.d_ref) cudaFree(m.d_ref);
    if (m.d_index) cudaFree(m.d_index);
    m = ResidentComponent();
    const u64 r = ref_index[n_species];
    CK(cudaMalloc(&m.d_ref, (size_t)(r ? r : 1) * 4));
    CK(cudaMalloc(&m.d_index, (size_t)(n_species + 1) * 8));
    if (r) CK(cudaMemcpyAsync(m.d_ref, ref_codes, (size_t)r * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(m.d_index, ref_index, (size_t)(n_species + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->prof.h2d_bytes += r * 4 + (u64)(n_species + 1) * 8;
    m.r = r;
    m.n_species = n_species;
    return MK_OK;
}

extern "C" int mk_markerdb_unload(mk_ctx *ctx)
{
    if (!ctx) return MK_ERR_ARG;
    for (ResidentComponent &m : ctx->mdb) {
        if (m.d_ref) cudaFree(m.d_ref);
        if (m.d_index) cudaFree(m.d_index);
    }
    ctx->mdb.clear();
    return MK_OK;
}

extern "C" int mk_composite_component_resident(mk_ctx *ctx, int component, const uint32_t *qry_codes,
                                               const uint16_t *qry_counts, uint64_t qry_lo, uint64_t qry_hi)
{
    if (!ctx || component < 0 || (size_t)component >= ctx->mdb.size() || qry_hi < qry_lo) return MK_ERR_ARG;
    const ResidentComponent &m = ctx->mdb[(size_t)component];
    if (!m.d_index || m.n_species != ctx->comp_species) return MK_ERR_ARG;
    return composite_component(ctx, nullptr, nullptr, m.d_ref, m.d_index, m.r, m.n_species, qry_codes, qry_counts, qry_lo,
                               qry_hi);
}

extern "C" int mk_composite_component_last(mk_ctx *ctx, int component)
{
    if (!ctx || component < 0 || (size_t)component >= ctx->mdb.size()) return MK_ERR_ARG;
    const ResidentComponent &m = ctx->mdb[(size_t)component];
    if (!m.d_index || m.n_species != ctx->comp_species) return MK_ERR_ARG;
    if (!ctx->last_out_code || (size_t)component + 1 >= ctx->last_seg.size()) {
        snprintf(ctx->err, sizeof(ctx->err), "no -A sketch of this context is resident (or it has fewer components)");
        return MK_ERR_ARG;
    }
    return composite_component(ctx, nullptr, nullptr, m.d_ref, m.d_index, m.r, m.n_species, nullptr, nullptr,
                               ctx->last_seg[(size_t)component], ctx->last_seg[(size_t)component + 1], ctx->last_out_code,
                               ctx->last_out_cnt);
}

// resident MarkerDB component against a query that already sits on the device (the merged runs of a code range,
// mk_comm.cu): the same intersection, no upload
int mk_composite_component_dev(mk_ctx *ctx, int component, const u32 *d_qry, const uint16_t *d_qcnt, u64 q)
{
    if (!ctx || component < 0 || (size_t)component >= ctx->mdb.size()) return MK_ERR_ARG;
    const ResidentComponent &m = ctx->mdb[(size_t)component];
    if (!m.d_index || m.n_species != ctx->comp_species) return MK_ERR_ARG;
    return composite_component(ctx, nullptr, nullptr, m.d_ref, m.d_index, m.r, m.n_species, nullptr, nullptr, 0, q, d_qry,
                               d_qcnt, true);
}

// ---- species_coverage lines (host): command_composite.c:582-624 ------------------------------------
// Species ordered by matched k-mers, descending (glibc's qsort is a stable merge sort: ties stay in
// index order), stopping at the first with fewer than 6; the two ratios are float divisions printed
// with %f.  Returns the number of bytes the text needs (excluding the NUL); writes at most cap bytes.
extern "C" size_t mk_format_species_coverage(const char *qry_name, const char *const *ref_names,
                                             const mk_species_stat *stats, int n_species, char *buf, size_t cap)
{
    if (!qry_name || !ref_names || !stats || n_species < 0) return 0;
    std::vector<int> order((size_t)n_species);
    for (int i = 0; i < n_species; i++) order[(size_t)i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return stats[a].n > stats[b].n; });
    size_t need = 0;
    char line[1024];
    for (int i = 0; i < n_species; i++) {
        const mk_species_stat *s = &stats[order[(size_t)i]];
        if (s->n < 6) break;
        int m = snprintf(line, sizeof(line), "%s\t%s\t%d\t%f\t%f\t%d\t%d\n", qry_name, ref_names[order[(size_t)i]], s->n,
                         (float)s->sum / s->n, (float)s->lastsum / s->lastn, s->median, s->max);
        if (m < 0) continue;
        if ((size_t)m >= sizeof(line)) m = (int)sizeof(line) - 1;
        if (buf && need + (size_t)m < cap) memcpy(buf + need, line, (size_t)m);
        need += (size_t)m;
    }
    if (buf && cap) buf[need < cap ? need : cap - 1] = 0;
    return need;
}

__global__ void __launch_bounds__(256)
k_cstat_keys(const u32 *__restrict__ store_s, const u32 *__restrict__ store_c, u64 n, u64 *__restrict__ keys,
             u64 *__restrict__ vals)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        keys[i] = ((u64)store_s[i] << 16) | (store_c[i] & 0xFFFFu);
        vals[i] = i;
    }
}

// element j of the reference's 1-based array: a[0] = n, a[j] = j-th smallest count
__device__ __forceinline__ int ref_elem(const u64 *__restrict__ keys, u64 lo, int n, int j)
{
    return j == 0 ? n : (int)(keys[lo + (u64)j - 1] & 0xFFFFu);
}

// one warp per species (a thread per species walked its whole hit list alone: 1 ms for 5 M hits at L2K11)
__global__ void __launch_bounds__(128)
k_cstat(const u64 *__restrict__ keys, u64 n, int n_species, mk_species_stat *__restrict__ out)
{
    const int s = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const u32 lane = threadIdx.x & 31u;
    if (s >= n_species) return;
    // [lo, hi) = entries of species s in the sorted keys
    u64 a = 0, b = n;
    while (a < b) { u64 m = (a + b) >> 1; if ((keys[m] >> 16) < (u64)s) a = m + 1; else b = m; }
    u64 lo = a;
    b = n;
    while (a < b) { u64 m = (a + b) >> 1; if ((keys[m] >> 16) <= (u64)s) a = m + 1; else b = m; }
    u64 hi = a;
    int cnt = (int)(hi - lo);
    u32 sum = 0;                                          // (32-bit wrap-around like the reference's int sum)
    for (u64 i = lo + lane; i < hi; i += 32) sum += (u32)(keys[i] & 0xFFFFu);
    sum = __reduce_add_sync(0xffffffffu, sum);
    u32 lastsum = 0, lastn = 0;
    int j0 = (int)(cnt * 0.98);
    for (int j = j0 + (int)lane; (double)j <= cnt * 0.99; j += 32) { lastsum += (u32)ref_elem(keys, lo, cnt, j); lastn++; }
    lastsum = __reduce_add_sync(0xffffffffu, lastsum);
    lastn = __reduce_add_sync(0xffffffffu, lastn);
    if (lane == 0) {
        mk_species_stat st;
        st.n = cnt;
        st.sum = (int32_t)sum;
        st.lastsum = (int32_t)lastsum;
        st.lastn = (int)lastn;
        st.median = cnt ? ref_elem(keys, lo, cnt, cnt / 2) : 0;
        st.max = cnt ? ref_elem(keys, lo, cnt, cnt) : 0;
        out[s] = st;
    }
}

extern "C" int mk_composite_stats(mk_ctx *ctx, mk_species_stat *stats)
{
    if (!ctx || !stats || ctx->comp_species <= 0) return MK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const int S = ctx->comp_species;
    const u64 n = ctx->comp_nhits;
    if (n == 0) {
        memset(stats, 0, sizeof(mk_species_stat) * (size_t)S);
        return MK_OK;
    }
    CK(cudaEventRecord(ctx->ev2, ctx->stream));
    u64 *k0, *v0, *k1, *v1;
    mk_species_stat *d_out;
    CKR(mk_scratch(ctx, SB_SORT_K0, (size_t)n, &k0));
    CKR(mk_scratch(ctx, SB_SORT_V0, (size_t)n, &v0));
    CKR(mk_scratch(ctx, SB_SORT_K1, (size_t)n, &k1));
    CKR(mk_scratch(ctx, SB_SORT_V1, (size_t)n, &v1));
    CKR(mk_scratch(ctx, SB_C_STATS, (size_t)S, &d_out));
    k_cstat_keys<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const u32 *)ctx->sb[SB_C_STORE_S].p,
                                                                      (const u32 *)ctx->sb[SB_C_STORE_C].p, n, k0, v0);
    LAUNCH_COUNT(ctx);
    int sb = 0;
    for (u64 v = (u64)(S - 1); v; v >>= 1) sb++;
    u64 *sk = k0, *sv = v0;
    CKR(mk_radix_sort_pairs(ctx, &sk, &sv, k1, v1, n, 0, 16 + sb));
    k_cstat<<<(unsigned)(((u64)S * 32 + 127) / 128), 128, 0, ctx->stream>>>(sk, n, S, d_out);
    LAUNCH_COUNT(ctx);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(stats, d_out, sizeof(mk_species_stat) * (size_t)S, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaEventRecord(ctx->ev3, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3));
    ctx->prof.composite_ms += ms;
    ctx->prof.d2h_bytes += sizeof(mk_species_stat) * (u64)S;
    return MK_OK;
}

extern "C" int mk_composite_hits(mk_ctx *ctx, const int32_t *const **lists)
{
    if (!ctx || !lists || ctx->comp_species <= 0) return MK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const int S = ctx->comp_species;
    const u64 n = ctx->comp_nhits;
    std::vector<u32> hs(n), hc(n);
    if (n) {
        CK(cudaMemcpyAsync(hs.data(), ctx->sb[SB_C_STORE_S].p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(hc.data(), ctx->sb[SB_C_STORE_C].p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->prof.d2h_bytes += n * 8;
    }
    std::vector<u64> cnt((size_t)S + 1, 0);
    for (u64 i = 0; i < n; i++) cnt[hs[i] + 1]++;
    // layout: for species s, [n_s, v_1 .. v_n_s]
    std::vector<u64> start((size_t)S + 1, 0);
    for (int s = 0; s < S; s++) start[s + 1] = start[s] + 1 + cnt[s + 1];
    ctx->comp_lists_flat.assign((size_t)start[S], 0);
    std::vector<u64> fill((size_t)S, 0);
    for (int s = 0; s < S; s++) ctx->comp_lists_flat[start[s]] = (int32_t)cnt[s + 1];
    for (u64 i = 0; i < n; i++) {
        u32 s = hs[i];
        ctx->comp_lists_flat[start[s] + 1 + fill[s]++] = (int32_t)hc[i];
    }
    ctx->comp_lists_ptr.resize((size_t)S);
    for (int s = 0; s < S; s++) ctx->comp_lists_ptr[s] = ctx->comp_lists_flat.data() + start[s];
    *lists = ctx->comp_lists_ptr.data();
    return MK_OK;
}
