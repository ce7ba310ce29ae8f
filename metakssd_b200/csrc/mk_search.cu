// mk_search.cu — `dist -r <ref> <qry>`: shared k-mer counts of every query sketch with every reference sketch
// (SURVEY.md §8(f)4).
//
// Replaces the inverted index of combco2mco() (/root/reference/co2mco.c:12-86: 2^32 row pointers and a 32 GiB
// offset table per component) and the probe loop of mco_cbdco_nobin_dist() (command_dist.c:1031-1046):
//     for every code of query q:  for every reference g that holds the code:  count[q][g]++
// The index is what it encodes — the (code, reference) pairs sorted by code — built with one radix sort per
// component; a query code finds its row by binary search and adds to the count matrix with atomics.  The distance
// table (Jaccard / containment, Mash distance, p-values, confidence intervals) stays host code: it is printf formatting
// of four integers per pair (host/mkssd_main.c, command_dist.c:1531-1680).
#include "mk_common.cuh"

// (code << 32 | reference) for every code of the reference component
__global__ void __launch_bounds__(256)
k_ref_pairs(const u32 *__restrict__ codes, const u64 *__restrict__ index, int n_ref, u64 n, u64 *__restrict__ keys, u64 *__restrict__ vals)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = 0, hi = n_ref;                         // last g with index[g] <= i
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (index[mid] <= i) lo = mid; else hi = mid; }
    keys[i] = ((u64)codes[i] << 32) | (u32)lo;
    vals[i] = 0;
}

__global__ void __launch_bounds__(256)
k_shared_counts(const u64 *__restrict__ pairs, u64 n_pairs, const u32 *__restrict__ qcodes, const u64 *__restrict__ qindex,
                int n_qry, u64 nq, const u32 *__restrict__ qry_ctx_ct, int n_ref, u32 *__restrict__ counts)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    int q = 0, hi = n_qry;
    while (hi - q > 1) { const int mid = (q + hi) >> 1; if (qindex[mid] <= i) q = mid; else hi = mid; }
    if (qry_ctx_ct && qry_ctx_ct[q] == 0) return;                 // command_dist.c:1033
    const u64 key = (u64)qcodes[i] << 32;
    u64 a = 0, b = n_pairs;
    while (a < b) { const u64 m = (a + b) >> 1; if (pairs[m] < key) a = m + 1; else b = m; }
    u32 *row = counts + (u64)q * (u64)n_ref;
    for (; a < n_pairs && (pairs[a] >> 32) == (u64)qcodes[i]; a++) atomicAdd(&row[(u32)pairs[a]], 1u);
}

extern "C" int mk_shared_counts(mk_ctx *ctx, const uint32_t *ref_codes, const uint64_t *ref_index, int n_ref,
                                const uint32_t *qry_codes, const uint64_t *qry_index, int n_qry, const uint32_t *qry_ctx_ct,
                                uint32_t *counts)
{
    if (!ctx || !ref_index || !qry_index || !counts || n_ref <= 0 || n_qry <= 0) return MK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const u64 R = ref_index[n_ref], Q = qry_index[n_qry];
    if ((R && !ref_codes) || (Q && !qry_codes)) return MK_ERR_ARG;
    if (R == 0 || Q == 0) return MK_OK;
    if (R >= 0xFFFFFFF0ull || Q >= 0xFFFFFFF0ull) return MK_ERR_UNSUPPORTED;
    const u64 cells = (u64)n_qry * (u64)n_ref;
    u32 *d_rc, *d_qc, *d_ct, *d_cnt;
    u64 *d_ri, *d_qi, *k0, *v0, *k1, *v1;
    CKR(mk_scratch(ctx, SB_OUT_CODE, (size_t)R, &d_rc));
    CKR(mk_scratch(ctx, SB_R_CNT, (size_t)Q, &d_qc));
    CKR(mk_scratch(ctx, SB_ACC_CNT, (size_t)n_qry + 1, &d_ct));
    CKR(mk_scratch(ctx, SB_IT_CNT, (size_t)cells, &d_cnt));
    CKR(mk_scratch(ctx, SB_IT_POS, (size_t)n_ref + 1, &d_ri));
    CKR(mk_scratch(ctx, SB_IT_CODE, (size_t)n_qry + 1, &d_qi));
    CKR(mk_scratch(ctx, SB_SORT_K0, (size_t)R, &k0));
    CKR(mk_scratch(ctx, SB_SORT_V0, (size_t)R, &v0));
    CKR(mk_scratch(ctx, SB_SORT_K1, (size_t)R, &k1));
    CKR(mk_scratch(ctx, SB_SORT_V1, (size_t)R, &v1));
    CK(cudaMemcpyAsync(d_rc, ref_codes, (size_t)R * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_ri, ref_index, (size_t)(n_ref + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_qc, qry_codes, (size_t)Q * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_qi, qry_index, (size_t)(n_qry + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (qry_ctx_ct) CK(cudaMemcpyAsync(d_ct, qry_ctx_ct, (size_t)n_qry * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(d_cnt, 0, (size_t)cells * 4, ctx->stream));
    ctx->prof.h2d_bytes += (R + Q) * 4 + (u64)(n_ref + n_qry + 2) * 8;
    k_ref_pairs<<<(unsigned)((R + 255) / 256), 256, 0, ctx->stream>>>(d_rc, d_ri, n_ref, R, k0, v0);
    LAUNCH_COUNT(ctx);
    u64 *sk = k0, *sv = v0;
    CKR(mk_radix_sort_pairs(ctx, &sk, &sv, k1, v1, R, 32, 64));           // by code: the rows of the inverted index
    k_shared_counts<<<(unsigned)((Q + 255) / 256), 256, 0, ctx->stream>>>(sk, R, d_qc, d_qi, n_qry, Q, qry_ctx_ct ? d_ct : nullptr,
                                                                         n_ref, d_cnt);
    LAUNCH_COUNT(ctx);
    CK(cudaGetLastError());
    void *stage = nullptr;
    CKR(mk_pinned(ctx, (size_t)cells * 4, &stage));
    CK(cudaMemcpyAsync(stage, d_cnt, (size_t)cells * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->prof.d2h_bytes += cells * 4;
    const u32 *h = (const u32 *)stage;
    for (u64 i = 0; i < cells; i++) counts[i] += h[i];                      // components accumulate (command_dist.c:1024)
    return MK_OK;
}
