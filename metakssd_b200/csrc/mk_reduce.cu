// mk_reduce.cu — from (code, position) candidates to the reference's on-disk sketch.
//
// Replaces the open-addressing table of mt_shortreads2koc()/fasta2co()
// (/root/reference/iseq2comem.c:701-717, :295-310) and the slot-order dump of
// write_fqkoc2files()/wrt_co2cmpn_use_inn_subctx() (iseq2comem.c:539-553, :637-645):
//
//   1. count accumulation: a global-memory hash (atomicCAS insert, atomicAdd count, atomicMin of
//      the first position) keyed by (file, code);
//   2. first-occurrence ranking: radix sort of the distinct codes by first position — this is the
//      order in which a single-threaded reference run inserts them;
//   3. slot reconstruction: the reference writes codes in ascending slot of a double-hashing table
//      of `hashsize` entries.  Sequential insertion "rank r takes the first slot of its probe
//      sequence not held by a smaller rank" is the unique fix-point of a parallel scheme where every
//      slot keeps the minimum rank that claimed it (atomicMin) and an evicted rank resumes probing
//      from its next step.  The table is kept sparse (slot -> rank map sized by the number of
//      codes), so cost does not depend on hashsize (268 MB / 4.3 GB dense in the reference);
//   4. radix sort by (file, component, slot) and emission of uint32 codes / uint16 counts.
#include "mk_common.cuh"
#include <limits.h>

#define EMPTY64 0xFFFFFFFFFFFFFFFFull
#define EMPTY32 0xFFFFFFFFu

__device__ __forceinline__ u64 mix64(u64 x)
{
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

__global__ void k_fill_u64(u64 *p, u64 n, u64 v)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_fill_u32(u32 *p, u64 n, u32 v)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) p[i] = v;
}

// ---- 1. accumulate -----------------------------------------------------------------------------
// cand_cnt == nullptr: every candidate counts 1 (fresh k-mer occurrences);
// otherwise candidates are partial runs carrying their own (already saturated) count.
__global__ void __launch_bounds__(256)
k_acc_insert(const u64 *__restrict__ cand_code, const u64 *__restrict__ cand_pos, const u32 *__restrict__ cand_cnt,
             u64 n, long long keep_below, const u64 *__restrict__ file_off, int n_files, int TL, int code_bits,
             u64 *__restrict__ keys, u32 *__restrict__ cnt, u64 *__restrict__ minpos, u64 mask, u64 *__restrict__ overflow)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 pos = cand_pos[i];
    if ((long long)pos >= keep_below) return;
    if (cand_code[i] == EMPTY64) return;    // hit rejected by k_verify
    u64 file = 0;
    if (file_off) {
        int lo = 0, hi = n_files - 1; // largest f with file_off[f] <= pos
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (file_off[mid] <= pos) lo = mid; else hi = mid - 1;
        }
        file = (u64)lo;
        if (pos + 1 < file_off[lo] + (u64)TL) return; // k-mer would start before this file's first base
    }
    u64 key = (file << code_bits) | cand_code[i];
    u64 h = mix64(key) & mask;
    for (u32 probes = 0;; probes++) {
        u64 old = atomicCAS((unsigned long long *)&keys[h], EMPTY64, key);
        if (old == EMPTY64 || old == key) break;
        if (probes > 2048u) { *overflow = 1; return; }   // table sized too small for this input: the host retries
        h = (h + 1) & mask;
    }
    u32 add = cand_cnt ? cand_cnt[i] : 1u;
    // counts saturate (the output clamps at 65535, iseq2comem.c:713): once a slot passes 2^31 it is pinned there
    // instead of wrapping after 2^32 occurrences of one code
    if (atomicAdd(&cnt[h], add) > 0x7FFFFFFFu) atomicExch(&cnt[h], 0x80000000u);
    atomicMin((unsigned long long *)&minpos[h], pos);
}

__global__ void __launch_bounds__(256)
k_acc_compact(const u64 *__restrict__ keys, const u32 *__restrict__ cnt, const u64 *__restrict__ minpos, u64 cap,
              int code_bits, int drop_zero_code, u64 *__restrict__ out_key, u32 *__restrict__ out_cnt,
              u64 *__restrict__ out_pos, u64 *__restrict__ out_n)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    bool occ = false;
    u64 key = 0;
    if (i < cap) {
        key = keys[i];
        occ = key != EMPTY64;
        if (occ && drop_zero_code && (key & ((1ull << code_bits) - 1)) == 0) occ = false;
    }
    u32 m = __ballot_sync(0xffffffffu, occ);
    if (!m) return;
    u32 lane = threadIdx.x & 31;
    u64 base = 0;
    if (lane == 0) base = atomicAdd((unsigned long long *)out_n, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (occ) {
        u64 o = base + __popc(m & ((1u << lane) - 1u));
        out_key[o] = key;
        out_cnt[o] = cnt[i];
        out_pos[o] = minpos[i];
    }
}

static inline u64 pow2_at_least(u64 v)
{
    u64 p = 1024;
    while (p < v) p <<= 1;
    return p;
}
static inline int bit_length(u64 v)
{
    int b = 0;
    while (v) { b++; v >>= 1; }
    return b;
}

static int accumulate(mk_ctx *ctx, const u64 *d_code, const u64 *d_pos, const u32 *d_cnt, u64 n, long long keep_below,
                      const u64 *d_file_off, int n_files, int code_bits, bool drop_zero, u64 **d_it_key, u32 **d_it_cnt,
                      u64 **d_it_pos, u64 *n_items)
{
    *n_items = 0;
    // Reads repeat their k-mers many times over: start with a table a quarter of the candidate count
    // (clearing and compacting the table is what this step costs) and fall back to 2n slots if a probe
    // sequence gets long.  Genome batches (file_off) are mostly distinct codes: full size at once.
    u64 cap_full = pow2_at_least(2 * n + 2);
    u64 cap = (d_file_off || d_cnt) ? cap_full : pow2_at_least(n / 4 + 2);   // (runs being merged are mostly distinct too)
    if (cap < (1ull << 16)) cap = cap_full < (1ull << 16) ? cap_full : (1ull << 16);
    if (cap > cap_full) cap = cap_full;
  retry:
    u64 *keys, *minpos, *it_key, *it_pos, *counters;
    u32 *cnt, *it_cnt;
    CKR(mk_scratch(ctx, SB_ACC_KEYS, (size_t)cap, &keys));
    CKR(mk_scratch(ctx, SB_ACC_CNT, (size_t)cap, &cnt));
    CKR(mk_scratch(ctx, SB_ACC_POS, (size_t)cap, &minpos));
    CKR(mk_scratch(ctx, SB_IT_CODE, (size_t)n + 1, &it_key));
    CKR(mk_scratch(ctx, SB_IT_CNT, (size_t)n + 1, &it_cnt));
    CKR(mk_scratch(ctx, SB_IT_POS, (size_t)n + 1, &it_pos));
    CKR(mk_scratch(ctx, SB_COUNTERS, 16, &counters));
    CK(cudaMemsetAsync(keys, 0xFF, (size_t)cap * 8, ctx->stream));
    CK(cudaMemsetAsync(minpos, 0xFF, (size_t)cap * 8, ctx->stream));
    CK(cudaMemsetAsync(cnt, 0, (size_t)cap * 4, ctx->stream));
    CK(cudaMemsetAsync(counters + 4, 0, 16, ctx->stream));
    if (n) {
        k_acc_insert<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_code, d_pos, d_cnt, n, keep_below, d_file_off,
                                                                         n_files, ctx->kp.TL, code_bits, keys, cnt,
                                                                         minpos, cap - 1, counters + 5);
        LAUNCH_COUNT(ctx);
    }
    k_acc_compact<<<(unsigned)((cap + 255) / 256), 256, 0, ctx->stream>>>(keys, cnt, minpos, cap, code_bits,
                                                                        drop_zero ? 1 : 0, it_key, it_cnt, it_pos,
                                                                        counters + 4);
    LAUNCH_COUNT(ctx);
    CK(cudaGetLastError());
    u64 h2[2];
    CK(cudaMemcpyAsync(h2, counters + 4, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->prof.d2h_bytes += 16;
    if (h2[1]) {        // a probe sequence ran past its limit: a candidate was dropped, the result is not usable
        if (cap < cap_full) cap = cap_full;                  // once more with the full-size table
        else if (cap_full < (16 * n + 1024)) cap = cap_full = cap_full * 2;   // (clustered keys: a sparser table)
        else {
            snprintf(ctx->err, sizeof(ctx->err), "count accumulation: probe limit reached with a table of %llu slots for %llu candidates",
                     (unsigned long long)cap, (unsigned long long)n);
            return MK_ERR_NOMEM;
        }
        goto retry;
    }
    *n_items = h2[0];
    *d_it_key = it_key;
    *d_it_cnt = it_cnt;
    *d_it_pos = it_pos;
    return MK_OK;
}

int mk_reduce_candidates(mk_ctx *ctx, const u64 *d_cand_code, const u64 *d_cand_pos, u64 n_cand, long long keep_below,
                         const u64 *d_file_off, int n_files, int code_bits, u64 **d_it_key, u32 **d_it_cnt,
                         u64 **d_it_pos, u64 *n_items)
{
    return accumulate(ctx, d_cand_code, d_cand_pos, nullptr, n_cand, keep_below, d_file_off, n_files, code_bits,
                      d_file_off != nullptr, d_it_key, d_it_cnt, d_it_pos, n_items);
}

// ---- 2./3. ranking and slot reconstruction -----------------------------------------------------
__global__ void __launch_bounds__(256)
k_make_rank_keys(const u64 *__restrict__ it_pos, u64 n, u64 *__restrict__ keys, u64 *__restrict__ vals)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = it_pos[i]; vals[i] = i; }
}

__global__ void __launch_bounds__(256)
k_gather_ranked(const u64 *__restrict__ perm, u64 n, const u64 *__restrict__ it_key, const u32 *__restrict__ it_cnt,
                u64 *__restrict__ r_key, u32 *__restrict__ r_cnt)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        u64 s = perm[i];
        r_key[i] = it_key[s];
        r_cnt[i] = it_cnt[s];
    }
}

__global__ void __launch_bounds__(256) k_count_files(const u64 *__restrict__ it_key, u64 n, int code_bits, u32 *__restrict__ fc)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&fc[it_key[i] >> code_bits], 1u);
}

__device__ __forceinline__ u32 probe_slot(u64 code, u64 i, u64 hs)
{
    // global_basic.h:282-284 — 64-bit arithmetic
    return (u32)((code % hs + i * (1ull + code % (hs - 1ull))) % hs);
}

__global__ void __launch_bounds__(256)
k_slot_assign(const u64 *__restrict__ r_key, u64 n, int code_bits, u64 hs, u64 *__restrict__ slot_keys,
              u32 *__restrict__ slot_vals, u64 smask, u32 *probe_i)
{
    u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const u64 cmask = (code_bits >= 64) ? ~0ull : ((1ull << code_bits) - 1);
    u32 carry = (u32)r;
    u64 i = 0;
    for (;;) {
        if (i >= hs) return;                 // cannot happen below hashlimit; never spin on a full table
        u64 key = r_key[carry];
        u64 file = key >> code_bits, code = key & cmask;
        u32 slot = probe_slot(code, i, hs);
        u64 skey = (file << 32) | slot;
        u64 h = mix64(skey) & smask;
        for (;;) {
            u64 old = atomicCAS((unsigned long long *)&slot_keys[h], EMPTY64, skey);
            if (old == EMPTY64 || old == skey) break;
            h = (h + 1) & smask;
        }
        ((volatile u32 *)probe_i)[carry] = (u32)i;
        __threadfence();
        u32 old = atomicMin(&slot_vals[h], carry);
        if (old == EMPTY32) break;           // free slot: placed
        if (old > carry) {                   // evicted a later rank: it resumes from its next probe step
            __threadfence();
            i = (u64)((volatile u32 *)probe_i)[old] + 1;
            carry = old;
        } else {
            i++;                             // held by an earlier rank
        }
    }
}

__global__ void __launch_bounds__(256)
k_make_out_keys(const u64 *__restrict__ r_key, const u32 *__restrict__ probe_i, u64 n, int code_bits, u64 hs,
                u32 comp_mask, int sbits, int cbits, u64 *__restrict__ keys, u64 *__restrict__ vals)
{
    u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const u64 cmask = (code_bits >= 64) ? ~0ull : ((1ull << code_bits) - 1);
    u64 key = r_key[r];
    u64 file = key >> code_bits, code = key & cmask;
    u32 slot = probe_slot(code, probe_i[r], hs);
    u64 comp = code & comp_mask;
    keys[r] = (file << (sbits + cbits)) | (comp << sbits) | slot;
    vals[r] = r;
}

__global__ void __launch_bounds__(256)
k_emit(const u64 *__restrict__ skeys, const u64 *__restrict__ perm, u64 n, const u64 *__restrict__ r_key,
       const u32 *__restrict__ r_cnt, int code_bits, int comp_code_bits, int sbits, u32 *__restrict__ out_code,
       uint16_t *__restrict__ out_cnt, u64 *__restrict__ seg_start)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 cmask = (code_bits >= 64) ? ~0ull : ((1ull << code_bits) - 1);
    u64 r = perm[i];
    u64 code = r_key[r] & cmask;
    out_code[i] = (u32)(code >> comp_code_bits);
    u32 c = r_cnt[r];
    out_cnt[i] = (uint16_t)(c > 65535u ? 65535u : c);
    u64 seg = skeys[i] >> sbits;
    if (i == 0 || (skeys[i - 1] >> sbits) != seg) seg_start[seg] = i;
}

int mk_order_and_emit(mk_ctx *ctx, u64 *d_it_key, u32 *d_it_cnt, u64 *d_it_pos, u64 n, int n_files, bool with_counts,
                      bool drop_zero_code, mk_sketch *out)
{
    (void)drop_zero_code;
    const mk_info &I = ctx->info;
    const int cn = I.component_num;
    const int code_bits = I.code_bits;
    // initialise outputs (also for n == 0)
    for (int f = 0; f < n_files; f++) {
        mk_sketch *s = &out[f];
        s->n_components = cn;
        s->n_total = 0;
        s->n = (uint64_t *)calloc((size_t)cn, sizeof(uint64_t));
        s->codes = (uint32_t **)calloc((size_t)cn, sizeof(uint32_t *));
        s->counts = with_counts ? (uint16_t **)calloc((size_t)cn, sizeof(uint16_t *)) : nullptr;
        if (!s->n || !s->codes || (with_counts && !s->counts)) return MK_ERR_NOMEM;
    }
    ctx->last_out_code = nullptr;
    ctx->last_out_cnt = nullptr;
    if (n == 0) return MK_OK;
    if (n >= 0xFFFFFFF0ull) return MK_ERR_UNSUPPORTED;
    const u64 nb = (n + 255) / 256;
    // The reference aborts as soon as a table holds more than hashlimit codes (iseq2comem.c:708,
    // :303); check before reconstructing slots (a table with more codes than slots has no layout).
    {
        bool crowded = false;
        if (n_files == 1) crowded = n > I.hashlimit;
        else if (n > I.hashlimit) {
            u32 *fc;
            CKR(mk_scratch(ctx, SB_SEG_COUNTS, (size_t)n_files * 2 + 2, &fc));
            CK(cudaMemsetAsync(fc, 0, (size_t)n_files * 4, ctx->stream));
            k_count_files<<<(unsigned)nb, 256, 0, ctx->stream>>>(d_it_key, n, code_bits, fc);
            LAUNCH_COUNT(ctx);
            std::vector<u32> h((size_t)n_files);
            CK(cudaMemcpyAsync(h.data(), fc, (size_t)n_files * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            for (int f = 0; f < n_files; f++) crowded |= h[f] > I.hashlimit;
        }
        if (crowded) {
            snprintf(ctx->err, sizeof(ctx->err),
                     "the context space is too crowd, try rerun the program using -k%d", I.k + 1);
            return MK_ERR_CROWDED;
        }
    }

    u64 *k0, *v0, *k1, *v1;
    CKR(mk_scratch(ctx, SB_SORT_K0, (size_t)n, &k0));
    CKR(mk_scratch(ctx, SB_SORT_V0, (size_t)n, &v0));
    CKR(mk_scratch(ctx, SB_SORT_K1, (size_t)n, &k1));
    CKR(mk_scratch(ctx, SB_SORT_V1, (size_t)n, &v1));

    // rank by first position
    k_make_rank_keys<<<(unsigned)nb, 256, 0, ctx->stream>>>(d_it_pos, n, k0, v0);
    LAUNCH_COUNT(ctx);
    u64 *sk = k0, *sv = v0;
    CKR(mk_radix_sort_pairs(ctx, &sk, &sv, k1, v1, n, 0, ctx->pos_bits)); // bits of the largest position
    u64 *r_key;
    u32 *r_cnt, *probe_i;
    CKR(mk_scratch(ctx, SB_R_CODE, (size_t)n, &r_key));
    CKR(mk_scratch(ctx, SB_R_CNT, (size_t)n, &r_cnt));
    CKR(mk_scratch(ctx, SB_R_PROBE, (size_t)n, &probe_i));
    k_gather_ranked<<<(unsigned)nb, 256, 0, ctx->stream>>>(sv, n, d_it_key, d_it_cnt, r_key, r_cnt);
    LAUNCH_COUNT(ctx);
    MkPhaseClock pc(ctx->stream);
    pc.mark("    (rank sort done)");

    // sparse slot map
    u64 scap = pow2_at_least(2 * n + 2);
    u64 *slot_keys;
    u32 *slot_vals;
    CKR(mk_scratch(ctx, SB_SLOT_KEYS, (size_t)scap, &slot_keys));
    CKR(mk_scratch(ctx, SB_SLOT_VALS, (size_t)scap, &slot_vals));
    CK(cudaMemsetAsync(slot_keys, 0xFF, (size_t)scap * 8, ctx->stream));
    CK(cudaMemsetAsync(slot_vals, 0xFF, (size_t)scap * 4, ctx->stream));
    k_slot_assign<<<(unsigned)nb, 256, 0, ctx->stream>>>(r_key, n, code_bits, (u64)I.hashsize, slot_keys, slot_vals,
                                                        scap - 1, probe_i);
    LAUNCH_COUNT(ctx);

    // order by (file, component, slot)
    int sbits = bit_length((u64)I.hashsize);
    int cbits = bit_length((u64)cn - 1);
    int fbits = bit_length((u64)(n_files > 1 ? n_files - 1 : 0));
    if (sbits + cbits + fbits > 64) return MK_ERR_UNSUPPORTED;
    k_make_out_keys<<<(unsigned)nb, 256, 0, ctx->stream>>>(r_key, probe_i, n, code_bits, (u64)I.hashsize, (u32)(cn - 1),
                                                          sbits, cbits, k0, v0);
    LAUNCH_COUNT(ctx);
    sk = k0; sv = v0;
    CKR(mk_radix_sort_pairs(ctx, &sk, &sv, k1, v1, n, 0, sbits + cbits + fbits));

    u32 *out_code;
    uint16_t *out_cnt;
    u64 *seg_start;
    const u64 nseg = (u64)n_files << cbits;
    CKR(mk_scratch(ctx, SB_OUT_CODE, (size_t)n, &out_code));
    CKR(mk_scratch(ctx, SB_OUT_CNT, (size_t)n, &out_cnt));
    CKR(mk_scratch(ctx, SB_SEG_COUNTS, (size_t)nseg + 1, &seg_start));
    CK(cudaMemsetAsync(seg_start, 0xFF, (size_t)(nseg + 1) * 8, ctx->stream));
    k_emit<<<(unsigned)nb, 256, 0, ctx->stream>>>(sk, sv, n, r_key, r_cnt, code_bits, I.comp_code_bits, sbits, out_code,
                                                 out_cnt, seg_start);
    LAUNCH_COUNT(ctx);
    CK(cudaGetLastError());

    pc.mark("    slots+sort+emit");
    // results come back through the context's pinned staging block
    const size_t seg_bytes = ((size_t)(nseg + 1) * 8 + 15) & ~(size_t)15, code_bytes = ((size_t)n * 4 + 15) & ~(size_t)15;
    void *stage = nullptr;
    CKR(mk_pinned(ctx, seg_bytes + code_bytes + (size_t)n * 2, &stage));
    u64 *h_seg = (u64 *)stage;
    uint32_t *h_code = (uint32_t *)((char *)stage + seg_bytes);
    uint16_t *h_cnt = (uint16_t *)((char *)stage + seg_bytes + code_bytes);
    CK(cudaMemcpyAsync(h_seg, seg_start, (size_t)nseg * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(h_code, out_code, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    // occurrence filter of the secondary sketchers (`dist -n M` keeps codes seen at least M times, `dist -u` codes seen
    // once): every distinct code took part in the slot reconstruction, like in the reference's table; the filter
    // only decides what is written (write_fqco2file(), wrt_co2cmpn_use_inn_subctx(): iseq2comem.c:611, :640)
    const bool filtered = ctx->emit_lo > 1 || ctx->emit_hi != 0xFFFFFFFFu;
    if (with_counts || filtered) CK(cudaMemcpyAsync(h_cnt, out_cnt, (size_t)n * 2, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->prof.d2h_bytes += nseg * 8 + n * 4 + ((with_counts || filtered) ? n * 2 : 0);
    pc.mark("    d2h");
    // empty segments inherit the start of the next non-empty one
    h_seg[nseg] = n;
    for (long long s = (long long)nseg - 1; s >= 0; s--)
        if (h_seg[s] == EMPTY64) h_seg[s] = h_seg[s + 1];
    // a single-file -A sketch stays usable on the device (mk_composite_component_last)
    ctx->last_out_code = nullptr;
    ctx->last_out_cnt = nullptr;
    if (n_files == 1 && with_counts) {
        ctx->last_out_code = out_code;
        ctx->last_out_cnt = out_cnt;
        ctx->last_seg.assign(h_seg, h_seg + nseg + 1);
    }
    for (int f = 0; f < n_files; f++) {
        mk_sketch *s = &out[f];
        for (int c = 0; c < cn; c++) {
            u64 seg = ((u64)f << cbits) | (u64)c;
            u64 lo = h_seg[seg], hi = h_seg[seg + 1];
            u64 m = hi - lo;
            if (filtered) {        // compact in place (segments are disjoint): keep codes whose count passes
                u64 w = lo;
                for (u64 i = lo; i < hi; i++) {
                    const u32 k = h_cnt[i];
                    if (k >= ctx->emit_lo && k <= ctx->emit_hi) { h_code[w] = h_code[i]; h_cnt[w] = h_cnt[i]; w++; }
                }
                m = w - lo;
            }
            s->n[c] = m;
            s->n_total += m;
            if (ctx->borrow_output) {      // views into the pinned staging block, valid until the next call on this context
                s->borrowed = 1;
                s->codes[c] = h_code + lo;
                if (with_counts) s->counts[c] = h_cnt + lo;
                continue;
            }
            s->codes[c] = (uint32_t *)malloc((size_t)(m ? m : 1) * 4);
            if (!s->codes[c]) return MK_ERR_NOMEM;
            memcpy(s->codes[c], h_code + lo, (size_t)m * 4);
            if (with_counts) {
                s->counts[c] = (uint16_t *)malloc((size_t)(m ? m : 1) * 2);
                if (!s->counts[c]) return MK_ERR_NOMEM;
                memcpy(s->counts[c], h_cnt + lo, (size_t)m * 2);
            }
        }
        if (s->n_total > I.hashlimit) {
            snprintf(ctx->err, sizeof(ctx->err),
                     "the context space is too crowd, try rerun the program using -k%d", I.k + 1);
            return MK_ERR_CROWDED;
        }
    }
    return MK_OK;
}

int mk_finalize_candidates(mk_ctx *ctx, const u64 *d_cand_code, const u64 *d_cand_pos, u64 n_cand, long long keep_below,
                           const u64 *d_file_off, int n_files, bool with_counts, mk_sketch *out)
{
    u64 *it_key, *it_pos, n_items;
    u32 *it_cnt;
    CK(cudaEventRecord(ctx->ev2, ctx->stream));
    MkPhaseClock pc(ctx->stream);
    CKR(mk_reduce_candidates(ctx, d_cand_code, d_cand_pos, n_cand, keep_below, d_file_off, n_files, ctx->info.code_bits,
                             &it_key, &it_cnt, &it_pos, &n_items));
    pc.mark("  accumulate");
    int rc = mk_order_and_emit(ctx, it_key, it_cnt, it_pos, n_items, n_files, with_counts, d_file_off != nullptr, out);
    pc.mark("  order+emit+copy");
    cudaEventRecord(ctx->ev3, ctx->stream);
    cudaEventSynchronize(ctx->ev3);
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3) == cudaSuccess) ctx->prof.reduce_ms += ms;
    return rc;
}

// ---- multi-GPU building blocks -----------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_split_runs(const u64 *__restrict__ perm, const u64 *__restrict__ skeys, u64 n, const u32 *__restrict__ it_cnt,
             const u64 *__restrict__ it_pos, u64 *__restrict__ o_code, u64 *__restrict__ o_pos, u32 *__restrict__ o_cnt)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 s = perm[i];
    o_code[i] = skeys[i];
    o_pos[i] = it_pos[s];
    u32 c = it_cnt[s];
    o_cnt[i] = c > 65535u ? 65535u : c;
}

static int runs_sorted_by_code(mk_ctx *ctx, u64 *it_key, u32 *it_cnt, u64 *it_pos, u64 n, mk_runs *runs)
{
    runs->n = n;
    runs->d_code = nullptr; runs->d_firstpos = nullptr; runs->d_count = nullptr;
    if (n == 0) return MK_OK;
    const u64 nb = (n + 255) / 256;
    u64 *k0, *v0, *k1, *v1, *o_code, *o_pos;
    u32 *o_cnt;
    CKR(mk_scratch(ctx, SB_SORT_K0, (size_t)n, &k0));
    CKR(mk_scratch(ctx, SB_SORT_V0, (size_t)n, &v0));
    CKR(mk_scratch(ctx, SB_SORT_K1, (size_t)n, &k1));
    CKR(mk_scratch(ctx, SB_SORT_V1, (size_t)n, &v1));
    CKR(mk_scratch(ctx, SB_RUN_CODE, (size_t)n, &o_code));
    CKR(mk_scratch(ctx, SB_RUN_POS, (size_t)n, &o_pos));
    CKR(mk_scratch(ctx, SB_RUN_CNT, (size_t)n, &o_cnt));
    k_make_rank_keys<<<(unsigned)nb, 256, 0, ctx->stream>>>(it_key, n, k0, v0);
    LAUNCH_COUNT(ctx);
    u64 *sk = k0, *sv = v0;
    CKR(mk_radix_sort_pairs(ctx, &sk, &sv, k1, v1, n, 0, ctx->info.code_bits));
    k_split_runs<<<(unsigned)nb, 256, 0, ctx->stream>>>(sv, sk, n, it_cnt, it_pos, o_code, o_pos, o_cnt);
    LAUNCH_COUNT(ctx);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    runs->d_code = (const uint64_t *)o_code; runs->d_firstpos = (const uint64_t *)o_pos; runs->d_count = o_cnt;
    return MK_OK;
}

extern "C" int mk_fastq_partial_device(mk_ctx *ctx, const void *d_text, size_t nbytes, uint64_t pos_base,
                                       uint64_t line_base, int is_last, mk_runs *runs)
{
    if (!ctx || !runs) return MK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    u64 *cc = nullptr, *cp = nullptr, n_cand = 0, nl = 0;
    MkPhaseClock pc(ctx->stream);
    CKR(mk_stream_fastq(ctx, (const uint8_t *)d_text, nbytes, pos_base, line_base, false, &cc, &cp, &n_cand, &nl));
    pc.mark("p: stream+verify");
    ctx->last_newlines = nl;
    long long keep_below = LLONG_MAX;
    if (is_last && nbytes) {
        CKR(mk_tail_cut(ctx, (const uint8_t *)d_text, nbytes, &keep_below));
        if (keep_below >= 0) keep_below += (long long)pos_base; else keep_below = (long long)pos_base;
    }
    pc.mark("p: tail cut");
    u64 *it_key, *it_pos, n_items;
    u32 *it_cnt;
    CK(cudaEventRecord(ctx->ev2, ctx->stream));
    CKR(mk_reduce_candidates(ctx, cc, cp, n_cand, keep_below, nullptr, 1, ctx->info.code_bits, &it_key, &it_cnt, &it_pos,
                             &n_items));
    int rc = runs_sorted_by_code(ctx, it_key, it_cnt, it_pos, n_items, runs);
    cudaEventRecord(ctx->ev3, ctx->stream);
    cudaEventSynchronize(ctx->ev3);
    pc.mark("p: reduce -> runs");
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3) == cudaSuccess) ctx->prof.reduce_ms += ms;
    return rc;
}

extern "C" int mk_fastq_partial_host(mk_ctx *ctx, const void *h_text, size_t nbytes, uint64_t pos_base,
                                     uint64_t line_base, int is_last, mk_runs *runs)
{
    if (!ctx || !runs || (!h_text && nbytes)) return MK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    uint8_t *d = nullptr;
    CKR(mk_scratch(ctx, SB_TEXT, nbytes + 256, &d));
    CK(cudaMemsetAsync(d + nbytes, 0, 64, ctx->stream));
    ctx->h_src = (const uint8_t *)h_text;      // consumed by the stream driver (pipelined upload)
    ctx->h_src_all = (const uint8_t *)h_text;
    int rc = mk_fastq_partial_device(ctx, d, nbytes, pos_base, line_base, is_last, runs);
    ctx->h_src = nullptr;
    ctx->h_src_all = nullptr;
    return rc;
}

extern "C" int mk_runs_merge_device(mk_ctx *ctx, const uint64_t *d_code, const uint64_t *d_firstpos,
                                    const uint32_t *d_count, uint64_t n, mk_runs *merged)
{
    if (!ctx || !merged) return MK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    u64 *it_key, *it_pos, n_items;
    u32 *it_cnt;
    CKR(accumulate(ctx, (const u64 *)d_code, (const u64 *)d_firstpos, d_count, n, LLONG_MAX, nullptr, 1,
                   ctx->info.code_bits, false, &it_key, &it_cnt, &it_pos, &n_items));
    return runs_sorted_by_code(ctx, it_key, it_cnt, it_pos, n_items, merged);
}

// largest value of an array (bounds the radix passes of the first-position sort)
__global__ void __launch_bounds__(256) k_max_u64(const u64 *__restrict__ v, u64 n, u64 *__restrict__ out)
{
    u64 m = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) m = v[i] > m ? v[i] : m;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        u64 x = __shfl_xor_sync(0xffffffffu, m, o);
        m = x > m ? x : m;
    }
    if ((threadIdx.x & 31) == 0 && m) atomicMax((unsigned long long *)out, (unsigned long long)m);
}

__global__ void __launch_bounds__(256)
k_widen_runs(const u64 *__restrict__ code, const u64 *__restrict__ pos, const u32 *__restrict__ cnt, u64 n,
             u64 *__restrict__ it_key, u32 *__restrict__ it_cnt, u64 *__restrict__ it_pos)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { it_key[i] = code[i]; it_cnt[i] = cnt[i]; it_pos[i] = pos[i]; }
}

static int runs_finalize(mk_ctx *ctx, const uint64_t *d_code, const uint64_t *d_firstpos, const uint32_t *d_count,
                         uint64_t n, bool distinct, mk_sketch *out);

extern "C" int mk_runs_finalize_device(mk_ctx *ctx, const uint64_t *d_code, const uint64_t *d_firstpos,
                                       const uint32_t *d_count, uint64_t n, mk_sketch *out)
{
    return runs_finalize(ctx, d_code, d_firstpos, d_count, n, false, out);
}

extern "C" int mk_runs_finalize_distinct_device(mk_ctx *ctx, const uint64_t *d_code, const uint64_t *d_firstpos,
                                                const uint32_t *d_count, uint64_t n, mk_sketch *out)
{
    return runs_finalize(ctx, d_code, d_firstpos, d_count, n, true, out);
}

static int runs_finalize(mk_ctx *ctx, const uint64_t *d_code, const uint64_t *d_firstpos, const uint32_t *d_count,
                         uint64_t n, bool distinct, mk_sketch *out)
{
    if (!ctx || !out) return MK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    memset(out, 0, sizeof(*out));
    u64 *it_key, *it_pos, n_items;
    u32 *it_cnt;
    CK(cudaEventRecord(ctx->ev2, ctx->stream));
    if (n) {   // (queued before accumulate's own read-back, which also synchronises this one)
        u64 *d_max;
        CKR(mk_scratch(ctx, SB_MISC, 64, &d_max));
        CK(cudaMemsetAsync(d_max + 4, 0, 8, ctx->stream));
        k_max_u64<<<ctx->sm_count, 256, 0, ctx->stream>>>((const u64 *)d_firstpos, n, d_max + 4);
        LAUNCH_COUNT(ctx);
        CK(cudaMemcpyAsync(&ctx->h_maxpos, d_max + 4, 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (distinct && n) {      // every code occurs once (ranges merged by their owners): nothing to accumulate
        CKR(mk_scratch(ctx, SB_IT_CODE, (size_t)n + 1, &it_key));
        CKR(mk_scratch(ctx, SB_IT_CNT, (size_t)n + 1, &it_cnt));
        CKR(mk_scratch(ctx, SB_IT_POS, (size_t)n + 1, &it_pos));
        k_widen_runs<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const u64 *)d_code, (const u64 *)d_firstpos,
                                                                         d_count, n, it_key, it_cnt, it_pos);
        LAUNCH_COUNT(ctx);
        CK(cudaStreamSynchronize(ctx->stream));       // (h_maxpos has arrived)
        n_items = n;
    } else {
        CKR(accumulate(ctx, (const u64 *)d_code, (const u64 *)d_firstpos, d_count, n, LLONG_MAX, nullptr, 1,
                       ctx->info.code_bits, false, &it_key, &it_cnt, &it_pos, &n_items));
    }
    if (n) {
        int b = 1;
        while (b < 64 && (ctx->h_maxpos >> b)) b++;
        ctx->pos_bits = b;
    }
    int rc = mk_order_and_emit(ctx, it_key, it_cnt, it_pos, n_items, 1, true, false, out);
    ctx->pos_bits = 64;
    cudaEventRecord(ctx->ev3, ctx->stream);
    cudaEventSynchronize(ctx->ev3);
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3) == cudaSuccess) ctx->prof.reduce_ms += ms;
    return rc;
}

// Blocks gathered from the owners of the code ranges (csrc/mk_comm.cu): block p holds h_counts[p] runs at p * cap, codes
// ascending inside a block and across blocks and every code once, so the runs are only packed next to each other
// (no second accumulate) before the slot order is reconstructed.
struct BlockOffsets { u64 off[65]; };

__global__ void __launch_bounds__(256)
k_gather_blocks(BlockOffsets B, int W, u64 cap, const u64 *__restrict__ code, const u64 *__restrict__ pos,
                const u32 *__restrict__ cnt, u64 *__restrict__ it_key, u32 *__restrict__ it_cnt, u64 *__restrict__ it_pos,
                u64 *__restrict__ maxpos)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 m = 0;
    if (i < (u64)W * cap) {
        const u64 p = i / cap, j = i - p * cap;
        if (j < B.off[p + 1] - B.off[p]) {
            const u64 o = B.off[p] + j;
            it_key[o] = code[i];
            it_cnt[o] = cnt[i];
            m = pos[i];
            it_pos[o] = m;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        u64 x = __shfl_xor_sync(0xffffffffu, m, o);
        m = x > m ? x : m;
    }
    if ((threadIdx.x & 31) == 0 && m) atomicMax((unsigned long long *)maxpos, (unsigned long long)m);
}

int mk_runs_finalize_blocks(mk_ctx *ctx, const u64 *d_code, const u64 *d_firstpos, const u32 *d_count, int W, u64 cap,
                            const u64 *h_counts, mk_sketch *out)
{
    if (!ctx || !out || W < 1 || W > 64) return MK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    memset(out, 0, sizeof(*out));
    BlockOffsets B;
    B.off[0] = 0;
    for (int p = 0; p < W; p++) B.off[p + 1] = B.off[p] + (h_counts[p] < cap ? h_counts[p] : cap);
    const u64 n = B.off[W];
    u64 *it_key, *it_pos, *d_max;
    u32 *it_cnt;
    CK(cudaEventRecord(ctx->ev2, ctx->stream));
    CKR(mk_scratch(ctx, SB_IT_CODE, (size_t)n + 1, &it_key));
    CKR(mk_scratch(ctx, SB_IT_CNT, (size_t)n + 1, &it_cnt));
    CKR(mk_scratch(ctx, SB_IT_POS, (size_t)n + 1, &it_pos));
    if (n) {
        CKR(mk_scratch(ctx, SB_MISC, 64, &d_max));
        CK(cudaMemsetAsync(d_max + 4, 0, 8, ctx->stream));
        const u64 tot = (u64)W * cap;
        k_gather_blocks<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(B, W, cap, d_code, d_firstpos, d_count, it_key,
                                                                              it_cnt, it_pos, d_max + 4);
        LAUNCH_COUNT(ctx);
        CK(cudaMemcpyAsync(&ctx->h_maxpos, d_max + 4, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        int b = 1;
        while (b < 64 && (ctx->h_maxpos >> b)) b++;
        ctx->pos_bits = b;
    }
    int rc = mk_order_and_emit(ctx, it_key, it_cnt, it_pos, n, 1, true, false, out);
    ctx->pos_bits = 64;
    cudaEventRecord(ctx->ev3, ctx->stream);
    cudaEventSynchronize(ctx->ev3);
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3) == cudaSuccess) ctx->prof.reduce_ms += ms;
    return rc;
}

// ================================================================================================
// `metakssd set -g / -q / -i` on the device: the MarkerDB build from genome sketches.
//   -g  grouping_genomes()    /root/reference/command_set.c:831-1003   union of a taxon's genome sketches
//   -q  uniq_sketch_union()   command_set.c:427-512                    codes found in exactly one taxon
//   -i  sketch_operate()      command_set.c:322-423 (-s: subtract)     every taxon's codes that are in the pan
// One component per call; the host keeps the file formats and the taxon table (organize_taxf()).
// ================================================================================================
// 32-bit wrap-around probe of grouping_genomes() (HASH() with an unsigned int key and int table size:
// int * unsigned -> unsigned, command_set.c:892; global_basic.h:282-284)
__device__ __forceinline__ u32 probe_slot32(u32 code, u32 i, u32 hs)
{
    return (code % hs + i * (1u + code % (hs - 1u))) % hs;
}

// sequential insertion order of a taxon's table as a parallel fix point (see k_slot_assign): every slot keeps the
// smallest rank that claimed it; here every taxon has its own table size
__global__ void __launch_bounds__(256)
k_set_slot_assign(const u64 *__restrict__ r_key /* taxon << 32 | code, in insertion order */, u64 n,
                  const u32 *__restrict__ hs_taxon, u64 *__restrict__ slot_keys, u32 *__restrict__ slot_vals, u64 smask,
                  u32 *probe_i)
{
    u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    u32 carry = (u32)r;
    u32 i = 0;
    for (u32 guard = 0; guard < 0x7FFFFFFFu; guard++) {
        u64 key = r_key[carry];
        u32 taxon = (u32)(key >> 32), code = (u32)key;
        u32 hs = hs_taxon[taxon];
        if (i >= hs) return;                 // "hashtable overflow" in the reference: the code is not stored
        u32 slot = probe_slot32(code, i, hs);
        u64 skey = ((u64)taxon << 32) | slot;
        u64 h = mix64(skey) & smask;
        for (;;) {
            u64 old = atomicCAS((unsigned long long *)&slot_keys[h], EMPTY64, skey);
            if (old == EMPTY64 || old == skey) break;
            h = (h + 1) & smask;
        }
        ((volatile u32 *)probe_i)[carry] = i;
        __threadfence();
        u32 old = atomicMin(&slot_vals[h], carry);
        if (old == EMPTY32) break;
        if (old > carry) {
            __threadfence();
            i = ((volatile u32 *)probe_i)[old] + 1;
            carry = old;
        } else {
            i++;
        }
    }
}

__global__ void __launch_bounds__(256)
k_set_keys(const u32 *__restrict__ codes, const u32 *__restrict__ taxon_of, u64 n, u64 *__restrict__ keys, u64 *__restrict__ vals)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // code 0 reads as an empty slot in the reference's table and is never stored; taxon 0xFFFFFFFF = ignored genome
    const bool drop = codes[i] == 0 || taxon_of[i] == 0xFFFFFFFFu;
    keys[i] = drop ? EMPTY64 : (((u64)taxon_of[i] << 32) | codes[i]);
    vals[i] = i;
}
// after a stable sort by key: the first entry of every run of equal keys survives, flagged for compaction
__global__ void __launch_bounds__(256) k_set_first_of_run(const u64 *__restrict__ keys, u64 n, u32 *__restrict__ flag)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flag[i] = keys[i] != EMPTY64 && (i == 0 || keys[i - 1] != keys[i]);
}
__global__ void __launch_bounds__(256)
k_set_compact(const u64 *__restrict__ keys, const u64 *__restrict__ vals, const u32 *__restrict__ flag, const u32 *__restrict__ pos,
              u64 n, u64 *__restrict__ o_keys /* original index */, u64 *__restrict__ o_vals /* taxon << 32 | code */)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flag[i]) return;
    o_keys[pos[i]] = vals[i];
    o_vals[pos[i]] = keys[i];
}
__global__ void __launch_bounds__(256)
k_set_out_keys(const u64 *__restrict__ r_key, const u32 *__restrict__ probe_i, const u32 *__restrict__ hs_taxon, u64 n,
               u64 *__restrict__ keys, u64 *__restrict__ vals)
{
    u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    u64 key = r_key[r];
    u32 taxon = (u32)(key >> 32), code = (u32)key;
    u32 hs = hs_taxon[taxon];
    u32 pi = probe_i[r];
    keys[r] = pi >= hs ? EMPTY64 : (((u64)taxon << 32) | probe_slot32(code, pi, hs));
    vals[r] = key;
}
__global__ void __launch_bounds__(256)
k_set_emit(const u64 *__restrict__ skeys, const u64 *__restrict__ svals, u64 n, u32 *__restrict__ out, u32 *__restrict__ per_taxon)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || skeys[i] == EMPTY64) return;
    out[i] = (u32)svals[i];
    atomicAdd(&per_taxon[(u32)(skeys[i] >> 32)], 1u);
}

// hash-table size of a taxon: primer[LOG2((ull)(total * 1.5)) - 7] (command_set.c:877-879)
static u32 set_group_table_size(u64 total_codes)
{
    unsigned long long x = (unsigned long long)((double)total_codes * 1.5);
    int lg = x ? 63 - __builtin_clzll(x) : -1;
    // LOG2(0) in the reference is clz(0) (undefined); an empty taxon stores nothing whatever the size
    int ind = lg > 7 ? lg - 7 : 0;
    if (ind > 24) ind = 24;
    // primer[i] = largest prime below 2^(8+i)
    u64 v = (1ull << (8 + ind)) - 1;
    for (;; v--) {
        bool prime = v >= 2 && (v == 2 || v % 2);
        for (u64 d = 3; prime && d * d <= v; d += 2) prime = v % d != 0;
        if (prime) break;
    }
    return (u32)v;
}

extern "C" int mk_set_group(mk_ctx *ctx, const uint32_t *codes, const uint64_t *index, int n_genomes, const int32_t *taxon_of_genome,
                            int n_taxa, uint32_t **out_codes, uint64_t *out_index)
{
    if (!ctx || !index || !taxon_of_genome || !out_codes || !out_index || n_genomes <= 0 || n_taxa <= 0) return MK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    *out_codes = nullptr;
    for (int t = 0; t <= n_taxa; t++) out_index[t] = 0;
    // insertion order of the reference: taxa in output order, a taxon's genomes in taxfile order, a genome's codes in
    // sketch order.  The host lays the codes out in that order (genomes of one taxon are contiguous).
    std::vector<std::vector<int>> members((size_t)n_taxa);
    for (int g = 0; g < n_genomes; g++)
        if (taxon_of_genome[g] >= 0 && taxon_of_genome[g] < n_taxa) members[(size_t)taxon_of_genome[g]].push_back(g);
    u64 n = 0;
    std::vector<u32> hs((size_t)n_taxa);
    for (int t = 0; t < n_taxa; t++) {
        u64 tot = 0;
        for (int g : members[(size_t)t]) tot += index[g + 1] - index[g];
        hs[(size_t)t] = set_group_table_size(tot);
        n += tot;
    }
    if (n == 0) { *out_codes = (uint32_t *)malloc(4); return *out_codes ? MK_OK : MK_ERR_NOMEM; }
    if (n >= 0xFFFFFFF0ull) return MK_ERR_UNSUPPORTED;
    std::vector<u32> h_codes((size_t)n), h_tax((size_t)n);
    {
        u64 o = 0;
        for (int t = 0; t < n_taxa; t++)
            for (int g : members[(size_t)t])
                for (u64 i = index[g]; i < index[g + 1]; i++, o++) { h_codes[(size_t)o] = codes[i]; h_tax[(size_t)o] = (u32)t; }
    }
    const u64 nb = (n + 255) / 256;
    u32 *d_codes, *d_tax, *d_hs, *flag, *pos, *probe_i, *slot_vals, *d_out, *d_cnt;
    u64 *k0, *v0, *k1, *v1, *slot_keys;
    CKR(mk_scratch(ctx, SB_OUT_CODE, (size_t)n, &d_codes));
    CKR(mk_scratch(ctx, SB_R_CNT, (size_t)n, &d_tax));
    CKR(mk_scratch(ctx, SB_SEG_COUNTS, (size_t)2 * n_taxa + 2, &d_hs));
    d_cnt = d_hs + n_taxa;
    CKR(mk_scratch(ctx, SB_ACC_CNT, (size_t)n, &flag));
    CKR(mk_scratch(ctx, SB_IT_CNT, (size_t)n, &pos));
    CKR(mk_scratch(ctx, SB_R_PROBE, (size_t)n, &probe_i));
    CKR(mk_scratch(ctx, SB_SORT_K0, (size_t)n, &k0));
    CKR(mk_scratch(ctx, SB_SORT_V0, (size_t)n, &v0));
    CKR(mk_scratch(ctx, SB_SORT_K1, (size_t)n, &k1));
    CKR(mk_scratch(ctx, SB_SORT_V1, (size_t)n, &v1));
    CK(cudaMemcpyAsync(d_codes, h_codes.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_tax, h_tax.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_hs, hs.data(), (size_t)n_taxa * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(d_cnt, 0, (size_t)n_taxa * 4, ctx->stream));
    ctx->prof.h2d_bytes += n * 8;
    // 1. first occurrence of every (taxon, code): stable sort by key, heads of runs
    k_set_keys<<<(unsigned)nb, 256, 0, ctx->stream>>>(d_codes, d_tax, n, k0, v0);
    LAUNCH_COUNT(ctx);
    int tbits = bit_length((u64)(n_taxa > 1 ? n_taxa - 1 : 1));
    u64 *sk = k0, *sv = v0;
    CKR(mk_radix_sort_pairs(ctx, &sk, &sv, k1, v1, n, 0, 64));       // (EMPTY keys sort last: full width)
    (void)tbits;
    k_set_first_of_run<<<(unsigned)nb, 256, 0, ctx->stream>>>(sk, n, flag);
    LAUNCH_COUNT(ctx);
    u64 m = 0;
    CKR(mk_exclusive_scan_u32(ctx, flag, pos, n, &m));
    if (m == 0) { *out_codes = (uint32_t *)malloc(4); return *out_codes ? MK_OK : MK_ERR_NOMEM; }
    u64 *ok = sk == k0 ? k1 : k0, *ov = sv == v0 ? v1 : v0;          // the other pair of buffers
    k_set_compact<<<(unsigned)nb, 256, 0, ctx->stream>>>(sk, sv, flag, pos, n, ok, ov);
    LAUNCH_COUNT(ctx);
    // 2. insertion order = original index order
    u64 *rk = ok, *rv = ov;
    CKR(mk_radix_sort_pairs(ctx, &rk, &rv, rk == k0 ? k1 : k0, rv == v0 ? v1 : v0, m, 0, bit_length(n)));
    // rv = taxon << 32 | code in insertion order
    // 3. slot of every code in its taxon's table
    const u64 mb = (m + 255) / 256;
    u64 scap = pow2_at_least(2 * m + 2);
    CKR(mk_scratch(ctx, SB_SLOT_KEYS, (size_t)scap, &slot_keys));
    CKR(mk_scratch(ctx, SB_SLOT_VALS, (size_t)scap, &slot_vals));
    CK(cudaMemsetAsync(slot_keys, 0xFF, (size_t)scap * 8, ctx->stream));
    CK(cudaMemsetAsync(slot_vals, 0xFF, (size_t)scap * 4, ctx->stream));
    CK(cudaMemsetAsync(probe_i, 0xFF, (size_t)m * 4, ctx->stream));
    k_set_slot_assign<<<(unsigned)mb, 256, 0, ctx->stream>>>(rv, m, d_hs, slot_keys, slot_vals, scap - 1, probe_i);
    LAUNCH_COUNT(ctx);
    // 4. order by (taxon, slot)
    u64 *fk = rk, *fv = rk == k0 ? k1 : k0;       // reuse: keys into the buffer rk, values into its twin
    u64 *src_keys = rv;
    // (rk holds original indices, no longer needed; rv must stay readable while the out keys are built)
    u64 *alt_k = (fk == k0) ? k1 : k0;
    (void)alt_k;
    u64 *tmp_vals = (rv == v0) ? v1 : v0;
    k_set_out_keys<<<(unsigned)mb, 256, 0, ctx->stream>>>(src_keys, probe_i, d_hs, m, fk, tmp_vals);
    LAUNCH_COUNT(ctx);
    u64 *ek = fk, *ev = tmp_vals;
    CKR(mk_radix_sort_pairs(ctx, &ek, &ev, fv, rv, m, 0, 64));
    CKR(mk_scratch(ctx, SB_OUT_CODE, (size_t)n, &d_out));
    k_set_emit<<<(unsigned)mb, 256, 0, ctx->stream>>>(ek, ev, m, d_out, d_cnt);
    LAUNCH_COUNT(ctx);
    CK(cudaGetLastError());
    std::vector<u32> cnt((size_t)n_taxa);
    CK(cudaMemcpyAsync(cnt.data(), d_cnt, (size_t)n_taxa * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    u64 total = 0;
    for (int t = 0; t < n_taxa; t++) { out_index[t] = total; total += cnt[(size_t)t]; }
    out_index[n_taxa] = total;
    *out_codes = (uint32_t *)malloc((size_t)(total ? total : 1) * 4);
    if (!*out_codes) return MK_ERR_NOMEM;
    CK(cudaMemcpy(*out_codes, d_out, (size_t)total * 4, cudaMemcpyDeviceToHost));     // (stored codes sort before dropped ones)
    ctx->prof.d2h_bytes += total * 4;
    return MK_OK;
}

// ---- -q: codes that occur exactly once in the pan (ascending, like the reference's bitmap scan) ----------
__global__ void __launch_bounds__(256) k_set_u32_keys(const u32 *__restrict__ codes, u64 n, u64 *__restrict__ keys, u64 *__restrict__ vals)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = codes[i]; vals[i] = i; }
}
__global__ void __launch_bounds__(256) k_set_single(const u64 *__restrict__ keys, u64 n, u32 *__restrict__ flag)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flag[i] = (i == 0 || keys[i - 1] != keys[i]) && (i + 1 == n || keys[i + 1] != keys[i]);
}
__global__ void __launch_bounds__(256)
k_set_take_u32(const u64 *__restrict__ keys, const u32 *__restrict__ flag, const u32 *__restrict__ pos, u64 n, u32 *__restrict__ out)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) out[pos[i]] = (u32)keys[i];
}

extern "C" int mk_set_uniq_union(mk_ctx *ctx, const uint32_t *codes, uint64_t n, uint32_t **out, uint64_t *n_out)
{
    if (!ctx || !out || !n_out || (!codes && n)) return MK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    *n_out = 0;
    *out = (uint32_t *)malloc((size_t)(n ? n : 1) * 4);
    if (!*out) return MK_ERR_NOMEM;
    if (n == 0) return MK_OK;
    if (n >= 0xFFFFFFF0ull) return MK_ERR_UNSUPPORTED;
    const u64 nb = (n + 255) / 256;
    u32 *d_codes, *flag, *pos, *d_out;
    u64 *k0, *v0, *k1, *v1;
    CKR(mk_scratch(ctx, SB_OUT_CODE, (size_t)n, &d_codes));
    CKR(mk_scratch(ctx, SB_ACC_CNT, (size_t)n, &flag));
    CKR(mk_scratch(ctx, SB_IT_CNT, (size_t)n, &pos));
    CKR(mk_scratch(ctx, SB_R_CNT, (size_t)n, &d_out));
    CKR(mk_scratch(ctx, SB_SORT_K0, (size_t)n, &k0));
    CKR(mk_scratch(ctx, SB_SORT_V0, (size_t)n, &v0));
    CKR(mk_scratch(ctx, SB_SORT_K1, (size_t)n, &k1));
    CKR(mk_scratch(ctx, SB_SORT_V1, (size_t)n, &v1));
    CK(cudaMemcpyAsync(d_codes, codes, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    ctx->prof.h2d_bytes += n * 4;
    k_set_u32_keys<<<(unsigned)nb, 256, 0, ctx->stream>>>(d_codes, n, k0, v0);
    LAUNCH_COUNT(ctx);
    u64 *sk = k0, *sv = v0;
    CKR(mk_radix_sort_pairs(ctx, &sk, &sv, k1, v1, n, 0, 32));
    k_set_single<<<(unsigned)nb, 256, 0, ctx->stream>>>(sk, n, flag);
    LAUNCH_COUNT(ctx);
    u64 m = 0;
    CKR(mk_exclusive_scan_u32(ctx, flag, pos, n, &m));
    k_set_take_u32<<<(unsigned)nb, 256, 0, ctx->stream>>>(sk, flag, pos, n, d_out);
    LAUNCH_COUNT(ctx);
    CK(cudaMemcpyAsync(*out, d_out, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->prof.d2h_bytes += m * 4;
    *n_out = m;
    return MK_OK;
}

// ---- -i / -s: keep every sketch's codes that are (intersect = 1) / are not (0) in the pan, order kept -------
__global__ void __launch_bounds__(256)
k_set_member(const u32 *__restrict__ codes, u64 n, const u32 *__restrict__ pan_sorted, u64 n_pan, int intersect, u32 *__restrict__ flag)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 c = codes[i];
    u64 a = 0, b = n_pan;
    while (a < b) { u64 m = (a + b) >> 1; if (pan_sorted[m] < c) a = m + 1; else b = m; }
    const bool in = a < n_pan && pan_sorted[a] == c;
    flag[i] = in == (intersect != 0);
}
__global__ void __launch_bounds__(256)
k_set_take_codes(const u32 *__restrict__ codes, const u32 *__restrict__ flag, const u32 *__restrict__ pos, u64 n, u32 *__restrict__ out)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) out[pos[i]] = codes[i];
}
__global__ void k_set_index(const u32 *__restrict__ pos, const u32 *__restrict__ flag, const u64 *__restrict__ index, int n_sketches, u64 n, u64 *__restrict__ out_index)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > n_sketches) return;
    const u64 i = index[s];
    out_index[s] = i < n ? pos[i] : (n ? pos[n - 1] + flag[n - 1] : 0);
}

__global__ void k_set_narrow(const u64 *__restrict__ k, u64 n, u32 *__restrict__ o);

extern "C" int mk_set_operate(mk_ctx *ctx, const uint32_t *pan, uint64_t n_pan, const uint32_t *codes, const uint64_t *index,
                              int n_sketches, int intersect, uint32_t **out_codes, uint64_t *out_index)
{
    if (!ctx || !index || !out_codes || !out_index || n_sketches <= 0 || (!pan && n_pan)) return MK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const u64 n = index[n_sketches];
    *out_codes = (uint32_t *)malloc((size_t)(n ? n : 1) * 4);
    if (!*out_codes) return MK_ERR_NOMEM;
    for (int s = 0; s <= n_sketches; s++) out_index[s] = 0;
    if (n == 0) return MK_OK;
    if (n >= 0xFFFFFFF0ull || n_pan >= 0xFFFFFFF0ull) return MK_ERR_UNSUPPORTED;
    const u64 nb = (n + 255) / 256;
    u32 *d_codes, *d_pan, *flag, *pos, *d_out;
    u64 *k0, *v0, *k1, *v1, *d_index, *d_oindex;
    CKR(mk_scratch(ctx, SB_OUT_CODE, (size_t)n, &d_codes));
    CKR(mk_scratch(ctx, SB_R_PROBE, (size_t)(n_pan ? n_pan : 1), &d_pan));
    CKR(mk_scratch(ctx, SB_ACC_CNT, (size_t)n, &flag));
    CKR(mk_scratch(ctx, SB_IT_CNT, (size_t)n, &pos));
    CKR(mk_scratch(ctx, SB_R_CNT, (size_t)n, &d_out));
    CKR(mk_scratch(ctx, SB_FILE_OFF, (size_t)2 * (n_sketches + 1), &d_index));
    d_oindex = d_index + n_sketches + 1;
    CK(cudaMemcpyAsync(d_codes, codes, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_index, index, (size_t)(n_sketches + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    ctx->prof.h2d_bytes += n * 4 + n_pan * 4;
    if (n_pan) {      // the pan as a sorted array (uniq_pan files are ascending already; pan files of -g are not)
        CKR(mk_scratch(ctx, SB_SORT_K0, (size_t)n_pan, &k0));
        CKR(mk_scratch(ctx, SB_SORT_V0, (size_t)n_pan, &v0));
        CKR(mk_scratch(ctx, SB_SORT_K1, (size_t)n_pan, &k1));
        CKR(mk_scratch(ctx, SB_SORT_V1, (size_t)n_pan, &v1));
        u32 *tmp;
        CKR(mk_scratch(ctx, SB_X_QCODE, (size_t)n_pan, &tmp));
        CK(cudaMemcpyAsync(tmp, pan, (size_t)n_pan * 4, cudaMemcpyHostToDevice, ctx->stream));
        const u64 pb = (n_pan + 255) / 256;
        k_set_u32_keys<<<(unsigned)pb, 256, 0, ctx->stream>>>(tmp, n_pan, k0, v0);
        LAUNCH_COUNT(ctx);
        u64 *sk = k0, *sv = v0;
        CKR(mk_radix_sort_pairs(ctx, &sk, &sv, k1, v1, n_pan, 0, 32));
        k_set_narrow<<<(unsigned)pb, 256, 0, ctx->stream>>>(sk, n_pan, d_pan);
        LAUNCH_COUNT(ctx);
    }
    k_set_member<<<(unsigned)nb, 256, 0, ctx->stream>>>(d_codes, n, d_pan, n_pan, intersect, flag);
    LAUNCH_COUNT(ctx);
    u64 m = 0;
    CKR(mk_exclusive_scan_u32(ctx, flag, pos, n, &m));
    k_set_take_codes<<<(unsigned)nb, 256, 0, ctx->stream>>>(d_codes, flag, pos, n, d_out);
    LAUNCH_COUNT(ctx);
    k_set_index<<<(n_sketches + 256) / 256, 256, 0, ctx->stream>>>(pos, flag, d_index, n_sketches, n, d_oindex);
    LAUNCH_COUNT(ctx);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(*out_codes, d_out, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(out_index, d_oindex, (size_t)(n_sketches + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->prof.d2h_bytes += m * 4;
    return MK_OK;
}

__global__ void k_set_narrow(const u64 *__restrict__ k, u64 n, u32 *__restrict__ o)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = (u32)k[i];
}

extern "C" void mk_free(void *p) { free(p); }
