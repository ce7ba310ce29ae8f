// mk_api.cu — extern "C" entry points of libmkssd_b200.so (see include/mkssd_b200.h).
#include "mk_common.cuh"
#include "mk_stream3.cuh"
#include <limits.h>
#include <errno.h>
#include <new>

// hash-table sizes of /root/reference/global_basic.c:75-82: the largest prime below 2^(8+i).
static bool is_prime64(u64 n)
{
    if (n < 2) return false;
    if (n % 2 == 0) return n == 2;
    for (u64 d = 3; d * d <= n; d += 2)
        if (n % d == 0) return false;
    return true;
}
static u32 prime_below_pow2(int e)
{
    u64 n = (1ull << e) - 1;
    while (!is_prime64(n)) n--;
    return (u32)n;
}

extern "C" const char *mk_strerror(int code)
{
    switch (code) {
    case MK_OK: return "ok";
    case MK_ERR_ARG: return "invalid argument";
    case MK_ERR_PARAM: return "k / dimension-reduction level outside the hash-size table (get_hashsz)";
    case MK_ERR_CUDA: return "CUDA error (no CPU fallback exists)";
    case MK_ERR_NOMEM: return "out of memory";
    case MK_ERR_CROWDED: return "the context space is too crowd";
    case MK_ERR_LONG_LINE: return "FASTQ line of 4095 bytes or more";
    case MK_ERR_IO: return "I/O error";
    case MK_ERR_EMPTY_QUERY: return "composite query component with exactly one code";
    case MK_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown error";
    }
}

extern "C" const char *mk_last_error(const mk_ctx *ctx) { return ctx ? ctx->err : "no context"; }

extern "C" int mk_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static u64 pair_reverse(u64 v, int pairs)
{
    u64 r = 0;
    for (int i = 0; i < pairs; i++) r |= ((v >> (2 * i)) & 3ull) << (2 * (pairs - 1 - i));
    return r;
}

extern "C" int mk_ctx_create(mk_ctx **out, const int32_t *shuf_perm, int k, int subk, int drlevel, int device)
{
    if (!out) return MK_ERR_ARG;
    *out = nullptr;
    if (k < 2 || k > 16 || subk < 1 || subk > 7 || subk > k || drlevel < 0 || drlevel > subk) return MK_ERR_PARAM;
    int primer_ind = 4 * (k - drlevel) - 8 - 7; // CTX_SPC_USE_L = 8
    if (primer_ind < 0 || primer_ind > 24) return MK_ERR_PARAM;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        return MK_ERR_CUDA;
    }
    mk_ctx *ctx = new (std::nothrow) mk_ctx();
    if (!ctx) return MK_ERR_NOMEM;
    ctx->err[0] = 0;
    memset(&ctx->prof, 0, sizeof(ctx->prof));
    ctx->device = device;
    auto fail = [&](int code) {
        mk_ctx_destroy(ctx);
        return code;
    };
    if (cudaSetDevice(device) != cudaSuccess) return fail(MK_ERR_CUDA);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(MK_ERR_CUDA);
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(MK_ERR_CUDA);
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) return fail(MK_ERR_CUDA);
    cudaEventCreate(&ctx->ev0); cudaEventCreate(&ctx->ev1); cudaEventCreate(&ctx->ev2); cudaEventCreate(&ctx->ev3);
    cudaEventCreateWithFlags(&ctx->copy_ev[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->copy_ev[1], cudaEventDisableTiming);

    mk_info &I = ctx->info;
    I.k = k; I.subk = subk; I.drlevel = drlevel;
    I.kmer_len = 2 * k;
    I.outctx = k - subk;
    u64 subspace = 1ull << (4 * (subk - drlevel));
    I.dim_end = (int)(subspace > 4096 ? subspace : 4096);
    I.hashsize = prime_below_pow2(primer_ind + 8);
    I.hashlimit = (u32)(I.hashsize * 0.6);
    I.component_num = (k - drlevel > 8) ? (int)(1ul << (4 * (k - drlevel - 8))) : 1;
    I.comp_code_bits = (k - drlevel > 8) ? 4 * (k - drlevel - 8) : 0;
    I.code_bits = 4 * (k - drlevel);
    I.device = device;
    I.sm_count = ctx->sm_count;

    KParams &K = ctx->kp;
    K.k = k; K.subk = subk; K.drlevel = drlevel; K.outctx = I.outctx; K.TL = 2 * k;
    K.crvs_shift = 4 * k - 2;
    K.dim_end = I.dim_end;
    K.tupmask = (4 * k >= 64) ? ~0ull : ((1ull << (4 * k)) - 1);
    K.domask = ((1ull << (4 * subk)) - 1) << (2 * K.outctx);
    K.undomask = ((1ull << (2 * K.outctx)) - 1) << (2 * (k + subk));
    K.lowmask = (1ull << (2 * K.outctx)) - 1;
    K.code_shift = 2 * K.TL - 4 * K.outctx;
    K.hashsize = I.hashsize;
    K.mw = 4 * subk;
    if (K.mw != 12 && K.mw != 16 && K.mw < 20) {
        snprintf(ctx->err, sizeof(ctx->err), "subk=%d not supported (need subk >= 3)", subk);
        return fail(MK_ERR_UNSUPPORTED);
    }
    K.pre = k + subk - 1;
    K.spare = K.mw < 22 ? 1 : 0;
    K.prew = (K.pre + K.spare + 15) / 16;
    if (K.prew < 1) K.prew = 1;
    if (K.prew > 2) return fail(MK_ERR_UNSUPPORTED);
    K.shift_s = 2 * (16 * K.prew - K.pre - K.spare);

    if (!shuf_perm) {          // composite-only context: no pass-set tables (sketching calls return MK_ERR_ARG)
        ctx->no_tables = true;
        *out = ctx;
        return MK_OK;
    }
    // pass set -> exact table (dim -> pf) and the probe bitmap over S ∪ revcomp(S)
    const u64 ndim = 1ull << (4 * subk);
    const u64 mwmask = ndim - 1;
    u64 npass = 0;
    for (u64 d = 0; d < ndim; d++) {
        int32_t pf = shuf_perm[d];
        if (pf >= 0 && pf < I.dim_end) npass++;
    }
    u64 tcap = 1024;
    while (tcap < 4 * npass) tcap <<= 1;
    K.ptab_mask = (u32)(tcap - 1);
    std::vector<u64> ptab(tcap, 0);
    u32 bm_words;
    if (K.mw >= 22) bm_words = 1u << MK_BLOOM_WBITS; // two-hash Bloom filter
    else bm_words = 1u << (K.mw - 5);               // exact bitmap (128 KB at mw = 20)
    if (bm_words < 4) bm_words = 4;
    std::vector<u32> bitmap(bm_words, 0);
    std::vector<u32> bitmap3((size_t)1 << mk_s3_word_bits(K.mw), 0);
    // mw >= 22: two-hash Bloom filter in 2^(MK_BLOOM_WBITS+5) bits; mw <= 20: exact bitmap.  Mirrors probe_block()/second_hash_hit().
    auto set_bit = [&](u64 q) {
        u32 word, bit;
        if (K.mw >= 22) {
            const u32 wm = (1u << MK_BLOOM_WBITS) - 1u;
            word = (u32)(q >> 2) & wm; bit = (u32)(q >> (MK_BLOOM_WBITS + 2)) & 31u;
            bitmap[word] |= 1u << (31 - bit);
            word = (u32)(q >> 6) & wm;
            bit = ((u32)(q >> 19) ^ (u32)q) & 31u;
            bitmap[word] |= 1u << bit;                     // (natural bit order for the second hash)
        } else {
            word = (u32)(q & ((1ull << (K.mw - 5)) - 1)); bit = (u32)(q >> (K.mw - 5)) & 31u;
            bitmap[word] |= 1u << (31 - bit);
        }
    };
    for (u64 d = 0; d < ndim; d++) {
        int32_t pf = shuf_perm[d];
        if (pf < 0 || pf >= I.dim_end) continue;
        u32 h = ((u32)d * 0x9E3779B1u) >> 11;
        for (;;) {
            h &= K.ptab_mask;
            if (ptab[h] == 0) { ptab[h] = ((u64)(u32)pf << 32) | (u64)((u32)d + 1u); break; }
            h++;
        }
        set_bit(pair_reverse(d, 2 * subk));   // forward strand is canonical
        set_bit((~d) & mwmask);               // reverse complement is canonical
        mk_s3_filter_add(bitmap3, K.mw, pair_reverse(d, 2 * subk));
        mk_s3_filter_add(bitmap3, K.mw, (~d) & mwmask);
    }
    ctx->bitmap_words = bm_words;
    if (cudaMalloc(&ctx->d_bitmap, (size_t)bm_words * 4) != cudaSuccess) return fail(MK_ERR_NOMEM);
    if (cudaMalloc(&ctx->d_ptab, tcap * 8) != cudaSuccess) return fail(MK_ERR_NOMEM);
    ctx->bitmap3_words = (u32)bitmap3.size();
    if (cudaMalloc(&ctx->d_bitmap3, bitmap3.size() * 4) != cudaSuccess) return fail(MK_ERR_NOMEM);
    if (cudaMemcpy(ctx->d_bitmap3, bitmap3.data(), bitmap3.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess)
        return fail(MK_ERR_CUDA);
    if (cudaMemcpy(ctx->d_bitmap, bitmap.data(), (size_t)bm_words * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(ctx->d_ptab, ptab.data(), tcap * 8, cudaMemcpyHostToDevice) != cudaSuccess)
        return fail(MK_ERR_CUDA);
    *out = ctx;
    return MK_OK;
}

extern "C" void mk_ctx_destroy(mk_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < SB_NUM; i++)
        if (ctx->sb[i].p) cudaFree(ctx->sb[i].p);
    if (ctx->d_bitmap) cudaFree(ctx->d_bitmap);
    if (ctx->d_bitmap3) cudaFree(ctx->d_bitmap3);
    if (ctx->d_ptab) cudaFree(ctx->d_ptab);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    mk_comm_destroy(ctx);
    mk_markerdb_unload(ctx);
    for (cudaEvent_t e : ctx->chunk_ev) cudaEventDestroy(e);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev2) cudaEventDestroy(ctx->ev2);
    if (ctx->ev3) cudaEventDestroy(ctx->ev3);
    if (ctx->evx[0]) cudaEventDestroy(ctx->evx[0]);
    if (ctx->evx[1]) cudaEventDestroy(ctx->evx[1]);
    if (ctx->copy_ev[0]) cudaEventDestroy(ctx->copy_ev[0]);
    if (ctx->copy_ev[1]) cudaEventDestroy(ctx->copy_ev[1]);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
}

extern "C" int mk_ctx_info(const mk_ctx *ctx, mk_info *info)
{
    if (!ctx || !info) return MK_ERR_ARG;
    *info = ctx->info;
    return MK_OK;
}

extern "C" int mk_ctx_profile(mk_ctx *ctx, mk_profile *prof, int reset)
{
    if (!ctx) return MK_ERR_ARG;
    if (prof) *prof = ctx->prof;
    if (reset) memset(&ctx->prof, 0, sizeof(ctx->prof));
    return MK_OK;
}

extern "C" int mk_ctx_synchronize(mk_ctx *ctx)
{
    if (!ctx) return MK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return MK_OK;
}

extern "C" void *mk_ctx_cuda_stream(mk_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

extern "C" void mk_sketch_free(mk_sketch *s)
{
    if (!s) return;
    for (int c = 0; c < s->n_components && !s->borrowed; c++) {
        if (s->codes) free(s->codes[c]);
        if (s->counts) free(s->counts[c]);
    }
    free(s->codes);
    free(s->counts);
    free(s->n);
    memset(s, 0, sizeof(*s));
}

static int bits_for(u64 v)
{
    int b = 1;
    while (b < 64 && (v >> b)) b++;
    return b;
}

// ---- FASTQ -A -------------------------------------------------------------------------------------
extern "C" int mk_fastq_koc_device(mk_ctx *ctx, const void *d_text, size_t nbytes, mk_sketch *out)
{
    if (!ctx || !out || (!d_text && nbytes)) return MK_ERR_ARG;
    memset(out, 0, sizeof(*out));
    CK(cudaSetDevice(ctx->device));
    u64 *cc = nullptr, *cp = nullptr, n_cand = 0;
    MkPhaseClock pc(ctx->stream);
    CKR(mk_stream_fastq(ctx, (const uint8_t *)d_text, nbytes, 0, 0, false, &cc, &cp, &n_cand, nullptr));
    pc.mark("stream+verify");
    long long keep_below = -1;
    if (nbytes) CKR(mk_tail_cut(ctx, (const uint8_t *)d_text, nbytes, &keep_below));
    pc.mark("tail cut");
    ctx->pos_bits = bits_for((u64)nbytes);
    int rc = mk_finalize_candidates(ctx, cc, cp, n_cand, keep_below, nullptr, 1, true, out);
    pc.mark("finalize");
    ctx->pos_bits = 64;
    if (rc != MK_OK) mk_sketch_free(out);
    return rc;
}

static int upload_text(mk_ctx *ctx, const void *h_text, size_t nbytes, uint8_t **d_text)
{
    uint8_t *d;
    CKR(mk_scratch(ctx, SB_TEXT, nbytes + 256, &d));
    // chunked so that pageable sources do not need one giant staging pass
    const size_t CH = (size_t)256 << 20;
    for (size_t o = 0; o < nbytes; o += CH) {
        size_t m = nbytes - o < CH ? nbytes - o : CH;
        CK(cudaMemcpyAsync(d + o, (const uint8_t *)h_text + o, m, cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(cudaMemsetAsync(d + nbytes, 0, 64, ctx->stream));
    ctx->prof.h2d_bytes += nbytes;
    *d_text = d;
    return MK_OK;
}

extern "C" int mk_fastq_koc_host(mk_ctx *ctx, const void *h_text, size_t nbytes, mk_sketch *out)
{
    if (!ctx || !out || (!h_text && nbytes)) return MK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    // the text is uploaded by the stream driver, chunk by chunk under the kernel of the chunk before
    uint8_t *d = nullptr;
    CKR(mk_scratch(ctx, SB_TEXT, nbytes + 256, &d));
    CK(cudaMemsetAsync(d + nbytes, 0, 64, ctx->stream));
    ctx->h_src = (const uint8_t *)h_text;
    ctx->h_src_all = (const uint8_t *)h_text;
    int rc = mk_fastq_koc_device(ctx, d, nbytes, out);
    ctx->h_src = nullptr;
    ctx->h_src_all = nullptr;
    return rc;
}

// ---- FASTQ without -A (`dist -Q q -n m`) ---------------------------------------------------------------
// fastq2co() + write_fqco2file() (iseq2comem.c:323-419, 596-621): the same k-mer arithmetic and slot table as the -A
// path with (a) a per-base quality threshold, (b) lines read with fgets(.., 20000, ..), (c) a record after the first
// used only if its fourth line ended with a newline, (d) only codes seen at least m times written, without counts.
struct FqCoScope {       // the per-call options live in the context only while the call runs
    mk_ctx *ctx;
    FqCoScope(mk_ctx *c, int q, int m) : ctx(c) { c->verify_quality = q; c->line_limit = 19999; c->emit_lo = (u32)m; c->emit_hi = 0xFFFFFFFFu; }
    ~FqCoScope() { ctx->verify_quality = INT_MIN; ctx->line_limit = 4095; ctx->emit_lo = 0; ctx->emit_hi = 0xFFFFFFFFu; }
};

extern "C" int mk_fastq_co_device(mk_ctx *ctx, const void *d_text, size_t nbytes, int quality, int min_occurrence, mk_sketch *out)
{
    if (!ctx || !out || (!d_text && nbytes)) return MK_ERR_ARG;
    if (min_occurrence < 1 || min_occurrence >= 15) {        // fastq2co(): "Occurence num should smaller than 15"
        snprintf(ctx->err, sizeof(ctx->err), "fastq2co(): Occurence num should smaller than 15");
        return MK_ERR_ARG;
    }
    memset(out, 0, sizeof(*out));
    CK(cudaSetDevice(ctx->device));
    FqCoScope scope(ctx, quality, min_occurrence);
    u64 *cc = nullptr, *cp = nullptr, n_cand = 0, n_newlines = 0;
    CKR(mk_stream_fastq(ctx, (const uint8_t *)d_text, nbytes, 0, 0, false, &cc, &cp, &n_cand, &n_newlines));
    long long keep_below = LLONG_MAX;
    CKR(mk_tail_cut_fq2co(ctx, (const uint8_t *)d_text, nbytes, n_newlines, &keep_below));
    ctx->pos_bits = bits_for((u64)nbytes);
    int rc = mk_finalize_candidates(ctx, cc, cp, n_cand, keep_below, nullptr, 1, false, out);
    ctx->pos_bits = 64;
    if (rc != MK_OK) mk_sketch_free(out);
    return rc;
}

extern "C" int mk_fastq_co_host(mk_ctx *ctx, const void *h_text, size_t nbytes, int quality, int min_occurrence, mk_sketch *out)
{
    if (!ctx || !out || (!h_text && nbytes)) return MK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    uint8_t *d = nullptr;
    CKR(mk_scratch(ctx, SB_TEXT, nbytes + 256, &d));
    CK(cudaMemsetAsync(d + nbytes, 0, 64, ctx->stream));
    ctx->h_src = (const uint8_t *)h_text;
    ctx->h_src_all = (const uint8_t *)h_text;
    int rc = mk_fastq_co_device(ctx, d, nbytes, quality, min_occurrence, out);
    ctx->h_src = nullptr;
    ctx->h_src_all = nullptr;
    return rc;
}

extern "C" int mk_ctx_set_borrowed_output(mk_ctx *ctx, int on)
{
    if (!ctx) return MK_ERR_ARG;
    ctx->borrow_output = on != 0;
    return MK_OK;
}

extern "C" int mk_ctx_set_dedup(mk_ctx *ctx, int on)
{
    if (!ctx) return MK_ERR_ARG;
    ctx->fasta_dedup = on != 0;
    return MK_OK;
}

// ---- FASTA ----------------------------------------------------------------------------------------
extern "C" int mk_fasta_co_device(mk_ctx *ctx, const void *d_text, const uint64_t *offsets, int n_files, mk_sketch *out)
{
    if (!ctx || !out || !offsets || n_files <= 0) return MK_ERR_ARG;
    for (int f = 0; f < n_files; f++) {
        memset(&out[f], 0, sizeof(out[f]));
        if (offsets[f + 1] < offsets[f]) return MK_ERR_ARG;
    }
    if (offsets[0] != 0) return MK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    size_t nbytes = (size_t)offsets[n_files];
    uint8_t *dense = nullptr;
    u64 dense_bytes = 0, *d_dense_off = nullptr;
    MkPhaseClock pc(ctx->stream);
    CKR(mk_fasta_compact(ctx, (const uint8_t *)d_text, nbytes, (const u64 *)offsets, n_files, &dense, &dense_bytes,
                         &d_dense_off));
    u64 *cc = nullptr, *cp = nullptr, n_cand = 0;
    pc.mark("fasta compact");
    CKR(mk_stream_fastq(ctx, dense, (size_t)dense_bytes, 0, 0, true, &cc, &cp, &n_cand, nullptr));
    pc.mark("stream+verify");
    ctx->pos_bits = bits_for(dense_bytes);
    if (ctx->fasta_dedup) { ctx->emit_lo = 1; ctx->emit_hi = 1; }      // `dist -u`: uniq_fasta2co() (iseq2comem.c:729-828)
    int rc = mk_finalize_candidates(ctx, cc, cp, n_cand, LLONG_MAX, d_dense_off, n_files, false, out);
    ctx->emit_lo = 0; ctx->emit_hi = 0xFFFFFFFFu;
    pc.mark("finalize");
    ctx->pos_bits = 64;
    if (rc != MK_OK)
        for (int f = 0; f < n_files; f++) mk_sketch_free(&out[f]);
    return rc;
}

extern "C" int mk_fasta_co_host(mk_ctx *ctx, const void *h_text, const uint64_t *offsets, int n_files, mk_sketch *out)
{
    if (!ctx || !out || !offsets || n_files <= 0) return MK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    uint8_t *d = nullptr;
    CKR(upload_text(ctx, h_text, (size_t)offsets[n_files], &d));
    return mk_fasta_co_device(ctx, d, offsets, n_files, out);
}

// development aid (not part of the public header): device buffer of 64 * warps * 8 u64 receiving
// clock64() stamps of CTA 0's first 64 iterations; nullptr disables.
extern "C" int mk_debug_set_trace(mk_ctx *ctx, void *d_buf)
{
    if (!ctx) return MK_ERR_ARG;
    ctx->d_trace = d_buf;
    return MK_OK;
}
