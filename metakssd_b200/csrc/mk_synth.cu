// mk_synth.cu — device-side synthetic workload generator (bench / tests only).
// Every byte is a pure function of (params, offset) defined in include/mkssd_synth.h, so the text
// is identical to what the C generator (oracle side) writes for the same parameters.
#include "mk_common.cuh"
#include "mkssd_synth.h"

__global__ void __launch_bounds__(256)
k_synth_fastq(mks_params P, const u32 *__restrict__ cdf32, const u32 *__restrict__ spc, u64 off0, u64 nbytes,
              uint8_t *__restrict__ out)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += (u64)gridDim.x * blockDim.x) {
        u64 off = off0 + i;
        u64 r = mks_fastq_record_of(&P, off);
        u32 j = (u32)(off - mks_fastq_offset(&P, r));
        out[i] = (uint8_t)mks_fastq_char(&P, cdf32, spc, r, j);
    }
}

__device__ __forceinline__ u32 dev_ndigits(u32 v)
{
    u32 n = 1;
    while (v >= 10) { v /= 10; n++; }
    return n;
}

__global__ void __launch_bounds__(256)
k_synth_fasta(mks_params P, u32 s0, u32 ns, const u64 *__restrict__ offsets, u64 nbytes, uint8_t *__restrict__ out)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += (u64)gridDim.x * blockDim.x) {
        // species whose file holds byte i
        u32 lo = 0, hi = ns - 1;
        while (lo < hi) {
            u32 mid = (lo + hi + 1) >> 1;
            if (offsets[mid] <= i) lo = mid; else hi = mid - 1;
        }
        u32 s = s0 + lo;
        u64 j = i - offsets[lo];
        u32 nd = dev_ndigits(s);
        u32 hdr = 3 + nd + 1; // ">sp" digits "\n"
        uint8_t c;
        if (j < hdr) {
            if (j == 0) c = '>';
            else if (j == 1) c = 's';
            else if (j == 2) c = 'p';
            else if (j < 3 + nd) {
                u32 kdig = nd - 1 - (u32)(j - 3);
                u32 v = s;
                while (kdig--) v /= 10;
                c = (uint8_t)('0' + v % 10);
            } else c = '\n';
        } else {
            u64 q = j - hdr;
            u64 line = q / 81, col = q % 81;
            u64 p = line * 80 + col;
            if (col == 80 || p >= P.genome_len) c = '\n';
            else c = (uint8_t)("ACGT"[mks_genome_base(&P, s, (u32)p)]);
        }
        out[i] = c;
    }
}

extern "C" int mk_synth_fastq_device(mk_ctx *ctx, const struct mks_params *P, const uint32_t *cdf32,
                                     const uint32_t *cdf_species, uint64_t r0, uint64_t r1, void *d_out,
                                     size_t capacity, size_t *written)
{
    if (!ctx || !P || !cdf32 || !cdf_species || !d_out || r1 < r0) return MK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    u64 off0 = mks_fastq_offset(P, r0), off1 = mks_fastq_offset(P, r1);
    u64 nbytes = off1 - off0;
    if (written) *written = (size_t)nbytes;
    if (nbytes > capacity) return MK_ERR_ARG;
    if (nbytes == 0) return MK_OK;
    u32 *d_cdf, *d_spc;
    CKR(mk_scratch(ctx, SB_SYN_CDF, (size_t)P->n_present, &d_cdf));
    CKR(mk_scratch(ctx, SB_SYN_SPC, (size_t)P->n_present, &d_spc));
    CK(cudaMemcpyAsync(d_cdf, cdf32, (size_t)P->n_present * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_spc, cdf_species, (size_t)P->n_present * 4, cudaMemcpyHostToDevice, ctx->stream));
    k_synth_fastq<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>(*P, d_cdf, d_spc, off0, nbytes, (uint8_t *)d_out);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return MK_OK;
}

extern "C" int mk_synth_fasta_device(mk_ctx *ctx, const struct mks_params *P, uint32_t s0, uint32_t s1, void *d_out,
                                     size_t capacity, uint64_t *offsets)
{
    if (!ctx || !P || !d_out || !offsets || s1 <= s0) return MK_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    u32 ns = s1 - s0;
    offsets[0] = 0;
    for (u32 i = 0; i < ns; i++) offsets[i + 1] = offsets[i] + mks_fasta_size(P, s0 + i);
    u64 nbytes = offsets[ns];
    if (nbytes > capacity) return MK_ERR_ARG;
    u64 *d_off;
    CKR(mk_scratch(ctx, SB_FILE_OFF, (size_t)ns + 1, &d_off));
    CK(cudaMemcpyAsync(d_off, offsets, (size_t)(ns + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_synth_fasta<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>(*P, s0, ns, d_off, nbytes, (uint8_t *)d_out);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return MK_OK;
}

// ---- host-side helpers of the generator (thin wrappers over include/mkssd_synth.h) --------------
extern "C" int mk_synth_build(struct mks_params *P, uint64_t seed, uint32_t n_species, uint32_t genome_len,
                              uint32_t read_len, uint32_t **cdf32, uint32_t **species)
{
    if (!P || !cdf32 || !species || n_species == 0 || genome_len < read_len) return MK_ERR_ARG;
    mks_default_params(P, seed, n_species, genome_len, read_len);
    return mks_build_cdf(P, cdf32, species) == 0 ? MK_OK : MK_ERR_NOMEM;
}
extern "C" void mk_synth_free(void *p) { free(p); }
extern "C" uint64_t mk_synth_fastq_bytes(const struct mks_params *P, uint64_t r0, uint64_t r1)
{
    return mks_fastq_offset(P, r1) - mks_fastq_offset(P, r0);
}
extern "C" uint64_t mk_synth_fasta_bytes(const struct mks_params *P, uint32_t s) { return mks_fasta_size(P, s); }
extern "C" void mk_synth_shuf_perm(uint64_t seed, int subk, int32_t *perm) { mks_make_shuf_perm(seed, subk, perm); }
extern "C" int32_t mk_synth_shuf_id(uint64_t seed) { return mks_shuf_id(seed); }
