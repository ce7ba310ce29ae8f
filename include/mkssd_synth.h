/*
 * mkssd_synth.h — deterministic synthetic metagenome generator shared by the CPU (C) and the
 * GPU (CUDA) sides.  Header-only; every function is counter-based (a pure function of
 * (seed, index)), so a CUDA kernel and a plain C loop produce byte-identical FASTQ/FASTA text.
 *
 * Workload definition: SURVEY.md §8(d) — species genomes are iid-uniform ACGT, species of one
 * genus share a common leading segment, reads are fixed-length, uniform start, 50/50 strand,
 * substitution errors and a small fraction of 'N' (exercises the window-reset path of
 * /root/reference/iseq2comem.c:682-688).  Abundances are log-normal-like, built with integer
 * arithmetic only so that the container and the GPU box agree bit for bit.
 *
 * Nothing here comes from the reference; it only produces inputs for it.
 */
#ifndef MKSSD_SYNTH_H
#define MKSSD_SYNTH_H
#include <stdint.h>
#include <stddef.h>

#ifdef __CUDACC__
#define MKS_HD __host__ __device__ __forceinline__
#else
#define MKS_HD static inline
#endif

typedef struct mks_params {
    uint64_t seed;
    uint32_t n_species;    /* species in the community / MarkerDB                         */
    uint32_t genome_len;   /* bases per species genome                                    */
    uint32_t read_len;     /* bases per read                                              */
    uint32_t genus_size;   /* species per genus (share the first shared_len bases)        */
    uint32_t shared_len;   /* length of the genus-common segment                          */
    uint32_t sub_thresh16; /* P(substitution) = sub_thresh16 / 65536 per base             */
    uint32_t n_thresh16;   /* P(base -> 'N')  = n_thresh16  / 65536 per base              */
    uint32_t n_present;    /* species with non-zero abundance (entries of the CDF)        */
} mks_params;

/* splitmix64 finaliser */
MKS_HD uint64_t mks_mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

MKS_HD uint64_t mks_hash2(uint64_t seed, uint64_t a, uint64_t b)
{
    return mks_mix64(mks_mix64(seed ^ (a * 0xD6E8FEB86659FD93ull)) + b);
}

/* 2-bit base (A=0,C=1,G=2,T=3) of species s at position p */
MKS_HD uint32_t mks_genome_base(const mks_params *P, uint32_t s, uint32_t p)
{
    uint64_t owner = (p < P->shared_len) ? (0x8000000000ull + (uint64_t)(s / P->genus_size))
                                         : (uint64_t)s;
    uint64_t h = mks_hash2(P->seed ^ 0x47454E4Full /* "GENO" */, owner, (uint64_t)(p >> 5));
    return (uint32_t)(h >> (2 * (p & 31))) & 3u;
}

/* per-read draw: species slot in the CDF, start, strand */
typedef struct mks_read_hdr {
    uint32_t species;
    uint32_t start;
    uint32_t strand; /* 1 = reverse complement */
} mks_read_hdr;

MKS_HD mks_read_hdr mks_read_header(const mks_params *P, const uint32_t *cdf32,
                                    const uint32_t *cdf_species, uint64_t r)
{
    mks_read_hdr H;
    uint64_t h = mks_hash2(P->seed ^ 0x52454144ull /* "READ" */, r, 0);
    uint32_t u = (uint32_t)(h >> 32);
    /* first slot with cdf32[slot] > u ; cdf32[n_present-1] == 0xFFFFFFFF */
    uint32_t lo = 0, hi = P->n_present - 1;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (cdf32[mid] > u) hi = mid; else lo = mid + 1;
    }
    H.species = cdf_species[lo];
    H.start = (uint32_t)((h & 0xFFFFFFFFull) % (uint64_t)(P->genome_len - P->read_len + 1));
    H.strand = (uint32_t)(mks_mix64(h) & 1u);
    return H;
}

/* ASCII character of base i of read r */
MKS_HD char mks_read_char(const mks_params *P, const mks_read_hdr *H, uint64_t r, uint32_t i)
{
    uint32_t b;
    if (H->strand) b = 3u - mks_genome_base(P, H->species, H->start + P->read_len - 1 - i);
    else           b = mks_genome_base(P, H->species, H->start + i);
    uint64_t e = mks_hash2(P->seed ^ 0x4552524Full /* "ERRO" */, r, (uint64_t)(i >> 2));
    uint32_t e16 = (uint32_t)(e >> (16 * (i & 3))) & 0xFFFFu;
    if (e16 < P->n_thresh16) return 'N';
    if (e16 < P->n_thresh16 + P->sub_thresh16) b = (b + 1u + (e16 % 3u)) & 3u;
    return (char)("ACGT"[b]);
}

/* ---- FASTQ record layout:  "@r<idx>\n" <seq> "\n+\n" <'I' x L> "\n" -------------------- */
MKS_HD uint32_t mks_ndigits(uint64_t v)
{
    uint32_t n = 1;
    while (v >= 10) { v /= 10; n++; }
    return n;
}

/* sum of decimal digit counts of 0 .. r-1 */
MKS_HD uint64_t mks_digits_below(uint64_t r)
{
    uint64_t total = 0, lo = 0, hi = 10;
    uint32_t d = 1;
    while (lo < r) {
        uint64_t top = r < hi ? r : hi;
        total += (top - lo) * d;
        lo = hi;
        hi = (hi > 1000000000000000000ull) ? ~0ull : hi * 10;
        d++;
    }
    return total;
}

/* byte offset of record r in the FASTQ text */
MKS_HD uint64_t mks_fastq_offset(const mks_params *P, uint64_t r)
{
    return r * (2ull * P->read_len + 7ull) + mks_digits_below(r);
}

/* record index containing byte offset `off` (inverse of mks_fastq_offset) */
MKS_HD uint64_t mks_fastq_record_of(const mks_params *P, uint64_t off)
{
    uint64_t lo = 0, hi = 10, base_off = 0;
    uint32_t d = 1;
    for (;;) {
        uint64_t rec = 2ull * P->read_len + 7ull + d;
        uint64_t span = (hi - lo) * rec;
        if (off < base_off + span || hi == ~0ull) return lo + (off - base_off) / rec;
        base_off += span;
        lo = hi;
        hi = (hi > 1000000000000000000ull) ? ~0ull : hi * 10;
        d++;
    }
}

/* character at offset j (0-based) inside record r */
MKS_HD char mks_fastq_char(const mks_params *P, const uint32_t *cdf32, const uint32_t *cdf_species,
                           uint64_t r, uint32_t j)
{
    uint32_t nd = mks_ndigits(r);
    uint32_t L = P->read_len;
    if (j == 0) return '@';
    if (j == 1) return 'r';
    if (j < 2 + nd) {
        uint32_t k = nd - 1 - (j - 2); /* digit weight index */
        uint64_t v = r;
        while (k--) v /= 10;
        return (char)('0' + (v % 10));
    }
    j -= 2 + nd;
    if (j == 0) return '\n';
    j -= 1;
    if (j < L) {
        mks_read_hdr H = mks_read_header(P, cdf32, cdf_species, r);
        return mks_read_char(P, &H, r, j);
    }
    j -= L;
    if (j == 0) return '\n';
    if (j == 1) return '+';
    if (j == 2) return '\n';
    j -= 3;
    if (j < L) return 'I';
    return '\n';
}

/* ---------------- host-only helpers (plain C; plain host functions under nvcc) ----------- */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

/* Integer-only heavy-tailed ("log2-normal") weight for species rank i: 12 summed 16-bit
 * uniforms ~ N(0,1)*65536, scaled so that ln(w) ~ N(0, 1.5^2). */
static inline uint64_t mks_weight(uint64_t seed, uint32_t i)
{
    int64_t t = 0;
    for (int j = 0; j < 3; j++) {
        uint64_t h = mks_hash2(seed ^ 0x41424E44ull /* "ABND" */, i, (uint64_t)j);
        t += (int64_t)(h & 0xFFFF) + (int64_t)((h >> 16) & 0xFFFF) + (int64_t)((h >> 32) & 0xFFFF) +
             (int64_t)((h >> 48) & 0xFFFF);
    }
    t -= 6 * 65535;
    int64_t lf = (t * 141822) / 65536; /* log2(w) in 16.16 fixed point: 1.5/ln2 = 2.164 */
    int64_t ip = lf >> 16;
    uint64_t fr = (uint64_t)(lf & 0xFFFF);
    if (ip < -14) ip = -14;
    if (ip > 14) ip = 14;
    return (65536ull + fr) << (ip + 14);
}

/* Build the abundance CDF: n_present = ceil(0.3 * n_species) species (chosen by hash order),
 * returns arrays (caller frees) and fills P->n_present. */
static inline int mks_build_cdf(mks_params *P, uint32_t **cdf32_out, uint32_t **species_out)
{
    uint32_t S = P->n_species;
    uint32_t np = (3 * S + 9) / 10;
    if (np < 1) np = 1;
    /* choose the np species with the smallest hash */
    uint64_t *key = (uint64_t *)malloc(sizeof(uint64_t) * S);
    uint32_t *idx = (uint32_t *)malloc(sizeof(uint32_t) * S);
    if (!key || !idx) return -1;
    for (uint32_t s = 0; s < S; s++) {
        key[s] = mks_hash2(P->seed ^ 0x50524553ull /* "PRES" */, s, 0);
        idx[s] = s;
    }
    /* selection by simple insertion sort on (key,idx): S is at most a few thousand */
    for (uint32_t a = 1; a < S; a++) {
        uint64_t k = key[a];
        uint32_t v = idx[a];
        uint32_t b = a;
        while (b > 0 && key[b - 1] > k) { key[b] = key[b - 1]; idx[b] = idx[b - 1]; b--; }
        key[b] = k; idx[b] = v;
    }
    uint32_t *cdf = (uint32_t *)malloc(sizeof(uint32_t) * np);
    uint32_t *spc = (uint32_t *)malloc(sizeof(uint32_t) * np);
    uint64_t *w = (uint64_t *)malloc(sizeof(uint64_t) * np);
    if (!cdf || !spc || !w) return -1;
    __uint128_t tot = 0;
    for (uint32_t i = 0; i < np; i++) { spc[i] = idx[i]; w[i] = mks_weight(P->seed, idx[i]); tot += w[i]; }
    __uint128_t acc = 0;
    for (uint32_t i = 0; i < np; i++) {
        acc += w[i];
        __uint128_t c = (acc << 32) / tot;
        cdf[i] = (c >= ((__uint128_t)1 << 32)) ? 0xFFFFFFFFu : (uint32_t)c;
    }
    cdf[np - 1] = 0xFFFFFFFFu;
    free(key); free(idx); free(w);
    P->n_present = np;
    *cdf32_out = cdf; *species_out = spc;
    return 0;
}

static inline void mks_default_params(mks_params *P, uint64_t seed, uint32_t n_species,
                                      uint32_t genome_len, uint32_t read_len)
{
    memset(P, 0, sizeof(*P));
    P->seed = seed;
    P->n_species = n_species;
    P->genome_len = genome_len;
    P->read_len = read_len;
    P->genus_size = 10;
    P->shared_len = genome_len / 5;
    P->sub_thresh16 = 328; /* 0.5 %  */
    P->n_thresh16 = 13;    /* 0.02 % */
    P->n_present = 0;
}

/* write reads [r0, r1) as FASTQ text into buf (must hold mks_fastq_offset(r1)-mks_fastq_offset(r0)) */
static inline size_t mks_write_fastq(const mks_params *P, const uint32_t *cdf32, const uint32_t *spc,
                                     uint64_t r0, uint64_t r1, char *buf)
{
    char *o = buf;
    uint32_t L = P->read_len;
    for (uint64_t r = r0; r < r1; r++) {
        o += sprintf(o, "@r%llu\n", (unsigned long long)r);
        mks_read_hdr H = mks_read_header(P, cdf32, spc, r);
        for (uint32_t i = 0; i < L; i++) *o++ = mks_read_char(P, &H, r, i);
        *o++ = '\n'; *o++ = '+'; *o++ = '\n';
        memset(o, 'I', L); o += L;
        *o++ = '\n';
    }
    return (size_t)(o - buf);
}

/* FASTA text of species s: ">sp<s>\n" then 80-column lines. Returns bytes written. */
static inline size_t mks_fasta_size(const mks_params *P, uint32_t s)
{
    char hdr[32];
    size_t h = (size_t)sprintf(hdr, ">sp%u\n", s);
    size_t G = P->genome_len;
    return h + G + (G + 79) / 80;
}
static inline size_t mks_write_fasta(const mks_params *P, uint32_t s, char *buf)
{
    char *o = buf;
    o += sprintf(o, ">sp%u\n", s);
    uint32_t G = P->genome_len;
    for (uint32_t p = 0; p < G; p++) {
        *o++ = "ACGT"[mks_genome_base(P, s, p)];
        if ((p % 80) == 79 || p == G - 1) *o++ = '\n';
    }
    return (size_t)(o - buf);
}

/* Deterministic .shuf permutation (file format of /root/reference/command_shuffle.c:164-211:
 * 16-byte header {id,k,subk,drlevel} + int32[16^subk]); Fisher-Yates driven by splitmix64
 * instead of the reference's srand(time(NULL)). perm must hold 1 << 4*subk ints. */
static inline void mks_make_shuf_perm(uint64_t seed, int subk, int32_t *perm)
{
    uint32_t n = 1u << (4 * subk);
    for (uint32_t i = 0; i < n; i++) perm[i] = (int32_t)i;
    uint64_t st = seed ^ 0x53485546ull; /* "SHUF" */
    for (uint32_t i = n - 1; i > 0; i--) {
        st = mks_mix64(st);
        uint32_t j = (uint32_t)(((__uint128_t)st * (uint64_t)(i + 1)) >> 64);
        int32_t t = perm[i]; perm[i] = perm[j]; perm[j] = t;
    }
}
static inline int32_t mks_shuf_id(uint64_t seed) { return (int32_t)(mks_mix64(seed ^ 0x4944ull) & 0x7FFFFFFF); }
#endif /* MKSSD_SYNTH_H */
