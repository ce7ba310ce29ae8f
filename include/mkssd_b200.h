/*
 * mkssd_b200.h — C ABI of libmkssd_b200.so: the B200 (sm_100a) implementation of MetaKSSD's
 * data-parallel hot path.  Plain pointers and sizes only; this is what the reference's host C
 * code (or any FFI: cgo / JNI / ctypes) binds instead of its CPU loops.
 *
 * Reference (yhg926/MetaKSSD v2.21) interface replaced by each entry point:
 *
 *   mk_ctx_create            seq2co_global_var_initial()           iseq2comem.c:54-86
 *                            + get_hashsz()                        command_dist.c:286-315
 *                            (the .shuf permutation is what read_dim_shuffle_file() returns,
 *                             command_shuffle.c:215-235)
 *   mk_fastq_koc_*           mt_shortreads2koc() + write_fqkoc2files()
 *                                                                  iseq2comem.c:657-727, 516-562
 *                            (call sites command_dist.c:380,382)
 *   mk_fasta_co_*            fasta2co() + wrt_co2cmpn_use_inn_subctx()
 *                                                                  iseq2comem.c:218-315, 625-652
 *                            (call sites command_dist.c:397-398)
 *   mk_composite_*           the dictionary build + probe loop of get_species_abundance()
 *                                                                  command_composite.c:535-566
 *                            and (mk_composite_stats) the per-species order statistics of
 *                                                                  command_composite.c:598-613
 *
 * Results are bit-identical to the reference run single-threaded (`-p 1`): same codes and
 * counts, same on-disk (hash-slot) order, same per-component split.
 *
 * There is NO CPU fallback: every call needs a CUDA device and fails with MK_ERR_CUDA otherwise.
 * All functions return 0 (MK_OK) or a negative MK_ERR_* code; mk_strerror() names it and
 * mk_last_error() returns the last detailed message of the calling context.
 * A context is bound to one device and must be used from one host thread at a time.
 */
#ifndef MKSSD_B200_H
#define MKSSD_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MK_OK 0
#define MK_ERR_ARG (-1)         /* bad argument (NULL, misaligned device pointer, ...)          */
#define MK_ERR_PARAM (-2)       /* k/subk/L outside the reference's table range (get_hashsz err) */
#define MK_ERR_CUDA (-3)        /* CUDA runtime failure / no device                              */
#define MK_ERR_NOMEM (-4)
#define MK_ERR_CROWDED (-5)     /* distinct codes > hashlimit: reference err() "the context
                                   space is too crowd, try rerun the program using -k<k+1>"     */
#define MK_ERR_LONG_LINE (-6)   /* FASTQ line >= 4095 bytes: fgets(…,4096) splitting in the
                                   reference is input-buffer dependent; refused                 */
#define MK_ERR_IO (-7)          /* popen/fopen/read failure                                      */
#define MK_ERR_EMPTY_QUERY (-8) /* composite query component of exactly 1 code (reference: modulo by
                                   zero, command_composite.c:535; 0 codes = no hits, not an error)  */
#define MK_ERR_UNSUPPORTED (-9)

typedef struct mk_ctx mk_ctx;

/* Derived sketch parameters, same values the reference computes (iseq2comem.c:54-86). */
typedef struct mk_info {
    int k, subk, drlevel;
    int kmer_len;            /* TL = 2k                                                     */
    int outctx;              /* half outer context = k - subk                               */
    int dim_end;             /* a k-mer passes iff shuf[inner substring] < dim_end          */
    uint32_t hashsize;       /* primer[4(k-L)-15]                                            */
    uint32_t hashlimit;      /* (uint32)(hashsize * 0.6)                                     */
    int component_num;       /* 16^(k-L-8) or 1                                              */
    int comp_code_bits;      /* 4(k-L-8) or 0                                                */
    int code_bits;           /* 4(k-L)                                                       */
    int device;
    int sm_count;
} mk_info;

/* A sketch in the reference's per-component on-disk layout: for component c,
 * codes[c][0..n[c]) is exactly the content of "<i>.co.<c>" (uint32 = code >> comp_code_bits)
 * and counts[c] the content of "<i>.co.<c>.a" (NULL for FASTA sketches).  Library-owned;
 * release with mk_sketch_free(). */
typedef struct mk_sketch {
    int n_components;
    uint64_t n_total;        /* sum of n[c]  (== write_fqkoc2files() return value)           */
    uint64_t *n;
    uint32_t **codes;
    uint16_t **counts;
    int borrowed;            /* 1: codes[c] / counts[c] point into the context's pinned staging block (valid until the
                                next call on the context; mk_ctx_set_borrowed_output), mk_sketch_free() leaves them */
} mk_sketch;

/* Accumulated device timings (CUDA events on the context's stream), for bench/roofline. */
typedef struct mk_profile {
    double stream_kernel_ms;     /* k_stream_* launches (the dominant kernel)               */
    uint64_t stream_kernel_launches;
    uint64_t stream_kernel_bytes;/* text bytes those launches scanned                       */
    double reduce_ms;            /* count accumulation + ordering + slot reconstruction     */
    double composite_ms;
    uint64_t kernel_launches;    /* every kernel launched by this context                   */
    uint64_t h2d_bytes, d2h_bytes;
    double exchange_ms;          /* multi-GPU: pack + NCCL exchange + owner merge + rank-local composite + gather */
    double exchange_wait_ms;     /* multi-GPU: the first grouped send/recv alone, i.e. mostly the wait for the slowest rank */
} mk_profile;

const char *mk_strerror(int code);
const char *mk_last_error(const mk_ctx *ctx);
int mk_device_count(void);

/* shuf_perm: the 16^subk int32 permutation of the .shuf file (host memory).  NULL creates a context without
 * pass-set tables, good for the composite entry points only (sketching calls then return MK_ERR_ARG). */
int mk_ctx_create(mk_ctx **out, const int32_t *shuf_perm, int k, int subk, int drlevel, int device);
void mk_ctx_destroy(mk_ctx *ctx);
int mk_ctx_info(const mk_ctx *ctx, mk_info *info);
int mk_ctx_profile(mk_ctx *ctx, mk_profile *prof, int reset);
int mk_ctx_synchronize(mk_ctx *ctx);
/* The CUDA stream (cudaStream_t) all kernels and copies of this context are issued on, so that a
 * caller can bracket calls with its own events. */
void *mk_ctx_cuda_stream(mk_ctx *ctx);

/* ---- FASTQ with k-mer counts (`dist -A`) ------------------------------------------------ */
/* d_text: device pointer (16-byte aligned, readable up to the next 16-byte boundary past
 * nbytes) holding the whole decompressed FASTQ text.  Input stays in HBM. */
int mk_fastq_koc_device(mk_ctx *ctx, const void *d_text, size_t nbytes, mk_sketch *out);
/* h_text: host memory (pinned memory makes the copy asynchronous and faster); the H2D copy
 * is chunked and overlapped with the kernels. */
int mk_fastq_koc_host(mk_ctx *ctx, const void *h_text, size_t nbytes, mk_sketch *out);
/* Same as the reference call (iseq2comem.c:664-673): the text of `<pipecmd or "zcat -fc"> <path>`.  Streamed:
 * line-aligned chunks (MK_INGEST_CHUNK_BYTES, default 16 MB) go through a small ring of pinned buffers, the
 * copy of chunk i+1 and the read of chunk i+2 run under the sketch of chunk i; host memory does not grow
 * with the input.  A plain (uncompressed) file is read directly with parallel pread() instead of a pipe. */
int mk_fastq_koc_file(mk_ctx *ctx, const char *path, const char *pipecmd, mk_sketch *out);

/* ---- FASTQ without -A (`dist -Q q -n m`): fastq2co() + write_fqco2file(), iseq2comem.c:323-419, 596-621
 * (call site command_dist.c:386-387).  A base counts iff it is ACGT and the quality byte in the same column of
 * the record's fourth line, compared as a signed char, is >= quality (reference default 0); codes seen at least
 * min_occurrence times (1..14, default 1) are written in slot order, without counts.  Lines are read with
 * fgets(.., 20000, ..) there: a line of 19999 bytes or more is refused (MK_ERR_LONG_LINE).  Undefined in the
 * reference and here: a first record without its four lines, a quality line shorter than its sequence line. */
int mk_fastq_co_device(mk_ctx *ctx, const void *d_text, size_t nbytes, int quality, int min_occurrence, mk_sketch *out);
int mk_fastq_co_host(mk_ctx *ctx, const void *h_text, size_t nbytes, int quality, int min_occurrence, mk_sketch *out);
int mk_fastq_co_file(mk_ctx *ctx, const char *path, const char *pipecmd, int quality, int min_occurrence, mk_sketch *out);

/* ---- FASTA genomes (`dist` without -A; MarkerDB-build sketching) ------------------------- */
/* mk_ctx_set_dedup(ctx, 1) = `dist -u`: the following mk_fasta_co_* calls follow uniq_fasta2co()
 * (iseq2comem.c:729-828) and write only the codes that occur ONCE in their genome. */
int mk_ctx_set_dedup(mk_ctx *ctx, int on);
/* A batch of n_files FASTA texts concatenated in one buffer; file i is bytes
 * [offsets[i], offsets[i+1]).  out[i] receives the sketch of file i (counts == NULL). */
int mk_fasta_co_device(mk_ctx *ctx, const void *d_text, const uint64_t *offsets, int n_files, mk_sketch *out);
int mk_fasta_co_host(mk_ctx *ctx, const void *h_text, const uint64_t *offsets, int n_files, mk_sketch *out);
int mk_fasta_co_file(mk_ctx *ctx, const char *path, const char *pipecmd, mk_sketch *out);
/* The file loop of run_stageI() (command_dist.c:365, 397-398) as one call: n_files genomes, read by a pool
 * of threads into a pinned batch buffer and sketched in batches of at most MK_FASTA_BATCH_BYTES (1 GB). */
int mk_fasta_co_files(mk_ctx *ctx, const char *const *paths, int n_files, const char *pipecmd, mk_sketch *out);

void mk_sketch_free(mk_sketch *s);
/* on: the sketches of the following calls are not copied out of the pinned block the results arrive in (no
 * per-component malloc + memcpy): what a caller wants that writes the arrays straight to files or only needs the
 * resident copy for mk_composite_component_last(). */
int mk_ctx_set_borrowed_output(mk_ctx *ctx, int on);

/* ---- composite --------------------------------------------------------------------------- */
/* One (query, component) step of get_species_abundance(): for every species s and every
 * MarkerDB code ref_codes[ref_index[s] .. ref_index[s+1]) present in
 * qry_codes[qry_lo .. qry_hi), append that code's query count.  Results are ACCUMULATED into a
 * library-owned per-context hit store so that components can be fed one after another
 * (call mk_composite_begin first).  All pointers are host memory. */
int mk_composite_begin(mk_ctx *ctx, int n_species);
int mk_composite_component(mk_ctx *ctx, const uint32_t *ref_codes, const uint64_t *ref_index, int n_species,
                           const uint32_t *qry_codes, const uint16_t *qry_counts, uint64_t qry_lo,
                           uint64_t qry_hi);
/* Resident MarkerDB (serving many samples against one database): mk_markerdb_load() uploads
 * component `component` once and keeps it on the device; mk_composite_component_resident() is
 * mk_composite_component() against that copy.  The reference re-reads the MarkerDB from disk on
 * every `composite` run (command_composite.c:500-530); this is the same intersection without the
 * repeated transfer.  mk_markerdb_unload() (or mk_ctx_destroy) releases the copies. */
int mk_markerdb_load(mk_ctx *ctx, int component, const uint32_t *ref_codes, const uint64_t *ref_index, int n_species);
int mk_markerdb_unload(mk_ctx *ctx);
int mk_composite_component_resident(mk_ctx *ctx, int component, const uint32_t *qry_codes,
                                    const uint16_t *qry_counts, uint64_t qry_lo, uint64_t qry_hi);
/* mk_composite_component_resident() with the query taken from the device: component `component` of the
 * -A sketch this context produced last (mk_fastq_koc_* or mk_runs_finalize_*): nothing is uploaded. */
int mk_composite_component_last(mk_ctx *ctx, int component);
/* Per-species order statistics over the accumulated hits, the integers the reference prints
 * (command_composite.c:598-624); the two float ratios are sum/n and lastsum/lastn. */
typedef struct mk_species_stat {
    int32_t n;          /* matched k-mers                                                   */
    int32_t sum;        /* int sum of counts (32-bit wrap like the reference's int)        */
    int32_t lastsum;    /* sum over the 98th..99th percentile window                        */
    int32_t lastn;
    int32_t median;     /* a[n/2] of the 1-based ascending array                            */
    int32_t max;        /* a[n]                                                             */
} mk_species_stat;
int mk_composite_stats(mk_ctx *ctx, mk_species_stat *stats /* [n_species] */);
/* The species_coverage lines exactly as command_composite.c:582-624 prints them (host code, no device
 * work): species by matched k-mers descending, ties in index order, stop below 6 matches.  Returns the
 * text length (excluding the NUL); at most cap bytes are written, call with buf = NULL to size it. */
size_t mk_format_species_coverage(const char *qry_name, const char *const *ref_names, const mk_species_stat *stats,
                                  int n_species, char *buf, size_t cap);
/* Raw hit lists in reference layout: lists[s][0] = n, lists[s][1..n] = counts in MarkerDB code
 * order (component-major).  Library-owned until the next mk_composite_begin/destroy. */
int mk_composite_hits(mk_ctx *ctx, const int32_t *const **lists);

/* ---- multi-GPU building blocks (one context per rank) ------------------------------------ */
/* Rank-local partial sketch of a contiguous shard of the FASTQ text.  pos_base = global byte
 * offset of the shard, line_base = number of '\n' bytes before it (its low 2 bits select the
 * record phase), is_last = shard ends the file (trailing-partial-record rule applies).
 * Produces device-resident runs sorted by code: (code u64, firstpos u64, count u32).
 * The d_* pointers are library-owned and valid until the next call on this context. */
typedef struct mk_runs {
    uint64_t n;
    const uint64_t *d_code;
    const uint64_t *d_firstpos;
    const uint32_t *d_count;
} mk_runs;
int mk_fastq_partial_device(mk_ctx *ctx, const void *d_text, size_t nbytes, uint64_t pos_base, uint64_t line_base,
                            int is_last, mk_runs *runs);
/* Same with the shard in HOST memory (pinned for full speed): uploaded and sketched chunk by chunk like
 * mk_fastq_koc_host, into a library-owned device buffer. */
int mk_fastq_partial_host(mk_ctx *ctx, const void *h_text, size_t nbytes, uint64_t pos_base, uint64_t line_base,
                          int is_last, mk_runs *runs);
/* Merge runs (possibly from several ranks, concatenated in device memory, any order): sum the
 * counts per code with saturation at 65535, keep the minimum firstpos, then reproduce the
 * reference's hash-slot order and split by component. */
int mk_runs_finalize_device(mk_ctx *ctx, const uint64_t *d_code, const uint64_t *d_firstpos, const uint32_t *d_count,
                            uint64_t n, mk_sketch *out);
/* Same for runs whose codes are already distinct (the concatenation of code ranges merged by their
 * owners): skips the accumulation.  Undefined result if a code occurs twice. */
int mk_runs_finalize_distinct_device(mk_ctx *ctx, const uint64_t *d_code, const uint64_t *d_firstpos,
                                     const uint32_t *d_count, uint64_t n, mk_sketch *out);
/* Merge only (no ordering): device-resident reduced runs sorted by code. */
int mk_runs_merge_device(mk_ctx *ctx, const uint64_t *d_code, const uint64_t *d_firstpos, const uint32_t *d_count,
                         uint64_t n, mk_runs *merged);
/* number of '\n' bytes in a device buffer (to derive line_base of the next shard) */
int mk_count_newlines_device(mk_ctx *ctx, const void *d_text, size_t nbytes, uint64_t *count);

/* ---- `set -g / -q / -i`: the MarkerDB build from genome sketches, one component per call ---------------
 * Replaces grouping_genomes() (command_set.c:831-1003), uniq_sketch_union() (:427-512) and sketch_operate()
 * (:322-423).  All pointers are host memory; outputs are malloc'ed by the library (release with mk_free()).
 * mk_set_group: genome g (codes[index[g] .. index[g+1])) belongs to output taxon taxon_of_genome[g] (< 0: skipped);
 *   every taxon gets the union of its genomes' codes in the order of the reference's per-taxon hash table
 *   (primer[LOG2(1.5 * codes) - 7] slots, 32-bit wrap-around double hashing, code 0 never stored);
 * mk_set_uniq_union: the codes that occur exactly once in `codes`, ascending;
 * mk_set_operate: every sketch keeps, in order, its codes that are (intersect != 0) / are not (0) in `pan`. */
int mk_set_group(mk_ctx *ctx, const uint32_t *codes, const uint64_t *index, int n_genomes, const int32_t *taxon_of_genome,
                 int n_taxa, uint32_t **out_codes, uint64_t *out_index /* [n_taxa + 1] */);
int mk_set_uniq_union(mk_ctx *ctx, const uint32_t *codes, uint64_t n, uint32_t **out, uint64_t *n_out);
int mk_set_operate(mk_ctx *ctx, const uint32_t *pan, uint64_t n_pan, const uint32_t *codes, const uint64_t *index,
                   int n_sketches, int intersect, uint32_t **out_codes, uint64_t *out_index /* [n_sketches + 1] */);
void mk_free(void *p);

/* ---- `dist -r <ref> <qry>`: shared k-mer counts -------------------------------------------
 * Replaces the inverted index of combco2mco() (co2mco.c:12-86) and the probe loop of mco_cbdco_nobin_dist()
 * (command_dist.c:1031-1046) for one component: counts[q * n_ref + r] += number of codes of query sketch q
 * (qry_codes[qry_index[q] .. qry_index[q+1])) that reference sketch r holds.  All pointers are host memory;
 * `counts` ([n_qry * n_ref], zeroed by the caller) accumulates over the components.  qry_ctx_ct (may be NULL):
 * queries whose code count is 0 are skipped like the reference does.  The distance table is printed by the
 * host from counts and the two ctx_ct lists (host/mkssd_main.c, command_dist.c:1531-1680). */
int mk_shared_counts(mk_ctx *ctx, const uint32_t *ref_codes, const uint64_t *ref_index, int n_ref,
                     const uint32_t *qry_codes, const uint64_t *qry_index, int n_qry, const uint32_t *qry_ctx_ct,
                     uint32_t *counts);

/* ---- the multi-GPU step inside the library (NCCL over NVLink; one context per rank / GPU) -------- */
/* mk_comm_unique_id(): 128 bytes from ncclGetUniqueId() on one rank, handed to all ranks by the host's own
 * means (MPI / torch.distributed / a file); mk_comm_init() joins the communicator (ncclCommInitRank).  NCCL is
 * loaded with dlopen("libnccl.so.2") on first use. */
int mk_comm_unique_id(void *id, size_t bytes);
int mk_comm_init(mk_ctx *ctx, const void *id, int rank, int world);
int mk_comm_destroy(mk_ctx *ctx);
/* This rank's slice of a MarkerDB component: the codes whose full code falls into its code range (the same
 * boundaries the runs are exchanged on), kept resident; also records every rank's slice size.  Every rank calls
 * it with the WHOLE component. */
int mk_markerdb_load_sharded(mk_ctx *ctx, int component, const uint32_t *ref_codes, const uint64_t *ref_index,
                             int n_species);
/* The whole sharded step for this rank's shard of the FASTQ text (collective: every rank calls it).  Runs go to
 * the owner of their code range in one grouped ncclSend/ncclRecv step (blocks of `max_runs` slots per pair, count
 * in a header), owners merge, probe their MarkerDB slice (if one is loaded) and send hits and merged runs to
 * rank 0.  A block that does not fit — one sent, or an owner's merged range — fails the step with MK_ERR_NOMEM on
 * EVERY rank (the flag is reduced inside the step; nothing is truncated).  The code ranges are quantiles of the
 * law the codes of unbiased sequence follow, so a block holds about 1/world of a shard's runs.
 * mk_comm_last_block_need(): the largest block this rank saw in the last step (also after MK_ERR_NOMEM): the
 * maximum over the ranks, plus headroom, is the `max_runs` of the next batch.
 * On rank 0: `out` = the sketch of the whole file in reference order, `stats` (may be NULL) = per-species
 * statistics; other ranks pass NULL / get nothing. */
int mk_comm_last_block_need(mk_ctx *ctx, uint64_t *need);
int mk_fastq_koc_sharded_device(mk_ctx *ctx, const void *d_text, size_t nbytes, uint64_t pos_base, uint64_t line_base,
                                int is_last, uint64_t max_runs, mk_sketch *out, mk_species_stat *stats);
int mk_fastq_koc_sharded_host(mk_ctx *ctx, const void *h_text, size_t nbytes, uint64_t pos_base, uint64_t line_base,
                              int is_last, uint64_t max_runs, mk_sketch *out, mk_species_stat *stats);

/* ---- synthetic workload generator on the device (bench / tests; mkssd_synth.h) ----------- */
struct mks_params;
/* Writes FASTQ records [r0, r1) into d_out (capacity bytes); *written = bytes produced. */
int mk_synth_fastq_device(mk_ctx *ctx, const struct mks_params *P, const uint32_t *cdf32, const uint32_t *cdf_species,
                          uint64_t r0, uint64_t r1, void *d_out, size_t capacity, size_t *written);
/* Writes the FASTA text of species [s0, s1) back to back; offsets[s1-s0+1] (host) receives the
 * file boundaries. */
int mk_synth_fasta_device(mk_ctx *ctx, const struct mks_params *P, uint32_t s0, uint32_t s1, void *d_out,
                          size_t capacity, uint64_t *offsets);

/* Host-side helpers of the generator: default parameters + abundance CDF (arrays are malloc'ed,
 * release with mk_synth_free), text sizes, and a deterministic .shuf permutation (same file
 * format as `metakssd shuffle`, seeded instead of srand(time)). */
int mk_synth_build(struct mks_params *P, uint64_t seed, uint32_t n_species, uint32_t genome_len, uint32_t read_len,
                   uint32_t **cdf32, uint32_t **species);
void mk_synth_free(void *p);
uint64_t mk_synth_fastq_bytes(const struct mks_params *P, uint64_t r0, uint64_t r1);
uint64_t mk_synth_fasta_bytes(const struct mks_params *P, uint32_t s);
void mk_synth_shuf_perm(uint64_t seed, int subk, int32_t *perm);
int32_t mk_synth_shuf_id(uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif /* MKSSD_B200_H */
