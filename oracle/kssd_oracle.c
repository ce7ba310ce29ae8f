/*
 * kssd_oracle.c — TEST INFRASTRUCTURE ONLY.  Not part of the product.
 *
 * A plain-C, single-threaded restatement of MetaKSSD's hot path (FASTQ -> KSSD sketch with
 * k-mer counts, FASTA -> KSSD sketch, and the `composite` MarkerDB intersection), written from
 * the behaviour of the reference (yhg926/MetaKSSD v2.21, mounted at /root/reference).  It is
 * the checker for the CUDA path: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product library (libmkssd_b200.so)
 * never links or calls it and has no CPU fallback.
 *
 * Parity status: PINNED against the reference binary itself (oracle/_ref/metakssd, built from
 * /root/reference by oracle/Makefile) run at `-p 1`; the comparison script is
 * tests/golden/make_golden.py and the resulting vectors are committed under tests/golden/.
 * The reference has no tests or golden vectors of its own.
 *
 * Each function cites the reference lines it follows.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include "mkssd_synth.h"

typedef unsigned long long u64;

/* hash-table sizes: the table of /root/reference/global_basic.c:75-82 (primes just below
 * 2^8 .. 2^32), regenerated here by primality search instead of being listed. */
static int is_prime_u64(u64 n)
{
    if (n < 2) return 0;
    if (n % 2 == 0) return n == 2;
    for (u64 d = 3; d * d <= n; d += 2)
        if (n % d == 0) return 0;
    return 1;
}
static uint32_t prime_below_pow2(int e)
{
    u64 n = (1ull << e) - 1;
    while (!is_prime_u64(n)) n--;
    return (uint32_t)n;
}

typedef struct ko_params {
    int k, subk, drlevel;
    int outctx;          /* half outer context length = k - subk                           */
    int TL;              /* k-mer length 2k                                                */
    int crvs_shift;      /* where a new base enters the reverse-complement register       */
    u64 tupmask;         /* 4k low bits                                                    */
    u64 domask;          /* inner substring (4*subk bits) at bit 2*outctx                  */
    u64 undomask;        /* left outer context                                             */
    u64 lowmask;         /* right outer context (2*outctx low bits)                        */
    int dim_end;         /* pass iff shuf[dim] < dim_end                                   */
    uint32_t hashsize;
    uint32_t hashlimit;
    int component_num;
    int comp_code_bits;
} ko_params;

/* iseq2comem.c:54-86 (seq2co_global_var_initial) + command_dist.c:286-315 (get_hashsz).
 * Returns 0, or -1 when the primer index falls outside 0..24 (reference: fatal err()). */
int ko_params_init(ko_params *p, int k, int subk, int drlevel)
{
    memset(p, 0, sizeof(*p));
    p->k = k; p->subk = subk; p->drlevel = drlevel;
    p->outctx = k - subk;
    p->TL = 2 * k;
    p->crvs_shift = 4 * k - 2;
    p->tupmask = (4 * k >= 64) ? ~0ull : ((1ull << (4 * k)) - 1);
    p->domask = ((1ull << (4 * subk)) - 1) << (2 * p->outctx);
    p->undomask = ((1ull << (2 * p->outctx)) - 1) << (2 * (k + subk));
    p->lowmask = (1ull << (2 * p->outctx)) - 1;
    u64 subspace = 1ull << (4 * (subk - drlevel));
    p->dim_end = (int)(subspace > 4096 ? subspace : 4096); /* MIN_SUBCTX_DIM_SMP_SZ */
    int primer_ind = 4 * (k - drlevel) - 8 /* CTX_SPC_USE_L */ - 7;
    if (primer_ind < 0 || primer_ind > 24) return -1;
    p->hashsize = prime_below_pow2(primer_ind + 8);
    p->hashlimit = (uint32_t)(p->hashsize * 0.6); /* LD_FCTR, double multiply then truncate */
    p->component_num = (k - drlevel > 8) ? (int)(1ul << (4 * (k - drlevel - 8))) : 1;
    p->comp_code_bits = (k - drlevel > 8) ? 4 * (k - drlevel - 8) : 0;
    return 0;
}

static inline int base_code(unsigned char c)
{
    /* global_basic.c:62-69 (Basemap): A/a C/c G/g T/t -> 0..3, everything else "default" */
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
    }
}

/* Shared by both sketchers: canonical k-mer -> sketch code, or -1 when filtered out.
 * iseq2comem.c:690-699 (FASTQ -A) and :283-293 (FASTA) are the same arithmetic. */
static inline long long kmer_to_code(const ko_params *p, const int32_t *shuf, u64 fwd, u64 rc)
{
    u64 u = fwd < rc ? fwd : rc;
    u64 dim = (u & p->domask) >> (2 * p->outctx);
    long long pf = shuf[dim];
    if (pf >= p->dim_end || pf < 0) return -1;
    u64 code = (((u & p->undomask) + ((u & p->lowmask) << (2 * p->TL - 4 * p->outctx))) >>
                (4 * p->drlevel)) + (u64)pf;
    return (long long)code;
}

/* ---------------- open-addressing table with the reference's probe sequence --------------
 * global_basic.h:282-284: slot_i = (K % hs + i * (1 + K % (hs-1))) % hs in 64-bit arithmetic.
 * The reference allocates hashsize*8 bytes and scans every slot when writing; here the table
 * is calloc'ed (untouched pages stay virtual) and occupied slots are remembered and sorted,
 * which yields the same ascending-slot order without the full scan. */
typedef struct {
    u64 *tab;        /* hashsize entries; 0 = empty                                         */
    uint32_t *used;  /* occupied slot list                                                  */
    size_t n_used, cap_used;
} ko_table;

static int table_init(ko_table *t, uint32_t hashsize)
{
    t->tab = (u64 *)calloc(hashsize, sizeof(u64));
    t->cap_used = 1 << 16;
    t->used = (uint32_t *)malloc(t->cap_used * sizeof(uint32_t));
    t->n_used = 0;
    return (t->tab && t->used) ? 0 : -1;
}
static void table_free(ko_table *t) { free(t->tab); free(t->used); }
static int table_note_used(ko_table *t, uint32_t slot)
{
    if (t->n_used == t->cap_used) {
        t->cap_used *= 2;
        t->used = (uint32_t *)realloc(t->used, t->cap_used * sizeof(uint32_t));
        if (!t->used) return -1;
    }
    t->used[t->n_used++] = slot;
    return 0;
}
static int cmp_u32(const void *a, const void *b)
{
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return x < y ? -1 : x > y;
}

typedef struct ko_sketch {
    size_t n;          /* number of codes                                                  */
    u64 *codes;        /* full 4(k-L)-bit codes, in reference on-disk (ascending slot) order */
    uint16_t *counts;  /* NULL for FASTA sketches                                          */
    uint32_t *slots;   /* slot of each code (diagnostic)                                   */
    int status;        /* 0 ok; 1 = "context space is too crowd" (keycount > hashlimit);
                          2 = a FASTQ line of >= 4095 bytes (reference behaviour undefined) */
} ko_sketch;

void ko_sketch_free(ko_sketch *s)
{
    if (!s) return;
    free(s->codes); free(s->counts); free(s->slots); free(s);
}

/* fgets(buf, 4096, fp) on an in-memory text: returns length consumed (0 = EOF) */
static size_t mem_fgets(const char *text, size_t n, size_t at, int *too_long)
{
    size_t i = at;
    size_t limit = at + 4095 < n ? at + 4095 : n;
    while (i < limit) {
        if (text[i++] == '\n') return i - at;
    }
    if (i - at == 4095) *too_long = 1; /* filled the buffer without meeting a newline */
    return i - at;
}

/* iseq2comem.c:657-727 (mt_shortreads2koc) run with one thread, followed by the slot-order
 * dump of iseq2comem.c:516-562 (write_fqkoc2files).  `text` is the whole decompressed FASTQ. */
ko_sketch *ko_fastq_koc(const ko_params *p, const int32_t *shuf, const char *text, size_t n)
{
    ko_sketch *out = (ko_sketch *)calloc(1, sizeof(ko_sketch));
    ko_table T;
    if (!out || table_init(&T, p->hashsize)) return NULL;
    uint32_t keycount = 0;
    size_t at = 0;
    int too_long = 0;
    while (at < n) {
        /* one record = four fgets() calls, all of which must succeed (:673) */
        size_t l1 = mem_fgets(text, n, at, &too_long); if (!l1) break;
        size_t s0 = at + l1;
        size_t l2 = mem_fgets(text, n, s0, &too_long); if (!l2) break;
        size_t l3 = mem_fgets(text, n, s0 + l2, &too_long); if (!l3) break;
        size_t l4 = mem_fgets(text, n, s0 + l2 + l3, &too_long); if (!l4) break;
        at = s0 + l2 + l3 + l4;
        if (too_long) { out->status = 2; break; }
        /* per read (:677-720): the scan stops at '\n' */
        int base = 1;
        u64 fwd = 0, rc = 0;
        for (size_t i = s0; i < s0 + l2 && text[i] != '\n'; i++) {
            int b = base_code((unsigned char)text[i]);
            if (b < 0) { base = 1; continue; }
            fwd = ((fwd << 2) | (u64)b) & p->tupmask;
            rc = (rc >> 2) + (((u64)b ^ 3ull) << p->crvs_shift);
            base++;
            if (base <= p->TL) continue;
            long long c = kmer_to_code(p, shuf, fwd, rc);
            if (c < 0) continue;
            u64 code = (u64)c;
            u64 h1 = code % p->hashsize, h2 = 1 + code % (p->hashsize - 1);
            for (u64 i2 = 0; i2 < p->hashsize; i2++) {
                uint32_t slot = (uint32_t)((h1 + i2 * h2) % p->hashsize);
                if (T.tab[slot] == 0) {
                    T.tab[slot] = (code << 16) + 1; /* OCCRC_BIT = 16 */
                    table_note_used(&T, slot);
                    if (++keycount > p->hashlimit) out->status = 1;
                    break;
                }
                if ((T.tab[slot] >> 16) == code) {
                    if ((T.tab[slot] & 0xFFFF) < 0xFFFF) T.tab[slot] += 1;
                    break;
                }
            }
            if (out->status) break;
        }
        if (out->status) break;
    }
    qsort(T.used, T.n_used, sizeof(uint32_t), cmp_u32);
    out->n = T.n_used;
    out->codes = (u64 *)malloc(sizeof(u64) * (T.n_used + 1));
    out->counts = (uint16_t *)malloc(sizeof(uint16_t) * (T.n_used + 1));
    out->slots = (uint32_t *)malloc(sizeof(uint32_t) * (T.n_used + 1));
    for (size_t i = 0; i < T.n_used; i++) {
        u64 e = T.tab[T.used[i]];
        out->codes[i] = e >> 16;
        out->counts[i] = (uint16_t)(e & 0xFFFF);
        out->slots[i] = T.used[i];
    }
    table_free(&T);
    return out;
}

/* iseq2comem.c:218-315 (fasta2co) + :625-652 (wrt_co2cmpn_use_inn_subctx): set semantics, newline
 * and CR transparent, '>' skips to end of line, any other byte resets the window; code 0 is
 * stored as "empty" and therefore never written.  The reference's 65536-byte refill logic is
 * stream-equivalent except for an out-of-bounds read when a header straddles a refill
 * (:261-271); that corner is restated as the intended "skip to end of line". */
ko_sketch *ko_fasta_co(const ko_params *p, const int32_t *shuf, const char *text, size_t n)
{
    ko_sketch *out = (ko_sketch *)calloc(1, sizeof(ko_sketch));
    ko_table T;
    if (!out || table_init(&T, p->hashsize)) return NULL;
    uint32_t keycount = 0;
    long long base = 1;
    u64 fwd = 0, rc = 0;
    for (size_t i = 0; i < n; i++) {
        unsigned char ch = (unsigned char)text[i];
        int b = base_code(ch);
        if (b >= 0) {
            fwd = ((fwd << 2) | (u64)b) & p->tupmask;
            rc = (rc >> 2) + (((u64)b ^ 3ull) << p->crvs_shift);
            base++;
        } else if (ch == '\n' || ch == '\r') {
            continue;
        } else if (ch == '>') {
            while (i < n && text[i] != '\n') i++;
            base = 1;
            continue;
        } else {
            base = 1;
            continue;
        }
        if (base <= p->TL) continue;
        long long c = kmer_to_code(p, shuf, fwd, rc);
        if (c < 0) continue;
        u64 code = (u64)c;
        u64 h1 = code % p->hashsize, h2 = 1 + code % (p->hashsize - 1);
        for (u64 i2 = 0; i2 < p->hashsize; i2++) {
            uint32_t slot = (uint32_t)((h1 + i2 * h2) % p->hashsize);
            if (T.tab[slot] == 0) {
                T.tab[slot] = code;
                if (code != 0) table_note_used(&T, slot);
                if (++keycount > p->hashlimit) out->status = 1;
                break;
            }
            if (T.tab[slot] == code) break;
        }
        if (out->status) break;
    }
    qsort(T.used, T.n_used, sizeof(uint32_t), cmp_u32);
    out->n = T.n_used;
    out->codes = (u64 *)malloc(sizeof(u64) * (T.n_used + 1));
    out->slots = (uint32_t *)malloc(sizeof(uint32_t) * (T.n_used + 1));
    for (size_t i = 0; i < T.n_used; i++) {
        out->codes[i] = T.tab[T.used[i]];
        out->slots[i] = T.used[i];
    }
    table_free(&T);
    return out;
}

/* ---------------- FASTQ without -A: fastq2co() + write_fqco2file() -------------------------------
 * iseq2comem.c:323-419: the first record is read unconditionally (four fgets(.., LEN = 20000, ..)); every further
 * record is read when the previous sequence line ends and is used only if feof() is still false after its four
 * fgets() calls, i.e. its fourth line ended with a newline.  A base counts iff it is ACGT and the quality byte in
 * the same column, compared as a (signed) char, is >= Q.  Table entries are code << 4 | count; CT_MAX (15) in the
 * low bits marks "occurred at least M times" (set at once when M == 1); write_fqco2file() (:596-621) writes the
 * marked codes in ascending slot order, no counts.  keycount is never incremented in the reference, so the
 * "too crowd" error cannot fire.  A line of LEN - 1 = 19999 bytes or more is split by fgets: status 2. */
typedef struct { const char *t; size_t n, at; int eof; } ko_reader;
static int rd_fgets(ko_reader *r, char *buf, int size, int *too_long)
{
    int len = 0;
    while (len < size - 1) {
        if (r->at >= r->n) { r->eof = 1; break; }
        char c = r->t[r->at++];
        buf[len++] = c;
        if (c == '\n') break;
    }
    if (len == 0) return 0;                 /* NULL: the buffer keeps its old content */
    if (len == size - 1 && buf[len - 1] != '\n') *too_long = 1;
    buf[len] = 0;
    return 1;
}

ko_sketch *ko_fastq_co(const ko_params *p, const int32_t *shuf, const char *text, size_t n, int Q, int M)
{
    enum { LEN = 20000 };
    ko_sketch *out = (ko_sketch *)calloc(1, sizeof(ko_sketch));
    ko_table T;
    if (!out || table_init(&T, p->hashsize)) return NULL;
    char *seq = (char *)calloc(LEN + 10, 1), *qual = (char *)calloc(LEN + 10, 1);
    ko_reader R = {text, n, 0, 0};
    int too_long = 0;
    rd_fgets(&R, seq, LEN, &too_long); rd_fgets(&R, seq, LEN, &too_long);
    rd_fgets(&R, qual, LEN, &too_long); rd_fgets(&R, qual, LEN, &too_long);
    int base = 1;
    u64 fwd = 0, rc = 0;
    int sl = (int)strlen(seq);
    for (int pos = 0; pos < sl; pos++) {
        if (too_long) { out->status = 2; break; }
        if (seq[pos] == '\n') {
            rd_fgets(&R, seq, LEN, &too_long); rd_fgets(&R, seq, LEN, &too_long);
            rd_fgets(&R, qual, LEN, &too_long); rd_fgets(&R, qual, LEN, &too_long);
            sl = (int)strlen(seq);
            if (!R.eof) { base = 1; pos = -1; continue; }
            break;
        }
        int b = base_code((unsigned char)seq[pos]);
        if (b >= 0 && qual[pos] >= Q) {        /* char comparison: bytes >= 0x80 are negative */
            fwd = ((fwd << 2) | (u64)b) & p->tupmask;
            rc = (rc >> 2) + (((u64)b ^ 3ull) << p->crvs_shift);
            base++;
        } else { base = 1; continue; }
        if (base <= p->TL) continue;
        long long c = kmer_to_code(p, shuf, fwd, rc);
        if (c < 0) continue;
        u64 code = (u64)c;
        u64 h1 = code % p->hashsize, h2 = 1 + code % (p->hashsize - 1);
        for (u64 i2 = 0; i2 < p->hashsize; i2++) {
            uint32_t slot = (uint32_t)((h1 + i2 * h2) % p->hashsize);
            if (T.tab[slot] == 0) {
                T.tab[slot] = M == 1 ? ((code << 4) | 15ull) : ((code << 4) + 1ull);
                table_note_used(&T, slot);
                break;
            }
            if ((T.tab[slot] >> 4) == code) {
                if ((T.tab[slot] & 15ull) == 15ull) break;
                T.tab[slot] += 1;
                if (!((T.tab[slot] & 15ull) < (u64)M)) T.tab[slot] |= 15ull;
                break;
            }
        }
    }
    if (too_long) out->status = 2;
    qsort(T.used, T.n_used, sizeof(uint32_t), cmp_u32);
    out->codes = (u64 *)malloc(sizeof(u64) * (T.n_used + 1));
    out->slots = (uint32_t *)malloc(sizeof(uint32_t) * (T.n_used + 1));
    size_t m = 0;
    for (size_t i = 0; i < T.n_used; i++) {
        u64 e = T.tab[T.used[i]];
        if ((e & 15ull) != 15ull) continue;
        out->codes[m] = e >> 4;
        out->slots[m] = T.used[i];
        m++;
    }
    out->n = m;
    free(seq); free(qual);
    table_free(&T);
    return out;
}

/* ---------------- FASTA with -u: uniq_fasta2co() + wrt_co2cmpn_use_inn_subctx() -------------------
 * iseq2comem.c:729-828: the tokenizer and arithmetic of fasta2co(); a code met again gets bit 63 set, and the
 * writer (:640) skips entries with that bit: only codes occurring ONCE in the file are written, in the slot
 * order all distinct codes produced.  Code 0 looks like an empty slot and is lost, as in fasta2co(). */
ko_sketch *ko_fasta_co_uniq(const ko_params *p, const int32_t *shuf, const char *text, size_t n)
{
    const u64 HI = 0x8000000000000000ull;
    ko_sketch *out = (ko_sketch *)calloc(1, sizeof(ko_sketch));
    ko_table T;
    if (!out || table_init(&T, p->hashsize)) return NULL;
    uint32_t keycount = 0;
    long long base = 1;
    u64 fwd = 0, rc = 0;
    for (size_t i = 0; i < n; i++) {
        unsigned char ch = (unsigned char)text[i];
        int b = base_code(ch);
        if (b >= 0) {
            fwd = ((fwd << 2) | (u64)b) & p->tupmask;
            rc = (rc >> 2) + (((u64)b ^ 3ull) << p->crvs_shift);
            base++;
        } else if (ch == '\n' || ch == '\r') {
            continue;
        } else if (ch == '>') {
            while (i < n && text[i] != '\n') i++;
            base = 1;
            continue;
        } else {
            base = 1;
            continue;
        }
        if (base <= p->TL) continue;
        long long c = kmer_to_code(p, shuf, fwd, rc);
        if (c < 0) continue;
        u64 code = (u64)c;
        u64 h1 = code % p->hashsize, h2 = 1 + code % (p->hashsize - 1);
        for (u64 i2 = 0; i2 < p->hashsize; i2++) {
            uint32_t slot = (uint32_t)((h1 + i2 * h2) % p->hashsize);
            if (T.tab[slot] == 0) {
                T.tab[slot] = code;
                if (code != 0) table_note_used(&T, slot);
                if (++keycount > p->hashlimit) out->status = 1;
                break;
            }
            if ((T.tab[slot] | HI) == (code | HI)) { T.tab[slot] |= HI; break; }
        }
        if (out->status) break;
    }
    qsort(T.used, T.n_used, sizeof(uint32_t), cmp_u32);
    out->codes = (u64 *)malloc(sizeof(u64) * (T.n_used + 1));
    out->slots = (uint32_t *)malloc(sizeof(uint32_t) * (T.n_used + 1));
    size_t m = 0;
    for (size_t i = 0; i < T.n_used; i++) {
        u64 e = T.tab[T.used[i]];
        if (e & HI) continue;
        out->codes[m] = e;
        out->slots[m] = T.used[i];
        m++;
    }
    out->n = m;
    table_free(&T);
    return out;
}

/* Split a sketch into the per-component file arrays of write_fqkoc2files / wrt_co2cmpn:
 * component = code % component_num, file value = (uint32)(code >> comp_code_bits).
 * comp_of[i] receives the component, filecode[i] the 32-bit value. */
void ko_split_components(const ko_params *p, const ko_sketch *s, uint32_t *comp_of, uint32_t *filecode)
{
    for (size_t i = 0; i < s->n; i++) {
        comp_of[i] = (uint32_t)(s->codes[i] % (u64)p->component_num);
        filecode[i] = (uint32_t)(s->codes[i] >> p->comp_code_bits);
    }
}

/* ---------------- `set -g / -q / -i`: the MarkerDB pipeline (command_set.c) ------------------------------
 * ko_organize_taxa: organize_taxf() (command_set.c:635-704) — taxa live in an open-addressing table of
 * nextPrime((int)(lines / 0.6)) slots probed with HASH(taxid, n, size) in int arithmetic; the output order of the
 * taxa is ascending slot.  taxon_of[g] receives the output position of genome g's taxon (or -1 for taxid 0, which
 * grouping_genomes() skips; its position is still counted).  Returns the number of taxa including a taxid-0 one. */
static int next_prime_int(int n);
int ko_organize_taxa(const int *taxid, int n_genomes, int *taxon_of, int *taxid_of_taxon)
{
    int hashsz = next_prime_int((int)((double)n_genomes / 0.6));
    int *tab = (int *)malloc(sizeof(int) * (size_t)(hashsz > 0 ? hashsz : 1));
    for (int i = 0; i < hashsz; i++) tab[i] = -1;
    int *slot_of = (int *)malloc(sizeof(int) * (size_t)n_genomes);
    for (int g = 0; g < n_genomes; g++) {
        slot_of[g] = -1;
        for (int n = 0; n < hashsz; n++) {
            int hv = (taxid[g] % hashsz + n * (1 + taxid[g] % (hashsz - 1))) % hashsz;
            if (tab[hv] == -1) { tab[hv] = taxid[g]; slot_of[g] = hv; break; }
            if (tab[hv] == taxid[g]) { slot_of[g] = hv; break; }
        }
    }
    int n_taxa = 0;
    int *pos_of_slot = (int *)malloc(sizeof(int) * (size_t)(hashsz > 0 ? hashsz : 1));
    for (int i = 0; i < hashsz; i++) {
        pos_of_slot[i] = -1;
        if (tab[i] != -1) { taxid_of_taxon[n_taxa] = tab[i]; pos_of_slot[i] = n_taxa++; }
    }
    for (int g = 0; g < n_genomes; g++) taxon_of[g] = slot_of[g] < 0 ? -1 : pos_of_slot[slot_of[g]];
    free(tab); free(slot_of); free(pos_of_slot);
    return n_taxa;
}

/* grouping_genomes() (command_set.c:831-1003), one component: per taxon an open-addressing table of
 * primer[LOG2((ull)(codes * 1.5)) - 7] slots (primer[0] below 2^8), HASH() in 32-bit wrap-around arithmetic,
 * 0 = empty (so code 0 is never stored); genomes in taxfile order, codes in sketch order; written in ascending
 * slot order.  taxon_of[g] < 0 skips the genome.  out_codes needs room for index[n_genomes] values. */
void ko_set_group(const uint32_t *codes, const size_t *index, int n_genomes, const int *taxon_of, int n_taxa,
                  uint32_t *out_codes, size_t *out_index)
{
    size_t o = 0;
    out_index[0] = 0;
    for (int t = 0; t < n_taxa; t++) {
        size_t total = 0;
        for (int g = 0; g < n_genomes; g++)
            if (taxon_of[g] == t) total += index[g + 1] - index[g];
        unsigned long long x = (unsigned long long)((double)total * 1.5);
        int lg = x ? 63 - __builtin_clzll(x) : -1;
        int ind = lg > 7 ? lg - 7 : 0;
        unsigned hs = prime_below_pow2(8 + ind);
        uint32_t *tab = (uint32_t *)calloc(hs, sizeof(uint32_t));
        for (int g = 0; g < n_genomes; g++) {
            if (taxon_of[g] != t) continue;
            for (size_t i = index[g]; i < index[g + 1]; i++) {
                unsigned key = codes[i];
                for (int xx = 0; xx < (int)hs; xx++) {
                    unsigned y = (key % hs + (unsigned)xx * (1u + key % (hs - 1u))) % hs;
                    if (tab[y] == 0) { tab[y] = key; break; }
                    if (tab[y] == key) break;
                }
            }
        }
        for (unsigned y = 0; y < hs; y++)
            if (tab[y] != 0) out_codes[o++] = tab[y];
        out_index[t + 1] = o;
        free(tab);
    }
}

/* uniq_sketch_union() (command_set.c:427-512): codes that occur exactly once in the whole pan, ascending */
static int cmp_u32v(const void *a, const void *b) { uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b; return x < y ? -1 : x > y; }
size_t ko_set_uniq_union(const uint32_t *codes, size_t n, uint32_t *out)
{
    uint32_t *s = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    memcpy(s, codes, sizeof(uint32_t) * n);
    qsort(s, n, sizeof(uint32_t), cmp_u32v);
    size_t m = 0;
    for (size_t i = 0; i < n; i++)
        if ((i == 0 || s[i - 1] != s[i]) && (i + 1 == n || s[i + 1] != s[i])) out[m++] = s[i];
    free(s);
    return m;
}

/* sketch_operate() (command_set.c:322-423): every sketch keeps the codes that are (intersect) / are not (subtract)
 * in the pan, order kept */
void ko_set_operate(const uint32_t *pan, size_t n_pan, const uint32_t *codes, const size_t *index, int n_sketches,
                    int intersect, uint32_t *out_codes, size_t *out_index)
{
    uint32_t *s = (uint32_t *)malloc(sizeof(uint32_t) * (n_pan ? n_pan : 1));
    memcpy(s, pan, sizeof(uint32_t) * n_pan);
    qsort(s, n_pan, sizeof(uint32_t), cmp_u32v);
    size_t o = 0;
    out_index[0] = 0;
    for (int k = 0; k < n_sketches; k++) {
        for (size_t i = index[k]; i < index[k + 1]; i++) {
            int in = bsearch(&codes[i], s, n_pan, sizeof(uint32_t), cmp_u32v) != NULL;
            if (in == (intersect != 0)) out_codes[o++] = codes[i];
        }
        out_index[k + 1] = o;
    }
    free(s);
}

/* ---------------- composite -------------------------------------------------------------- */
/* global_basic.c:453-475 (nextPrime): smallest prime >= n by trial division up to (int)sqrt(n) */
static int next_prime_int(int n)
{
    for (;;) {
        int composite = 0;
        for (int j = 2; (long long)j * j <= n; j++)
            if (n % j == 0) { composite = 1; break; }
        if (!composite) return n;
        n++;
    }
}

/* command_composite.c:535-566 for ONE (query, component): builds the query dictionary with the
 * reference's 32-bit wrap-around probe arithmetic and appends, per species, the query count of
 * every MarkerDB code found.  hits[s] must have room for (ref_index[s+1]-ref_index[s]) values;
 * nhits[s] is incremented.  Returns 0, or -1 for a query component of exactly one code (reference:
 * modulo by zero); an empty component yields no hits like in the reference. */
int ko_composite_component(const uint32_t *ref_codes, const size_t *ref_index, int n_species,
                           const uint32_t *qry_codes, const uint16_t *qry_counts, size_t q_lo,
                           size_t q_hi, int32_t **hits, int32_t *nhits)
{
    int hash_sz = next_prime_int((int)((double)(q_hi - q_lo) / 0.6));
    if (hash_sz == 0) return 0;    /* empty query component: both loops of the reference run zero times */
    if (hash_sz == 1) return -1;   /* one code: HASH() takes K % (hash_sz - 1) = K % 0, the reference dies with SIGFPE */
    size_t *dict = (size_t *)calloc((size_t)hash_sz, sizeof(size_t));
    if (!dict) return -2;
    for (size_t idx = q_lo; idx < q_hi; idx++) {
        unsigned int key = qry_codes[idx];
        for (int i = 0; i < hash_sz; i++) {
            /* int * unsigned -> unsigned: 32-bit wrap, as in HASH() applied to 32-bit operands */
            unsigned int hv = (key % (unsigned)hash_sz + (unsigned)i * (1u + key % (unsigned)(hash_sz - 1))) %
                              (unsigned)hash_sz;
            if (dict[hv] == 0) { dict[hv] = idx + 1; break; }
        }
    }
    for (int s = 0; s < n_species; s++) {
        for (size_t ri = ref_index[s]; ri < ref_index[s + 1]; ri++) {
            unsigned int key = ref_codes[ri];
            for (int i = 0; i < hash_sz; i++) {
                unsigned int hv = (key % (unsigned)hash_sz + (unsigned)i * (1u + key % (unsigned)(hash_sz - 1))) %
                                  (unsigned)hash_sz;
                if (dict[hv] == 0) break;
                if (qry_codes[dict[hv] - 1] == key) {
                    hits[s][nhits[s]++] = qry_counts[dict[hv] - 1];
                    break;
                }
            }
        }
    }
    free(dict);
    return 0;
}

static int32_t *g_nhits_for_sort;
static int cmp_species_desc(const void *a, const void *b)
{
    return g_nhits_for_sort[*(const int *)b] - g_nhits_for_sort[*(const int *)a];
}
static int cmp_i32(const void *a, const void *b) { return *(const int32_t *)a - *(const int32_t *)b; }

/* command_composite.c:577-626: order species by hit count (glibc qsort, descending), stop at
 * the first with fewer than 6 hits, and per species print
 *   qry \t ref \t n \t mean \t mean(98..99th pct) \t median \t max
 * Writes the lines into `out` (size cap), returns bytes written.  The two ratios are computed
 * in float and printed with %f exactly as the reference does. */
size_t ko_composite_report(const char *qry_name, const char *const *ref_names, int n_species,
                           int32_t **hits, int32_t *nhits, char *out, size_t cap)
{
    int *order = (int *)malloc(sizeof(int) * (size_t)n_species);
    for (int i = 0; i < n_species; i++) order[i] = i;
    g_nhits_for_sort = nhits;
    qsort(order, (size_t)n_species, sizeof(int), cmp_species_desc);
    size_t w = 0;
    for (int r = 0; r < n_species; r++) {
        int s = order[r];
        int n = nhits[s];
        if (n < 6) break; /* MIN_KM_S */
        /* the reference keeps n in a[0] and the values in a[1..n] */
        int32_t *a = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + 1));
        a[0] = n;
        memcpy(a + 1, hits[s], sizeof(int32_t) * (size_t)n);
        qsort(a + 1, (size_t)n, sizeof(int32_t), cmp_i32);
        int sum = 0;
        for (int j = 1; j <= n; j++) sum += a[j];
        int median_idx = n / 2;
        int lo = (int)(n * 0.98);
        int lastsum = 0, lastn = 0;
        for (int j = lo; j <= n * 0.99; j++) { lastsum += a[j]; lastn++; }
        int m = snprintf(out + w, cap > w ? cap - w : 0, "%s\t%s\t%d\t%f\t%f\t%d\t%d\n", qry_name,
                         ref_names[s], n, (float)sum / n, (float)lastsum / lastn, a[median_idx], a[n]);
        if (m > 0) w += (size_t)m;
        free(a);
    }
    free(order);
    return w;
}

/* ---------------- synthetic data (thin wrappers so ctypes can reach the header) ---------- */
/* =================================================================================================
 * `dist -r <ref> <qry>`: shared k-mer counts of every query sketch with every reference sketch and the
 * distance table.
 *   ko_shared_counts()  mco_cbdco_nobin_dist() core, /root/reference/command_dist.c:1031-1046, with the inverted
 *                       index of combco2mco() (co2mco.c:12-86) replaced by what it encodes: for every code the
 *                       reference sketches that hold it.  counts[q * n_ref + r] += 1 for every code of query q
 *                       that reference r holds (one component per call, counts accumulate over components).
 *   ko_distance_out()   dist_print_nobin() + output_ctrl(), command_dist.c:1531-1680: header, per query either
 *                       every reference in order or the N with the largest metric (insertion order of :1595-1604),
 *                       lines farther than max_dist dropped.
 * ================================================================================================= */
static int cmp_u64(const void *a, const void *b) { uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b; return x < y ? -1 : x > y; }

void ko_shared_counts(const uint32_t *ref_codes, const size_t *ref_index, int n_ref, const uint32_t *qry_codes,
                      const size_t *qry_index, int n_qry, const uint32_t *qry_ctx_ct, uint32_t *counts)
{
    /* (code, ref) pairs sorted by code = the rows of the reference's mco index */
    size_t R = ref_index[n_ref];
    uint64_t *pair = malloc(sizeof(uint64_t) * (R ? R : 1));
    size_t m = 0;
    for (int r = 0; r < n_ref; r++)
        for (size_t i = ref_index[r]; i < ref_index[r + 1]; i++) pair[m++] = ((uint64_t)ref_codes[i] << 32) | (uint32_t)r;
    qsort(pair, m, sizeof(uint64_t), cmp_u64);
    for (int q = 0; q < n_qry; q++) {
        if (qry_ctx_ct && qry_ctx_ct[q] == 0) continue;                      /* command_dist.c:1033 */
        for (size_t i = qry_index[q]; i < qry_index[q + 1]; i++) {
            const uint64_t key = (uint64_t)qry_codes[i] << 32;
            size_t lo = 0, hi = m;
            while (lo < hi) { size_t mid = (lo + hi) >> 1; if (pair[mid] < key) lo = mid + 1; else hi = mid; }
            for (; lo < m && (pair[lo] >> 32) == qry_codes[i]; lo++) counts[(size_t)q * n_ref + (uint32_t)pair[lo]]++;
        }
    }
    free(pair);
}

typedef struct { int metric, outfields, correction, n_max; double max_dist; int kmerlen, dim_rd_len; } ko_dist_opt;

/* one line of output_ctrl() (command_dist.c:1637-1680); returns its length, 0 if the line is dropped */
static int ko_dist_line(char *line, size_t cap, const ko_dist_opt *o, const char *qname, const char *rname, unsigned X,
                        unsigned Y, unsigned XnY, double n_cmp)
{
    double rs = 0;
    if (o->correction) {
        unsigned a = X - XnY, b = Y - XnY;
        double pa = 1 - pow((1 - 1 / pow(4, (o->kmerlen - o->dim_rd_len))), a);
        double pb = 1 - pow((1 - 1 / pow(4, (o->kmerlen - o->dim_rd_len))), b);
        rs = pa * pb * (a + b) / (pa + pb - 2 * pa * pb);
    }
    unsigned tmp = o->metric == 0 ? X + Y - XnY : (X < Y ? X : Y);
    double metric = ((double)XnY - rs) / tmp;
    double dist = log(o->metric == 0 ? 1 / (2 * metric) + 0.5 : 1 / metric) / o->kmerlen;
    if (dist > 1) dist = 1;
    if (dist > o->max_dist) return 0;
    int len = snprintf(line, cap, "%s\t%s\t%u-%u|%u|%u\t%.6lf\t%.6lf", qname, rname, XnY, (unsigned)rs, X, Y, metric, dist);
    if (o->outfields > 0) {
        double sd = pow(metric * (1 - metric) / tmp, 0.5);
        double pv = 0.5 * erfc(metric / sd * pow(0.5, 0.5));
        len += snprintf(line + len, cap - len, "\t%E\t%E", pv, pv * n_cmp);
        if (o->outfields > 1) {
            double m1 = metric - 1.96 * sd, m2 = metric + 1.96 * sd;
            double d1 = log(o->metric == 0 ? 1 / (2 * m2) + 0.5 : 1 / m2) / o->kmerlen;
            double d2 = log(o->metric == 0 ? 1 / (2 * m1) + 0.5 : 1 / m1) / o->kmerlen;
            len += snprintf(line + len, cap - len, "\t[%.6lf,%.6lf]\t[%.6lf,%.6lf]", m1, m2, d1, d2);
        }
    }
    len += snprintf(line + len, cap - len, "\n");
    return len;
}

size_t ko_distance_out(const uint32_t *counts, int n_ref, int n_qry, const uint32_t *ref_ctx_ct, const uint32_t *qry_ctx_ct,
                       const char *const *ref_names, const char *const *qry_names, int metric, int outfields,
                       int correction, int n_max, double max_dist, int kmerlen, int dim_rd_len, char *out, size_t cap)
{
    static const char *hdr[2][3] = {{"Jaccard\tMashD", "P-value(J)\tFDR(J)", "Jaccard_CI\tMashD_CI"},
                                    {"ContainmentM\tAafD", "P-value(C)\tFDR(C)", "ContainmentM_CI\tAafD_CI"}};
    ko_dist_opt o = {metric, outfields, correction, n_max, max_dist, kmerlen, dim_rd_len};
    size_t w = 0;
    char line[2048];
#define KO_PUT(str, n) do { if (out && w + (n) <= cap) memcpy(out + w, (str), (n)); w += (n); } while (0)
    int n = snprintf(line, sizeof line, "Qry\tRef\tShared_k|Ref_s|Qry_s");
    for (int i = 0; i <= outfields; i++) n += snprintf(line + n, sizeof line - n, "\t%s", hdr[metric][i]);
    n += snprintf(line + n, sizeof line - n, "\n");
    KO_PUT(line, (size_t)n);
    const double n_cmp = (double)((unsigned)n_ref * (unsigned)n_qry);          /* cmprsn_num: unsigned product, :1560 */
    typedef struct { double metric; int rid; } best_t;
    best_t *best = malloc(sizeof(best_t) * (size_t)(n_max + 2));
    for (int q = 0; q < n_qry; q++) {
        const unsigned Y = qry_ctx_ct[q];
        const uint32_t *row = counts + (size_t)q * n_ref;
        if (n_max) {
            for (int i = 0; i < n_max; i++) { best[i].metric = 0; best[i].rid = -1; }
            for (int r = 0; r < n_ref; r++) {
                unsigned X = ref_ctx_ct[r], XnY = row[r];
                double m = metric == 1 ? (double)XnY / (X < Y ? X : Y) : (double)XnY / (X + Y - XnY);
                for (int i = n_max - 1; i >= 0; i--) {
                    if (m > best[i].metric) { best[i + 1] = best[i]; best[i].metric = m; best[i].rid = r; }
                    else break;
                }
            }
            for (int i = 0; i < n_max; i++) {
                if (best[i].rid < 0) continue;
                int len = ko_dist_line(line, sizeof line, &o, qry_names[q], ref_names[best[i].rid], ref_ctx_ct[best[i].rid], Y,
                                       row[best[i].rid], n_cmp);
                if (len > 1) KO_PUT(line, (size_t)len);
            }
        } else {
            for (int r = 0; r < n_ref; r++) {
                int len = ko_dist_line(line, sizeof line, &o, qry_names[q], ref_names[r], ref_ctx_ct[r], Y, row[r], n_cmp);
                if (len > 1) KO_PUT(line, (size_t)len);
            }
        }
    }
#undef KO_PUT
    free(best);
    return w;
}

int ko_synth_params(mks_params *P, uint64_t seed, uint32_t n_species, uint32_t genome_len, uint32_t read_len,
                    uint32_t **cdf32, uint32_t **species)
{
    mks_default_params(P, seed, n_species, genome_len, read_len);
    return mks_build_cdf(P, cdf32, species);
}
uint64_t ko_fastq_bytes(const mks_params *P, uint64_t r0, uint64_t r1)
{
    return mks_fastq_offset(P, r1) - mks_fastq_offset(P, r0);
}
size_t ko_write_fastq(const mks_params *P, const uint32_t *cdf32, const uint32_t *spc, uint64_t r0, uint64_t r1,
                      char *buf)
{
    return mks_write_fastq(P, cdf32, spc, r0, r1, buf);
}
char ko_fastq_char_at(const mks_params *P, const uint32_t *cdf32, const uint32_t *spc, uint64_t off)
{
    uint64_t r = mks_fastq_record_of(P, off);
    return mks_fastq_char(P, cdf32, spc, r, (uint32_t)(off - mks_fastq_offset(P, r)));
}
size_t ko_fasta_bytes(const mks_params *P, uint32_t s) { return mks_fasta_size(P, s); }
size_t ko_write_fasta(const mks_params *P, uint32_t s, char *buf) { return mks_write_fasta(P, s, buf); }
void ko_make_shuf_perm(uint64_t seed, int subk, int32_t *perm) { mks_make_shuf_perm(seed, subk, perm); }
int32_t ko_shuf_id(uint64_t seed) { return mks_shuf_id(seed); }
void ko_free(void *p) { free(p); }
