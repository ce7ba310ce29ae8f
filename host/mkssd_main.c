/*
 * mkssd_main.c — C host program above the C ABI (include/mkssd_b200.h): the `dist`, `composite` and
 * `shuffle` sub-commands of MetaKSSD restricted to the hot path, with the reference's flags and
 * on-disk formats, calling the CUDA library instead of the CPU loops.
 *
 * Mirrors (behaviour, not code): cmd_dist/dist_dispatch/run_stageI (command_dist_wrapper.c:309,
 * command_dist.c:49,341), cmd_composite/get_species_abundance (command_composite.c:145,446) and
 * write_dim_shuffle_file (command_shuffle.c:164).  Written from scratch.
 *
 *   metakssd-b200 shuffle -k 11 -s 6 -l 3 -o L3K11 [--seed N]
 *   metakssd-b200 dist -L L3K11.shuf -A -o sketch reads.fq [more inputs ...]
 *   metakssd-b200 dist -L L3K11.shuf -o gsk [-u] sp1.fasta sp2.fasta ...
 *   metakssd-b200 dist -L L3K11.shuf -o sk [-Q 20] [-n 2] reads.fq        (no -A: fastq2co)
 *   metakssd-b200 set -g group_name.txt -o pan gsk ; set -q -o union_sp pan ; set -i union_sp -o markerdb pan
 *   metakssd-b200 composite -r markerdb -q sketch > species_coverage.tsv
 *
 * Differences from the reference, all outside the sketch content: input files keep their command
 * line order (the reference shuffles them with srand(time)); -p is accepted and ignored (the GPU
 * library is driven from one host thread).
 */
#define _GNU_SOURCE
#include <dirent.h>
#include <errno.h>
#include <math.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <sys/stat.h>
#include <time.h>
#include "mkssd_b200.h"

#define PATHLEN 256
#define MIN_KM_S 6

typedef struct {           /* co_dstat_t, global_basic.h:116-126 */
    unsigned int shuf_id;
    bool koc;
    int kmerlen, dim_rd_len, comp_num, infile_num;
    unsigned long long all_ctx_ct;
} co_dstat_t;

/* MK_TIMING=1: wall clock of the host program's phases on stderr (development aid) */
static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static double g_t0;
static void phase(const char *what)
{
    if (!getenv("MK_TIMING")) return;
    double t = now_s();
    fprintf(stderr, "[host timing] %-28s %8.1f ms\n", what, (t - g_t0) * 1e3);
    g_t0 = t;
}

static void die(const char *what, const char *arg)
{
    fprintf(stderr, "metakssd-b200: %s%s%s\n", what, arg ? ": " : "", arg ? arg : "");
    exit(1);
}
static void ck(mk_ctx *ctx, int rc, const char *where)
{
    if (rc == MK_OK) return;
    fprintf(stderr, "metakssd-b200: %s: %s (%s)\n", where, mk_strerror(rc), ctx ? mk_last_error(ctx) : "");
    exit(rc == MK_ERR_CROWDED ? 2 : 1);
}
static bool has_ext(const char *path, const char *const *exts)
{
    char buf[1024];
    snprintf(buf, sizeof buf, "%s", path);
    size_t n = strlen(buf);
    if (n > 3 && !strcmp(buf + n - 3, ".gz")) buf[n - 3] = 0;
    else if (n > 4 && !strcmp(buf + n - 4, ".bz2")) buf[n - 4] = 0;
    n = strlen(buf);
    for (; *exts; exts++) {
        size_t m = strlen(*exts);
        if (n > m && buf[n - m - 1] == '.' && !strcasecmp(buf + n - m, *exts)) return true;
    }
    return false;
}
static void *slurp(const char *path, size_t *bytes)
{
    FILE *f = fopen(path, "rb");
    if (!f) die("cannot open", path);
    struct stat st;
    if (stat(path, &st)) die("cannot stat", path);
    void *p = malloc(st.st_size ? st.st_size : 1);
    if (!p || fread(p, 1, st.st_size, f) != (size_t)st.st_size) die("cannot read", path);
    fclose(f);
    *bytes = st.st_size;
    return p;
}

/* ---- shuffle ------------------------------------------------------------------------------ */
static int cmd_shuffle(int argc, char **argv)
{
    int k = 8, s = 5, l = 2;
    const char *out = "default";
    uint64_t seed = (uint64_t)time(NULL);
    for (int i = 0; i < argc; i++) {
        if (!strcmp(argv[i], "-k") && i + 1 < argc) k = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-s") && i + 1 < argc) s = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-l") && i + 1 < argc) l = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-o") && i + 1 < argc) out = argv[++i];
        else if (!strcmp(argv[i], "--seed") && i + 1 < argc) seed = strtoull(argv[++i], NULL, 0);
    }
    if (k < s) die("half-context length must not be shorter than the half-subcontext length", NULL);
    if (s >= 8) die("subk should be smaller than 8", NULL);
    size_t n = (size_t)1 << (4 * s);
    int32_t *perm = malloc(n * sizeof *perm);
    if (!perm) die("out of memory", NULL);
    mk_synth_shuf_perm(seed, s, perm);
    int hdr[4] = {mk_synth_shuf_id(seed), k, s, l};
    char path[PATHLEN + 8];
    snprintf(path, sizeof path, "%s.shuf", out);
    FILE *f = fopen(path, "wb");
    if (!f) die("cannot create", path);
    fwrite(hdr, sizeof hdr, 1, f);
    fwrite(perm, sizeof *perm, n, f);
    fclose(f);
    printf("kssd shuffle: shuf_id=%d, k = %d, halfCtxLen = %d, level= %d\n", hdr[0], k, s, l);
    free(perm);
    return 0;
}

/* ---- dist ------------------------------------------------------------------------------------ */
/* Input arguments like the reference takes them (organize_infile_frm_arg() / organize_infile_list(), global_basic.c:169-330):
 * a directory stands for the sequence files in it (one level, accepted extensions fna fas fasta fa fq fastq, optionally
 * .gz / .bz2), `-l <file>` names one input per line.  The reference then processes the files in a time-seeded random order;
 * here directory entries are taken in name order, so that two runs write the same sketch directory. */
static const char *const seq_ext[] = {"fna", "fas", "fasta", "fa", "fq", "fastq", NULL};
static int by_name(const void *a, const void *b) { return strcmp(*(char *const *)a, *(char *const *)b); }
static void add_input(char ***list, int *n, int *cap, const char *path)
{
    if (*n + 1 >= *cap) { *cap = *cap ? *cap * 2 : 64; *list = realloc(*list, sizeof(char *) * (size_t)*cap); }
    (*list)[(*n)++] = strdup(path);
}
static void expand_input(char ***list, int *n, int *cap, const char *arg)
{
    struct stat st;
    if (stat(arg, &st) == 0 && S_ISDIR(st.st_mode)) {
        DIR *d = opendir(arg);
        if (!d) die("cannot open directory", arg);
        int first = *n;
        struct dirent *e;
        while ((e = readdir(d)) != NULL) {
            char full[PATHLEN * 2];
            snprintf(full, sizeof full, "%s/%s", arg, e->d_name);
            if (strlen(full) >= PATHLEN) die("path exceeds the maximal path length", full);
            if (stat(full, &st) == 0 && S_ISREG(st.st_mode) && has_ext(full, seq_ext)) add_input(list, n, cap, full);
        }
        closedir(d);
        qsort(*list + first, (size_t)(*n - first), sizeof(char *), by_name);
    } else {
        add_input(list, n, cap, arg);
    }
}
static void expand_list_file(char ***list, int *n, int *cap, const char *path)
{
    FILE *f = fopen(path, "r");
    if (!f) die("can't open file", path);
    char line[PATHLEN * 4];
    while (fgets(line, sizeof line, f)) {
        char *p = line;
        while (*p == ' ' || *p == '\t') p++;
        p[strcspn(p, "\r\n")] = 0;
        if (!*p) continue;
        if (strlen(p) >= PATHLEN) die("a line of the input list exceeds the maximal path length", path);
        struct stat st;
        if (stat(p, &st) || !S_ISREG(st.st_mode)) die("not a file (input list)", p);
        if (!has_ext(p, seq_ext)) die("wrong format in the input list (supported: .fna .fas .fasta .fq .fastq .fa)", p);
        add_input(list, n, cap, p);
    }
    fclose(f);
}

/* ---- dist -r <ref> -o <out> <qry>: shared k-mer counts + distance table ---------------------------------
 * mco_cbdco_nobin_dist() (command_dist.c:902-1079): the counts come from mk_shared_counts() (one call per component),
 * the table is dist_print_nobin() / output_ctrl() (command_dist.c:1531-1680) — four integers per pair through the same
 * libm expressions and printf formats.  <ref> is a sketch directory (cofiles.stat, or mcofiles.stat next to the combco
 * files when the reference has already indexed it in place); the reference's own 32 GiB mco.index.N is neither
 * needed nor written. */
typedef struct { unsigned int shuf_id; int kmerlen, dim_rd_len, comp_num, infile_num; } mco_dstat_t;   /* command_dist.h:67-75 */
static mk_ctx *plain_ctx(const co_dstat_t *st);
typedef struct { int metric, outfields, correction, n_max, keep; double max_dist; const char *dump_ref, *skf; } search_opt;

static int dist_line(char *line, size_t cap, const search_opt *o, int kmerlen, int dim_rd_len, const char *qname,
                     const char *rname, unsigned X, unsigned Y, unsigned XnY, double n_cmp)
{
#define GET_METRIC(M, Y_) ((M) == 0 ? 1 / (2 * (Y_)) + 0.5 : 1 / (Y_))
    double rs = 0;
    if (o->correction) {
        unsigned a = X - XnY, b = Y - XnY;
        double pa = 1 - pow((1 - 1 / pow(4, (kmerlen - dim_rd_len))), a);
        double pb = 1 - pow((1 - 1 / pow(4, (kmerlen - dim_rd_len))), b);
        rs = pa * pb * (a + b) / (pa + pb - 2 * pa * pb);
    }
    unsigned tmp = o->metric == 0 ? X + Y - XnY : (X < Y ? X : Y);
    double metric = ((double)XnY - rs) / tmp;
    double dist = log(GET_METRIC(o->metric, metric)) / kmerlen;
    if (dist > 1) dist = 1;
    if (dist > o->max_dist) return 0;
    int len = snprintf(line, cap, "%s\t%s\t%u-%u|%u|%u\t%.6lf\t%.6lf", qname, rname, XnY, (unsigned)rs, X, Y, metric, dist);
    if (o->outfields > 0) {
        double sd = pow(metric * (1 - metric) / tmp, 0.5);
        double pv = 0.5 * erfc(metric / sd * pow(0.5, 0.5));
        len += snprintf(line + len, cap - len, "\t%E\t%E", pv, pv * n_cmp);
        if (o->outfields > 1) {
            double m1 = metric - 1.96 * sd, m2 = metric + 1.96 * sd;
            double d1 = log(GET_METRIC(o->metric, m2)) / kmerlen, d2 = log(GET_METRIC(o->metric, m1)) / kmerlen;
            len += snprintf(line + len, cap - len, "\t[%.6lf,%.6lf]\t[%.6lf,%.6lf]", m1, m2, d1, d2);
        }
    }
    len += snprintf(line + len, cap - len, "\n");
    return len;
#undef GET_METRIC
}

/* One component of a database the reference built (`dist -L x.shuf -r <genomes> -o <db>`: only mco files are left):
 * mco.index.<c> holds, for every code s = 0 .. 2^32-1, the END offset of row s in mco.<c>, the list of references
 * that hold code s (combco2mco(), co2mco.c:56-79; COMPONENT_SZ = 8, global_basic.h:36).  Turned back into what
 * mk_shared_counts() takes: every reference's codes (ascending) and the index over the references.  One sequential
 * pass over the 32 GiB index file. */
static void read_mco_component(const char *dir, int c, int n_ref, uint32_t **codes_out, uint64_t **index_out)
{
    char path[PATHLEN * 2];
    size_t gbytes;
    snprintf(path, sizeof path, "%s/mco.%d", dir, c);
    uint32_t *gid = slurp(path, &gbytes);
    const uint64_t G = gbytes / 4;
    uint32_t *code_of = malloc((size_t)(G ? G : 1) * 4);           /* code of every (code, reference) pair, in file order */
    snprintf(path, sizeof path, "%s/mco.index.%d", dir, c);
    FILE *f = fopen(path, "rb");
    if (!f) die("cannot open", path);
    const size_t CH = (size_t)8 << 20;                             /* entries per read */
    uint64_t *buf = malloc(CH * 8);
    if (!code_of || !buf) die("out of memory reading", path);
    uint64_t s = 0, prev = 0;
    for (;;) {
        size_t got = fread(buf, 8, CH, f);
        if (!got) break;
        for (size_t i = 0; i < got; i++, s++) {
            const uint64_t end = buf[i];
            if (end != prev) {
                if (end < prev || end > G) die("malformed mco index", path);
                for (uint64_t g = prev; g < end; g++) code_of[g] = (uint32_t)s;
                prev = end;
            }
        }
    }
    fclose(f);
    free(buf);
    if (prev != G) die("mco index and mco file disagree under", dir);
    if (n_ref <= 0) die("no reference sketches in", dir);
    uint64_t *index = calloc((size_t)n_ref + 1, 8);
    for (uint64_t g = 0; g < G; g++) {
        if (gid[g] >= (uint32_t)n_ref) die("reference id out of range in", path);
        index[gid[g] + 1]++;
    }
    for (int r = 0; r < n_ref; r++) index[r + 1] += index[r];
    uint64_t *pos = malloc(((size_t)n_ref + 1) * 8);
    for (int r = 0; r < n_ref; r++) pos[r] = index[r];
    uint32_t *codes = malloc((size_t)(G ? G : 1) * 4);
    for (uint64_t g = 0; g < G; g++) codes[pos[gid[g]]++] = code_of[g];      /* codes ascend inside every reference */
    free(pos); free(code_of); free(gid);
    *codes_out = codes;
    *index_out = index;
}

static int dist_search(const char *refdir, const char *qrydir, const char *outdir, const search_opt *o)
{
    char path[PATHLEN * 2];
    size_t n;
    /* reference side: mcofiles.stat if the reference binary has indexed the directory, else cofiles.stat */
    mco_dstat_t R;
    unsigned int *r_ct;
    char (*r_names)[PATHLEN];
    snprintf(path, sizeof path, "%s/mcofiles.stat", refdir);
    struct stat sb;
    char *raw;
    if (stat(path, &sb) == 0) {
        raw = slurp(path, &n);
        memcpy(&R, raw, sizeof R);
        r_ct = (unsigned int *)(raw + sizeof R);
    } else {
        snprintf(path, sizeof path, "%s/cofiles.stat", refdir);
        raw = slurp(path, &n);
        co_dstat_t c;
        memcpy(&c, raw, sizeof c);
        R.shuf_id = c.shuf_id; R.kmerlen = c.kmerlen; R.dim_rd_len = c.dim_rd_len; R.comp_num = c.comp_num; R.infile_num = c.infile_num;
        r_ct = (unsigned int *)(raw + sizeof c);
    }
    r_names = (char (*)[PATHLEN])((char *)r_ct + (size_t)R.infile_num * 4);
    snprintf(path, sizeof path, "%s/cofiles.stat", qrydir);
    char *qraw = slurp(path, &n);
    co_dstat_t Q;
    memcpy(&Q, qraw, sizeof Q);
    unsigned int *q_ct = (unsigned int *)(qraw + sizeof Q);
    char (*q_names)[PATHLEN] = (char (*)[PATHLEN])((char *)q_ct + (size_t)Q.infile_num * 4);
    if (Q.shuf_id != R.shuf_id) {
        fprintf(stderr, "metakssd-b200: qry shuf_id: %d not match ref shuf_id: %d\n", (int)Q.shuf_id, (int)R.shuf_id);
        exit(1);
    }
    if (Q.comp_num != R.comp_num) {
        fprintf(stderr, "metakssd-b200: qry comp_num: %d not match ref comp_num: %d\n", Q.comp_num, R.comp_num);
        exit(1);
    }
    const int n_ref = R.infile_num, n_qry = Q.infile_num;
    if (o->n_max > 1024 || o->n_max > n_ref) {
        fprintf(stderr, "metakssd-b200: neighborN_max %d should smaller than NREF %d and ref_num %d\n", o->n_max, 1024, n_ref);
        exit(1);
    }
    g_t0 = now_s();
    if (o->dump_ref) {          /* no device needed: write the reference side back as combco files + cofiles.stat and leave */
        mkdir(o->dump_ref, 0777);
        co_dstat_t st;
        memset(&st, 0, sizeof st);
        st.shuf_id = R.shuf_id; st.kmerlen = R.kmerlen; st.dim_rd_len = R.dim_rd_len; st.comp_num = R.comp_num; st.infile_num = n_ref;
        for (int r = 0; r < n_ref; r++) st.all_ctx_ct += r_ct[r];
        for (int c = 0; c < R.comp_num; c++) {
            uint32_t *rc;
            uint64_t *ri;
            read_mco_component(refdir, c, n_ref, &rc, &ri);
            snprintf(path, sizeof path, "%s/combco.%d", o->dump_ref, c);
            FILE *fo = fopen(path, "wb");
            if (!fo || fwrite(rc, 4, ri[n_ref], fo) != ri[n_ref]) die("cannot write", path);
            fclose(fo);
            snprintf(path, sizeof path, "%s/combco.index.%d", o->dump_ref, c);
            fo = fopen(path, "wb");
            if (!fo || fwrite(ri, 8, (size_t)n_ref + 1, fo) != (size_t)n_ref + 1) die("cannot write", path);
            fclose(fo);
            free(rc); free(ri);
        }
        snprintf(path, sizeof path, "%s/cofiles.stat", o->dump_ref);
        FILE *fo = fopen(path, "wb");
        if (!fo) die("cannot write", path);
        fwrite(&st, sizeof st, 1, fo);
        fwrite(r_ct, 4, (size_t)n_ref, fo);
        fwrite(r_names, PATHLEN, (size_t)n_ref, fo);
        fclose(fo);
        phase("dump reference side");
        return 0;
    }
    mk_ctx *ctx = NULL;
    uint32_t *counts;
    if (o->skf) {               /* the counts of an earlier run (--keepskf): only the table is printed, no device needed */
        size_t b;
        counts = slurp(o->skf, &b);
        if (b != (size_t)n_ref * (size_t)n_qry * sizeof *counts) die("the shared k-mer count file does not fit these directories", o->skf);
    } else {
        co_dstat_t geo = Q;
        ctx = plain_ctx(&geo);
        phase("mk_ctx_create");
        counts = calloc((size_t)n_ref * (size_t)n_qry, sizeof *counts);
        if (!counts) die("out of memory for the shared k-mer count matrix", NULL);
        for (int c = 0; c < R.comp_num; c++) {
            size_t b;
            uint32_t *rc;
            uint64_t *ri;
            snprintf(path, sizeof path, "%s/combco.%d", refdir, c);
            if (stat(path, &sb) == 0) {
                rc = slurp(path, &b);
                snprintf(path, sizeof path, "%s/combco.index.%d", refdir, c);
                ri = slurp(path, &b);
            } else {                              /* a database built by the reference: only its inverted index is there */
                read_mco_component(refdir, c, n_ref, &rc, &ri);
                phase("read mco index");
            }
            snprintf(path, sizeof path, "%s/combco.%d", qrydir, c);
            uint32_t *qc = slurp(path, &b);
            snprintf(path, sizeof path, "%s/combco.index.%d", qrydir, c);
            uint64_t *qi = slurp(path, &b);
            ck(ctx, mk_shared_counts(ctx, rc, ri, n_ref, qc, qi, n_qry, q_ct, counts), "mk_shared_counts");
            free(rc); free(ri); free(qc); free(qi);
        }
    }
    phase("shared k-mer counts");
    mkdir(outdir, 0700);
    if (o->keep) {
        snprintf(path, sizeof path, "%s/sharedk_ct.dat", outdir);
        FILE *fk = fopen(path, "wb");
        if (!fk || fwrite(counts, sizeof *counts, (size_t)n_ref * (size_t)n_qry, fk) != (size_t)n_ref * (size_t)n_qry) die("cannot write", path);
        fclose(fk);
    }
    /* distance.out (dist_print_nobin(), command_dist.c:1531-1634) */
    static const char *hdr[2][3] = {{"Jaccard\tMashD", "P-value(J)\tFDR(J)", "Jaccard_CI\tMashD_CI"},
                                    {"ContainmentM\tAafD", "P-value(C)\tFDR(C)", "ContainmentM_CI\tAafD_CI"}};
    snprintf(path, sizeof path, "%s/distance.out", outdir);
    FILE *f = fopen(path, "w");
    if (!f) die("cannot write", path);
    fprintf(f, "Qry\tRef\tShared_k|Ref_s|Qry_s");
    for (int i = 0; i <= o->outfields; i++) fprintf(f, "\t%s", hdr[o->metric][i]);
    fprintf(f, "\n");
    const double n_cmp = (double)(long long)((unsigned)n_ref * (unsigned)n_qry);     /* outfield.cmprsn_num, :1560 */
    typedef struct { double metric; int rid; } best_t;
    best_t *best = malloc(sizeof(best_t) * (size_t)(o->n_max + 2));
    char line[2048];
    for (int q = 0; q < n_qry; q++) {
        const unsigned Y = q_ct[q];
        const uint32_t *row = counts + (size_t)q * (size_t)n_ref;
        if (o->n_max) {                       /* the N references with the largest metric, ties by insertion (:1586-1605) */
            for (int i = 0; i < o->n_max; i++) { best[i].metric = 0; best[i].rid = -1; }
            for (int r = 0; r < n_ref; r++) {
                unsigned X = r_ct[r], XnY = row[r];
                double m = o->metric == 1 ? (double)XnY / (X < Y ? X : Y) : (double)XnY / (X + Y - XnY);
                for (int i = o->n_max - 1; i >= 0; i--) {
                    if (m > best[i].metric) { best[i + 1] = best[i]; best[i].metric = m; best[i].rid = r; }
                    else break;
                }
            }
            for (int i = 0; i < o->n_max; i++) {
                if (best[i].rid < 0) continue;
                int len = dist_line(line, sizeof line, o, Q.kmerlen, Q.dim_rd_len, q_names[q], r_names[best[i].rid],
                                    r_ct[best[i].rid], Y, row[best[i].rid], n_cmp);
                if (len > 1) fwrite(line, 1, (size_t)len, f);
            }
        } else {
            for (int r = 0; r < n_ref; r++) {
                int len = dist_line(line, sizeof line, o, Q.kmerlen, Q.dim_rd_len, q_names[q], r_names[r], r_ct[r], Y, row[r], n_cmp);
                if (len > 1) fwrite(line, 1, (size_t)len, f);
            }
        }
    }
    fclose(f);
    phase("distance.out");
    if (ctx) mk_ctx_destroy(ctx);
    free(best); free(counts); free(raw); free(qraw);
    return 0;
}

static int cmd_dist(int argc, char **argv)
{
    const char *shuf = NULL, *outdir = "./", *pipecmd = "", *refpath = NULL;
    search_opt so = {0, 2, 0, 0, 0, 1.0, NULL, NULL};      /* command_dist_wrapper.c:83-92: Jaccard, all fields, every reference, D <= 1 */
    bool abundance = false, dedup = false;
    int kmerqlty = 0, kmerocrs = 1;            /* command_dist_wrapper.c:79-80 */
    char **inputs = NULL;
    int n_in = 0, cap_in = 0;
    const char *listfile = NULL;
    bool list_only = false;
    char **raw = malloc(sizeof(char *) * (size_t)(argc + 1));
    int n_raw = 0;
    for (int i = 0; i < argc; i++) {
        if (!strcmp(argv[i], "-L") && i + 1 < argc) shuf = argv[++i];
        else if (!strcmp(argv[i], "-o") && i + 1 < argc) outdir = argv[++i];
        else if (!strcmp(argv[i], "-p") && i + 1 < argc) ++i;
        else if (!strcmp(argv[i], "-P") && i + 1 < argc) pipecmd = argv[++i];
        else if (!strcmp(argv[i], "-A")) abundance = true;
        else if (!strcmp(argv[i], "-u")) dedup = true;
        else if (!strcmp(argv[i], "-r") && i + 1 < argc) refpath = argv[++i];
        else if (!strcmp(argv[i], "-M") && i + 1 < argc) so.metric = atoi(argv[++i]) ? 1 : 0;
        else if (!strcmp(argv[i], "-O") && i + 1 < argc) { so.outfields = atoi(argv[++i]); if (so.outfields < 0 || so.outfields > 2) die("-O takes 0, 1 or 2", NULL); }
        else if (!strcmp(argv[i], "-N") && i + 1 < argc) so.n_max = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-D") && i + 1 < argc) so.max_dist = atof(argv[++i]);
        else if (!strcmp(argv[i], "--correction") && i + 1 < argc) so.correction = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--keepskf")) so.keep = 1;
        else if (!strcmp(argv[i], "-f") && i + 1 < argc) so.skf = argv[++i];      /* print from a kept sharedk_ct.dat (command_dist.c:984-987) */
        else if (!strcmp(argv[i], "--dump-ref") && i + 1 < argc) so.dump_ref = argv[++i];   /* (tool: the reference side as a sketch directory) */
        else if (!strcmp(argv[i], "-Q") && i + 1 < argc) kmerqlty = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-n") && i + 1 < argc) {        /* clamped to 1..7 like command_dist_wrapper.c:169-179 */
            int v = atoi(argv[++i]);
            if (v > 7) { fprintf(stderr, "metakssd-b200: -n argument is larger than Max, it has been set to 7, ignorned -n %d \n", v); v = 7; }
            if (v < 1) { fprintf(stderr, "metakssd-b200: -n argument is smaller than Min, it has been set to 1, ignorned -n %d \n", v); v = 1; }
            kmerocrs = v;
        }
        else if (!strcmp(argv[i], "-l") && i + 1 < argc) listfile = argv[++i];
        else if (!strcmp(argv[i], "--list-inputs")) list_only = true;        /* (tool: print the expanded inputs and leave) */
        else if (argv[i][0] == '-') die("option not on the hot path", argv[i]);
        else raw[n_raw++] = argv[i];
    }
    for (int i = 0; i < n_raw; i++) {
        if (refpath) add_input(&inputs, &n_in, &cap_in, raw[i]);            /* (the query of dist -r is a sketch directory) */
        else expand_input(&inputs, &n_in, &cap_in, raw[i]);
    }
    if (listfile) expand_list_file(&inputs, &n_in, &cap_in, listfile);
    if (list_only) {
        for (int i = 0; i < n_in; i++) printf("%s\n", inputs[i]);
        return 0;
    }
    if (refpath) {            /* database search: the query is a sketch directory (command_dist.c:152-171) */
        if (n_in != 1) die("dist -r takes one query sketch directory", NULL);
        return dist_search(refpath, inputs[0], outdir, &so);
    }
    if (!shuf) die("-L <file.shuf> is required", NULL);
    if (!n_in) die("no input sequence files", NULL);
    size_t sb;
    g_t0 = now_s();
    int *sh = slurp(shuf, &sb);
    int shuf_id = sh[0], k = sh[1], subk = sh[2], drl = sh[3];
    if (sb != 16 + ((size_t)4 << (4 * subk))) die("malformed .shuf file", shuf);
    phase("read .shuf");

    mk_ctx *ctx = NULL;
    ck(NULL, mk_ctx_create(&ctx, sh + 4, k, subk, drl, 0), "mk_ctx_create");
    phase("mk_ctx_create");
    mk_info info;
    mk_ctx_info(ctx, &info);
    printf("rand_id=%d\thalf_ctx_len=%d\thashsize=%u\thashlimit=%u\n", shuf_id, k, info.hashsize, info.hashlimit);
    mkdir(outdir, 0777);

    static const char *const fq_ext[] = {"fq", "fastq", NULL};
    mk_sketch *sk = calloc((size_t)n_in, sizeof *sk);
    for (int i = 0; i < n_in;) {
        bool fq = has_ext(inputs[i], fq_ext) || pipecmd[0];
        if (fq && abundance) {
            printf("running mt_shortreads2koc()\n");
            ck(ctx, mk_fastq_koc_file(ctx, inputs[i], pipecmd, &sk[i]), inputs[i]);
            printf("%d/%d decomposing %s\r", i + 1, n_in, inputs[i]);
            i++;
        } else if (fq) {            /* fastq2co() + write_fqco2file() (command_dist.c:386-387) */
            ck(ctx, mk_fastq_co_file(ctx, inputs[i], pipecmd, kmerqlty, kmerocrs, &sk[i]), inputs[i]);
            printf("%d/%d decomposing %s\r", i + 1, n_in, inputs[i]);
            i++;
        } else {
            if (abundance) {
                abundance = false;
                printf("Warning: close abundance mode (-A) since non-fastq file input.\n");
            }
            ck(ctx, mk_ctx_set_dedup(ctx, dedup), "mk_ctx_set_dedup");      /* -u: uniq_fasta2co() (command_dist.c:394-395) */
            /* the file loop of run_stageI() (command_dist.c:365) as one batched call over the run of genomes */
            int j = i;
            while (j < n_in && !(has_ext(inputs[j], fq_ext) || pipecmd[0])) j++;
            ck(ctx, mk_fasta_co_files(ctx, (const char *const *)(inputs + i), j - i, pipecmd, &sk[i]), inputs[i]);
            printf("%d/%d decomposing %s\r", j, n_in, inputs[j - 1]);
            i = j;
        }
    }
    printf("\n");
    phase("sketching");
    /* combco.<c>, combco.<c>.a, combco.index.<c>  (command_dist.c:408-470) */
    unsigned long long all_ct = 0;
    char path[PATHLEN * 2];
    for (int c = 0; c < info.component_num; c++) {
        snprintf(path, sizeof path, "%s/combco.%d", outdir, c);
        FILE *fc = fopen(path, "wb");
        snprintf(path, sizeof path, "%s/combco.index.%d", outdir, c);
        FILE *fi = fopen(path, "wb");
        FILE *fa = NULL;
        if (abundance) {
            snprintf(path, sizeof path, "%s/combco.%d.a", outdir, c);
            fa = fopen(path, "wb");
        }
        if (!fc || !fi || (abundance && !fa)) die("cannot write into", outdir);
        size_t off = 0;
        fwrite(&off, sizeof off, 1, fi);
        for (int i = 0; i < n_in; i++) {
            fwrite(sk[i].codes[c], 4, sk[i].n[c], fc);
            if (fa) fwrite(sk[i].counts[c], 2, sk[i].n[c], fa);
            off += sk[i].n[c];
            fwrite(&off, sizeof off, 1, fi);
        }
        fclose(fc); fclose(fi);
        if (fa) fclose(fa);
    }
    /* cofiles.stat (command_dist.c:477-500) */
    co_dstat_t st;
    memset(&st, 0, sizeof st);
    st.shuf_id = (unsigned)shuf_id; st.koc = abundance; st.kmerlen = 2 * k; st.dim_rd_len = 2 * drl;
    st.comp_num = info.component_num; st.infile_num = n_in;
    for (int i = 0; i < n_in; i++) all_ct += sk[i].n_total;
    st.all_ctx_ct = all_ct;
    snprintf(path, sizeof path, "%s/cofiles.stat", outdir);
    FILE *fs = fopen(path, "wb");
    if (!fs) die("cannot write", path);
    fwrite(&st, sizeof st, 1, fs);
    for (int i = 0; i < n_in; i++) { unsigned int ct = (unsigned)sk[i].n_total; fwrite(&ct, 4, 1, fs); }
    for (int i = 0; i < n_in; i++) {
        char name[PATHLEN];
        memset(name, 0, sizeof name);
        snprintf(name, sizeof name, "%s", inputs[i]);
        fwrite(name, PATHLEN, 1, fs);
    }
    fclose(fs);
    phase("write sketch directory");
    for (int i = 0; i < n_in; i++) mk_sketch_free(&sk[i]);
    mk_ctx_destroy(ctx);
    free(sk); free(sh); free(inputs);
    phase("mk_ctx_destroy");
    return 0;
}

/* ---- set -g / -q / -i ----------------------------------------------------------------------- */
/* The MarkerDB pipeline of the reference (README: dist -> set -g -> set -q -> set -i), directory formats and taxon
 * order as command_set.c writes them; the unions / filters run on the device (mk_set_group, mk_set_uniq_union,
 * mk_set_operate). */
static int next_prime(int n)                         /* global_basic.c:453-475 */
{
    for (;;) {
        int composite = 0;
        for (int j = 2; (long long)j * j <= n; j++)
            if (n % j == 0) { composite = 1; break; }
        if (!composite) return n;
        n++;
    }
}
typedef struct { int taxid; char *name; int first_line; } taxon_t;

/* organize_taxf() (command_set.c:635-704): taxa in ascending slot of a table of nextPrime(lines / 0.6) slots */
static int read_taxfile(const char *path, int **taxon_of_line, taxon_t **taxa, int *n_lines)
{
    size_t nb;
    char *raw = slurp(path, &nb);
    int ln = 0;
    for (size_t i = 0; i < nb; i++) ln += raw[i] == '\n';
    int hashsz = next_prime((int)((double)ln / 0.6));
    taxon_t *tab = calloc((size_t)(hashsz > 0 ? hashsz : 1), sizeof *tab);
    for (int i = 0; i < hashsz; i++) tab[i].taxid = -1;
    int *slot_of = malloc(sizeof(int) * (size_t)(ln > 0 ? ln : 1));
    size_t at = 0;
    for (int i = 0; i < ln; i++) {
        size_t e = at;
        while (raw[e] != '\n') e++;
        if (e - at >= PATHLEN - 1) die("organize_taxf(): taxfile line exceeds PATHLEN", path);
        raw[e] = 0;
        char *line = raw + at;
        at = e + 1;
        char *tab1 = strchr(line, '\t');
        char *name = NULL;
        if (tab1) { *tab1 = 0; name = tab1 + 1; char *t2 = strchr(name, '\t'); if (t2) *t2 = 0; if (!*name) name = NULL; }
        int taxid = atoi(line);
        slot_of[i] = -1;
        for (int n = 0; n < hashsz; n++) {
            int hv = (taxid % hashsz + n * (1 + taxid % (hashsz - 1))) % hashsz;
            if (tab[hv].taxid == -1) { tab[hv].taxid = taxid; tab[hv].name = name ? strdup(name) : NULL; tab[hv].first_line = i; slot_of[i] = hv; break; }
            if (tab[hv].taxid == taxid) {
                if ((tab[hv].name == NULL) != (name == NULL) || (name && strcmp(tab[hv].name, name)))
                    die("organize_taxf() abort!: a taxid has different taxnames", path);
                slot_of[i] = hv;
                break;
            }
        }
    }
    int n_taxa = 0;
    int *pos_of_slot = malloc(sizeof(int) * (size_t)(hashsz > 0 ? hashsz : 1));
    *taxa = calloc((size_t)(ln > 0 ? ln : 1), sizeof **taxa);
    for (int i = 0; i < hashsz; i++)
        if (tab[i].taxid != -1) { (*taxa)[n_taxa] = tab[i]; pos_of_slot[i] = n_taxa++; }
    *taxon_of_line = malloc(sizeof(int) * (size_t)(ln > 0 ? ln : 1));
    for (int i = 0; i < ln; i++) (*taxon_of_line)[i] = slot_of[i] < 0 ? -1 : pos_of_slot[slot_of[i]];
    *n_lines = ln;
    free(tab); free(slot_of); free(pos_of_slot); free(raw);
    return n_taxa;
}

static mk_ctx *plain_ctx(const co_dstat_t *st)
{
    int k = st->kmerlen / 2, drl = st->dim_rd_len / 2;
    int subk = drl + 3 > k ? k : drl + 3;
    mk_ctx *ctx = NULL;
    ck(NULL, mk_ctx_create(&ctx, NULL, k, subk, drl, 0), "mk_ctx_create");
    return ctx;
}
static void write_file(const char *dir, const char *name, int c, const void *p, size_t bytes)
{
    char path[PATHLEN * 2];
    if (c >= 0) snprintf(path, sizeof path, "%s/%s.%d", dir, name, c); else snprintf(path, sizeof path, "%s/%s", dir, name);
    FILE *f = fopen(path, "wb");
    if (!f || (bytes && fwrite(p, bytes, 1, f) != 1)) die("cannot write", path);
    fclose(f);
}

static int cmd_set(int argc, char **argv)
{
    const char *taxfile = NULL, *outdir = "./", *pan = NULL, *in = NULL;
    int op = -1;                     /* 3 = -q, 1 = -i, 0 = -s, 5 = -g */
    for (int i = 0; i < argc; i++) {
        if (!strcmp(argv[i], "-g") && i + 1 < argc) { taxfile = argv[++i]; if (op == -1) op = 5; }
        else if (!strcmp(argv[i], "-q")) { if (op == -1) op = 3; }
        else if (!strcmp(argv[i], "-i") && i + 1 < argc) { pan = argv[++i]; if (op == -1) op = 1; }
        else if (!strcmp(argv[i], "-s") && i + 1 < argc) { pan = argv[++i]; if (op == -1) op = 0; }
        else if (!strcmp(argv[i], "-o") && i + 1 < argc) outdir = argv[++i];
        else if (!strcmp(argv[i], "-p") && i + 1 < argc) ++i;
        else if (argv[i][0] == '-') die("option not on the hot path", argv[i]);
        else in = argv[i];
    }
    if (!in || op == -1) die("set operation use : -g, -q, -i or -s", NULL);
    char path[PATHLEN * 2];
    snprintf(path, sizeof path, "%s/cofiles.stat", in);
    size_t stat_bytes;
    char *stat_raw = slurp(path, &stat_bytes);
    co_dstat_t st;
    memcpy(&st, stat_raw, sizeof st);
    mk_ctx *ctx = plain_ctx(&st);
    mkdir(outdir, 0777);
    size_t b;
    if (op == 5) {                                       /* grouping_genomes(), command_set.c:831-1003 */
        int *taxon_of_line, n_lines;
        taxon_t *taxa;
        int n_all = read_taxfile(taxfile, &taxon_of_line, &taxa, &n_lines);
        if (n_lines != st.infile_num) die("grouping_genomes(): genome number of the sketch does not match the taxfile", taxfile);
        /* taxid 0 is ignored: its genomes are skipped and its position drops out of the output */
        int *out_pos = malloc(sizeof(int) * (size_t)(n_all > 0 ? n_all : 1)), n_taxa = 0;
        for (int t = 0; t < n_all; t++) out_pos[t] = taxa[t].taxid == 0 ? -1 : n_taxa++;
        int32_t *taxon_of = malloc(sizeof(int32_t) * (size_t)n_lines);
        for (int g = 0; g < n_lines; g++) taxon_of[g] = taxon_of_line[g] < 0 ? -1 : out_pos[taxon_of_line[g]];
        unsigned int *ctx_ct = calloc((size_t)(n_taxa > 0 ? n_taxa : 1), sizeof *ctx_ct);
        unsigned long long all_ct = 0;
        uint64_t *oi = malloc(sizeof(uint64_t) * (size_t)(n_taxa + 1));
        for (int c = 0; c < st.comp_num; c++) {
            snprintf(path, sizeof path, "%s/combco.%d", in, c);
            uint32_t *codes = slurp(path, &b);
            snprintf(path, sizeof path, "%s/combco.index.%d", in, c);
            uint64_t *index = slurp(path, &b);
            uint32_t *oc = NULL;
            if (n_taxa > 0) ck(ctx, mk_set_group(ctx, codes, index, n_lines, taxon_of, n_taxa, &oc, oi), "mk_set_group");
            else oi[0] = 0;
            write_file(outdir, "combco", c, oc, (size_t)oi[n_taxa] * 4);
            write_file(outdir, "combco.index", c, oi, sizeof(uint64_t) * (size_t)(n_taxa + 1));
            for (int t = 0; t < n_taxa; t++) ctx_ct[t] += (unsigned)(oi[t + 1] - oi[t]);
            all_ct += oi[n_taxa];
            mk_free(oc); free(codes); free(index);
        }
        st.infile_num = n_taxa; st.koc = 0; st.all_ctx_ct = all_ct;
        snprintf(path, sizeof path, "%s/cofiles.stat", outdir);
        FILE *f = fopen(path, "wb");
        if (!f) die("cannot write", path);
        memcpy(stat_raw, &st, sizeof st);                 /* (keeps the input's padding bytes like the reference's fread/fwrite) */
        fwrite(stat_raw, sizeof st, 1, f);
        fwrite(ctx_ct, sizeof *ctx_ct, (size_t)n_taxa, f);
        for (int t = 0; t < n_all; t++) {
            if (taxa[t].taxid == 0) continue;
            char name[PATHLEN];
            memset(name, 0, sizeof name);
            if (taxa[t].name) snprintf(name, sizeof name, "%d_%s", taxa[t].taxid, taxa[t].name);
            else snprintf(name, sizeof name, "%d", taxa[t].taxid);
            fwrite(name, PATHLEN, 1, f);
        }
        fclose(f);
    } else if (op == 3) {                                /* uniq_sketch_union(), command_set.c:427-512 */
        write_file(outdir, "cofiles.stat", -1, stat_raw, sizeof st);
        for (int c = 0; c < st.comp_num; c++) {
            snprintf(path, sizeof path, "%s/combco.%d", in, c);
            uint32_t *codes = slurp(path, &b);
            uint32_t *out = NULL;
            uint64_t n_out = 0;
            ck(ctx, mk_set_uniq_union(ctx, codes, b / 4, &out, &n_out), "mk_set_uniq_union");
            write_file(outdir, "uniq_pan", c, out, (size_t)n_out * 4);
            mk_free(out); free(codes);
        }
    } else {                                             /* sketch_operate(), command_set.c:322-423 */
        snprintf(path, sizeof path, "%s/cofiles.stat", pan);
        size_t pb;
        char *pan_raw = slurp(path, &pb);
        co_dstat_t pst;
        memcpy(&pst, pan_raw, sizeof pst);
        if (pst.shuf_id != st.shuf_id) die("sketcing id not match", pan);
        unsigned int *ctx_ct = (unsigned int *)(stat_raw + sizeof st);
        memset(ctx_ct, 0, sizeof(unsigned int) * (size_t)st.infile_num);
        uint64_t *oi = malloc(sizeof(uint64_t) * (size_t)(st.infile_num + 1));
        for (int c = 0; c < pst.comp_num; c++) {
            snprintf(path, sizeof path, "%s/pan.%d", pan, c);
            struct stat fs;
            if (stat(path, &fs) != 0) snprintf(path, sizeof path, "%s/uniq_pan.%d", pan, c);
            uint32_t *pcodes = slurp(path, &pb);
            snprintf(path, sizeof path, "%s/combco.%d", in, c);
            uint32_t *codes = slurp(path, &b);
            snprintf(path, sizeof path, "%s/combco.index.%d", in, c);
            uint64_t *index = slurp(path, &b);
            uint32_t *oc = NULL;
            ck(ctx, mk_set_operate(ctx, pcodes, pb / 4, codes, index, st.infile_num, op, &oc, oi), "mk_set_operate");
            write_file(outdir, "combco", c, oc, (size_t)oi[st.infile_num] * 4);
            write_file(outdir, "combco.index", c, oi, sizeof(uint64_t) * (size_t)(st.infile_num + 1));
            for (int i = 0; i < st.infile_num; i++) ctx_ct[i] += (unsigned)(oi[i + 1] - oi[i]);
            mk_free(oc); free(pcodes); free(codes); free(index);
        }
        write_file(outdir, "cofiles.stat", -1, stat_raw, stat_bytes);     /* (all_ctx_ct stays the input's, as in the reference) */
    }
    mk_ctx_destroy(ctx);
    return 0;
}

/* ---- composite ---------------------------------------------------------------------------- */
static const mk_species_stat *g_stats;
static int by_hits_desc(const void *a, const void *b) { return g_stats[*(const int *)b].n - g_stats[*(const int *)a].n; }

typedef struct { co_dstat_t st; char (*names)[PATHLEN]; } sketch_dir;
static sketch_dir read_stat(const char *dir)
{
    char path[PATHLEN * 2];
    snprintf(path, sizeof path, "%s/cofiles.stat", dir);
    size_t n;
    char *raw = slurp(path, &n);
    sketch_dir d;
    memcpy(&d.st, raw, sizeof d.st);
    if (n < sizeof d.st + (size_t)d.st.infile_num * (4 + PATHLEN)) die("malformed cofiles.stat under", dir);
    d.names = malloc((size_t)d.st.infile_num * PATHLEN);
    memcpy(d.names, raw + sizeof d.st + (size_t)d.st.infile_num * 4, (size_t)d.st.infile_num * PATHLEN);
    free(raw);
    return d;
}

static int cmd_composite(int argc, char **argv)
{
    const char *refdir = NULL, *qrydir = NULL;
    for (int i = 0; i < argc; i++) {
        if (!strcmp(argv[i], "-r") && i + 1 < argc) refdir = argv[++i];
        else if (!strcmp(argv[i], "-q") && i + 1 < argc) qrydir = argv[++i];
        else if (!strcmp(argv[i], "-p") && i + 1 < argc) ++i;
        else die("option not on the hot path", argv[i]);
    }
    if (!refdir || !qrydir || !strcmp(refdir, qrydir)) die("refdir or qrydir is not initialized", NULL);
    g_t0 = now_s();
    sketch_dir R = read_stat(refdir), Q = read_stat(qrydir);
    if (!Q.st.koc) die("get_species_abundance(): query has not abundance", NULL);
    if (Q.st.shuf_id != R.st.shuf_id)
        printf("get_species_abundance(): qry shuf_id %u not match ref shuf_id: %u\n", Q.st.shuf_id, R.st.shuf_id);
    /* composite is geometry free: a context without pass-set tables */
    int k = R.st.kmerlen / 2, drl = R.st.dim_rd_len / 2;
    int subk = drl + 3 > k ? k : drl + 3;
    mk_ctx *ctx = NULL;
    ck(NULL, mk_ctx_create(&ctx, NULL, k, subk, drl, 0), "mk_ctx_create");
    phase("mk_ctx_create");
    int S = R.st.infile_num;
    mk_species_stat *stats = malloc(sizeof *stats * (size_t)S);
    int *order = malloc(sizeof(int) * (size_t)S);
    char path[PATHLEN * 2];
    for (int qn = 0; qn < Q.st.infile_num; qn++) {
        ck(ctx, mk_composite_begin(ctx, S), "mk_composite_begin");
        for (int c = 0; c < R.st.comp_num; c++) {
            size_t b;
            snprintf(path, sizeof path, "%s/combco.%d", refdir, c);
            uint32_t *rc = slurp(path, &b);
            snprintf(path, sizeof path, "%s/combco.index.%d", refdir, c);
            uint64_t *ri = slurp(path, &b);
            snprintf(path, sizeof path, "%s/combco.%d", qrydir, c);
            uint32_t *qc = slurp(path, &b);
            snprintf(path, sizeof path, "%s/combco.index.%d", qrydir, c);
            uint64_t *qi = slurp(path, &b);
            snprintf(path, sizeof path, "%s/combco.%d.a", qrydir, c);
            uint16_t *qa = slurp(path, &b);
            ck(ctx, mk_composite_component(ctx, rc, ri, S, qc, qa, qi[qn], qi[qn + 1]), "mk_composite_component");
            free(rc); free(ri); free(qc); free(qi); free(qa);
        }
        ck(ctx, mk_composite_stats(ctx, stats), "mk_composite_stats");
        for (int i = 0; i < S; i++) order[i] = i;
        g_stats = stats;
        qsort(order, (size_t)S, sizeof(int), by_hits_desc);         /* command_composite.c:584 */
        for (int i = 0; i < S; i++) {
            const mk_species_stat *s = &stats[order[i]];
            if (s->n < MIN_KM_S) break;
            printf("%s\t%s\t%d\t%f\t%f\t%d\t%d\n", Q.names[qn], R.names[order[i]], s->n, (float)s->sum / s->n,
                   (float)s->lastsum / s->lastn, s->median, s->max);    /* command_composite.c:624 */
        }
    }
    phase("composite");
    mk_ctx_destroy(ctx);
    phase("mk_ctx_destroy");
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 2) {
        fprintf(stderr, "usage: %s <shuffle|dist|set|composite> [options] [arguments]\n", argv[0]);
        return 1;
    }
    if (!strcmp(argv[1], "shuffle")) return cmd_shuffle(argc - 2, argv + 2);
    if (!strcmp(argv[1], "dist")) return cmd_dist(argc - 2, argv + 2);
    if (!strcmp(argv[1], "composite")) return cmd_composite(argc - 2, argv + 2);
    if (!strcmp(argv[1], "set")) return cmd_set(argc - 2, argv + 2);
    die("sub-command outside the accelerated path", argv[1]);
    return 1;
}
