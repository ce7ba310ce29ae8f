// mk_ingest.cu — bounded, overlapped ingest behind the file entry points (mk_fastq_koc_file,
// mk_fasta_co_file, mk_fasta_co_files).
//
// Replaces the reader of mt_shortreads2koc() (/root/reference/iseq2comem.c:664-673: popen("zcat -fc"),
// 65 536 reads per batch through fgets) and of fasta2co() (iseq2comem.c:225-245: 64 KiB fread buffer).
//
//   source      a plain file is read directly (`zcat -fc` copies it unchanged), with several pread()
//               threads per chunk; gzip / bzip2 / xz input and an explicit pipe command go through popen()
//               exactly like the reference and are drained with read(2)
//   chunks      the text is cut at line starts into chunks of MK_INGEST_CHUNK_BYTES (16 MB): no k-mer and no
//               FASTQ line spans two chunks, the line count carries over as `line_base`
//   overlap     a ring of pinned host buffers: while chunk i is sketched (mk_fastq_partial_device) the
//               H2D copy of chunk i+1 runs on the copy stream and the reader threads fill chunk i+2
//   memory      host: MK_INGEST_BUFFERS x (chunk + slack) pinned bytes, whatever the input size; device:
//               two chunk buffers + the runs collected so far (merged whenever they exceed 8 M entries)
//   result      per-chunk runs (code, first global position, count) are merged at the end by
//               mk_runs_finalize_device(): the same sketch as one pass over the whole text
#include "mk_common.cuh"
#include <errno.h>
#include <fcntl.h>
#include <limits.h>
#include <unistd.h>
#include <sys/stat.h>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <new>
#include <string>
#include <thread>

#define MK_LINE_MAX 4095      // fgets(buf, 4096): a longer line is split by the reference (refused here)

namespace {

struct Source {
    int fd = -1;
    FILE *pipe = nullptr;
    bool direct = false;      // plain file: pread() at arbitrary offsets
    u64 size = 0;             // direct only
    std::string what;

    int open_path(mk_ctx *ctx, const char *path, const char *pipecmd)
    {
        what = path;
        bool compressed = false;
        if (!(pipecmd && pipecmd[0])) {
            int f = ::open(path, O_RDONLY);
            if (f < 0) {
                snprintf(ctx->err, sizeof(ctx->err), "open(%s): %s", path, strerror(errno));
                return MK_ERR_IO;
            }
            unsigned char magic[6] = {0};
            ssize_t g = ::pread(f, magic, sizeof magic, 0);
            struct stat st;
            // what `zcat -f` decompresses: gzip, (old) compress / pack; anything else is copied through
            if (g >= 2 && magic[0] == 0x1f && (magic[1] == 0x8b || magic[1] == 0x9d || magic[1] == 0x1e || magic[1] == 0xa0))
                compressed = true;
            if (!compressed && fstat(f, &st) == 0 && S_ISREG(st.st_mode)) {
                fd = f; direct = true; size = (u64)st.st_size;
                return MK_OK;
            }
            ::close(f);
        }
        char cmd[PATH_MAX + 256];
        if (pipecmd && pipecmd[0]) snprintf(cmd, sizeof(cmd), "%s %s", pipecmd, path);
        else snprintf(cmd, sizeof(cmd), "zcat -fc %s", path);       // iseq2comem.c:216, :664-669
        pipe = popen(cmd, "r");
        if (!pipe) {
            snprintf(ctx->err, sizeof(ctx->err), "popen(%s): %s", cmd, strerror(errno));
            return MK_ERR_IO;
        }
        fd = fileno(pipe);
        what = cmd;
        return MK_OK;
    }
    void close_all()
    {
        if (pipe) { pclose(pipe); pipe = nullptr; fd = -1; }
        if (fd >= 0) { ::close(fd); fd = -1; }
    }
    ~Source() { close_all(); }
};

// read exactly n bytes at `off` (direct) with `threads` pread() workers; returns bytes read (< n at EOF)
size_t read_direct(const Source &s, uint8_t *dst, size_t n, u64 off, int threads)
{
    if (off >= s.size) return 0;
    if (off + n > s.size) n = (size_t)(s.size - off);
    if (threads < 1) threads = 1;
    const size_t per = ((n + threads - 1) / threads + 4095) & ~(size_t)4095;
    std::atomic<bool> ok{true};
    auto work = [&](size_t lo, size_t hi) {
        while (lo < hi) {
            ssize_t g = ::pread(s.fd, dst + lo, hi - lo, (off_t)(off + lo));
            if (g < 0 && errno == EINTR) continue;
            if (g <= 0) { ok = false; return; }
            lo += (size_t)g;
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; t++) {
        size_t lo = per * t, hi = lo + per < n ? lo + per : n;
        if (lo < n) th.emplace_back(work, lo, hi);
    }
    work(0, per < n ? per : n);
    for (auto &t : th) t.join();
    return ok ? n : 0;
}

// read up to n bytes from a pipe (blocks until n bytes or EOF)
size_t read_pipe(const Source &s, uint8_t *dst, size_t n)
{
    size_t got = 0;
    while (got < n) {
        ssize_t g = ::read(s.fd, dst + got, n - got);
        if (g < 0 && errno == EINTR) continue;
        if (g <= 0) break;
        got += (size_t)g;
    }
    return got;
}

struct Pinned {
    uint8_t *p = nullptr;
    size_t cap = 0;
    int alloc(mk_ctx *ctx, size_t bytes)
    {
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e != cudaSuccess) {
            p = nullptr;
            cudaGetLastError();
            snprintf(ctx->err, sizeof(ctx->err), "cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e));
            return MK_ERR_NOMEM;
        }
        cap = bytes;
        return MK_OK;
    }
    ~Pinned() { if (p) cudaFreeHost(p); }
};

// device array that keeps its content when it grows
template <class T>
struct Growing {
    T *p = nullptr;
    u64 n = 0, cap = 0;
    int append(mk_ctx *ctx, const T *d_src, u64 m)
    {
        if (n + m > cap) {
            u64 want = (n + m) + (n + m) / 2 + 4096;
            T *q = nullptr;
            if (cudaMalloc(&q, want * sizeof(T)) != cudaSuccess) {
                cudaGetLastError();
                snprintf(ctx->err, sizeof(ctx->err), "cudaMalloc(%llu) failed (collected runs)", (unsigned long long)(want * sizeof(T)));
                return MK_ERR_NOMEM;
            }
            if (n) CK(cudaMemcpyAsync(q, p, n * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            if (p) cudaFree(p);
            p = q; cap = want;
        }
        if (m) CK(cudaMemcpyAsync(p + n, d_src, m * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
        n += m;
        return MK_OK;
    }
    ~Growing() { if (p) cudaFree(p); }
};

struct Chunk {
    size_t nbytes = 0;        // line-aligned payload
    bool last = false;
    int rc = MK_OK;
};

int default_threads()
{
    unsigned n = std::thread::hardware_concurrency();
    return n == 0 ? 8 : (n > 32 ? 32 : (int)n);
}

int env_int(const char *name, int dflt, int lo, int hi)
{
    const char *e = getenv(name);
    if (!e || !*e) return dflt;
    long v = atol(e);
    if (v < lo) v = lo;
    if (v > hi) v = hi;
    return (int)v;
}

}   // namespace

// ---- FASTQ -A from a file / pipe ---------------------------------------------------------------------
extern "C" int mk_fastq_koc_file(mk_ctx *ctx, const char *path, const char *pipecmd, mk_sketch *out)
{
    if (!ctx || !path || !out) return MK_ERR_ARG;
    memset(out, 0, sizeof(*out));
    CK(cudaSetDevice(ctx->device));
    try {
        Source src;
        CKR(src.open_path(ctx, path, pipecmd));
        size_t chunk = (size_t)16 << 20;       // (measured on a page-cached 2.5 GB file: 16 MB 146 ms, 64 MB 210 ms, 256 MB 566 ms — pinning the ring is the fixed cost)
        if (const char *e = getenv("MK_INGEST_CHUNK_BYTES")) chunk = (size_t)atoll(e);
        if (chunk < 2 * (MK_LINE_MAX + 1)) chunk = 2 * (MK_LINE_MAX + 1);
        chunk = (chunk + 4095) & ~(size_t)4095;
        const int NB = env_int("MK_INGEST_BUFFERS", 4, 3, 16);
        const int threads = env_int("MK_INGEST_THREADS", default_threads(), 1, 64);
        const size_t cap = chunk + 2 * (MK_LINE_MAX + 1) + 256;        // + a carried partial line + the final one

        std::vector<Pinned> hbuf((size_t)NB);
        for (auto &b : hbuf) CKR(b.alloc(ctx, cap));
        uint8_t *dbuf[2] = {nullptr, nullptr};
        {
            uint8_t *d;
            CKR(mk_scratch(ctx, SB_TEXT, 2 * (cap + 256), &d));
            dbuf[0] = d; dbuf[1] = d + cap + 256;            // (cap is a multiple of 16: both aligned)
        }

        // ---- reader thread: fills hbuf[i % NB] with chunk i (line-aligned), one chunk ahead of need ----
        std::mutex mu;
        std::condition_variable cv;
        std::vector<Chunk> chunks;            // produced so far
        u64 consumed = 0;                     // chunks whose pinned buffer may be reused (H2D done)
        bool stop = false;
        std::thread reader([&]() {
            const size_t SLACK = 2 * (MK_LINE_MAX + 1);
            // cut after the last newline of b[0, have); returns the cut (0 = a line of MK_LINE_MAX+ bytes)
            auto cut_at_line = [&](const uint8_t *b, size_t have) -> size_t {
                size_t cut = have;
                while (cut > 0 && b[cut - 1] != '\n') {
                    cut--;
                    if (have - cut > MK_LINE_MAX) return 0;
                }
                return cut;
            };
            auto wait_buffer = [&](u64 i) -> bool {    // buffers i and i+1 are free (i+1 receives the carried line)
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || i + 1 < consumed + (u64)NB; });
                return !stop;
            };
            auto publish = [&](const Chunk &c) {
                { std::lock_guard<std::mutex> lk(mu); chunks.push_back(c); }
                cv.notify_all();
            };
            if (src.direct) {
                u64 off = 0;
                for (u64 i = 0;; i++) {
                    if (!wait_buffer(i)) return;
                    uint8_t *b = hbuf[i % NB].p;
                    Chunk c;
                    size_t have = read_direct(src, b, chunk, off, threads);
                    if (have == 0 && off < src.size) { c.rc = MK_ERR_IO; publish(c); return; }
                    if (off + have >= src.size) { c.nbytes = have; c.last = true; publish(c); return; }
                    size_t cut = cut_at_line(b, have);
                    if (cut == 0) { c.rc = MK_ERR_LONG_LINE; publish(c); return; }
                    if (src.size - (off + cut) <= SLACK) {       // a short rest: it joins this chunk (the last chunk is never tiny)
                        size_t more = (size_t)(src.size - (off + have));
                        if (read_direct(src, b + have, more, off + have, 1) != more) { c.rc = MK_ERR_IO; publish(c); return; }
                        c.nbytes = have + more; c.last = true; publish(c); return;
                    }
                    c.nbytes = cut;
                    off += cut;
                    publish(c);
                }
            } else {
                // pipe: whether chunk i is the last one is known only after the next read, so a chunk is published
                // one read late
                size_t carry = 0;                 // bytes of a partial line already at the start of the buffer
                Chunk pending;
                size_t pending_have = 0;          // bytes in the pending chunk's buffer (payload + carried line)
                bool have_pending = false;
                for (u64 i = 0;; i++) {
                    if (!wait_buffer(i)) return;
                    uint8_t *b = hbuf[i % NB].p;
                    const size_t want = chunk - carry;
                    const size_t g = read_pipe(src, b + carry, want);
                    const size_t have = carry + g;
                    const bool eof = g < want;
                    if (eof && have_pending && have <= SLACK) {
                        // a short rest: the previous chunk takes it (its buffer still holds the carried line)
                        uint8_t *pb = hbuf[(i - 1) % NB].p;
                        memcpy(pb + pending_have, b + carry, g);
                        pending.nbytes = pending_have + g;
                        pending.last = true;
                        publish(pending);
                        return;
                    }
                    if (have_pending) publish(pending);
                    Chunk c;
                    if (eof) { c.nbytes = have; c.last = true; publish(c); return; }
                    size_t cut = cut_at_line(b, have);
                    if (cut == 0) { c.rc = MK_ERR_LONG_LINE; publish(c); return; }
                    c.nbytes = cut;
                    carry = have - cut;
                    memcpy(hbuf[(i + 1) % NB].p, b + cut, carry);
                    pending = c; pending_have = have; have_pending = true;
                }
            }
        });
        struct Joiner {
            std::thread &t; std::mutex &mu; std::condition_variable &cv; bool &stop;
            ~Joiner() { { std::lock_guard<std::mutex> lk(mu); stop = true; } cv.notify_all(); if (t.joinable()) t.join(); }
        } joiner{reader, mu, cv, stop};
        auto wait_chunk = [&](u64 i) -> Chunk {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return chunks.size() > i; });
            return chunks[i];
        };
        auto release_upto = [&](u64 n) {
            { std::lock_guard<std::mutex> lk(mu); if (n > consumed) consumed = n; }
            cv.notify_all();
        };

        // ---- pipeline: copy(i+1) on the copy stream under partial(i) --------------------------------
        Growing<u64> acc_code, acc_pos;
        Growing<u32> acc_cnt;
        cudaEvent_t copied[2] = {ctx->copy_ev[0], ctx->copy_ev[1]};
        u64 pos_base = 0, line_base = 0;
        Chunk cur = wait_chunk(0);
        if (cur.rc != MK_OK) {
            if (cur.rc == MK_ERR_LONG_LINE) snprintf(ctx->err, sizeof(ctx->err), "FASTQ line of 4095 bytes or more (fgets(…, 4096) would split it)");
            else snprintf(ctx->err, sizeof(ctx->err), "read error on %s", src.what.c_str());
            return cur.rc;
        }
        auto start_copy = [&](u64 i, const Chunk &c) -> int {
            uint8_t *d = dbuf[i & 1];
            // (the device buffer was last read by the partial sketch of chunk i-2, which has returned)
            if (c.nbytes) CK(cudaMemcpyAsync(d, hbuf[i % NB].p, c.nbytes, cudaMemcpyHostToDevice, ctx->copy_stream));
            CK(cudaMemsetAsync(d + c.nbytes, 0, 64, ctx->copy_stream));
            CK(cudaEventRecord(copied[i & 1], ctx->copy_stream));
            ctx->prof.h2d_bytes += c.nbytes;
            return MK_OK;
        };
        CKR(start_copy(0, cur));
        for (u64 i = 0;; i++) {
            Chunk nxt;
            if (!cur.last) {
                nxt = wait_chunk(i + 1);
                if (nxt.rc != MK_OK) {
                    if (nxt.rc == MK_ERR_LONG_LINE) snprintf(ctx->err, sizeof(ctx->err), "FASTQ line of 4095 bytes or more (fgets(…, 4096) would split it)");
                    else snprintf(ctx->err, sizeof(ctx->err), "read error on %s", src.what.c_str());
                    return nxt.rc;
                }
                CKR(start_copy(i + 1, nxt));
            }
            CK(cudaStreamWaitEvent(ctx->stream, copied[i & 1], 0));
            mk_runs r;
            CKR(mk_fastq_partial_device(ctx, dbuf[i & 1], cur.nbytes, pos_base, line_base, cur.last ? 1 : 0, &r));
            line_base = ctx->last_newlines;           // (newlines in front of the next chunk)
            pos_base += cur.nbytes;
            CKR(acc_code.append(ctx, (const u64 *)r.d_code, r.n));
            CKR(acc_pos.append(ctx, (const u64 *)r.d_firstpos, r.n));
            CKR(acc_cnt.append(ctx, r.d_count, r.n));
            CK(cudaStreamSynchronize(ctx->stream));      // (the library-owned runs are overwritten by the next call)
            release_upto(i + 1);                          // pinned buffer i: its copy finished before the sketch ran
            if (acc_code.n > (8u << 20) && !cur.last) {   // keep the collected runs bounded
                mk_runs m;
                CKR(mk_runs_merge_device(ctx, (const uint64_t *)acc_code.p, (const uint64_t *)acc_pos.p, acc_cnt.p, acc_code.n, &m));
                acc_code.n = acc_pos.n = acc_cnt.n = 0;
                CKR(acc_code.append(ctx, (const u64 *)m.d_code, m.n));
                CKR(acc_pos.append(ctx, (const u64 *)m.d_firstpos, m.n));
                CKR(acc_cnt.append(ctx, m.d_count, m.n));
                CK(cudaStreamSynchronize(ctx->stream));
            }
            if (cur.last) break;
            cur = nxt;
        }
        int st = 0;
        if (src.pipe) { st = pclose(src.pipe); src.pipe = nullptr; src.fd = -1; }
        if (st != 0 && pos_base == 0) {
            snprintf(ctx->err, sizeof(ctx->err), "%s: exit status %d and no data", src.what.c_str(), st);
            return MK_ERR_IO;
        }
        return mk_runs_finalize_device(ctx, (const uint64_t *)acc_code.p, (const uint64_t *)acc_pos.p, acc_cnt.p, acc_code.n, out);
    } catch (const std::bad_alloc &) {
        snprintf(ctx->err, sizeof(ctx->err), "out of host memory while reading %s", path);
        return MK_ERR_NOMEM;
    } catch (const std::exception &e) {
        snprintf(ctx->err, sizeof(ctx->err), "%s: %s", path, e.what());
        return MK_ERR_IO;
    }
}

// ---- FASTQ without -A from a file / pipe ---------------------------------------------------------------
// The quality line of a record must sit in the same buffer as its sequence line, so this secondary path reads the
// whole (decompressed) text into host memory first and uploads it chunk by chunk under the kernel.
extern "C" int mk_fastq_co_file(mk_ctx *ctx, const char *path, const char *pipecmd, int quality, int min_occurrence, mk_sketch *out)
{
    if (!ctx || !path || !out) return MK_ERR_ARG;
    try {
        Source src;
        CKR(src.open_path(ctx, path, pipecmd));
        std::vector<uint8_t> buf;
        size_t n = 0;
        if (src.direct) {
            buf.resize((size_t)src.size + 64);
            n = read_direct(src, buf.data(), (size_t)src.size, 0, default_threads());
            if (n != src.size) {
                snprintf(ctx->err, sizeof(ctx->err), "read error on %s", path);
                return MK_ERR_IO;
            }
        } else {
            for (;;) {
                if (buf.size() < n + (1u << 24)) buf.resize(buf.size() ? buf.size() * 2 : (size_t)1 << 26);
                size_t g = read_pipe(src, buf.data() + n, buf.size() - n);
                n += g;
                if (g == 0) break;
            }
            int st = pclose(src.pipe);
            src.pipe = nullptr; src.fd = -1;
            if (st != 0 && n == 0) {
                snprintf(ctx->err, sizeof(ctx->err), "%s: exit status %d and no data", src.what.c_str(), st);
                return MK_ERR_IO;
            }
        }
        return mk_fastq_co_host(ctx, buf.data(), n, quality, min_occurrence, out);
    } catch (const std::bad_alloc &) {
        snprintf(ctx->err, sizeof(ctx->err), "out of host memory while reading %s", path);
        return MK_ERR_NOMEM;
    }
}

// ---- FASTA genomes from files ----------------------------------------------------------------------
// Whole files (a genome is a few MB) are read into one pinned batch buffer by a pool of reader threads and
// sketched with one mk_fasta_co_device call per batch of at most MK_FASTA_BATCH_BYTES of text.
static int read_whole(mk_ctx *ctx, const char *path, const char *pipecmd, std::vector<uint8_t> &tmp, uint8_t *dst,
                      size_t room, size_t *got, bool *fits)
{
    Source s;
    CKR(s.open_path(ctx, path, pipecmd));
    *fits = true;
    if (s.direct) {
        if (s.size > room) { *fits = false; *got = (size_t)s.size; return MK_OK; }
        size_t g = read_direct(s, dst, (size_t)s.size, 0, 1);
        if (g != s.size) {
            snprintf(ctx->err, sizeof(ctx->err), "read error on %s", path);
            return MK_ERR_IO;
        }
        *got = g;
        return MK_OK;
    }
    // pipe: size unknown — drain into a growing temporary (a genome, not a metagenome)
    tmp.clear();
    size_t n = 0;
    for (;;) {
        if (tmp.size() < n + (1u << 22)) tmp.resize(tmp.size() ? tmp.size() * 2 : (size_t)1 << 24);
        size_t g = read_pipe(s, tmp.data() + n, tmp.size() - n);
        n += g;
        if (g == 0) break;
    }
    int st = pclose(s.pipe);
    s.pipe = nullptr; s.fd = -1;
    if (st != 0 && n == 0) {
        snprintf(ctx->err, sizeof(ctx->err), "%s: exit status %d and no data", s.what.c_str(), st);
        return MK_ERR_IO;
    }
    *got = n;
    if (n > room) { *fits = false; return MK_OK; }
    memcpy(dst, tmp.data(), n);
    return MK_OK;
}

extern "C" int mk_fasta_co_files(mk_ctx *ctx, const char *const *paths, int n_files, const char *pipecmd, mk_sketch *out)
{
    if (!ctx || !paths || !out || n_files <= 0) return MK_ERR_ARG;
    for (int f = 0; f < n_files; f++) memset(&out[f], 0, sizeof(out[f]));
    CK(cudaSetDevice(ctx->device));
    try {
        size_t batch = (size_t)1 << 30;
        if (const char *e = getenv("MK_FASTA_BATCH_BYTES")) batch = (size_t)atoll(e);
        if (batch < (1u << 20)) batch = 1u << 20;
        const int threads = env_int("MK_INGEST_THREADS", default_threads(), 1, 64);
        Pinned buf;
        CKR(buf.alloc(ctx, batch + 256));
        std::vector<uint8_t> tmp;
        int f0 = 0;
        while (f0 < n_files) {
            // pass 1 (sizes): plain files are sized with stat(); anything else is read right away, one by one
            std::vector<u64> off(1, 0);
            int f1 = f0;
            std::vector<int> direct_files;
            while (f1 < n_files) {
                struct stat st;
                bool plain = !(pipecmd && pipecmd[0]) && stat(paths[f1], &st) == 0 && S_ISREG(st.st_mode);
                if (plain) {
                    int fd = ::open(paths[f1], O_RDONLY);
                    unsigned char magic[2] = {0, 0};
                    if (fd >= 0) { if (::pread(fd, magic, 2, 0) < 0) magic[0] = 0; ::close(fd); }
                    if (magic[0] == 0x1f && (magic[1] == 0x8b || magic[1] == 0x9d || magic[1] == 0x1e || magic[1] == 0xa0)) plain = false;
                }
                size_t got = 0;
                bool fits = true;
                if (plain) {
                    got = (size_t)st.st_size;
                    fits = off.back() + got <= batch;
                    if (fits) direct_files.push_back(f1);
                } else {
                    CKR(read_whole(ctx, paths[f1], pipecmd, tmp, buf.p + off.back(), batch - off.back(), &got, &fits));
                }
                if (!fits) {
                    if (f1 == f0) {                       // a single file larger than the batch buffer: grow it once
                        if (got + 256 > buf.cap) {
                            Pinned bigger;
                            CKR(bigger.alloc(ctx, got + 256));
                            std::swap(buf.p, bigger.p); std::swap(buf.cap, bigger.cap);
                            batch = got;
                            continue;                     // retry this file with the larger buffer
                        }
                    }
                    break;
                }
                off.push_back(off.back() + got);
                f1++;
            }
            // pass 2: the plain files of the batch, read in parallel
            {
                std::atomic<size_t> next{0};
                std::atomic<int> bad{-1};
                auto work = [&]() {
                    for (;;) {
                        size_t i = next.fetch_add(1);
                        if (i >= direct_files.size()) return;
                        int f = direct_files[i];
                        int fd = ::open(paths[f], O_RDONLY);
                        size_t lo = (size_t)off[(size_t)(f - f0)], hi = (size_t)off[(size_t)(f - f0) + 1];
                        bool ok = fd >= 0;
                        while (ok && lo < hi) {
                            ssize_t g = ::pread(fd, buf.p + lo, hi - lo, (off_t)(lo - off[(size_t)(f - f0)]));
                            if (g < 0 && errno == EINTR) continue;
                            if (g <= 0) ok = false; else lo += (size_t)g;
                        }
                        if (fd >= 0) ::close(fd);
                        if (!ok) bad = f;
                    }
                };
                std::vector<std::thread> th;
                int nt = threads < (int)direct_files.size() ? threads : (int)direct_files.size();
                for (int t = 1; t < nt; t++) th.emplace_back(work);
                work();
                for (auto &t : th) t.join();
                if (bad >= 0) {
                    snprintf(ctx->err, sizeof(ctx->err), "read error on %s", paths[bad.load()]);
                    return MK_ERR_IO;
                }
            }
            int rc = mk_fasta_co_host(ctx, buf.p, (const uint64_t *)off.data(), f1 - f0, out + f0);
            if (rc != MK_OK) {
                for (int f = 0; f < f0; f++) mk_sketch_free(&out[f]);
                return rc;
            }
            f0 = f1;
        }
        return MK_OK;
    } catch (const std::bad_alloc &) {
        snprintf(ctx->err, sizeof(ctx->err), "out of host memory while reading FASTA input");
        return MK_ERR_NOMEM;
    }
}

extern "C" int mk_fasta_co_file(mk_ctx *ctx, const char *path, const char *pipecmd, mk_sketch *out)
{
    const char *p[1] = {path};
    if (!path) return MK_ERR_ARG;
    return mk_fasta_co_files(ctx, p, 1, pipecmd, out);
}
