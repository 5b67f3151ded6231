/* imcvt_cli.c -- batch front-end of the ImCvt command line on the B200 engine.
 *
 * Same command line, same stdout and same exit codes as the reference's CLI (src/main.c): usage text (:8-28), switch and
 * file-pair parsing (:87-132), per-file progress / error lines (:135-140, :177-208), default ".png" output name (:171-175),
 * summary (:214-221), exit code = number of failed files (:223).  What is different is WHEN the work happens: the
 * reference encodes one file at a time (:162-211); here consecutive `.jls` / `.h265` outputs of the same shape are read
 * straight into one page-locked staging arena and encoded by ONE call of the batch C-ABI (imcvtEncodeBatchJLS /
 * imcvtEncodeBatchHEVC / imcvtEncodeBatchHEVCGreen, include/imcvt_b200.h), which is what keeps the GPU busy.  Each file's
 * lines are collected and printed in file order, so the text is the reference's byte for byte.
 *
 * A pending batch is flushed before a file is looked at whenever that file's source or destination is the destination of
 * a pending job: the reference's sequential semantics (an output used as a later input, the "already exist" guard on a
 * name written earlier in the same call) stay observable.
 *
 * Formats: every input / output format of the reference, each written for this repo: PNM (pnm_io.c), BMP (fmt_bmp.c), QOI
 * (fmt_qoi.c), PNG (fmt_png.c: the reference's stored-deflate writer, and an inflate + unfilter reader in place of its
 * third-party uPNG) -- SURVEY section 8(f) row f4; tests/test_cli_formats.py holds them against the reference's CLI byte for
 * byte.  Loaders are tried in the reference's order (PNM, PNG, BMP, QOI: src/main.c:142-160); their stdout lines are
 * captured and printed in file order.
 *
 * Deliberate deviation: `-o <file>` in front of any input makes the reference write dst_fnames[-1];
 * here that output name is ignored.
 *
 * Several GPUs: a batch is sharded over the visible devices in contiguous frame blocks by the multi-device entry points
 * (imcvtEncodeBatch*Multi: one engine and one host thread per device, no collective) -- the reference's file loop
 * (src/main.c:162-211) spread over the box.  A device only joins when its block is worth a CUDA context: at least 8 frames per
 * device for JPEG-LS, 2 for H.265.  The files are the same bytes whatever the number of devices.
 *
 * The streams of a batch are written to their files by a writer thread while the main thread reads and encodes the next batch
 * (the reference's fwrite per file, src/imageio_jls.c:467-472 / src/imageio_hevc.c:43-48, taken off the critical path); the
 * files' lines are still printed in file order, and a file that reads or guards a name still being written waits for it.
 *
 * Environment: IMCVT_DEVICE = CUDA device ordinal, that device only; IMCVT_DEVICES = "all" (default) | a count k (devices
 * 0..k-1) | a comma list of ordinals; IMCVT_BATCH = most frames per batch (default 256); IMCVT_TIMING = 1 prints where the
 * wall-clock time of the call went (stderr).
 */
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <stdarg.h>
#include <limits.h>
#include <unistd.h>
#include <errno.h>
#include <sys/stat.h>
#include <sys/resource.h>
#include <sys/types.h>
#include <time.h>
#include <pthread.h>
#include "imcvt_b200.h"
#include "imcvt_pnm.h"

int      writePNMImageFile(const char *, const uint8_t *, int, uint32_t, uint32_t);
uint8_t *loadPNGImageFile(const char *, int *, uint32_t *, uint32_t *);        /* fmt_png.c */
uint8_t *loadBMPImageFile(const char *, int *, uint32_t *, uint32_t *);        /* fmt_bmp.c */
uint8_t *loadQOIImageFile(const char *, int *, uint32_t *, uint32_t *);        /* fmt_qoi.c */
int writePNGImageFile(const char *, const uint8_t *, int, uint32_t, uint32_t);
int writeBMPImageFile(const char *, const uint8_t *, int, uint32_t, uint32_t);
int writeQOIImageFile(const char *, const uint8_t *, int, uint32_t, uint32_t);

static const char *const USAGE_LINES[] = {
    "|------------------------------------------------------------------------------------|",
    "| ImageConverter (v0.5)  by https://github.com/WangXuan95/                           |",
    "|------------------------------------------------------------------------------------|",
    "| Usage:                                                                             |",
    "|   ImCvt [-switches]  <in1> -o <out1>  [<in2> -o <out2]  ...                        |",
    "|                                                                                    |",
    "| Where <in> and <out> can be:                                                       |",
    "|   .pnm (Portable Any Map)          : gray 8-bit or RGB 24-bit                      |",
    "|   .pgm (Portable Gray Map)         : gray 8-bit                                    |",
    "|   .ppm (Portable Pix Map)          : RGB 24-bit                                    |",
    "|   .png (Portable Network Graphics) : gray 8-bit or RGB 24-bit                      |",
    "|   .bmp (Bitmap Image File)         : gray 8-bit or RGB 24-bit                      |",
    "|   .qoi (Quite OK Image)            : RGB 24-bit                                    |",
    "|   .jls (JPEG-LS Image)             : gray 8-bit or RGB 24-bit, can be <out> only!  |",
    "|   .h265 (H.265/HEVC Image)         : gray 8-bit              , can be <out> only!  |",
    "|                                                                                    |",
    "| switches:    -f                    : force overwrite of output file                |",
    "|              -0, -1, -2, -3, -4    : JPEG-LS near value or H.265 (qp-4)/6 value    |",
    "|------------------------------------------------------------------------------------|",
    "",
};

#define MAX_FILES 999                   /* src/main.c:84 */
#define HEVC_RGB_WARNING "   warning: this HEVCencoder currently only support gray 8-bit image instead of RGB image. Only compress the green channel of this image.\n"

enum { OUT_NONE = 0, OUT_PNM, OUT_PNG, OUT_BMP, OUT_QOI, OUT_JLS, OUT_HEVC };

typedef struct {
    const char *src, *dst;
    char *dst_own;                      /* default "<src>.png" name when no -o was given */
    char *canon_dst;                    /* canonical destination path, for the hazard check */
    char *text;                         /* this file's stdout lines */
    size_t text_len, text_cap;
    int resolved;                       /* nothing more will be appended to text */
    int ok;
} Job;

typedef struct {
    int kind, is_rgb, n, max_n;
    uint32_t h, w;
    size_t frame_bytes;
    int idx[MAX_FILES];                 /* job index of each frame of the batch */
    uint8_t *arena;                     /* page-locked when the engine is up, else malloc'd (the encode call then fails) */
    size_t arena_cap;
    int arena_pinned;
} Batch;

/* wall-clock phases of the call (IMCVT_TIMING=1) */
enum { T_STAGE_ALLOC = 0, T_READ, T_ENCODE, T_WRITE, T_SHUTDOWN, T_COUNT };
static double g_phase[T_COUNT];
static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static Job g_jobs[MAX_FILES];
static int g_njobs, g_printed, g_setting, g_batch_limit = 256;
static Batch g_batch;
/* devices a batch is sharded over: g_ndev < 0 = the engine's own device only (IMCVT_DEVICE), 0 = all visible, else the list */
static int g_devs[64], g_ndev = 0;
static int g_resident;                  /* this process serves requests (imcvt_serve.c): the engine and the arenas outlive a command */

static void parse_devices(const char *env)
{
    if (env == NULL || !strcmp(env, "all") || *env == 0) { g_ndev = 0; return; }
    if (strchr(env, ',') == NULL) {
        int k = atoi(env), i;
        if (k < 1) k = 1;
        if (k > 64) k = 64;
        for (i = 0; i < k; i++) g_devs[i] = i;
        g_ndev = k;
        return;
    }
    g_ndev = 0;
    while (*env && g_ndev < 64) {
        g_devs[g_ndev++] = atoi(env);
        env = strchr(env, ',');
        if (env == NULL) break;
        env++;
    }
}

/* the devices of this batch: no more than the batch is worth (a CUDA context per device costs more than a few frames) */
static int batch_devices(int n, int min_per_device, int *list)
{
    int visible = imcvtDeviceCount(), want, i, have;
    if (g_ndev < 0 || visible <= 1)
        return -1;
    have = g_ndev > 0 ? g_ndev : (visible < 64 ? visible : 64);
    want = n / min_per_device;
    if (want < 1) want = 1;
    if (want > have) want = have;
    for (i = 0; i < want; i++) list[i] = g_ndev > 0 ? g_devs[i] : i;
    return want;
}

static void say(Job *j, const char *fmt, ...)
{
    va_list ap;
    int need;
    va_start(ap, fmt);
    need = vsnprintf(NULL, 0, fmt, ap);
    va_end(ap);
    if (j->text_len + (size_t)need + 1 > j->text_cap) {
        j->text_cap = (j->text_len + (size_t)need + 1) * 2;
        j->text = (char *)realloc(j->text, j->text_cap);
        if (j->text == NULL) { fprintf(stderr, "out of memory\n"); exit(-1); }
    }
    va_start(ap, fmt);
    vsnprintf(j->text + j->text_len, (size_t)need + 1, fmt, ap);
    va_end(ap);
    j->text_len += (size_t)need;
}

/* a file's lines go out as soon as every file in front of it is resolved */
static void print_resolved(void)
{
    while (g_printed < g_njobs && g_jobs[g_printed].resolved) {
        Job *j = &g_jobs[g_printed++];
        if (j->text_len)
            fwrite(j->text, 1, j->text_len, stdout);
        free(j->text);
        j->text = NULL;
        j->text_len = j->text_cap = 0;
    }
    fflush(stdout);
}

static void fail(Job *j, const char *fmt, const char *name)
{
    say(j, "   ***ERROR: ");
    say(j, fmt, name);
    say(j, "\n");
    j->resolved = 1;
}

static int ends_with_nocase(const char *s, const char *suffix)
{
    size_t n = strlen(s), m = strlen(suffix), i;
    if (m > n)
        return 0;
    for (i = 0; i < m; i++) {
        int a = s[n - m + i], b = suffix[i];
        if (a >= 'A' && a <= 'Z') a += 32;
        if (b >= 'A' && b <= 'Z') b += 32;
        if (a != b)
            return 0;
    }
    return 1;
}

static int out_kind(const char *dst)            /* suffix dispatch of src/main.c:189-204 (no dot required there either) */
{
    if (ends_with_nocase(dst, "pnm") || ends_with_nocase(dst, "ppm") || ends_with_nocase(dst, "pgm")) return OUT_PNM;
    if (ends_with_nocase(dst, "png")) return OUT_PNG;
    if (ends_with_nocase(dst, "bmp")) return OUT_BMP;
    if (ends_with_nocase(dst, "qoi")) return OUT_QOI;
    if (ends_with_nocase(dst, "jls")) return OUT_JLS;
    if (ends_with_nocase(dst, "h265") || ends_with_nocase(dst, "265") || ends_with_nocase(dst, "hevc")) return OUT_HEVC;
    return OUT_NONE;
}

/* "<src minus its last suffix>.png" (src/main.c:50-74: the suffix is looked for after the last path separator only) */
static char *default_png_name(const char *src)
{
    size_t n = strlen(src), cut = n, i;
    char *s = (char *)malloc(n + 5);
    if (s == NULL) { fprintf(stderr, "out of memory\n"); exit(-1); }
    for (i = n; i-- > 0;) {
        if (src[i] == '/' || src[i] == '\\')
            break;
        if (src[i] == '.') { cut = i; break; }
    }
    memcpy(s, src, cut);
    memcpy(s + cut, ".png", 5);
    return s;
}

static int file_exists(const char *name)
{
    FILE *fp = fopen(name, "rb");
    if (fp)
        fclose(fp);
    return fp != NULL;
}

/* canonical form of a path that may not exist yet: realpath of its directory + its last component */
static char *canon_path(const char *path)
{
    char buf[PATH_MAX], dir[PATH_MAX];
    const char *slash = strrchr(path, '/'), *base = slash ? slash + 1 : path;
    char *out;
    size_t n;
    if (realpath(path, buf))
        return strdup(buf);
    if (slash == NULL) {
        strcpy(dir, ".");
    } else {
        n = (size_t)(slash - path);
        if (n == 0) n = 1;                                          /* "/name" */
        if (n >= sizeof dir)
            return strdup(path);
        memcpy(dir, path, n);
        dir[n] = 0;
    }
    if (!realpath(dir, buf))
        return strdup(path);
    out = (char *)malloc(strlen(buf) + strlen(base) + 2);
    if (out)
        sprintf(out, "%s/%s", buf, base);
    return out;
}

static int dump(const char *name, const uint8_t *buf, size_t len)
{
    int failed = 1;
    FILE *fp = fopen(name, "wb");
    if (fp) {
        failed = (len != fwrite(buf, 1, len, fp));
        fclose(fp);
    }
    return failed;
}

/* ---- the pending batch ------------------------------------------------------------------------------------------- */

/* the batch whose streams are being written (at most one in flight) */
typedef struct {
    pthread_t thread;
    int active, n, next, encode_failed, out_kept;     /* out_kept: `out` is one of the resident process's stream arenas, not to be freed */
    int idx[MAX_FILES];
    uint8_t *out;
    size_t *len, bound;
    unsigned char write_failed[MAX_FILES];
    double seconds;
} Writer;
static Writer g_writer;

/* the files of a batch are independent (no two pending jobs share a destination: convert() flushes first), so a large batch is
 * written by a few threads that take files off one counter */
static void *writer_worker(void *arg)
{
    Writer *Wr = (Writer *)arg;
    for (;;) {
        const int i = __atomic_fetch_add(&Wr->next, 1, __ATOMIC_RELAXED);
        const Job *j;
        if (i >= Wr->n)
            break;
        j = &g_jobs[Wr->idx[i]];
        Wr->write_failed[i] = (unsigned char)(Wr->encode_failed || Wr->len[i] == 0 || dump(j->dst, Wr->out + (size_t)i * Wr->bound, Wr->len[i]));
    }
    return NULL;
}

static void *writer_main(void *arg)
{
    Writer *Wr = (Writer *)arg;
    const double t0 = now_s();
    pthread_t helper[3];
    int nh = 0, k;
    Wr->next = 0;
    if (Wr->n >= 8)
        for (k = 0; k < 3; k++)
            if (pthread_create(&helper[nh], NULL, writer_worker, Wr) == 0)
                nh++;
    writer_worker(Wr);
    for (k = 0; k < nh; k++)
        pthread_join(helper[k], NULL);
    Wr->seconds = now_s() - t0;
    return NULL;
}

/* waits for the writer and settles its files' lines */
static void writer_join(void)
{
    Writer *Wr = &g_writer;
    int i;
    if (!Wr->active)
        return;
    if (Wr->active == 1) {
        const double t0 = now_s();
        pthread_join(Wr->thread, NULL);
        g_phase[T_WRITE] += now_s() - t0;           /* only what the main thread had to wait */
    }
    for (i = 0; i < Wr->n; i++) {
        Job *j = &g_jobs[Wr->idx[i]];
        if (Wr->write_failed[i]) {
            fail(j, "write %s failed", j->dst);
        } else {
            j->ok = 1;
            j->resolved = 1;
        }
    }
    if (!Wr->out_kept)
        free(Wr->out);
    free(Wr->len);
    Wr->out = NULL; Wr->len = NULL;
    Wr->active = 0;
    print_resolved();
}

/* Raw PGM / PPM pixels of codec jobs are not read where the file is looked at: the job only notes (file, offset, slot), and the
 * notes of a batch are carried out together by a few threads (pread straight into the page-locked arena) -- the copy out of the page
 * cache is what a staged file costs, and one thread moves ~6 GB/s.  A read is only postponed when the file is long enough for it to
 * succeed, so the outcome of the job is known at once; and every postponed read is carried out before this process writes any file
 * (read_pending_pixels() in front of the direct writers and of the batch's encode), so a later job can never change what an earlier
 * one was going to read. */
typedef struct { FILE *fp; long offset; size_t bytes; uint8_t *dst; int job; } PendingRead;
static PendingRead g_reads[MAX_FILES];
static int g_nreads, g_read_next;

static void *read_worker(void *arg)
{
    (void)arg;
    for (;;) {
        const int i = __atomic_fetch_add(&g_read_next, 1, __ATOMIC_RELAXED);
        PendingRead *r;
        size_t done = 0;
        if (i >= g_nreads)
            break;
        r = &g_reads[i];
        while (done < r->bytes) {
            const ssize_t k = pread(fileno(r->fp), r->dst + done, r->bytes - done, (off_t)r->offset + (off_t)done);
            if (k <= 0)
                break;
            done += (size_t)k;
        }
        if (done != r->bytes)
            memset(r->dst + done, 0, r->bytes - done);      /* the file shrank under us: the frame is coded as far as it was there */
        fclose(r->fp);
    }
    return NULL;
}

static void read_pending_pixels(void)
{
    const double t0 = now_s();
    pthread_t helper[3];
    int nh = 0, k;
    if (g_nreads == 0)
        return;
    g_read_next = 0;
    if (g_nreads >= 4)
        for (k = 0; k < 3; k++)
            if (pthread_create(&helper[nh], NULL, read_worker, NULL) == 0)
                nh++;
    read_worker(NULL);
    for (k = 0; k < nh; k++)
        pthread_join(helper[k], NULL);
    g_nreads = 0;
    g_phase[T_READ] += now_s() - t0;
}

/* 1 = the pixels of this raw PNM file will be read later (fp is kept); 0 = read them now */
static int postpone_read(FILE *fp, const ImcvtPNMHeader *hd, uint8_t *slot, int job)
{
    struct stat st;
    long at;
    if ((hd->type != 5 && hd->type != 6) || g_nreads >= MAX_FILES)
        return 0;
    at = ftell(fp);
    if (at < 0 || fstat(fileno(fp), &st) != 0 || !S_ISREG(st.st_mode) || (uint64_t)st.st_size < (uint64_t)at + hd->bytes)
        return 0;
    g_reads[g_nreads].fp = fp; g_reads[g_nreads].offset = at; g_reads[g_nreads].bytes = hd->bytes;
    g_reads[g_nreads].dst = slot; g_reads[g_nreads].job = job;
    g_nreads++;
    {
        /* each note holds a file open: carried out in lots that stay far below RLIMIT_NOFILE */
        static int lot;
        if (lot == 0) {
            struct rlimit rl;
            lot = 32;
            if (getrlimit(RLIMIT_NOFILE, &rl) == 0 && rl.rlim_cur != RLIM_INFINITY && rl.rlim_cur < 24 + 2 * 32)
                lot = rl.rlim_cur > 26 ? (int)(rl.rlim_cur - 24) / 2 : 1;
        }
        if (g_nreads >= lot)
            read_pending_pixels();
    }
    return 1;
}

/* A resident process (imcvt_serve.c) keeps two PAGE-LOCKED stream arenas and alternates between them: the writer thread of batch k
 * writes files out of one while batch k + 1 is encoded into the other.  With frames and slots both page-locked a gray NEAR = 0 batch is
 * coded in place over PCIe (no staging copies); a one-shot process does not pin hundreds of megabytes of slots
 * for a single use and takes pageable memory. */
static struct { uint8_t *p; size_t cap; } g_out_arena[2];
static int g_out_next;
static uint8_t *stream_arena(size_t need)
{
    uint8_t *p;
    const int k = g_out_next;
    if (!g_resident)
        return NULL;
    if (g_out_arena[k].cap < need) {
        if (g_out_arena[k].p) imcvtHostFree(g_out_arena[k].p);
        g_out_arena[k].p = (uint8_t *)imcvtHostAlloc(need);
        g_out_arena[k].cap = g_out_arena[k].p ? need : 0;
    }
    p = g_out_arena[k].p;
    if (p) g_out_next = k ^ 1;
    return p;
}
static void stream_arenas_release(void)
{
    int k;
    for (k = 0; k < 2; k++) {
        if (g_out_arena[k].p) imcvtHostFree(g_out_arena[k].p);
        g_out_arena[k].p = NULL;
        g_out_arena[k].cap = 0;
    }
}

static void flush_batch(void)
{
    Batch *B = &g_batch;
    Writer *Wr = &g_writer;
    size_t bound, *len;
    uint8_t *out;
    int failed, i, out_kept;
    read_pending_pixels();
    if (B->n == 0)
        return;
    bound = B->kind == OUT_JLS ? imcvtJLSBound(B->is_rgb, B->h, B->w) : imcvtHEVCBound(B->h, B->w);
    {
        const double t0 = now_s();
        out = stream_arena((size_t)B->n * bound);
        g_phase[T_STAGE_ALLOC] += now_s() - t0;
    }
    out_kept = out != NULL;
    if (out == NULL)
        out = (uint8_t *)malloc((size_t)B->n * bound);
    len = (size_t *)calloc((size_t)B->n, sizeof *len);
    failed = (out == NULL || len == NULL);
    if (!failed) {
        const double t0 = now_s();
        int devs[64];
        const int nd = batch_devices(B->n, B->kind == OUT_JLS ? 8 : 2, devs);
        if (nd > 1 || (nd == 1 && g_ndev > 0)) {
            if (B->kind == OUT_JLS)
                failed = imcvtEncodeBatchJLSMulti(B->arena, B->frame_bytes, B->n, B->is_rgb, B->h, B->w, g_setting, out, bound, len, devs, nd);
            else if (B->is_rgb)
                failed = imcvtEncodeBatchHEVCGreenMulti(B->arena, B->frame_bytes, B->n, B->h, B->w, g_setting, out, bound, len, devs, nd);
            else
                failed = imcvtEncodeBatchHEVCMulti(B->arena, B->frame_bytes, B->n, B->h, B->w, g_setting, out, bound, len, NULL, 0, devs, nd);
        } else if (B->kind == OUT_JLS)
            failed = imcvtEncodeBatchJLS(B->arena, B->frame_bytes, B->n, B->is_rgb, B->h, B->w, g_setting, out, bound, len);
        else if (B->is_rgb)
            failed = imcvtEncodeBatchHEVCGreen(B->arena, B->frame_bytes, B->n, B->h, B->w, g_setting, out, bound, len);
        else
            failed = imcvtEncodeBatchHEVC(B->arena, B->frame_bytes, B->n, B->h, B->w, g_setting, out, bound, len, NULL, 0);
        if (failed)
            fprintf(stderr, "imcvt_b200: %s\n", imcvtLastError());
        g_phase[T_ENCODE] += now_s() - t0;
    }
    /* the streams go to the writer thread; the arena is free for the next batch at once */
    writer_join();
    Wr->n = B->n; Wr->out = out; Wr->len = len; Wr->bound = bound; Wr->encode_failed = failed; Wr->out_kept = out_kept;
    for (i = 0; i < B->n; i++) Wr->idx[i] = B->idx[i];
    B->n = 0;
    if (pthread_create(&Wr->thread, NULL, writer_main, Wr) == 0) {
        Wr->active = 1;
    } else {                                        /* no thread to be had: write here; writer_join() has nothing to wait for */
        writer_main(Wr);
        Wr->active = 2;
        writer_join();
    }
}

static void arena_release(void)
{
    Batch *B = &g_batch;
    if (B->arena) {
        if (B->arena_pinned) imcvtHostFree(B->arena); else free(B->arena);
    }
    B->arena = NULL;
    B->arena_cap = 0;
}

/* opens a batch for frames of this key unless the pending one already has it; returns the slot of the next frame */
static uint8_t *batch_slot(int kind, int is_rgb, uint32_t h, uint32_t w)
{
    Batch *B = &g_batch;
    const size_t frame_bytes = (size_t)h * w * (is_rgb ? 3 : 1);
    if (B->n && (B->kind != kind || B->is_rgb != is_rgb || B->h != h || B->w != w || B->n >= B->max_n))
        flush_batch();
    if (B->n == 0) {
        /* frames per batch: the limit, at most about 2 GiB of staged pixels, and at most about 4 GiB of worst-case stream
         * slots (the slots exist once in host memory and once in device memory; real streams are ~10x smaller) */
        size_t fit = ((size_t)2 << 30) / frame_bytes, need;
        const size_t slot = kind == OUT_JLS ? imcvtJLSBound(is_rgb, h, w) : imcvtHEVCBound(h, w);
        if (slot && fit > ((size_t)4 << 30) / slot)
            fit = ((size_t)4 << 30) / slot;
        B->kind = kind; B->is_rgb = is_rgb; B->h = h; B->w = w; B->frame_bytes = frame_bytes;
        B->max_n = (int)(fit < 1 ? 1 : fit > (size_t)g_batch_limit ? (size_t)g_batch_limit : fit);
        /* the arena only grows, and in whole batches of the CURRENT shape as long as files of it may follow */
        need = frame_bytes * (size_t)B->max_n;
        if (need > B->arena_cap) {
            int remaining = g_njobs, cap_n;
            cap_n = remaining < B->max_n ? remaining : B->max_n;
            need = frame_bytes * (size_t)cap_n;
            B->max_n = cap_n;
        }
        if (need > B->arena_cap) {
            const double t0 = now_s();
            arena_release();
            B->arena = (uint8_t *)imcvtHostAlloc(need);
            B->arena_pinned = B->arena != NULL;
            if (B->arena == NULL)
                B->arena = (uint8_t *)malloc(need);     /* no device: the encode call reports it, the job fails like a write */
            g_phase[T_STAGE_ALLOC] += now_s() - t0;
            if (B->arena == NULL)
                return NULL;
            B->arena_cap = need;
        }
    }
    return B->arena + (size_t)B->n * frame_bytes;
}

/* The linked-in format files print their warnings straight to stdout (src/imageio_png.c:170); while they run, stdout is
 * pointed at a temporary file whose content is then appended to the job's own lines, so the order stays the reference's. */
static FILE *g_cap_file;
static int g_cap_saved = -1;

static void capture_begin(void)
{
    fflush(stdout);
    g_cap_file = tmpfile();
    if (g_cap_file == NULL)
        return;
    g_cap_saved = dup(1);
    if (g_cap_saved < 0 || dup2(fileno(g_cap_file), 1) < 0) {
        if (g_cap_saved >= 0) close(g_cap_saved);
        g_cap_saved = -1;
        fclose(g_cap_file);
        g_cap_file = NULL;
    }
}

static void capture_end(Job *j)
{
    char buf[1024];
    size_t n;
    if (g_cap_file == NULL)
        return;
    fflush(stdout);
    dup2(g_cap_saved, 1);
    close(g_cap_saved);
    g_cap_saved = -1;
    rewind(g_cap_file);
    while ((n = fread(buf, 1, sizeof buf - 1, g_cap_file)) > 0) {
        buf[n] = 0;
        say(j, "%s", buf);
    }
    fclose(g_cap_file);
    g_cap_file = NULL;
}

/* ---- one file ----------------------------------------------------------------------------------------------------- */

static uint8_t *load_other_formats(Job *j, const char *src, int *is_rgb, uint32_t *h, uint32_t *w)
{
    uint8_t *img = NULL;
    capture_begin();
    if (img == NULL) img = loadPNGImageFile(src, is_rgb, h, w);
    if (img == NULL) img = loadBMPImageFile(src, is_rgb, h, w);
    if (img == NULL) img = loadQOIImageFile(src, is_rgb, h, w);
    capture_end(j);
    return img;
}

static void convert(int i_file, int force_write)
{
    Job *j = &g_jobs[i_file];
    const int kind = out_kind(j->dst);
    uint8_t *img = NULL, *slot = NULL;
    uint32_t h = 0, w = 0;
    int is_rgb = 0, failed = 0, i, staged = 0;
    FILE *fp;

    /* sequential semantics: nothing this file reads or guards may still be owed by the pending batch or by the writer thread.  The
     * batch first: flushing it hands its files to the writer, and the writer check below then waits for them. */
    if (g_batch.n) {
        char *cs = canon_path(j->src);
        for (i = 0; i < g_batch.n; i++) {
            const char *pd = g_jobs[g_batch.idx[i]].canon_dst;
            if (pd && ((cs && !strcmp(cs, pd)) || (j->canon_dst && !strcmp(j->canon_dst, pd)))) {
                flush_batch();
                break;
            }
        }
        free(cs);
    }
    if (g_writer.active) {
        char *cs = canon_path(j->src);
        for (i = 0; i < g_writer.n; i++) {
            const char *pd = g_jobs[g_writer.idx[i]].canon_dst;
            if (pd && ((cs && !strcmp(cs, pd)) || (j->canon_dst && !strcmp(j->canon_dst, pd)))) {
                writer_join();
                break;
            }
        }
        free(cs);
    }

    say(j, "(%d/%d)  %s -> %s\n", i_file + 1, g_njobs, j->src, j->dst);
    if (!file_exists(j->src)) { fail(j, "%s not exist", j->src); return; }
    if (!force_write && file_exists(j->dst)) { fail(j, "%s already exist", j->dst); return; }

    /* load: PNM first (src/main.c:183); pixels of a codec job land in the staging arena, not in a buffer of their own */
    fp = fopen(j->src, "rb");
    if (fp == NULL && (errno == EMFILE || errno == ENFILE) && g_nreads) {
        read_pending_pixels();                      /* out of descriptors: the postponed reads hold some */
        fp = fopen(j->src, "rb");
    }
    if (fp != NULL) {
        ImcvtPNMHeader hd;
        if (!imcvtPNMReadHeader(fp, &hd)) {
            const int jls_dims_bad = hd.width > 32767 || hd.height > 32767;                 /* src/imageio_jls.c:436 */
            if ((kind == OUT_JLS && !jls_dims_bad) || kind == OUT_HEVC) {
                double t0;
                slot = batch_slot(kind, hd.is_rgb, hd.height, hd.width);
                t0 = now_s();
                if (slot && postpone_read(fp, &hd, slot, i_file)) {
                    fp = NULL;                                      /* closed by the read */
                    staged = 1;
                    is_rgb = hd.is_rgb; h = hd.height; w = hd.width;
                } else if (slot && !imcvtPNMReadPixels(fp, &hd, slot)) {
                    g_phase[T_READ] += now_s() - t0;
                    staged = 1;
                    is_rgb = hd.is_rgb; h = hd.height; w = hd.width;
                }
            } else if ((img = (uint8_t *)malloc(hd.bytes + 8)) != NULL) {
                if (imcvtPNMReadPixels(fp, &hd, img)) {
                    free(img);
                    img = NULL;
                } else {
                    is_rgb = hd.is_rgb; h = hd.height; w = hd.width;
                }
            }
        }
        if (fp)
            fclose(fp);
    }
    if (!staged && img == NULL)
        img = load_other_formats(j, j->src, &is_rgb, &h, &w);
    if (!staged && img == NULL) { fail(j, "open %s failed", j->src); return; }

    if (kind == OUT_NONE) {
        free(img);
        fail(j, "unsupported output suffix: %s", j->dst);
        return;
    }
    if (kind == OUT_JLS || kind == OUT_HEVC) {
        if (kind == OUT_HEVC && is_rgb)
            say(j, HEVC_RGB_WARNING);                                                       /* src/imageio_hevc.c:22 */
        if (!staged) {
            /* decoded by one of the other loaders (or JPEG-LS dimensions out of range): copy into the arena */
            if ((kind == OUT_JLS && (w < 1 || w > 32767 || h < 1 || h > 32767)) || w < 1 || h < 1 ||
                (slot = batch_slot(kind, is_rgb, h, w)) == NULL) {
                free(img);
                fail(j, "write %s failed", j->dst);
                return;
            }
            memcpy(slot, img, (size_t)h * w * (is_rgb ? 3 : 1));
            free(img);
        }
        g_batch.idx[g_batch.n++] = i_file;          /* resolved by flush_batch() */
        return;
    }
    read_pending_pixels();                          /* nothing is written while an earlier job still has pixels to read */
    capture_begin();
    switch (kind) {
    case OUT_PNM: failed = writePNMImageFile(j->dst, img, is_rgb, h, w); break;
    case OUT_PNG: failed = writePNGImageFile(j->dst, img, is_rgb, h, w); break;
    case OUT_BMP: failed = writeBMPImageFile(j->dst, img, is_rgb, h, w); break;
    default:      failed = writeQOIImageFile(j->dst, img, is_rgb, h, w); break;
    }
    capture_end(j);
    free(img);
    if (failed) { fail(j, "write %s failed", j->dst); return; }
    j->ok = 1;
    j->resolved = 1;
}

/* imcvt_serve.c: the resident engine (IMCVT_SERVE / IMCVT_SERVER) */
int imcvtClientRun(const char *path, int argc, char **argv, int *rc);
int imcvtServe(const char *path, int (*run)(int, char **));

/* one command line from argv to exit code (src/main.c:87-223) */
static int run_command(int argc, char **argv)
{
    int switches[128] = {0}, i, next_is_dst = 0, n_file = -1, force_write, n_success = 0, n_failed;
    const char *env;
    const double t_start = now_s();

    memset(g_jobs, 0, sizeof(Job) * (size_t)(g_njobs > 0 ? g_njobs : 0));      /* a resident process runs many commands */
    memset(g_phase, 0, sizeof g_phase);
    g_njobs = g_printed = 0;
    g_batch_limit = 256;
    g_batch.n = 0;

    for (i = 1; i < argc; i++) {                                    /* src/main.c:87-132 */
        const char *arg = argv[i];
        if (arg[0] == '-') {
            for (arg++; *arg; arg++) {
                if ((unsigned char)*arg < 128)
                    switches[(unsigned char)*arg] = 1;
                if (*arg == 'o')
                    next_is_dst = 1;
            }
        } else if (n_file + 1 < MAX_FILES) {
            if (next_is_dst) {
                next_is_dst = 0;
                if (n_file >= 0)
                    g_jobs[n_file].dst = arg;
            } else {
                g_jobs[++n_file].src = arg;
            }
        }
    }
    g_njobs = n_file + 1;
    force_write = switches['F'] || switches['f'];
    g_setting = switches['4'] ? 4 : switches['3'] ? 3 : switches['2'] ? 2 : switches['1'] ? 1 : 0;
    if (g_njobs <= 0) {
        for (i = 0; i < (int)(sizeof USAGE_LINES / sizeof USAGE_LINES[0]); i++)
            printf("%s\n", USAGE_LINES[i]);
        return -1;
    }
    if ((env = getenv("IMCVT_BATCH")) != NULL && atoi(env) > 0)
        g_batch_limit = atoi(env) < MAX_FILES ? atoi(env) : MAX_FILES;
    if ((env = getenv("IMCVT_DEVICE")) != NULL) {
        imcvtInit(atoi(env));                                       /* a failure shows up at the first encode call */
        g_ndev = -1;
    } else {
        parse_devices(getenv("IMCVT_DEVICES"));
    }

    for (i = 0; i < g_njobs; i++) {
        Job *j = &g_jobs[i];
        if (j->dst == NULL)
            j->dst = j->dst_own = default_png_name(j->src);
        j->canon_dst = canon_path(j->dst);
    }
    for (i = 0; i < g_njobs; i++) {
        convert(i, force_write);
        print_resolved();
    }
    flush_batch();
    writer_join();
    print_resolved();
    if (!g_resident) {
        const double t0 = now_s();
        arena_release();
        imcvtShutdown();
        g_phase[T_SHUTDOWN] = now_s() - t0;
    }
    if ((env = getenv("IMCVT_TIMING")) != NULL && atoi(env) > 0)
        fprintf(stderr, "imcvt_b200 timing: engine start + staging arena %.3f s, pixel reads %.3f s, encode calls %.3f s, "
                        "stream writes %.3f s, shutdown %.3f s, total %.3f s\n",
                g_phase[T_STAGE_ALLOC], g_phase[T_READ], g_phase[T_ENCODE], g_phase[T_WRITE], g_phase[T_SHUTDOWN], now_s() - t_start);

    for (i = 0; i < g_njobs; i++) {
        n_success += g_jobs[i].ok;
        free(g_jobs[i].dst_own);
        free(g_jobs[i].canon_dst);
    }
    n_failed = g_njobs - n_success;
    if (g_njobs > 1) {
        printf("\nsummary:");
        if (n_success) printf("  %d file converted", n_success);
        if (n_failed)  printf("  %d failed", n_failed);
        printf("\n");
    }
    return n_failed;
}

int main(int argc, char **argv)
{
    const char *env;
    if ((env = getenv("IMCVT_SERVE")) != NULL && *env) {
        /* resident engine: the CUDA context is created once, here, and every later command line finds it warm */
        const char *dev = getenv("IMCVT_DEVICE");
        int rc;
        g_resident = 1;
        imcvtInit(dev ? atoi(dev) : 0);                             /* without a device the server still converts between file formats */
        rc = imcvtServe(env, run_command);
        arena_release();
        stream_arenas_release();
        imcvtShutdown();
        return rc;
    }
    if ((env = getenv("IMCVT_SERVER")) != NULL && *env) {
        int rc;
        if (imcvtClientRun(env, argc, argv, &rc) == 0)
            return rc;                                              /* no server on that path: run here, as always */
    }
    return run_command(argc, argv);
}
