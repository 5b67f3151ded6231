/* imcvt_serve.c -- resident engine for the batch command line (what a one-shot process pays for).
 *
 * A cold `ImCvt_b200` run spends 0.7-1.1 s creating the CUDA context and page-locking its staging arena before the first
 * pixel is read -- for 64 full-HD files that is 20 x the encode call.  The reference has no such cost
 * (no device), so a drop-in that is called file by file from a script would lose to it.  This file keeps the engine alive
 * between invocations without touching the command line's surface:
 *
 *   IMCVT_SERVE=<socket path>  ImCvt_b200          starts a server: engine up, then requests one after the other
 *   IMCVT_SERVER=<socket path> ImCvt_b200 <args>   the usual command line; the work is done by the server's process
 *
 * Both are environment variables: argv means exactly what it means to the reference (src/main.c:87-132).  A client passes
 * its argument vector, working directory, the IMCVT_* variables that shape a call, and -- as SCM_RIGHTS file descriptors --
 * its own stdout and stderr; the server runs the request with those as fd 1 / fd 2, so text appears where and when it
 * would have appeared, and sends back the exit code.  When no server answers on the path the client runs in-process as usual.
 * Requests are served one at a time (the engine is one resource); the staging arena, the device workspaces and the contexts
 * of every GPU the batches were sharded over survive from one request to the next.
 */
#include <errno.h>
#include <signal.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <sys/un.h>
#include <unistd.h>

#define SERVE_MAGIC 0x31564D49u         /* "IMV1" */
static const char *const k_forwarded[] = { "IMCVT_BATCH", "IMCVT_DEVICES", "IMCVT_TIMING", NULL };

static int write_all(int fd, const void *p, size_t n)
{
    const char *c = (const char *)p;
    while (n) {
        const ssize_t k = write(fd, c, n);
        if (k < 0) { if (errno == EINTR) continue; return 1; }
        c += k; n -= (size_t)k;
    }
    return 0;
}
static int read_all(int fd, void *p, size_t n)
{
    char *c = (char *)p;
    while (n) {
        const ssize_t k = read(fd, c, n);
        if (k < 0) { if (errno == EINTR) continue; return 1; }
        if (k == 0) return 1;
        c += k; n -= (size_t)k;
    }
    return 0;
}

static int open_socket(const char *path, struct sockaddr_un *sa)
{
    if (strlen(path) >= sizeof sa->sun_path)
        return -1;
    memset(sa, 0, sizeof *sa);
    sa->sun_family = AF_UNIX;
    strcpy(sa->sun_path, path);
    return socket(AF_UNIX, SOCK_STREAM, 0);
}

/* ---- client ------------------------------------------------------------------------------------------------------ */

typedef struct { char *p; size_t n, cap; } Buf;
static void put(Buf *b, const void *src, size_t n)
{
    if (b->n + n > b->cap) {
        b->cap = (b->n + n) * 2 + 256;
        b->p = (char *)realloc(b->p, b->cap);
        if (b->p == NULL) { fprintf(stderr, "out of memory\n"); exit(-1); }
    }
    memcpy(b->p + b->n, src, n);
    b->n += n;
}
static void put_u32(Buf *b, uint32_t v) { put(b, &v, 4); }
static void put_str(Buf *b, const char *s) { const uint32_t n = (uint32_t)strlen(s); put_u32(b, n); put(b, s, n); }

/* 0 = the server ran the command, *rc is its exit code; 1 = no server (run in-process) */
int imcvtClientRun(const char *path, int argc, char **argv, int *rc)
{
    struct sockaddr_un sa;
    const int fd = open_socket(path, &sa);
    if (fd < 0)
        return 1;
    if (connect(fd, (struct sockaddr *)&sa, sizeof sa) != 0) {
        close(fd);
        return 1;
    }
    fflush(stdout);
    fflush(stderr);
    /* first message: the magic word, with this process's stdout and stderr attached */
    {
        uint32_t magic = SERVE_MAGIC;
        struct iovec iov = { &magic, sizeof magic };
        union { char buf[CMSG_SPACE(2 * sizeof(int))]; struct cmsghdr align; } u;
        struct msghdr msg;
        struct cmsghdr *cm;
        int fds[2] = { 1, 2 };
        memset(&msg, 0, sizeof msg);
        memset(&u, 0, sizeof u);
        msg.msg_iov = &iov; msg.msg_iovlen = 1;
        msg.msg_control = u.buf; msg.msg_controllen = sizeof u.buf;
        cm = CMSG_FIRSTHDR(&msg);
        cm->cmsg_level = SOL_SOCKET; cm->cmsg_type = SCM_RIGHTS; cm->cmsg_len = CMSG_LEN(sizeof fds);
        memcpy(CMSG_DATA(cm), fds, sizeof fds);
        if (sendmsg(fd, &msg, 0) != (ssize_t)sizeof magic) { close(fd); return 1; }
    }
    /* second message: working directory, arguments, environment */
    {
        Buf b = { NULL, 0, 0 };
        char cwd[4096];
        uint32_t total, nenv = 0;
        int i;
        if (getcwd(cwd, sizeof cwd) == NULL) { close(fd); return 1; }
        put_str(&b, cwd);
        put_u32(&b, (uint32_t)argc);
        for (i = 0; i < argc; i++) put_str(&b, argv[i]);
        for (i = 0; k_forwarded[i]; i++) nenv += getenv(k_forwarded[i]) != NULL;
        put_u32(&b, nenv);
        for (i = 0; k_forwarded[i]; i++)
            if (getenv(k_forwarded[i])) { put_str(&b, k_forwarded[i]); put_str(&b, getenv(k_forwarded[i])); }
        total = (uint32_t)b.n;
        if (write_all(fd, &total, 4) || write_all(fd, b.p, b.n)) { free(b.p); close(fd); return 1; }
        free(b.p);
    }
    {
        int32_t code;
        if (read_all(fd, &code, 4)) {           /* the server went away in the middle of the request: nothing sensible to fall back to */
            fprintf(stderr, "ImCvt_b200: the server at %s closed the connection\n", path);
            close(fd);
            *rc = -1;
            return 0;
        }
        *rc = code;
    }
    close(fd);
    return 0;
}

/* ---- server ------------------------------------------------------------------------------------------------------ */

static volatile sig_atomic_t g_stop;
static void on_signal(int sig) { (void)sig; g_stop = 1; }

static int get_u32(const char **p, const char *end, uint32_t *v)
{
    if (end - *p < 4) return 1;
    memcpy(v, *p, 4);
    *p += 4;
    return 0;
}
static int serve_one(int conn, int (*run)(int, char **))
{
    uint32_t magic = 0, total = 0;
    int fds[2] = { -1, -1 }, code = -1, ok = 0;
    char *payload = NULL;
    {
        struct iovec iov = { &magic, sizeof magic };
        union { char buf[CMSG_SPACE(2 * sizeof(int))]; struct cmsghdr align; } u;
        struct msghdr msg;
        struct cmsghdr *cm;
        memset(&msg, 0, sizeof msg);
        msg.msg_iov = &iov; msg.msg_iovlen = 1;
        msg.msg_control = u.buf; msg.msg_controllen = sizeof u.buf;
        if (recvmsg(conn, &msg, 0) != (ssize_t)sizeof magic || magic != SERVE_MAGIC)
            return 1;
        for (cm = CMSG_FIRSTHDR(&msg); cm; cm = CMSG_NXTHDR(&msg, cm))
            if (cm->cmsg_level == SOL_SOCKET && cm->cmsg_type == SCM_RIGHTS && cm->cmsg_len == CMSG_LEN(sizeof fds))
                memcpy(fds, CMSG_DATA(cm), sizeof fds);
        if (fds[0] < 0 || fds[1] < 0)
            return 1;
    }
    if (!read_all(conn, &total, 4) && total < (64u << 20) && (payload = (char *)malloc((size_t)total + 1)) != NULL && !read_all(conn, payload, total)) {
        const char *p = payload, *end = payload + total;
        char **argv = NULL, *cwd = NULL;
        uint32_t n = 0, argc = 0, nenv = 0, i;
        int bad = 0;
        payload[total] = 0;
        /* the fields are copied out: each string needs its terminator */
        #define TAKE_STR(dst) do { if (get_u32(&p, end, &n) || (size_t)(end - p) < n) { bad = 1; } else { (dst) = (char *)malloc((size_t)n + 1); \
                                   if ((dst) == NULL) bad = 1; else { memcpy((dst), p, n); (dst)[n] = 0; p += n; } } } while (0)
        TAKE_STR(cwd);
        if (!bad && !get_u32(&p, end, &argc) && argc < (1u << 20) && (argv = (char **)calloc((size_t)argc + 1, sizeof *argv)) != NULL) {
            for (i = 0; i < argc && !bad; i++) TAKE_STR(argv[i]);
            if (!bad && !get_u32(&p, end, &nenv) && nenv < 64) {
                int k;
                for (k = 0; k_forwarded[k]; k++) unsetenv(k_forwarded[k]);
                for (i = 0; i < nenv && !bad; i++) {
                    char *key = NULL, *val = NULL;
                    TAKE_STR(key);
                    if (!bad) TAKE_STR(val);
                    if (!bad)
                        for (k = 0; k_forwarded[k]; k++)
                            if (!strcmp(key, k_forwarded[k])) setenv(key, val, 1);
                    free(key); free(val);
                }
                if (!bad && chdir(cwd) == 0) {
                    /* the request runs with the client's stdout / stderr as fd 1 / fd 2 */
                    const int keep1 = dup(1), keep2 = dup(2);
                    fflush(stdout); fflush(stderr);
                    dup2(fds[0], 1); dup2(fds[1], 2);
                    code = run((int)argc, argv);
                    fflush(stdout); fflush(stderr);
                    dup2(keep1, 1); dup2(keep2, 2);
                    close(keep1); close(keep2);
                    ok = 1;
                }
            }
        }
        #undef TAKE_STR
        if (argv) { for (i = 0; i < argc; i++) free(argv[i]); free(argv); }
        free(cwd);
    }
    free(payload);
    close(fds[0]); close(fds[1]);
    if (ok) {
        const int32_t c = code;
        write_all(conn, &c, 4);
    }
    return !ok;
}

/* runs until SIGTERM / SIGINT; returns the process's exit code */
int imcvtServe(const char *path, int (*run)(int, char **))
{
    struct sockaddr_un sa;
    struct sigaction act;
    const int fd = open_socket(path, &sa);
    if (fd < 0) { fprintf(stderr, "ImCvt_b200: bad socket path %s\n", path); return 1; }
    {
        /* a socket file may be left over from a server that died -- or belong to one that is alive: ask before taking the path */
        struct sockaddr_un probe_sa;
        const int probe = open_socket(path, &probe_sa);
        if (probe >= 0) {
            const int alive = connect(probe, (struct sockaddr *)&probe_sa, sizeof probe_sa) == 0;
            close(probe);
            if (alive) { fprintf(stderr, "ImCvt_b200: a server is already listening on %s\n", path); close(fd); return 1; }
        }
    }
    unlink(path);
    {
        const mode_t old = umask(0077);                 /* the socket is the owner's */
        const int rc = bind(fd, (struct sockaddr *)&sa, sizeof sa);
        umask(old);
        if (rc != 0 || listen(fd, 64) != 0) { fprintf(stderr, "ImCvt_b200: cannot listen on %s: %s\n", path, strerror(errno)); close(fd); return 1; }
    }
    memset(&act, 0, sizeof act);
    act.sa_handler = on_signal;                         /* no SA_RESTART: accept() returns and the loop ends */
    sigaction(SIGTERM, &act, NULL);
    sigaction(SIGINT, &act, NULL);
    signal(SIGPIPE, SIG_IGN);                           /* a client that went away must not take the server with it */
    fprintf(stderr, "ImCvt_b200: serving on %s\n", path);
    while (!g_stop) {
        const int conn = accept(fd, NULL, NULL);
        if (conn < 0) {
            if (errno == EINTR) continue;
            break;
        }
        {
            /* a client that connects and then says nothing must not hold up the queue */
            struct timeval tv = { 5, 0 };
            setsockopt(conn, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
        }
        serve_one(conn, run);
        close(conn);
    }
    close(fd);
    unlink(path);
    return 0;
}
