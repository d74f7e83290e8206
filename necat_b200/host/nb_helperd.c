/* nb_helperd -- the resident helper of oc2pmov / oc2cns (nb_helper.h): nb_helperd <socket> <device>.
 * One job at a time, in the caller's directory, writing to the caller's stdout / stderr; leaves after
 * NECAT_B200_HELPER_IDLE seconds (default 30) without a call, or when a job fails hard (the tools exit on errors). */
#define _GNU_SOURCE
#include <errno.h>
#include <getopt.h>
#include <limits.h>
#include <signal.h>
#include <string.h>
#include <sys/select.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>

#include "nb_helper.h"

static char g_sock[256];
static void cleanup(void) { if (g_sock[0]) unlink(g_sock); }
static void on_signal(int s) { (void)s; cleanup(); _exit(1); }

static int read_all(int fd, void* p, size_t n)
{
    char* c = (char*)p;
    while (n) {
        const ssize_t r = read(fd, c, n);
        if (r < 0) { if (errno == EINTR) continue; return -1; }
        if (r == 0) return -1;
        c += r; n -= (size_t)r;
    }
    return 0;
}

int main(int argc, char** argv)
{
    if (argc != 3) { fprintf(stderr, "usage: nb_helperd <socket> <device>   (started by oc2pmov / oc2cns with NECAT_B200_HELPER=1)\n"); return 2; }
    snprintf(g_sock, sizeof g_sock, "%s", argv[1]);
    const int device = atoi(argv[2]);
    const char* idle_s = getenv("NECAT_B200_HELPER_IDLE");
    const int idle = idle_s ? atoi(idle_s) : 30;
    signal(SIGPIPE, SIG_IGN);
    signal(SIGTERM, on_signal);
    signal(SIGINT, on_signal);

    const int ls = socket(AF_UNIX, SOCK_STREAM, 0);
    if (ls < 0) return 1;
    struct sockaddr_un sa;
    memset(&sa, 0, sizeof sa);
    sa.sun_family = AF_UNIX;
    if (strlen(g_sock) >= sizeof sa.sun_path) return 1;
    strcpy(sa.sun_path, g_sock);
    {
        /* a socket file nobody answers on is left over from a helper that was killed */
        const int probe = socket(AF_UNIX, SOCK_STREAM, 0);
        if (probe >= 0 && connect(probe, (struct sockaddr*)&sa, sizeof sa) == 0) { close(probe); g_sock[0] = 0; return 0; }   /* another helper serves it */
        if (probe >= 0) close(probe);
        unlink(g_sock);
    }
    const mode_t old = umask(0077);
    if (bind(ls, (struct sockaddr*)&sa, sizeof sa) != 0) { g_sock[0] = 0; return 1; }
    umask(old);
    atexit(cleanup);
    if (listen(ls, 16) != 0) return 1;

    nbh_in_helper = 1;
    {
        /* the whole point: the context exists before the first caller asks (it is what they wait for) */
        nb_ctx* ctx = NULL;
        setenv("NECAT_B200_DEVICE", argv[2], 1);
        if (nbh_ctx_get(nbh_pick_device(), &ctx) != 0) return 1;
    }
    (void)device;
    for (;;) {
        fd_set rf;
        FD_ZERO(&rf);
        FD_SET(ls, &rf);
        struct timeval tv = {idle > 0 ? idle : 1, 0};
        const int r = select(ls + 1, &rf, NULL, NULL, &tv);
        if (r < 0 && errno == EINTR) continue;
        if (r <= 0) break;                                  /* idle: leave */
        const int c = accept(ls, NULL, NULL);
        if (c < 0) continue;
        uint32_t hdr[4];
        int fds[2] = {-1, -1};
        {
            struct msghdr msg;
            struct iovec iov;
            char ctl[CMSG_SPACE(2 * sizeof(int))];
            memset(&msg, 0, sizeof msg);
            iov.iov_base = hdr; iov.iov_len = sizeof hdr;
            msg.msg_iov = &iov; msg.msg_iovlen = 1;
            msg.msg_control = ctl; msg.msg_controllen = sizeof ctl;
            const ssize_t n = recvmsg(c, &msg, MSG_WAITALL);
            struct cmsghdr* cm = CMSG_FIRSTHDR(&msg);
            if (n == (ssize_t)sizeof hdr && cm && cm->cmsg_level == SOL_SOCKET && cm->cmsg_type == SCM_RIGHTS && cm->cmsg_len == CMSG_LEN(2 * sizeof(int)))
                memcpy(fds, CMSG_DATA(cm), sizeof fds);
            if (n != (ssize_t)sizeof hdr || hdr[0] != 0x4e424831u || fds[0] < 0 || hdr[2] < 1 || hdr[2] > 4096 || hdr[3] > (1u << 20)) {
                if (fds[0] >= 0) close(fds[0]);
                if (fds[1] >= 0) close(fds[1]);
                close(c);
                continue;
            }
        }
        char* buf = (char*)malloc(hdr[3] + 1);
        if (!buf || read_all(c, buf, hdr[3]) != 0) { free(buf); close(fds[0]); close(fds[1]); close(c); continue; }
        buf[hdr[3]] = 0;
        const int jargc = (int)hdr[2];
        char** jargv = (char**)calloc((size_t)jargc + 1, sizeof(char*));
        const char* cwd = buf;
        size_t at = strlen(buf) + 1;
        int ok = jargv != NULL;
        for (int i = 0; ok && i < jargc; ++i) {
            if (at >= hdr[3]) { ok = 0; break; }
            jargv[i] = buf + at;
            at += strlen(buf + at) + 1;
        }
        int32_t st = 1;
        if (ok && chdir(cwd) == 0) {
            fflush(stdout); fflush(stderr);
            const int so = dup(1), se = dup(2);
            dup2(fds[0], 1); dup2(fds[1], 2);
            optind = 1;
            st = hdr[1] == NBH_TOOL_PMOV ? oc2pmov_main(jargc, jargv) : oc2cns_main(jargc, jargv);
            fflush(stdout); fflush(stderr);
            dup2(so, 1); dup2(se, 2);
            close(so); close(se);
        }
        close(fds[0]); close(fds[1]);
        if (write(c, &st, sizeof st) != (ssize_t)sizeof st) { /* the caller has gone */ }
        close(c);
        free(jargv); free(buf);
    }
    return 0;
}
