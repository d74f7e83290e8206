/*
 * pigz -- drop-in for the way necat.pl calls the pigz it vendors (tool/pigz-2.4.tar.gz; necat.pl:124,352-353, Necat.pm:25):
 *     pigz -f -p <threads> <file>      ->  <file>.gz, <file> removed
 * (after a 0.3 s correction iteration the single-threaded passes around it are what a user waits for; the
 * vendored pigz 2.4 also refuses to start against zlib 1.3: "zlib version less than 1.2.3".)
 * The file is cut into blocks of 4 MiB, every block becomes a gzip member of its own on one of <threads> workers, the
 * members are written in order: a valid gzip file for zlib's gzread (oc2mkdb, fsa_rd_tools), gzip -d and zcat.  Options:
 * -p N, -f (overwrite), -k (keep the input), -1 .. -9, -c (to stdout); several files.  No decompression: gzip -d does that.
 */
#include <errno.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

enum { kBlock = 4 << 20 };

typedef struct {
    unsigned char* in; size_t n_in;
    unsigned char* out; size_t n_out, cap_out;
    int level, rc;
} job_t;

static void* work(void* arg)
{
    job_t* j = (job_t*)arg;
    z_stream z;
    memset(&z, 0, sizeof z);
    j->rc = deflateInit2(&z, j->level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY);       /* + 16: gzip wrapper */
    if (j->rc != Z_OK) return NULL;
    z.next_in = j->in; z.avail_in = (uInt)j->n_in;
    z.next_out = j->out; z.avail_out = (uInt)j->cap_out;
    j->rc = deflate(&z, Z_FINISH) == Z_STREAM_END ? Z_OK : Z_BUF_ERROR;
    j->n_out = j->cap_out - z.avail_out;
    deflateEnd(&z);
    return NULL;
}

static int compress_file(const char* path, int threads, int level, int force, int keep, int to_stdout)
{
    FILE* in = fopen(path, "rb");
    if (!in) { fprintf(stderr, "pigz: %s: %s\n", path, strerror(errno)); return 1; }
    char* opath = NULL;
    FILE* out = stdout;
    if (!to_stdout) {
        opath = (char*)malloc(strlen(path) + 4);
        sprintf(opath, "%s.gz", path);
        if (!force && access(opath, F_OK) == 0) { fprintf(stderr, "pigz: %s exists -- use -f to overwrite\n", opath); fclose(in); free(opath); return 1; }
        out = fopen(opath, "wb");
        if (!out) { fprintf(stderr, "pigz: %s: %s\n", opath, strerror(errno)); fclose(in); free(opath); return 1; }
    }
    job_t* jobs = (job_t*)calloc((size_t)threads, sizeof(job_t));
    pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    const size_t cap = compressBound(kBlock) + 64;
    for (int t = 0; t < threads; ++t) {
        jobs[t].in = (unsigned char*)malloc(kBlock); jobs[t].out = (unsigned char*)malloc(cap); jobs[t].cap_out = cap; jobs[t].level = level;
        if (!jobs[t].in || !jobs[t].out) { fprintf(stderr, "pigz: out of memory\n"); return 1; }
    }
    int rc = 0, wrote = 0;
    for (;;) {
        int n = 0;
        for (; n < threads; ++n) {
            jobs[n].n_in = fread(jobs[n].in, 1, kBlock, in);
            if (jobs[n].n_in == 0) break;
        }
        if (n == 0) break;
        for (int t = 1; t < n; ++t) if (pthread_create(&th[t], NULL, work, &jobs[t])) { work(&jobs[t]); th[t] = 0; }
        work(&jobs[0]);
        for (int t = 1; t < n; ++t) if (th[t]) pthread_join(th[t], NULL);
        for (int t = 0; t < n; ++t) {
            if (jobs[t].rc != Z_OK) { fprintf(stderr, "pigz: deflate failed on %s\n", path); rc = 1; }
            else if (fwrite(jobs[t].out, 1, jobs[t].n_out, out) != jobs[t].n_out) { fprintf(stderr, "pigz: write error on %s\n", opath ? opath : "stdout"); rc = 1; }
        }
        wrote = 1;
        if (rc || n < threads) break;
    }
    if (!wrote && !rc) {                    /* an empty file is one empty member */
        jobs[0].n_in = 0; work(&jobs[0]);
        if (jobs[0].rc != Z_OK || fwrite(jobs[0].out, 1, jobs[0].n_out, out) != jobs[0].n_out) rc = 1;
    }
    if (ferror(in)) { fprintf(stderr, "pigz: read error on %s\n", path); rc = 1; }
    fclose(in);
    if (!to_stdout && fclose(out)) { fprintf(stderr, "pigz: write error on %s\n", opath); rc = 1; }
    if (to_stdout) fflush(stdout);
    if (!rc && !to_stdout) {
        struct stat st;
        if (stat(path, &st) == 0) chmod(opath, st.st_mode & 07777);
        if (!keep) unlink(path);
    } else if (rc && opath) unlink(opath);
    for (int t = 0; t < threads; ++t) { free(jobs[t].in); free(jobs[t].out); }
    free(jobs); free(th); free(opath);
    return rc;
}

int main(int argc, char** argv)
{
    int threads = (int)sysconf(_SC_NPROCESSORS_ONLN), level = 6, force = 0, keep = 0, to_stdout = 0, c;
    if (threads < 1) threads = 1;
    while ((c = getopt(argc, argv, "p:fkc123456789q")) != -1) {
        if (c == 'p') threads = atoi(optarg);
        else if (c == 'f') force = 1;
        else if (c == 'k') keep = 1;
        else if (c == 'c') to_stdout = keep = 1;
        else if (c >= '1' && c <= '9') level = c - '0';
        else if (c == 'q') continue;
        else { fprintf(stderr, "usage: pigz [-f] [-k] [-c] [-1..-9] [-p threads] file...   (compression only; gzip -d decompresses)\n"); return 1; }
    }
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if (optind >= argc) { fprintf(stderr, "pigz: no file given (this drop-in compresses named files)\n"); return 1; }
    int rc = 0;
    for (int i = optind; i < argc; ++i) rc |= compress_file(argv[i], threads, level, force, keep, to_stdout);
    return rc;
}
