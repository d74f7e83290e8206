/*
 * fsa_rd_tools -- drop-in replacement of the two commands of the reference's fsa_rd_tools that `necat.pl correct` calls
 * (fsa/read_tools.cpp, fsa/prog/fsa_rd_tools.cpp):
 *
 *     fsa_rd_tools longest --ifname X --ofname Y [--base_size B] [--min_length L] [--discard_illegal_read]
 *         (necat.pl:122, Necat.pm:22)  the longest reads of X that together hold at least B bases, not shorter than L, in
 *         input order, as FASTA (ReadTools::Longest, read_tools.cpp:196-242; FindLongestXHeap, utility.cpp:25-49)
 *     fsa_rd_tools n50 --ifname X [--genome_size G]
 *         (necat.pl:1495)  Count / Total / Max / Min / N25 L25 N50 L50 N75 L75 (ReadTools::StatN50, read_tools.cpp:43-76)
 *
 * Host only, plain C + zlib.  The reference decompresses and parses its input twice, one line object at a time; here the
 * input is inflated once through a 4 MiB buffer and, while it fits kKeepBytes, kept in memory for the second pass.
 * Record boundaries follow the reference's readers to the letter (seq_io.cpp, file_io.cpp:105-116): every line is stripped
 * of white space at both ends, empty lines do not exist, a FASTA record is a '>' line and the lines up to the next '>' line,
 * a FASTQ record is four lines; the name is the first word of the header, the rest of the line is written back behind one
 * space (FastaWriter, seq_io.hpp:86-94: the space is written even when nothing follows it).
 * The other commands of the reference's tool (split, split_name, check) are not needed by the correct stage and end with
 * an error here.
 */
#include <ctype.h>
#include <getopt.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <zlib.h>

static void logmsg(const char* level, const char* fmt, const char* a, long long b)
{
    time_t t = time(NULL);
    struct tm tmv;
    localtime_r(&t, &tmv);
    char ts[64];
    strftime(ts, sizeof ts, "%Y-%m-%d %H:%M:%S", &tmv);
    fprintf(stderr, "%s [%s] ", ts, level);
    if (a) fprintf(stderr, fmt, a); else fprintf(stderr, fmt, b);
    fputc('\n', stderr);
}
#define DIE(fmt, a) do { logmsg("FATAL", fmt, a, 0); exit(1); } while (0)

/* ---- growing byte buffer */
typedef struct { char* p; size_t n, cap; } buf_t;
static void buf_need(buf_t* b, size_t extra)
{
    if (b->n + extra <= b->cap) return;
    size_t c = b->cap ? b->cap : 1 << 16;
    while (c < b->n + extra) c += c / 2 + 4096;
    b->p = (char*)realloc(b->p, c);
    if (!b->p) { fprintf(stderr, "fsa_rd_tools: out of memory\n"); exit(1); }
    b->cap = c;
}
static void buf_add(buf_t* b, const char* s, size_t n) { buf_need(b, n + 1); memcpy(b->p + b->n, s, n); b->n += n; b->p[b->n] = 0; }

/* ---- stripped, non-empty lines of a (possibly gzipped) file: GzFileReader::GetStrippedLine */
typedef struct {
    gzFile f;
    unsigned char* io;
    int io_n, io_at, eof;
    buf_t line;                 /* the current stripped line */
    int have;                   /* a line was read ahead (QueryStrippedLine) */
} lines_t;
enum { kIo = 4 << 20 };

static int lines_open(lines_t* L, const char* path)
{
    memset(L, 0, sizeof *L);
    L->f = gzopen(path, "rb");
    if (!L->f) return -1;
    gzbuffer(L->f, 1 << 20);
    L->io = (unsigned char*)malloc(kIo);
    return L->io ? 0 : -1;
}
static void lines_close(lines_t* L) { if (L->f) gzclose(L->f); free(L->io); free(L->line.p); }

/* one raw line (without its '\n') into L->line; 0 at the end of the file with nothing read */
static int raw_line(lines_t* L)
{
    L->line.n = 0;
    int got = 0;
    for (;;) {
        if (L->io_at == L->io_n) {
            if (L->eof) break;
            L->io_n = gzread(L->f, L->io, kIo);
            L->io_at = 0;
            if (L->io_n <= 0) { L->io_n = 0; L->eof = 1; break; }
        }
        unsigned char* s = L->io + L->io_at;
        unsigned char* nl = (unsigned char*)memchr(s, '\n', (size_t)(L->io_n - L->io_at));
        size_t n = nl ? (size_t)(nl - s) : (size_t)(L->io_n - L->io_at);
        buf_add(&L->line, (const char*)s, n);
        got = 1;
        L->io_at += (int)n + (nl ? 1 : 0);
        if (nl) break;
    }
    return got;
}
/* the next non-empty stripped line; "" (n = 0) at the end of the file */
static void next_stripped(lines_t* L)
{
    for (;;) {
        if (!raw_line(L)) { L->line.n = 0; if (L->line.p) L->line.p[0] = 0; return; }
        char* p = L->line.p;
        size_t s = 0, e = L->line.n;
        while (s < e && isspace((unsigned char)p[s])) ++s;
        while (e > s && isspace((unsigned char)p[e - 1])) --e;
        if (e > s) {
            if (s) memmove(p, p + s, e - s);
            L->line.n = e - s;
            p[L->line.n] = 0;
            return;
        }
    }
}
static const char* query_line(lines_t* L) { if (!L->have) { next_stripped(L); L->have = 1; } return L->line.n ? L->line.p : ""; }
static void consume_line(lines_t* L) { L->have = 0; }
static int at_end(const lines_t* L) { return L->eof && L->io_at == L->io_n && !(L->have && L->line.n); }

/* ---- records */
typedef struct { buf_t head, sub, seq; } rec_t;

/* FastaReader::GetHead / FastqReader::GetHead on a stripped line starting with `mark` */
static int parse_head(const char* line, size_t n, char mark, rec_t* r)
{
    if (n == 0 || line[0] != mark) return 0;
    size_t s = 1;
    while (s < n && isspace((unsigned char)line[s])) ++s;
    size_t e = s + 1 < n ? s + 1 : n;
    while (e < n && !isspace((unsigned char)line[e])) ++e;
    if (e <= s) return 0;
    r->head.n = 0; buf_add(&r->head, line + s, e - s);
    while (e < n && isspace((unsigned char)line[e])) ++e;
    r->sub.n = 0; buf_add(&r->sub, line + e, n - e);
    return 1;
}

typedef void (*rec_fn)(const rec_t* r, void* arg);

static void each_fasta(const char* path, rec_fn fn, void* arg)
{
    lines_t L;
    if (lines_open(&L, path)) DIE("Failed to open file: %s", path);
    rec_t r;
    memset(&r, 0, sizeof r);
    for (;;) {
        query_line(&L);
        const size_t n = L.line.n;
        consume_line(&L);
        if (!parse_head(L.line.p ? L.line.p : "", n, '>', &r)) break;
        r.seq.n = 0; buf_add(&r.seq, "", 0);
        for (;;) {
            const char* s = query_line(&L);
            if (L.line.n == 0 || s[0] == '>') break;
            buf_add(&r.seq, s, L.line.n);
            consume_line(&L);
        }
        fn(&r, arg);
    }
    if (!at_end(&L)) { logmsg("WARNING", "No all reads in file are loaded: %s", path, 0); }
    lines_close(&L);
    free(r.head.p); free(r.sub.p); free(r.seq.p);
}

static void each_fastq(const char* path, rec_fn fn, void* arg)
{
    lines_t L;
    if (lines_open(&L, path)) DIE("Failed to open file: %s", path);
    rec_t r;
    memset(&r, 0, sizeof r);
    for (;;) {
        next_stripped(&L);
        if (!parse_head(L.line.p ? L.line.p : "", L.line.n, '@', &r)) break;
        next_stripped(&L);
        r.seq.n = 0; buf_add(&r.seq, L.line.n ? L.line.p : "", L.line.n);
        next_stripped(&L);
        if (L.line.n == 0 || L.line.p[0] != '+') break;
        next_stripped(&L);                      /* quality */
        fn(&r, arg);
    }
    if (!(L.eof && L.io_at == L.io_n)) { logmsg("WARNING", "No all reads in file are loaded: %s", path, 0); }
    lines_close(&L);
    free(r.head.p); free(r.sub.p); free(r.seq.p);
}

static int ends_with(const char* s, const char* suf)
{
    const size_t n = strlen(s), m = strlen(suf);
    return n >= m && strcmp(s + n - m, suf) == 0;
}

/* LoadReadFile (read_store.hpp:122-136) with DetectFileType (read_store.cpp:136-158) */
static void each_read(const char* path, rec_fn fn, void* arg)
{
    if (ends_with(path, ".fasta") || ends_with(path, ".fasta.gz") || ends_with(path, ".fa") || ends_with(path, ".fa.gz")) each_fasta(path, fn, arg);
    else if (ends_with(path, ".fastq") || ends_with(path, ".fastq.gz")) each_fastq(path, fn, arg);
    else if (ends_with(path, ".fofn") || ends_with(path, ".txt")) {
        FILE* f = fopen(path, "r");
        if (!f) DIE("Failed to open file: %s", path);
        char* line = NULL;
        size_t cap = 0;
        ssize_t n;
        while ((n = getline(&line, &cap, f)) >= 0) {
            if (n && line[n - 1] == '\n') line[--n] = 0;
            int blank = 1;
            for (ssize_t i = 0; i < n; ++i) if (!isspace((unsigned char)line[i])) { blank = 0; break; }
            if (!blank) each_read(line, fn, arg);
        }
        free(line);
        fclose(f);
    } else {
        logmsg("WARNING", "Unrecognized the suffix of %s", path, 0);
        each_fasta(path, fn, arg);
    }
}

static int legal(const buf_t* seq)          /* DnaSeq::Check, sequence.hpp:122 */
{
    for (size_t i = 0; i < seq->n; ++i) {
        switch (seq->p[i]) {
            case 'A': case 'C': case 'G': case 'T': case 'a': case 'c': case 'g': case 't': break;
            default: return 0;
        }
    }
    return 1;
}

/* ---- longest */
static const size_t kKeepBytes = (size_t)3 << 30;   /* records kept in memory for the second pass up to this size */

typedef struct {
    int discard;
    /* FindLongestXHeap: a min-heap of the lengths kept so far */
    int* heap; size_t hn, hcap; long long accu, base;
    size_t n_len;
    /* the records of pass 1, for pass 2 */
    buf_t keep; int keeping;
    uint64_t* at; size_t n_at, cap_at;
} longest_t;

static void heap_push(longest_t* S, int v)
{
    if (S->hn == S->hcap) { S->hcap = S->hcap ? S->hcap * 2 : 1024; S->heap = (int*)realloc(S->heap, sizeof(int) * S->hcap); }
    size_t i = S->hn++;
    while (i > 0) { size_t p = (i - 1) / 2; if (S->heap[p] <= v) break; S->heap[i] = S->heap[p]; i = p; }
    S->heap[i] = v;
}
static void heap_pop(longest_t* S)
{
    const int v = S->heap[--S->hn];
    size_t i = 0;
    for (;;) {
        size_t c = 2 * i + 1;
        if (c >= S->hn) break;
        if (c + 1 < S->hn && S->heap[c + 1] < S->heap[c]) ++c;
        if (S->heap[c] >= v) break;
        S->heap[i] = S->heap[c]; i = c;
    }
    if (S->hn) S->heap[i] = v;
}

static void write_rec(FILE* fp, gzFile gz, const char* head, size_t hn, const char* sub, size_t sn, const char* seq, size_t qn)
{
    if (fp) {
        fputc('>', fp); fwrite(head, 1, hn, fp); fputc(' ', fp); fwrite(sub, 1, sn, fp); fputc('\n', fp);
        fwrite(seq, 1, qn, fp); fputc('\n', fp);
    } else {
        gzwrite(gz, ">", 1); gzwrite(gz, head, (unsigned)hn); gzwrite(gz, " ", 1); if (sn) gzwrite(gz, sub, (unsigned)sn); gzwrite(gz, "\n", 1);
        for (size_t o = 0; o < qn; o += 1u << 30) gzwrite(gz, seq + o, (unsigned)(qn - o < (1u << 30) ? qn - o : (1u << 30)));
        gzwrite(gz, "\n", 1);
    }
}

static void pass1(const rec_t* r, void* arg)
{
    longest_t* S = (longest_t*)arg;
    const int ok = !S->discard || legal(&r->seq);
    if (S->base > 0 && ok) {
        const int len = (int)r->seq.n;
        ++S->n_len;
        if (S->accu < S->base || S->heap[0] < len) { S->accu += len; heap_push(S, len); }      /* (accu = 0 < base on the first read) */
        while (S->hn && S->accu - S->heap[0] >= S->base) { S->accu -= S->heap[0]; heap_pop(S); }
    }
    if (S->keeping) {
        if (S->keep.n + r->head.n + r->sub.n + r->seq.n + 32 > kKeepBytes) { S->keeping = 0; free(S->keep.p); memset(&S->keep, 0, sizeof S->keep); return; }
        if (!ok) return;                                        /* never written */
        if (S->n_at + 4 > S->cap_at) { S->cap_at = S->cap_at ? S->cap_at * 2 : 4096; S->at = (uint64_t*)realloc(S->at, 8 * S->cap_at); }
        S->at[S->n_at++] = S->keep.n; buf_add(&S->keep, r->head.p, r->head.n);
        S->at[S->n_at++] = S->keep.n; buf_add(&S->keep, r->sub.p, r->sub.n);
        S->at[S->n_at++] = S->keep.n; buf_add(&S->keep, r->seq.p, r->seq.n);
    }
}

typedef struct { FILE* fp; gzFile gz; int min_length, discard; } pass2_t;
static void pass2(const rec_t* r, void* arg)
{
    pass2_t* W = (pass2_t*)arg;
    if ((int)r->seq.n >= W->min_length && (!W->discard || legal(&r->seq)))
        write_rec(W->fp, W->gz, r->head.p, r->head.n, r->sub.p, r->sub.n, r->seq.p, r->seq.n);
}

static int cmd_longest(const char* ifname, const char* ofname, long long base_size, int min_length, int discard)
{
    longest_t S;
    memset(&S, 0, sizeof S);
    S.discard = discard; S.base = base_size; S.keeping = 1;
    each_read(ifname, pass1, &S);
    if (base_size > 0) {
        logmsg("INFO", "length size = %lld", NULL, (long long)S.n_len);
        if (S.hn == 0) { fprintf(stderr, "fsa_rd_tools: no read in %s (the reference's tool stops on an assertion here, utility.cpp:47)\n", ifname); return 1; }
        if (S.heap[0] > min_length) min_length = S.heap[0];
    }
    logmsg("INFO", "min_length = %lld", NULL, (long long)min_length);
    pass2_t W;
    memset(&W, 0, sizeof W);
    W.min_length = min_length; W.discard = discard;
    if (ends_with(ofname, ".gz")) { W.gz = gzopen(ofname, "wb"); if (!W.gz) DIE("Failed to open file: %s", ofname); }
    else { W.fp = fopen(ofname, "wb"); if (!W.fp) DIE("Failed to open file: %s", ofname); setvbuf(W.fp, NULL, _IOFBF, 1 << 22); }
    if (S.keeping) {
        S.at = (uint64_t*)realloc(S.at, 8 * (S.n_at + 1));
        S.at[S.n_at] = S.keep.n;
        for (size_t i = 0; i + 3 <= S.n_at; i += 3) {
            const size_t hn = (size_t)(S.at[i + 1] - S.at[i]), sn = (size_t)(S.at[i + 2] - S.at[i + 1]), qn = (size_t)(S.at[i + 3] - S.at[i + 2]);
            if ((int)qn >= min_length) write_rec(W.fp, W.gz, S.keep.p + S.at[i], hn, S.keep.p + S.at[i + 1], sn, S.keep.p + S.at[i + 2], qn);
        }
    } else {
        each_read(ifname, pass2, &W);
    }
    if (W.fp && fclose(W.fp)) DIE("Failed to write file: %s", ofname);
    if (W.gz && gzclose(W.gz) != Z_OK) DIE("Failed to write file: %s", ofname);
    free(S.heap); free(S.keep.p); free(S.at);
    return 0;
}

/* ---- n50 */
typedef struct { int* v; size_t n, cap; } lens_t;
static void add_len(const rec_t* r, void* arg)
{
    lens_t* L = (lens_t*)arg;
    if (L->n == L->cap) { L->cap = L->cap ? L->cap * 2 : 4096; L->v = (int*)realloc(L->v, sizeof(int) * L->cap); }
    L->v[L->n++] = (int)r->seq.n;
}
static int cmp_desc(const void* a, const void* b) { const int x = *(const int*)a, y = *(const int*)b; return x < y ? 1 : x > y ? -1 : 0; }

static int cmd_n50(const char* ifname, long long genome_size)
{
    logmsg("INFO", "Load read file %s", ifname, 0);
    lens_t L;
    memset(&L, 0, sizeof L);
    each_read(ifname, add_len, &L);
    if (L.n == 0) { fprintf(stderr, "fsa_rd_tools: no read in %s\n", ifname); return 1; }
    qsort(L.v, L.n, sizeof(int), cmp_desc);
    long long total = 0;
    for (size_t i = 0; i < L.n; ++i) total += L.v[i];
    if (genome_size != 0) total = genome_size;
    printf("Count: %zu\nTotal: %lld\nMax: %d\nMin: %d\n", L.n, total, L.v[0], L.v[L.n - 1]);
    static const int ns[3] = {25, 50, 75};
    long long accu = 0;
    size_t ins = 0;
    for (size_t i = 0; i < L.n && ins < 3; ++i) {
        accu += L.v[i];
        for (; ins < 3 && (double)accu > (double)((long long)ns[ins] * 1.0 / 100 * (double)total); ++ins)
            printf("N%d: %d\nL%d: %zu\n", ns[ins], L.v[i], ns[ins], i + 1);
    }
    free(L.v);
    return 0;
}

int main(int argc, char** argv)
{
    /* ArgumentParser (argument_parser.cpp:57-120): getopt_long over --name value / --name=value, one positional command */
    static const struct option opts[] = {
        {"ifname", required_argument, 0, 'i'}, {"size", required_argument, 0, 1}, {"block_size", required_argument, 0, 2},
        {"base_size", required_argument, 0, 'b'}, {"ofname", required_argument, 0, 'o'}, {"thread_size", required_argument, 0, 3},
        {"names", required_argument, 0, 4}, {"min_length", required_argument, 0, 'm'}, {"genome_size", required_argument, 0, 'g'},
        {"id2name", required_argument, 0, 'n'}, {"discard_illegal_read", no_argument, 0, 'd'}, {0, 0, 0, 0}};
    const char *ifname = "", *ofname = "", *id2name = "";
    long long base_size = 0, genome_size = 0;
    int min_length = 0, discard = 0, c;
    opterr = 0;
    while ((c = getopt_long(argc, argv, "", opts, NULL)) != -1) {
        switch (c) {
            case 'i': ifname = optarg; break;
            case 'o': ofname = optarg; break;
            case 'b': base_size = atoll(optarg); break;
            case 'm': min_length = atoi(optarg); break;
            case 'g': genome_size = atoll(optarg); break;
            case 'n': id2name = optarg; break;
            case 'd': discard = 1; break;
            default: break;
        }
    }
    if (argc - optind != 1) {
        fprintf(stderr, "tools about reads\n\nUsage: fsa_rd_tools [options] cmd \n  cmd\n      command (longest, n50)\n"
                        "  --ifname=<file>  --ofname=<file>  --base_size=<n>  --min_length=<n>  --genome_size=<n>  --discard_illegal_read\n");
        return 1;
    }
    const char* cmd = argv[optind];
    logmsg("INFO", "%s", "Start", 0);
    int rc;
    if (strcmp(cmd, "longest") == 0) {
        if (id2name[0]) { fprintf(stderr, "fsa_rd_tools: --id2name is not supported by this drop-in (necat.pl does not use it)\n"); return 1; }
        if (!ifname[0] || !ofname[0]) { fprintf(stderr, "fsa_rd_tools longest: --ifname and --ofname are needed\n"); return 1; }
        rc = cmd_longest(ifname, ofname, base_size, min_length, discard);
    } else if (strcmp(cmd, "n50") == 0) {
        rc = cmd_n50(ifname, genome_size);
    } else {
        logmsg("ERROR", "Not recognizing the command: %s (this drop-in carries longest and n50)", cmd, 0);
        rc = 1;
    }
    logmsg("INFO", "%s", "End", 0);
    return rc;
}
