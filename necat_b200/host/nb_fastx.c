/* nb_fastx.c -- see nb_fastx.h */
#include "nb_fastx.h"

#include <ctype.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#define FX_BLOCK   (4u << 20)
#define FX_NBLOCKS 3

typedef struct { uint8_t* data; int len; } fx_block;

struct nbh_fastx {
    gzFile gz;
    /* ring of inflated blocks: the reader thread fills, the parser drains */
    fx_block blk[FX_NBLOCKS];
    int head, tail, count;              /* tail = next block the parser takes */
    int eof_seen, read_err;
    pthread_mutex_t mu;
    pthread_cond_t  not_empty, not_full;
    pthread_t th;
    int th_started;
    /* parser state */
    const uint8_t* buf;
    int begin, end;
    int have_block;                     /* the parser holds blk[tail] */
    int at_eof;
    int last_char, done;
    char*    name; size_t name_l, name_m;
    uint8_t* seq;  size_t seq_l,  seq_m;
    uint8_t* qual; size_t qual_l, qual_m;
};

static void* fx_reader(void* arg)
{
    nbh_fastx* fx = (nbh_fastx*)arg;
    for (;;) {
        pthread_mutex_lock(&fx->mu);
        while (fx->count == FX_NBLOCKS && !fx->done) pthread_cond_wait(&fx->not_full, &fx->mu);
        if (fx->done) { pthread_mutex_unlock(&fx->mu); return NULL; }
        fx_block* b = &fx->blk[fx->head];
        pthread_mutex_unlock(&fx->mu);
        const int n = gzread(fx->gz, b->data, FX_BLOCK);
        pthread_mutex_lock(&fx->mu);
        if (n <= 0) {
            fx->eof_seen = 1;
            int zerr = Z_OK;                   /* a gzip stream cut short reads as end of file with Z_BUF_ERROR set */
            gzerror(fx->gz, &zerr);
            if (n < 0 || (zerr != Z_OK && zerr != Z_STREAM_END)) fx->read_err = 1;
            pthread_cond_signal(&fx->not_empty);
            pthread_mutex_unlock(&fx->mu);
            return NULL;
        }
        b->len = n;
        fx->head = (fx->head + 1) % FX_NBLOCKS;
        fx->count++;
        pthread_cond_signal(&fx->not_empty);
        pthread_mutex_unlock(&fx->mu);
    }
}

nbh_fastx* nbh_fastx_open(const char* path)
{
    gzFile gz = gzopen(path, "r");
    if (!gz) return NULL;
    gzbuffer(gz, 1u << 20);
    nbh_fastx* fx = (nbh_fastx*)calloc(1, sizeof *fx);
    if (!fx) { gzclose(gz); return NULL; }
    fx->gz = gz;
    for (int i = 0; i < FX_NBLOCKS; ++i) {
        fx->blk[i].data = (uint8_t*)malloc(FX_BLOCK);
        if (!fx->blk[i].data) { nbh_fastx_close(fx); return NULL; }
    }
    fx->seq_m = 1 << 16; fx->seq = (uint8_t*)malloc(fx->seq_m);
    fx->qual_m = 1 << 16; fx->qual = (uint8_t*)malloc(fx->qual_m);
    fx->name_m = 256; fx->name = (char*)malloc(fx->name_m);
    if (!fx->seq || !fx->qual || !fx->name) { nbh_fastx_close(fx); return NULL; }
    fx->name[0] = 0;
    pthread_mutex_init(&fx->mu, NULL);
    pthread_cond_init(&fx->not_empty, NULL);
    pthread_cond_init(&fx->not_full, NULL);
    if (pthread_create(&fx->th, NULL, fx_reader, fx)) { nbh_fastx_close(fx); return NULL; }
    fx->th_started = 1;
    return fx;
}

void nbh_fastx_close(nbh_fastx* fx)
{
    if (!fx) return;
    if (fx->th_started) {
        pthread_mutex_lock(&fx->mu);
        fx->done = 1;
        pthread_cond_signal(&fx->not_full);
        pthread_mutex_unlock(&fx->mu);
        pthread_join(fx->th, NULL);
        pthread_mutex_destroy(&fx->mu);
        pthread_cond_destroy(&fx->not_empty);
        pthread_cond_destroy(&fx->not_full);
    }
    if (fx->gz) gzclose(fx->gz);
    for (int i = 0; i < FX_NBLOCKS; ++i) free(fx->blk[i].data);
    free(fx->seq); free(fx->qual); free(fx->name);
    free(fx);
}

/* make buf[begin, end) non-empty; 0 at end of input */
static int fx_fill(nbh_fastx* fx)
{
    if (fx->begin < fx->end) return 1;
    if (fx->at_eof) return 0;
    pthread_mutex_lock(&fx->mu);
    if (fx->have_block) {                     /* hand the drained block back */
        fx->tail = (fx->tail + 1) % FX_NBLOCKS;
        fx->count--;
        fx->have_block = 0;
        pthread_cond_signal(&fx->not_full);
    }
    while (fx->count == 0 && !fx->eof_seen) pthread_cond_wait(&fx->not_empty, &fx->mu);
    if (fx->count == 0) {
        pthread_mutex_unlock(&fx->mu);
        fx->at_eof = 1;
        return 0;
    }
    fx->buf = fx->blk[fx->tail].data;
    fx->begin = 0;
    fx->end = fx->blk[fx->tail].len;
    fx->have_block = 1;
    pthread_mutex_unlock(&fx->mu);
    return 1;
}

static inline int fx_getc(nbh_fastx* fx)
{
    if (fx->begin >= fx->end && !fx_fill(fx)) return -1;
    return fx->buf[fx->begin++];
}

static int fx_reserve(uint8_t** p, size_t* m, size_t need)
{
    if (need <= *m) return 0;
    size_t nm = *m ? *m : 256;
    while (nm < need) nm *= 2;
    uint8_t* q = (uint8_t*)realloc(*p, nm);
    if (!q) return -1;
    *p = q; *m = nm;
    return 0;
}

/* append the rest of the current line (its '\n' is consumed, not stored) to (*p)[*l..]; returns 1 if any input was
 * available (even an empty rest), 0 if the input had already ended, -1 out of memory.  A '\r' that now ends a string
 * longer than one character is dropped -- only when input was available, as in the parser this follows. */
static int fx_append_line(nbh_fastx* fx, uint8_t** p, size_t* l, size_t* m)
{
    int got = 0;
    for (;;) {
        if (!fx_fill(fx)) break;
        got = 1;
        const uint8_t* s = fx->buf + fx->begin;
        const size_t avail = (size_t)(fx->end - fx->begin);
        const uint8_t* nl = (const uint8_t*)memchr(s, '\n', avail);
        const size_t n = nl ? (size_t)(nl - s) : avail;
        if (fx_reserve(p, m, *l + n + 1)) return -1;
        memcpy(*p + *l, s, n);
        *l += n;
        fx->begin += (int)n + (nl ? 1 : 0);
        if (nl) break;
    }
    if (got && *l > 1 && (*p)[*l - 1] == '\r') --*l;
    return got;
}

static void fx_skip_line(nbh_fastx* fx)
{
    for (;;) {
        if (!fx_fill(fx)) return;
        const uint8_t* s = fx->buf + fx->begin;
        const size_t avail = (size_t)(fx->end - fx->begin);
        const uint8_t* nl = (const uint8_t*)memchr(s, '\n', avail);
        if (nl) { fx->begin += (int)(nl - s) + 1; return; }
        fx->begin = fx->end;
    }
}

int nbh_fastx_next(nbh_fastx* fx)
{
    if (fx->done) return fx->read_err ? -1 : 0;
    int c;
    if (fx->last_char == 0) {                                   /* look for the next header character */
        for (;;) {
            c = fx_getc(fx);
            if (c < 0) goto finished;
            if (c == '>' || c == '@') break;
        }
        fx->last_char = c;
    }
    fx->name_l = fx->seq_l = fx->qual_l = 0;
    {   /* name: up to the first white space */
        int got = 0, delim = 0;
        for (;;) {
            if (!fx_fill(fx)) break;
            got = 1;
            int i = fx->begin;
            while (i < fx->end && !isspace(fx->buf[i])) ++i;
            const size_t n = (size_t)(i - fx->begin);
            if (fx->name_l + n + 1 > fx->name_m) {
                size_t nm = fx->name_m;
                while (nm < fx->name_l + n + 1) nm *= 2;
                char* q = (char*)realloc(fx->name, nm);
                if (!q) { fx->read_err = 1; goto finished; }
                fx->name = q; fx->name_m = nm;
            }
            memcpy(fx->name + fx->name_l, fx->buf + fx->begin, n);
            fx->name_l += n;
            if (i < fx->end) { delim = fx->buf[i]; fx->begin = i + 1; break; }
            fx->begin = fx->end;
        }
        if (!got) goto finished;
        fx->name[fx->name_l] = 0;
        if (delim != '\n') fx_skip_line(fx);                    /* the comment */
    }
    for (;;) {                                                  /* sequence lines */
        c = fx_getc(fx);
        if (c < 0 || c == '>' || c == '+' || c == '@') break;
        if (c == '\n') continue;
        if (fx_reserve(&fx->seq, &fx->seq_m, fx->seq_l + 2)) { fx->read_err = 1; goto finished; }
        fx->seq[fx->seq_l++] = (uint8_t)c;
        if (fx_append_line(fx, &fx->seq, &fx->seq_l, &fx->seq_m) < 0) { fx->read_err = 1; goto finished; }
    }
    if (c == '>' || c == '@') fx->last_char = c;
    if (c != '+') return 1;                                     /* FASTA */
    while ((c = fx_getc(fx)) >= 0 && c != '\n') { }             /* rest of the '+' line */
    if (c < 0) goto finished;                                   /* no quality string: the record is dropped */
    for (;;) {                                                  /* quality, by length */
        const int got = fx_append_line(fx, &fx->qual, &fx->qual_l, &fx->qual_m);
        if (got < 0) { fx->read_err = 1; goto finished; }
        if (!got || fx->qual_l >= fx->seq_l) break;
    }
    fx->last_char = 0;
    if (fx->qual_l != fx->seq_l) goto finished;                 /* truncated: the record is dropped and reading stops */
    return 1;
finished:
    pthread_mutex_lock(&fx->mu);
    fx->done = 1;
    pthread_cond_signal(&fx->not_full);
    const int err = fx->read_err;
    pthread_mutex_unlock(&fx->mu);
    return err ? -1 : 0;
}

const char* nbh_fastx_name(const nbh_fastx* fx, size_t* len) { if (len) *len = fx->name_l; return fx->name; }
const uint8_t* nbh_fastx_seq(const nbh_fastx* fx, size_t* len) { if (len) *len = fx->seq_l; return fx->seq; }

/* ------------------------------------------------------------------------------------------------------------------ */

#define N4 4
const uint8_t nbh_nt4[256] = {
    N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4,
    N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, 5,  N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4,
    N4, 0,  N4, 1,  N4, N4, N4, 2,  N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, 3,  N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4,
    N4, 0,  N4, 1,  N4, N4, N4, 2,  N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, 3,  N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4,
    N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4,
    N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4,
    N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4,
    N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4, N4,
};

int nbh_pac_append(nbh_pac* p, const uint8_t* s, size_t n)
{
    const uint64_t need = ((p->size + n + 3) >> 2) + 8;
    if (need > p->cap_bytes) {
        uint64_t nc = p->cap_bytes ? p->cap_bytes : (1u << 20);
        while (nc < need) nc *= 2;
        uint8_t* q = (uint8_t*)realloc(p->pac, nc);
        if (!q) return -1;
        memset(q + p->cap_bytes, 0, nc - p->cap_bytes);
        p->pac = q; p->cap_bytes = nc;
    }
    uint64_t l = p->size;
    size_t i = 0;
    /* head: up to the next byte boundary; body: four characters per byte; tail.  Codes are OR-ed in unmasked and the
     * byte store truncates, which is what the reference's macro does on its u8 array. */
    for (; i < n && (l & 3); ++i, ++l) p->pac[l >> 2] |= (uint8_t)(nbh_nt4[s[i]] << (((~l) & 3) << 1));
    uint8_t* out = p->pac + (l >> 2);
    for (; i + 4 <= n; i += 4, l += 4)
        *out++ |= (uint8_t)((nbh_nt4[s[i]] << 6) | (nbh_nt4[s[i + 1]] << 4) | (nbh_nt4[s[i + 2]] << 2) | nbh_nt4[s[i + 3]]);
    for (; i < n; ++i, ++l) p->pac[l >> 2] |= (uint8_t)(nbh_nt4[s[i]] << (((~l) & 3) << 1));
    p->size = l;
    return 0;
}

void nbh_pac_clear(nbh_pac* p)
{
    if (p->pac && p->size) memset(p->pac, 0, (p->size + 3) >> 2);
    p->size = 0;
}

void nbh_pac_free(nbh_pac* p)
{
    free(p->pac);
    memset(p, 0, sizeof *p);
}

static char decode_lut[256][4];
static void decode_lut_init(void)
{
    for (int b = 0; b < 256; ++b)
        for (int k = 0; k < 4; ++k) decode_lut[b][k] = "ACGT"[(b >> ((3 - k) << 1)) & 3];
}

void nbh_pac_decode(const nbh_pac* p, uint64_t from, uint64_t n, char* out)
{
    static pthread_once_t once = PTHREAD_ONCE_INIT;
    pthread_once(&once, decode_lut_init);
    const char (*lut)[4] = decode_lut;
    uint64_t l = from, e = from + n;
    for (; l < e && (l & 3); ++l) *out++ = "ACGT"[(p->pac[l >> 2] >> (((~l) & 3) << 1)) & 3];
    for (; l + 4 <= e; l += 4, out += 4) memcpy(out, lut[p->pac[l >> 2]], 4);
    for (; l < e; ++l) *out++ = "ACGT"[(p->pac[l >> 2] >> (((~l) & 3) << 1)) & 3];
}
