/* nb_fastx.h -- streaming FASTA / FASTQ reader (plain or gzip) for the ingest-side hosts (oc2mkdb, oc2ReorderCnsReads).
 *
 * Record boundaries follow the parser the reference reads its inputs with (klib kseq, src/klib/kseq.h:178-218), because
 * they decide which bytes become bases: headers are recognised by '>' or '@', the name ends at the first white space, a
 * sequence runs over any number of lines up to a line that starts with '>', '@' or '+', a '\r' that ends a line is
 * dropped once the sequence is longer than one character, a FASTQ quality string is read by LENGTH (so quality lines may
 * start with '@'), and a record whose quality length differs from its sequence length ends the file.
 *
 * Inflation runs on its own thread, two blocks ahead of the parser.
 */
#ifndef NB_FASTX_H
#define NB_FASTX_H

#include <stddef.h>
#include <stdint.h>

typedef struct nbh_fastx nbh_fastx;

nbh_fastx* nbh_fastx_open(const char* path);            /* NULL if the file cannot be opened */
/* next record: 1 = a record, 0 = end of input (or a truncated FASTQ record, like the reference), -1 = read error */
int nbh_fastx_next(nbh_fastx* fx);
const char*    nbh_fastx_name(const nbh_fastx* fx, size_t* len);     /* first token of the header, NUL terminated */
const uint8_t* nbh_fastx_seq(const nbh_fastx* fx, size_t* len);      /* raw sequence characters */
void nbh_fastx_close(nbh_fastx* fx);

/* character -> code of common/nst_nt4_table.c:3 (ACGT any case -> 0..3, '-' -> 5, everything else -> 4) */
extern const uint8_t nbh_nt4[256];

/* Growable 2-bit store with the reference's packing rule (common/ontcns_aux.h:118: the code is OR-ed, unmasked, into its
 * byte, first base in the top two bits).  Codes > 3 therefore leak into the neighbouring base of the same byte exactly as
 * they do there; inputs are expected to be ACGT only (necat.pl:122 discards other reads upstream). */
typedef struct {
    uint8_t* pac;
    uint64_t size;        /* bases */
    uint64_t cap_bytes;
} nbh_pac;
int  nbh_pac_append(nbh_pac* p, const uint8_t* chars, size_t n);     /* 0 ok, -1 out of memory */
void nbh_pac_clear(nbh_pac* p);                                      /* zero the used bytes, size = 0 */
void nbh_pac_free(nbh_pac* p);
/* bases [from, from + n) as ASCII "ACGT" */
void nbh_pac_decode(const nbh_pac* p, uint64_t from, uint64_t n, char* out);

#endif
