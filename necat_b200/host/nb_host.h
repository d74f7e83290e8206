/* nb_host.h -- file-format helpers shared by the drop-in executables (plain C). */
#ifndef NB_HOST_H
#define NB_HOST_H

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>

#include "../../include/necat_b200.h"

/* PackedDB volume file (common/packed_db.c:291-348). */
typedef struct {
    uint8_t*  pac;
    uint64_t  db_size, nseq;
    uint64_t* offset;
    uint64_t* size;
    uint64_t* hdr_offset;
    char*     hdr;
    uint64_t  hdr_size;
} nbh_volume;

int  nbh_volume_load(const char* path, nbh_volume* v);
int  nbh_volume_append(nbh_volume* dst, const nbh_volume* src);    /* pdb_merge, packed_db.c:419-448 */
void nbh_volume_free(nbh_volume* v);

/* wrk_dir/reads_info.txt + volume_names.txt (common/makedb_aux.c:78-118). */
typedef struct {
    int    num_volumes, num_reads;
    char** names;
    int*   start_id;
    int*   count;
} nbh_volumes_info;
int  nbh_volumes_info_load(const char* wrk_dir, nbh_volumes_info* vi);
void nbh_volumes_info_free(nbh_volumes_info* vi);

/* order[0..n) = indices of recs in ascending read id, records of one id in their given order (stable).  oc2cns writes its
 * FASTA in this order so that the reorder step after it (reorder_cns_reads/main.c, necat.pl:351) finds sorted input.
 * Returns -1 when out of memory. */
int  nbh_order_by_id(const nb_cns_record* recs, uint64_t n, uint64_t* order);

/* "[time] LEVEL: msg (file, line)" banners of common/ontcns_aux.c:29-46 */
void nbh_log(const char* level, const char* file, int line, const char* fmt, ...);
#define NBH_LOG(...)   nbh_log("INFO", __FILE__, __LINE__, __VA_ARGS__)
#define NBH_ERROR(...) do { nbh_log("ERROR", __FILE__, __LINE__, __VA_ARGS__); exit(1); } while (0)
double nbh_now(void);
/* NB_TRACE=1: wall-clock of the phases of an executable on stderr (start-up, I/O and tear-down dominate small jobs) */
/* The GPU of this process: NECAT_B200_DEVICE (default 0).  Where the user has not restricted the visible devices the
 * process restricts them to that one before the first CUDA call: initialising the driver for every GPU of an 8-GPU
 * box costs each short-lived process several hundred ms it has no use for. */
static inline int nbh_pick_device(void)
{
    static int picked = -1;                 /* decided once per process (the resident helper runs many jobs) */
    if (picked >= 0) return picked;
    const char* devs = getenv("NECAT_B200_DEVICE");
    int dev = devs ? atoi(devs) : 0;
    if (!getenv("CUDA_VISIBLE_DEVICES") && dev >= 0) {
        char b[16];
        snprintf(b, sizeof b, "%d", dev);
        setenv("CUDA_VISIBLE_DEVICES", b, 1);
        dev = 0;
    }
    picked = dev;
    return dev;
}
/* End of a drop-in executable whose outputs are closed: the process image goes away without the device blocks, pinned
 * buffers and streams being handed back one by one (with 25 GB of cached device blocks that took 0.8 s); the
 * driver reclaims them with the process.  NB_FULL_TEARDOWN=1 keeps the orderly tear-down (leak checks). */
static inline int nbh_fast_exit(void) { return getenv("NB_FULL_TEARDOWN") == NULL; }
#define NBH_PHASE(name) do { if (getenv("NB_TRACE")) { const double now__ = nbh_now(); fprintf(stderr, "[phase] %-28s %8.3f s\n", name, now__ - nbh_phase_t0); nbh_phase_t0 = now__; } } while (0)
extern double nbh_phase_t0;

#endif
