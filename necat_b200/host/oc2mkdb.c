/* oc2mkdb -- FASTA / FASTQ (plain or gzip) -> PackedDB volumes.  Drop-in for the reference's makedb (makedb/main.c:10-126):
 *
 *   oc2mkdb wrk_dir file-list [file-list ...]
 *
 * writes wrk_dir/vol<N>, wrk_dir/volume_names.txt ("<path>\t<first id>\t<count>") and wrk_dir/reads_info.txt
 * ("<volumes>\t<reads>"), byte for byte what the reference writes (apart from the 4 padding bytes of each 32-byte
 * sequence record, which are uninitialised stack bytes there and zero here).  A volume is closed after the read that
 * takes it to 2 000 000 000 bases or more (makedb/main.c:8,29-38).  No GPU involved: this is the ingest side of the hot
 * path; inflation runs on its own thread, packing handles four bases per byte.
 */
#include "nb_fastx.h"
#include "nb_host.h"

#include <errno.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

static const char kMagic[] = "ontcns_pac_header_hofuwhogfuewo";
static uint64_t g_vol_size = 2000000000ull;

typedef struct { uint64_t offset, size, hdr_offset; int32_t platform, pad; } seq_rec;

typedef struct {
    nbh_pac  pac;
    seq_rec* recs;  size_t n, m;
    char*    hdr;   size_t hdr_l, hdr_m;
} volume;

static void vol_add(volume* v, const char* name, size_t name_l, const uint8_t* seq, size_t seq_l)
{
    if (v->n == v->m) {
        v->m = v->m ? v->m * 2 : 4096;
        v->recs = (seq_rec*)realloc(v->recs, v->m * sizeof(seq_rec));
        if (!v->recs) NBH_ERROR("out of memory");
    }
    if (v->hdr_l + name_l + 1 > v->hdr_m) {
        while (v->hdr_l + name_l + 1 > v->hdr_m) v->hdr_m = v->hdr_m ? v->hdr_m * 2 : (1 << 16);
        v->hdr = (char*)realloc(v->hdr, v->hdr_m);
        if (!v->hdr) NBH_ERROR("out of memory");
    }
    seq_rec r;
    memset(&r, 0, sizeof r);
    r.offset = v->pac.size; r.size = seq_l; r.hdr_offset = v->hdr_l; r.platform = 0;   /* TECH_PACBIO, makedb/main.c:25 */
    v->recs[v->n++] = r;
    memcpy(v->hdr + v->hdr_l, name, name_l);
    v->hdr[v->hdr_l + name_l] = 0;
    v->hdr_l += name_l + 1;
    if (nbh_pac_append(&v->pac, seq, seq_l)) NBH_ERROR("out of memory");
}

static void vol_dump(volume* v, const char* path)           /* packed_db.c:291-314 */
{
    FILE* f = fopen(path, "wb");
    if (!f) NBH_ERROR("cannot open '%s' for writing: %s", path, strerror(errno));
    const uint64_t ns = v->n, db = v->pac.size, hs = v->hdr_l, nb = (db + 3) >> 2;
    int ok = fwrite(kMagic, 1, sizeof kMagic - 1, f) == sizeof kMagic - 1;
    ok &= fwrite(&ns, 8, 1, f) == 1 && fwrite(&db, 8, 1, f) == 1;
    ok &= fwrite(v->recs, sizeof(seq_rec), ns, f) == ns;
    ok &= fwrite(&hs, 8, 1, f) == 1;
    if (hs) ok &= fwrite(v->hdr, 1, hs, f) == hs;
    if (nb) ok &= fwrite(v->pac.pac, 1, nb, f) == nb;
    if (fclose(f) || !ok) NBH_ERROR("writing '%s' failed", path);
}

static void vol_clear(volume* v)
{
    v->n = 0; v->hdr_l = 0;
    nbh_pac_clear(&v->pac);
}

static void volume_path(char* out, size_t cap, const char* wrk_dir, const char* leaf, int vid)
{
    const size_t wl = strlen(wrk_dir);
    const char* sep = (wl && wrk_dir[wl - 1] == '/') ? "" : "/";
    if (vid >= 0) snprintf(out, cap, "%s%s%s%d", wrk_dir, sep, leaf, vid);
    else snprintf(out, cap, "%s%s%s", wrk_dir, sep, leaf);
}

int main(int argc, char** argv)
{
    if (argc < 3) {
        fprintf(stderr, "USAGE:\n%s wrk-dir file-list [file-list]\n", argv[0]);
        return 1;
    }
    const char* wrk_dir = argv[1];
    const char* vs = getenv("NB_MKDB_VOLUME_SIZE");          /* tests only: the volume cap is a constant in the reference */
    if (vs && atoll(vs) > 0) g_vol_size = (uint64_t)atoll(vs);
    if (access(wrk_dir, F_OK) == -1 && mkdir(wrk_dir, 0755) == -1) NBH_ERROR("failed to create folder %s", wrk_dir);

    char path[4096], vname[4096];
    volume_path(path, sizeof path, wrk_dir, "volume_names.txt", -1);
    FILE* vn_out = fopen(path, "w");
    if (!vn_out) NBH_ERROR("cannot open '%s' for writing", path);
    volume v;
    memset(&v, 0, sizeof v);
    int vid = 0, num_reads = 0, read_start_id = 0;
    for (int c = 2; c < argc; ++c) {
        fprintf(stdout, "file_list: %s\n", argv[c]);
        FILE* lf = fopen(argv[c], "r");
        if (!lf) NBH_ERROR("cannot open '%s'", argv[c]);
        char line[2048];
        while (fgets(line, sizeof line, lf)) {                /* one input file per line, ontcns_aux.c:196-202 */
            size_t ll = strlen(line);
            if (ll && line[ll - 1] == '\n') line[ll - 1] = 0;
            nbh_fastx* fx = nbh_fastx_open(line);
            if (!fx) NBH_ERROR("cannot open '%s'", line);
            int file_reads = 0, rc;
            uint64_t file_bps = 0;
            while ((rc = nbh_fastx_next(fx)) == 1) {
                size_t nl, sl;
                const char* name = nbh_fastx_name(fx, &nl);
                const uint8_t* seq = nbh_fastx_seq(fx, &sl);
                vol_add(&v, name, nl, seq, sl);
                ++file_reads;
                file_bps += sl;
                if (v.pac.size >= g_vol_size) {
                    volume_path(vname, sizeof vname, wrk_dir, "vol", vid);
                    vol_dump(&v, vname);
                    fprintf(vn_out, "%s\t%d\t%lu\n", vname, read_start_id, (unsigned long)v.n);
                    read_start_id += (int)v.n;
                    vol_clear(&v);
                    ++vid;
                }
            }
            if (rc < 0) NBH_ERROR("read error in '%s' (truncated or corrupt input)", line);
            nbh_fastx_close(fx);
            NBH_LOG("pack %s: %d reads, %lu bps", line, file_reads, (unsigned long)file_bps);
            num_reads += file_reads;
        }
        fclose(lf);
    }
    if (v.pac.size) {
        volume_path(vname, sizeof vname, wrk_dir, "vol", vid);
        vol_dump(&v, vname);
        fprintf(vn_out, "%s\t%d\t%lu\n", vname, read_start_id, (unsigned long)v.n);
        ++vid;
    }
    if (fclose(vn_out)) NBH_ERROR("writing volume_names.txt failed");
    volume_path(path, sizeof path, wrk_dir, "reads_info.txt", -1);
    FILE* ri = fopen(path, "w");
    if (!ri) NBH_ERROR("cannot open '%s' for writing", path);
    fprintf(ri, "%d\t%d\n", vid, num_reads);
    if (fclose(ri)) NBH_ERROR("writing reads_info.txt failed");
    nbh_pac_free(&v.pac);
    free(v.recs); free(v.hdr);
    return 0;
}
