// nb_core.cu -- context, error handling, volume upload (C-ABI part 1).
#include <stdarg.h>

#include <mutex>

#include "nb_common.cuh"

char g_nb_err[512] = "";

struct NbDevBlock { void* p; size_t bytes; bool used; };
struct NbDevCache { std::vector<NbDevBlock> blocks; std::mutex mu; };

int nb_dev_alloc(nb_ctx* ctx, void** p, size_t bytes)
{
    if (!ctx->dev_cache) ctx->dev_cache = new NbDevCache();
    NbDevCache* C = (NbDevCache*)ctx->dev_cache;
    std::lock_guard<std::mutex> lock(C->mu);
    bytes = (bytes + 511) & ~(size_t)511;
    int best = -1;
    for (size_t i = 0; i < C->blocks.size(); ++i) {
        const NbDevBlock& b = C->blocks[i];
        if (b.used || b.bytes < bytes || b.bytes > bytes * 2 + (8u << 20)) continue;
        if (best < 0 || b.bytes < C->blocks[(size_t)best].bytes) best = (int)i;
    }
    if (best >= 0) { C->blocks[(size_t)best].used = true; *p = C->blocks[(size_t)best].p; return NB_OK; }
    void* q = NULL;
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        // give the cached, unused blocks back and retry once
        for (size_t i = 0; i < C->blocks.size();) {
            if (!C->blocks[i].used) { cudaFree(C->blocks[i].p); C->blocks.erase(C->blocks.begin() + (long)i); }
            else ++i;
        }
        e = cudaMalloc(&q, bytes);
        if (e != cudaSuccess) return nb_fail(ctx, NB_ENOMEM, "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    NbDevBlock b; b.p = q; b.bytes = bytes; b.used = true;
    C->blocks.push_back(b);
    *p = q;
    return NB_OK;
}

void nb_dev_release(nb_ctx* ctx, void* p)
{
    if (!p || !ctx || !ctx->dev_cache) return;
    NbDevCache* C = (NbDevCache*)ctx->dev_cache;
    std::lock_guard<std::mutex> lock(C->mu);
    for (auto& b : C->blocks) if (b.p == p) { b.used = false; return; }
    cudaFree(p);
}

int nb_pinned_reserve(nb_ctx* ctx, size_t bytes)
{
    if (bytes <= ctx->pinned_bytes) return NB_OK;
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx->pinned = NULL; ctx->pinned_bytes = 0;
    const size_t want = bytes + bytes / 4 + (1u << 20);
    NB_CUDA(ctx, cudaHostAlloc(&ctx->pinned, want, cudaHostAllocDefault));
    ctx->pinned_bytes = want;
    return NB_OK;
}

void nb_dev_cache_destroy(nb_ctx* ctx)
{
    if (!ctx || !ctx->dev_cache) return;
    NbDevCache* C = (NbDevCache*)ctx->dev_cache;
    for (auto& b : C->blocks) cudaFree(b.p);
    delete C;
    ctx->dev_cache = NULL;
}

int nb_fail(nb_ctx* ctx, int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_nb_err, sizeof g_nb_err, fmt, ap);
    va_end(ap);
    if (ctx) memcpy(ctx->err, g_nb_err, sizeof g_nb_err);
    return code;
}

extern "C" const char* nb_version(void) { return "necat_b200 0.1 (sm_100a)"; }

extern "C" const char* nb_last_error(const nb_ctx* ctx) { return ctx ? ctx->err : g_nb_err; }

static int nb_ctx_init_streams(nb_ctx* c);

extern "C" int nb_ctx_create(int device, nb_ctx** out)
{
    if (!out) return nb_fail(NULL, NB_EINVAL, "nb_ctx_create: out is NULL");
    *out = NULL;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return nb_fail(NULL, NB_ENODEVICE, "no CUDA device available (%s): libnecat_b200 has no CPU fallback",
                       cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return nb_fail(NULL, NB_EINVAL, "device %d out of range (%d present)", device, ndev);
    NB_CUDA(NULL, cudaSetDevice(device));
    cudaDeviceProp prop;
    NB_CUDA(NULL, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return nb_fail(NULL, NB_ENODEVICE, "device %d is sm_%d%d; this library only carries sm_100a code", device,
                       prop.major, prop.minor);
    nb_ctx* c = (nb_ctx*)calloc(1, sizeof(nb_ctx));
    if (!c) return nb_fail(NULL, NB_ENOMEM, "out of host memory");
    c->device = device;
    { const int rc = nb_ctx_init_streams(c); if (rc) { free(c); return rc; } }
    *out = c;
    return NB_OK;
}

static int nb_ctx_init_streams(nb_ctx* c)
{
    NB_CUDA(NULL, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    NB_CUDA(NULL, cudaEventCreate(&c->ev0));
    NB_CUDA(NULL, cudaEventCreate(&c->ev1));
    NB_CUDA(NULL, cudaEventCreate(&c->ev2));
    NB_CUDA(NULL, cudaEventCreate(&c->ev3));
    for (int i = 0; i < NB_AUX; ++i) NB_CUDA(NULL, cudaStreamCreateWithFlags(&c->aux[i], cudaStreamNonBlocking));
    for (int i = 0; i < NB_AUX + 1; ++i) NB_CUDA(NULL, cudaEventCreateWithFlags(&c->aux_ev[i], cudaEventDisableTiming));
    for (int i = 0; i < NB_AUX; ++i) NB_CUDA(NULL, cudaEventCreateWithFlags(&c->link_ev[i], cudaEventDisableTiming));
    return NB_OK;
}

int nb_worker_ctx(nb_ctx* parent, int w, nb_ctx** out)
{
    if (w < 1 || w >= NB_MAX_WORKERS) return nb_fail(parent, NB_EINVAL, "worker index %d out of range", w);
    if (!parent->workers[w]) {
        if (!parent->dev_cache) parent->dev_cache = new NbDevCache();
        nb_ctx* c = (nb_ctx*)calloc(1, sizeof(nb_ctx));
        if (!c) return nb_fail(parent, NB_ENOMEM, "out of host memory");
        c->device = parent->device;
        c->is_worker = 1;
        c->parent = parent;
        c->dev_cache = parent->dev_cache;
        NB_CUDA(parent, cudaSetDevice(parent->device));
        const int rc = nb_ctx_init_streams(c);
        if (rc) { free(c); memcpy(parent->err, g_nb_err, sizeof g_nb_err); return rc; }
        parent->workers[w] = c;
    }
    *out = parent->workers[w];
    return NB_OK;
}

extern "C" void nb_ctx_destroy(nb_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    for (int w = 1; w < NB_MAX_WORKERS; ++w) {
        nb_ctx* k = c->workers[w];
        if (!k) continue;
        cudaStreamSynchronize(k->stream);
        cudaEventDestroy(k->ev0); cudaEventDestroy(k->ev1); cudaEventDestroy(k->ev2); cudaEventDestroy(k->ev3);
        cudaStreamDestroy(k->stream);
        for (int i = 0; i < NB_AUX; ++i) cudaStreamDestroy(k->aux[i]);
        for (int i = 0; i < NB_AUX + 1; ++i) cudaEventDestroy(k->aux_ev[i]);
        for (int i = 0; i < NB_AUX; ++i) cudaEventDestroy(k->link_ev[i]);
        if (k->ext_mat) cudaFree(k->ext_mat);
        if (k->pinned) cudaFreeHost(k->pinned);
        free(k);
    }
    cudaStreamSynchronize(c->stream);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    cudaEventDestroy(c->ev2);
    cudaEventDestroy(c->ev3);
    cudaStreamDestroy(c->stream);
    for (int i = 0; i < NB_AUX; ++i) cudaStreamDestroy(c->aux[i]);
    for (int i = 0; i < NB_AUX + 1; ++i) cudaEventDestroy(c->aux_ev[i]);
    for (int i = 0; i < NB_AUX; ++i) cudaEventDestroy(c->link_ev[i]);
    if (c->ext_mat) cudaFree(c->ext_mat);
    nb_dev_cache_destroy(c);
    if (c->pinned) cudaFreeHost(c->pinned);
    for (int k = 0; k < 4; ++k) if (c->pin_cache[k]) cudaFreeHost(c->pin_cache[k]);
    free(c);
}

extern "C" uint64_t nb_launch_count(const nb_ctx* c) { return c ? c->launches : 0; }
extern "C" void nb_get_timings(const nb_ctx* c, nb_timings* out) { if (c && out) *out = c->tm; }
extern "C" void nb_get_map_counters(const nb_ctx* c, nb_map_counters* out) { if (c && out) *out = c->mc; }
extern "C" void nb_get_cns_counters(const nb_ctx* c, nb_cns_counters* out) { if (c && out) *out = c->cc; }
extern "C" void nb_free(void* p) { free(p); }

extern "C" int nb_volume_upload(nb_ctx* ctx, const uint8_t* pac, uint64_t db_size, const uint64_t* seq_offset,
                                const uint64_t* seq_size, uint64_t nseq, nb_volume** out)
{
    if (!ctx || !out || (!pac && db_size) || (nseq && (!seq_offset || !seq_size)))
        return nb_fail(ctx, NB_EINVAL, "nb_volume_upload: bad arguments");
    // (no size limit here: offsets are 64-bit everywhere.  A volume that gets an index has to stay below 2^32 bases --
    //  nb_index_build checks that -- the merged read store of oc2cns may be as large as device memory allows.)
    NB_CUDA(ctx, cudaSetDevice(ctx->device));
    nb_volume* v = (nb_volume*)calloc(1, sizeof(nb_volume));
    if (!v) return nb_fail(ctx, NB_ENOMEM, "out of host memory");
    v->db_size = db_size;
    v->nseq = nseq;
    v->owner = ctx;
    cudaStream_t st = ctx->stream;
    const uint64_t nb = (db_size + 3) >> 2;
    const uint64_t padded = ((nb + 15) & ~15ull) + 256;
    int rc;
    auto drop = [&](int code) {             // error exit: nothing of the half-built volume stays behind
        if (v->d_pac) nb_dev_release(ctx, v->d_pac);
        if (v->d_off) nb_dev_release(ctx, v->d_off);
        if (v->d_len) nb_dev_release(ctx, v->d_len);
        free(v->h_off); free(v->h_len);
        if (v->h_pac) cudaFreeHost(v->h_pac);
        free(v);
        return code;
    };
    if ((rc = nb_dev_alloc(ctx, (void**)&v->d_pac, padded))) return drop(rc);
    if ((rc = nb_dev_alloc(ctx, (void**)&v->d_off, sizeof(uint64_t) * (nseq + 1)))) return drop(rc);
    if ((rc = nb_dev_alloc(ctx, (void**)&v->d_len, sizeof(uint32_t) * (nseq + 1)))) return drop(rc);
    v->h_off = (uint64_t*)malloc(sizeof(uint64_t) * (nseq + 1));
    v->h_len = (uint32_t*)malloc(sizeof(uint32_t) * (nseq + 1));
    if (!v->h_off || !v->h_len) return drop(nb_fail(ctx, NB_ENOMEM, "out of host memory"));
    v->max_len = 0;
    for (uint64_t i = 0; i < nseq; ++i) {
        if (seq_size[i] >= (1ull << 31) || seq_offset[i] + seq_size[i] > db_size) {
            return drop(nb_fail(ctx, NB_EINVAL, "sequence %llu out of range", (unsigned long long)i));
        }
        v->h_off[i] = seq_offset[i];
        v->h_len[i] = (uint32_t)seq_size[i];
        if (v->h_len[i] > v->max_len) v->max_len = v->h_len[i];
    }
    // the library's host copy of the packed bases is pinned: one memcpy from the caller's buffer, then the upload
    // runs from it at full link speed
    for (int k = 0; k < 4 && !v->h_pac; ++k)
        if (ctx->pin_cache[k] && ctx->pin_cache_bytes[k] >= padded && ctx->pin_cache_bytes[k] <= padded * 2 + (1u << 20)) {
            v->h_pac = (uint8_t*)ctx->pin_cache[k]; v->h_pac_bytes = ctx->pin_cache_bytes[k];
            ctx->pin_cache[k] = NULL; ctx->pin_cache_bytes[k] = 0;
        }
    if (!v->h_pac) {
        if (cudaHostAlloc((void**)&v->h_pac, padded, cudaHostAllocDefault) != cudaSuccess) { v->h_pac = NULL; return drop(nb_fail(ctx, NB_ENOMEM, "out of pinned host memory (%llu bytes)", (unsigned long long)padded)); }
        v->h_pac_bytes = padded;
    }
    memcpy(v->h_pac, pac, nb);
    memset(v->h_pac + nb, 0, padded - nb);
    v->h_off[nseq] = db_size;
    v->h_len[nseq] = 0;
    cudaEventRecord(ctx->ev0, st);
    NB_CUDA(ctx, cudaMemcpyAsync(v->d_pac, v->h_pac, padded, cudaMemcpyHostToDevice, st));
    NB_CUDA(ctx, cudaMemcpyAsync(v->d_off, v->h_off, sizeof(uint64_t) * (nseq + 1), cudaMemcpyHostToDevice, st));
    NB_CUDA(ctx, cudaMemcpyAsync(v->d_len, v->h_len, sizeof(uint32_t) * (nseq + 1), cudaMemcpyHostToDevice, st));
    cudaEventRecord(ctx->ev1, st);
    NB_CUDA(ctx, cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    ctx->tm.h2d_ms += ms;
    ctx->tm.h2d_bytes += nb + 12 * (nseq + 1);
    *out = v;
    return NB_OK;
}

extern "C" void nb_volume_free(nb_ctx* ctx, nb_volume* v)
{
    if (!v) return;
    nb_ctx* c = v->owner ? v->owner : ctx;
    if (c) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        nb_dev_release(c, v->d_pac);
        nb_dev_release(c, v->d_off);
        nb_dev_release(c, v->d_len);
        int slot = -1;
        for (int k = 0; k < 4; ++k) if (!c->pin_cache[k]) { slot = k; break; }
        if (slot >= 0) { c->pin_cache[slot] = v->h_pac; c->pin_cache_bytes[slot] = v->h_pac_bytes; }
        else cudaFreeHost(v->h_pac);
    }
    free(v->h_off);
    free(v->h_len);
    free(v);
}

extern "C" void nb_map_options_default(nb_map_options* o)
{
    o->kmer_size = 15; o->scan_window = 10; o->kmer_cnt_cutoff = 500; o->block_size = 2000;
    o->block_score_cutoff = 3; o->num_candidates = 500; o->align_size_cutoff = 500;
    o->ddfs_cutoff = 0.25; o->error = 0.5; o->num_output = 500; o->num_threads = 1;
    o->job = 1; o->binary_output = 0; o->use_hdr_as_id = 1;
}

extern "C" void nb_cns_options_default(nb_cns_options* o)
{
    o->min_align_size = 400; o->min_cov = 4; o->max_cov = 12; o->min_size = 500; o->full_consensus = 0;
    o->error = 0.5; o->mapping_ratio = 0.8; o->num_threads = 1; o->rescue_long_indels = 0;
    o->use_fixed_ident_cutoff = 0; o->small_memory = 0;
}

extern "C" void nb_pack_candidates(const nb_candidate* c, uint64_t n, nb_packed_candidate* out)
{
    for (uint64_t i = 0; i < n; ++i) nb_pack_candidate(c + i, out + i);
}

extern "C" void nb_swap_candidate_roles(const nb_packed_candidate* in, uint64_t n, nb_packed_candidate* out)
{
    for (uint64_t i = 0; i < n; ++i) {
        const nb_packed_candidate a = in[i];
        const uint32_t w = a.item[0];
        nb_packed_candidate b;
        b.item[0] = (w & ((1u << 30) - 1)) | (((w >> 31) & 1u) << 30) | (((w >> 30) & 1u) << 31);
        b.item[1] = a.item[4]; b.item[2] = a.item[5]; b.item[3] = a.item[6];
        b.item[4] = a.item[1]; b.item[5] = a.item[2]; b.item[6] = a.item[3];
        out[i] = b;
    }
}

extern "C" void nb_pack_candidate(const nb_candidate* c, nb_packed_candidate* p)
{
    uint32_t w = 0;
    if (c->sdir == 1) w |= 1u << 31;
    if (c->qdir == 1) w |= 1u << 30;
    if (c->qoff == c->qbeg) w |= 1u << 29;
    w |= (uint32_t)(c->score > 1000000 ? 1000000 : c->score);
    p->item[0] = w;
    p->item[1] = (uint32_t)c->sid;
    p->item[2] = (uint32_t)c->sbeg;
    p->item[3] = (uint32_t)c->send;
    p->item[4] = (uint32_t)c->qid;
    p->item[5] = (uint32_t)c->qbeg;
    p->item[6] = (uint32_t)c->qend;
}
