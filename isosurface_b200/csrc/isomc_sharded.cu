/*
 * isomc_sharded.cu -- z-slab sharding over the GPUs of one box INSIDE the library: the entry points a
 * single-process host (the Rust shim, the C++ mirror) uses; a one-process-per-GPU host (torchrun) drives the isomc_slab_*
 * calls itself and brings its own collective.
 *
 *   rank g = device devices[g] owns cell layers [z_g, z_g+1) of the N cell layers and is given sample layers
 *   [z_g - (g > 0), z_g+1]; it counts its ghost layer too, so its local numbering is a window of the global one.
 *   One exchange: ncclAllGather of {V, V before the last layer, T} (3 x u64 per rank) on the extraction streams, grouped
 *   over the ranks of this process.  The id offsets are derived on the device (k_slab_bases) and added while the indices
 *   are written: no re-index pass, and no host synchronisation between the count and the emission of a step.
 *
 * NCCL is resolved with dlopen at the first isomc_sharded_create (libnccl.so.2: the system's, or the one a framework
 * already loaded), so the library itself carries no link-time dependency on it.  If the same device is listed more than
 * once (a single-GPU box exercising the sharded path) the exchange is done with stream-ordered device copies instead:
 * NCCL does not allow one device twice in a communicator.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/isomc.h"

namespace {

typedef struct ncclComm *nccl_comm_t;
typedef int nccl_result_t;
constexpr int NCCL_UINT64 = 5; /* ncclUint64, nccl.h */

struct NcclApi {
    void *lib = nullptr;
    nccl_result_t (*CommInitAll)(nccl_comm_t *, int, const int *) = nullptr;
    nccl_result_t (*CommDestroy)(nccl_comm_t) = nullptr;
    nccl_result_t (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    nccl_result_t (*GroupStart)() = nullptr;
    nccl_result_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(nccl_result_t) = nullptr;
    std::string why;
    bool load() {
        if (lib) return true;
        if (!why.empty()) return false;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { why = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
        auto sym = [&](const char *n) { void *p = dlsym(lib, n); if (!p && why.empty()) why = std::string("libnccl lacks ") + n; return p; };
        CommInitAll = (decltype(CommInitAll))sym("ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
        AllGather = (decltype(AllGather))sym("ncclAllGather");
        GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
        if (!why.empty()) { lib = nullptr; return false; }
        return true;
    }
};
NcclApi g_nccl;
thread_local std::string g_sharded_create_error;

}  // namespace

struct isomc_sharded {
    uint32_t size = 0, n = 0;
    std::vector<int> dev;
    std::vector<isomc_t *> h;
    std::vector<cudaStream_t> stream;
    std::vector<uint32_t> z0, z1;
    std::vector<unsigned long long *> gathered; /* per rank, on its device: 3 * n u64 */
    std::vector<const uint64_t *> totals;       /* per rank: device pointer to its {V, V before last layer, T} */
    std::vector<nccl_comm_t> comm;
    std::vector<cudaEvent_t> ev;                /* single-device mode: count of rank r is enqueued */
    bool use_nccl = false, use_mailbox = false, have_result = false;
    std::vector<uint64_t> nv, nt, na;
    std::string err;
};

namespace {

int32_t sfail(isomc_sharded *s, int32_t code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (s) s->err = buf; else g_sharded_create_error = buf;
    return code;
}

#define SCU(s, call)                                                                                              \
    do {                                                                                                          \
        cudaError_t e_ = (call);                                                                                  \
        if (e_ != cudaSuccess) return sfail((s), ISOMC_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)
#define SNC(s, call)                                                                                                    \
    do {                                                                                                                \
        nccl_result_t r_ = (call);                                                                                      \
        if (r_ != 0) return sfail((s), ISOMC_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?"); \
    } while (0)
#define SRC(s, r, call)                                                                                        \
    do {                                                                                                       \
        int32_t rc_ = (call);                                                                                  \
        if (rc_) return sfail((s), rc_, "rank %u: %s", (unsigned)(r), isomc_last_error((s)->h[(r)]));          \
    } while (0)

/* contiguous cell-layer ranges balanced to +-1 layer (the same split bench.py and sharded.py use) */
void slab_range(uint32_t size, uint32_t rank, uint32_t world, uint32_t *z0, uint32_t *z1) {
    const uint32_t base = size / world, rem = size % world;
    *z0 = rank * base + (rank < rem ? rank : rem);
    *z1 = *z0 + base + (rank < rem ? 1u : 0u);
}

/* count on every rank, exchange the totals, emit on every rank: everything enqueued, nothing synchronised */
int32_t exchange_and_emit(isomc_sharded *s, bool already_enqueued = false) {
    if (already_enqueued) {
    } else if (s->use_mailbox) { /* totals as peer stores over NVLink: one tiny kernel per rank publishes, waits and derives the offset */
        for (uint32_t r = 0; r < s->n; ++r) SRC(s, r, isomc_slab_enqueue_emit_exchanged(s->h[r]));
    } else if (s->use_nccl) {
        SNC(s, g_nccl.GroupStart());
        for (uint32_t r = 0; r < s->n; ++r)
            SNC(s, g_nccl.AllGather(s->totals[r], s->gathered[r], 3, NCCL_UINT64, s->comm[r], s->stream[r]));
        SNC(s, g_nccl.GroupEnd());
    } else {
        for (uint32_t r = 0; r < s->n; ++r) {
            SCU(s, cudaSetDevice(s->dev[r]));
            SCU(s, cudaEventRecord(s->ev[r], s->stream[r]));
        }
        for (uint32_t r = 0; r < s->n; ++r) {
            SCU(s, cudaSetDevice(s->dev[r]));
            for (uint32_t q = 0; q < s->n; ++q) {
                if (q != r) SCU(s, cudaStreamWaitEvent(s->stream[r], s->ev[q], 0));
                SCU(s, cudaMemcpyAsync(s->gathered[r] + 3 * q, s->totals[q], 3 * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s->stream[r]));
            }
        }
    }
    if (!s->use_mailbox)
        for (uint32_t r = 0; r < s->n; ++r) SRC(s, r, isomc_slab_enqueue_emit_gathered(s->h[r], (const uint64_t *)s->gathered[r], r, s->n));
    /* first extract of a handle, or a larger mesh: finish() grows the buffers and re-runs the emission of that rank */
    for (uint32_t r = 0; r < s->n; ++r) SRC(s, r, isomc_finish(s->h[r]));
    uint64_t vsum = 0;
    for (uint32_t r = 0; r < s->n; ++r) {
        SRC(s, r, isomc_counts(s->h[r], &s->nv[r], &s->nt[r], &s->na[r]));
        vsum += s->nv[r];
    }
    if (vsum >= (1ull << 32)) return sfail(s, ISOMC_ERR_INDEX_OVERFLOW, "sharded mesh has %llu vertices: does not fit u32 indices", (unsigned long long)vsum);
    s->have_result = true;
    return ISOMC_OK;
}

}  // namespace

extern "C" {

int32_t isomc_sharded_create(uint32_t size, uint32_t n_gpus, const int32_t *devices, isomc_sharded_t **out) {
    if (!out) return sfail(nullptr, ISOMC_ERR_BAD_ARG, "out == NULL");
    *out = nullptr;
    if (n_gpus < 1 || n_gpus > 64 || n_gpus > size) return sfail(nullptr, ISOMC_ERR_BAD_ARG, "bad number of slabs %u for size %u", n_gpus, size);
    isomc_sharded *s = new (std::nothrow) isomc_sharded();
    if (!s) return sfail(nullptr, ISOMC_ERR_OOM, "host allocation failed");
    s->size = size; s->n = n_gpus;
    s->dev.resize(n_gpus); s->h.assign(n_gpus, nullptr); s->stream.assign(n_gpus, nullptr);
    s->z0.resize(n_gpus); s->z1.resize(n_gpus); s->gathered.assign(n_gpus, nullptr); s->totals.assign(n_gpus, nullptr);
    s->ev.assign(n_gpus, nullptr); s->nv.assign(n_gpus, 0); s->nt.assign(n_gpus, 0); s->na.assign(n_gpus, 0);
    bool distinct = true;
    for (uint32_t r = 0; r < n_gpus; ++r) {
        s->dev[r] = devices ? devices[r] : (int)r;
        for (uint32_t q = 0; q < r; ++q) distinct &= s->dev[q] != s->dev[r];
    }
    auto body = [&]() -> int32_t {
        for (uint32_t r = 0; r < n_gpus; ++r) {
            slab_range(size, r, n_gpus, &s->z0[r], &s->z1[r]);
            int32_t rc = isomc_slab_create(size, s->z0[r], s->z1[r], s->dev[r], &s->h[r]);
            if (rc) return sfail(s, rc, "rank %u: %s", r, isomc_last_error(nullptr));
            void *st = nullptr;
            SRC(s, r, isomc_get_stream(s->h[r], &st));
            s->stream[r] = (cudaStream_t)st;
            SRC(s, r, isomc_slab_totals_device(s->h[r], &s->totals[r]));
            SCU(s, cudaSetDevice(s->dev[r]));
            SCU(s, cudaMalloc(&s->gathered[r], 3 * n_gpus * sizeof(unsigned long long)));
            SCU(s, cudaMemset(s->gathered[r], 0, 3 * n_gpus * sizeof(unsigned long long)));
            SCU(s, cudaEventCreateWithFlags(&s->ev[r], cudaEventDisableTiming));
        }
        /* distinct devices that reach each other as peers exchange through mailboxes in peer memory (ISOMC_EXCHANGE=nccl: NCCL) */
        const char *xe = getenv("ISOMC_EXCHANGE");
        bool peers_ok = distinct && n_gpus > 1 && n_gpus <= 64 && !(xe && strcmp(xe, "nccl") == 0);
        for (uint32_t r = 0; r < n_gpus && peers_ok; ++r)
            for (uint32_t q = 0; q < n_gpus && peers_ok; ++q) {
                int can = 1;
                if (q != r && (cudaDeviceCanAccessPeer(&can, s->dev[r], s->dev[q]) != cudaSuccess || !can)) peers_ok = false;
            }
        if (peers_ok) {
            std::vector<void *> boxes(n_gpus, nullptr);
            for (uint32_t r = 0; r < n_gpus; ++r) SRC(s, r, isomc_slab_mailbox(s->h[r], &boxes[r]));
            for (uint32_t r = 0; r < n_gpus; ++r) SRC(s, r, isomc_slab_connect(s->h[r], r, n_gpus, boxes.data()));
            s->use_mailbox = true;
        } else if (distinct && n_gpus > 1) {
            if (!g_nccl.load()) return sfail(s, ISOMC_ERR_NCCL, "%s", g_nccl.why.c_str());
            s->comm.assign(n_gpus, nullptr);
            SNC(s, g_nccl.CommInitAll(s->comm.data(), (int)n_gpus, s->dev.data()));
            s->use_nccl = true;
        }
        return ISOMC_OK;
    };
    const int32_t rc = body();
    if (rc) {
        g_sharded_create_error = s->err;
        isomc_sharded_destroy(s);
        return rc;
    }
    *out = s;
    return ISOMC_OK;
}

int32_t isomc_sharded_destroy(isomc_sharded_t *s) {
    if (!s) return ISOMC_OK;
    for (uint32_t r = 0; r < s->n; ++r) {
        cudaSetDevice(s->dev[r]);
        if (r < s->comm.size() && s->comm[r]) g_nccl.CommDestroy(s->comm[r]);
        if (s->h[r]) isomc_destroy(s->h[r]);
        cudaFree(s->gathered[r]);
        if (s->ev[r]) cudaEventDestroy(s->ev[r]);
    }
    delete s;
    return ISOMC_OK;
}

const char *isomc_sharded_last_error(const isomc_sharded_t *s) { return s ? s->err.c_str() : g_sharded_create_error.c_str(); }

int32_t isomc_sharded_uses_nccl(const isomc_sharded_t *s) { return s && s->use_nccl ? 1 : 0; }
int32_t isomc_sharded_uses_peer_memory(const isomc_sharded_t *s) { return s && s->use_mailbox ? 1 : 0; }

int32_t isomc_sharded_slab(const isomc_sharded_t *s, uint32_t rank, uint32_t *z_begin, uint32_t *z_end, uint32_t *first_sample_layer,
                           uint32_t *n_sample_layers) {
    if (!s || rank >= s->n) return ISOMC_ERR_BAD_ARG;
    const uint32_t ghost = s->z0[rank] > 0 ? 1u : 0u;
    if (z_begin) *z_begin = s->z0[rank];
    if (z_end) *z_end = s->z1[rank];
    if (first_sample_layer) *first_sample_layer = s->z0[rank] - ghost;
    if (n_sample_layers) *n_sample_layers = s->z1[rank] - s->z0[rank] + ghost + 1;
    return ISOMC_OK;
}

int32_t isomc_sharded_handle(isomc_sharded_t *s, uint32_t rank, isomc_t **h) {
    if (!s || rank >= s->n || !h) return ISOMC_ERR_BAD_ARG;
    *h = s->h[rank];
    return ISOMC_OK;
}

int32_t isomc_sharded_extract_grid(isomc_sharded_t *s, const float *const *d_slabs) {
    if (!s) return ISOMC_ERR_BAD_ARG;
    if (!d_slabs) return sfail(s, ISOMC_ERR_BAD_ARG, "d_slabs == NULL");
    s->have_result = false;
    for (uint32_t r = 0; r < s->n; ++r) {
        if (!d_slabs[r]) return sfail(s, ISOMC_ERR_BAD_ARG, "d_slabs[%u] == NULL", r);
        /* peer-memory exchange: count, exchange and emission of a rank are one launch sequence (one graph launch per rank) */
        if (s->use_mailbox) SRC(s, r, isomc_slab_enqueue_extract_grid_exchanged(s->h[r], d_slabs[r]));
        else SRC(s, r, isomc_slab_count_grid_device(s->h[r], d_slabs[r]));
    }
    return exchange_and_emit(s, s->use_mailbox);
}

int32_t isomc_sharded_extract_sdf(isomc_sharded_t *s, const isomc_sdf_node *prog, uint32_t n_nodes) {
    if (!s) return ISOMC_ERR_BAD_ARG;
    s->have_result = false;
    for (uint32_t r = 0; r < s->n; ++r) SRC(s, r, isomc_slab_count_sdf(s->h[r], prog, n_nodes));
    return exchange_and_emit(s);
}

int32_t isomc_sharded_extract_sdf_directed(isomc_sharded_t *s, const isomc_sdf_node *prog, uint32_t n_nodes) {
    if (!s) return ISOMC_ERR_BAD_ARG;
    s->have_result = false;
    for (uint32_t r = 0; r < s->n; ++r) SRC(s, r, isomc_slab_count_sdf_directed(s->h[r], prog, n_nodes));
    return exchange_and_emit(s);
}

int32_t isomc_sharded_counts(isomc_sharded_t *s, uint64_t *n_vertices, uint64_t *n_triangles, uint64_t *n_active_cells) {
    if (!s) return ISOMC_ERR_BAD_ARG;
    if (!s->have_result) return sfail(s, ISOMC_ERR_NO_RESULT, "no sharded extract has completed");
    uint64_t v = 0, t = 0, a = 0;
    for (uint32_t r = 0; r < s->n; ++r) { v += s->nv[r]; t += s->nt[r]; a += s->na[r]; }
    if (n_vertices) *n_vertices = v;
    if (n_triangles) *n_triangles = t;
    if (n_active_cells) *n_active_cells = a;
    return ISOMC_OK;
}

int32_t isomc_sharded_rank_counts(isomc_sharded_t *s, uint32_t rank, uint64_t *n_vertices, uint64_t *n_triangles, uint64_t *n_active_cells) {
    if (!s || rank >= s->n) return ISOMC_ERR_BAD_ARG;
    if (!s->have_result) return sfail(s, ISOMC_ERR_NO_RESULT, "no sharded extract has completed");
    if (n_vertices) *n_vertices = s->nv[rank];
    if (n_triangles) *n_triangles = s->nt[rank];
    if (n_active_cells) *n_active_cells = s->na[rank];
    return ISOMC_OK;
}

/* the ranks' parts concatenated in rank order = the unsharded mesh (indices are global already) */
int32_t isomc_sharded_copy_out(isomc_sharded_t *s, float *xyz, uint32_t *idx) {
    if (!s) return ISOMC_ERR_BAD_ARG;
    if (!s->have_result) return sfail(s, ISOMC_ERR_NO_RESULT, "no sharded extract has completed");
    uint64_t vo = 0, to = 0;
    for (uint32_t r = 0; r < s->n; ++r) {
        SRC(s, r, isomc_copy_out(s->h[r], xyz ? xyz + 3 * vo : nullptr, idx ? idx + 3 * to : nullptr));
        vo += s->nv[r]; to += s->nt[r];
    }
    return ISOMC_OK;
}

} /* extern "C" */
