/*
 * isomc_api.cu -- the C ABI of include/isomc.h: handle management, stream-ordered pipeline,
 * result delivery.  Mirrors `MarchingCubes::new(size)` / `.extract(&source, &mut extractor)`
 * (reference src/marching_cubes.rs:46-50,59-82) and the IndexedVertices sink
 * (src/extractor.rs:72-93).  No CPU fallback: every compute entry point needs a CUDA device.
 */
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <utility>
#include <vector>
#include <cstdlib>

#include "../../include/isomc.h"
#include "isomc_cell.cuh"
#include "isomc_device.cuh"
#include "isomc_kernels.h"
#include "isomc_tables.h"
#include "isomc_tile.cuh"

namespace {

thread_local std::string g_create_error;

enum SrcKind { SRC_NONE = 0, SRC_GRID = 1, SRC_SDF = 2, SRC_SDF_BATCH = 3 };
constexpr int MAX_CHUNKS = 16;
/* u32 after layerTot: emit tickets [MAX_CHUNKS], list block counter, list marks [MAX_CHUNKS + 1], chunk ends [2 * MAX_CHUNKS],
 * tile path: block counters [2] (+ 2 pad), tickets of pass 1 [MAX_CHUNKS] and of pass 2 [MAX_CHUNKS] */
constexpr int AUX_WORDS = 6 * MAX_CHUNKS + 8;
constexpr int N_TOTALS = 16;

}  // namespace

struct isomc {
    uint32_t size = 0, z_begin = 0, z_end = 0;
    int device = 0, sms = 148;
    Geo g{};
    cudaStream_t own_stream = nullptr, stream = nullptr;
    uint32_t n_chunks = 1, chunk_l[MAX_CHUNKS + 1] = {}; /* cell layers [chunk_l[c], chunk_l[c+1]): one chunk, or the z-chunks of the streamed host extract */
    /* ISOMC_TIMELINE=1: timing events after every kernel, printed at finish (debugging the stream overlap) */
    bool timeline = false;
    std::vector<std::pair<std::string, cudaEvent_t>> tl;
    /* scratch */
    uint32_t *signs = nullptr, *rowV = nullptr, *rowT = nullptr, *rowA = nullptr;
    unsigned long long *layerTot = nullptr, *totals = nullptr; /* totals: 12 u64 */
    uint32_t *vofs = nullptr, *ticket = nullptr;
    size_t zero_bytes = 0; /* what the memset at the start of an extract clears: layerTot, the auxiliary words, (list path) rowV/rowT/rowA */
    /* active-cell-list kernels (isomc_cell.cuh) by default; ISOMC_PATH=tile selects the TMA-staged tile path (isomc_tile.cuh):
     * parity-green and profiled, but measured slower on every workload (profiles/r02_tile_path.md) */
    bool tile_mode = false;
    TileGeo tg{};
    TileBufs TB{};
    uint32_t *tile_tickets = nullptr; /* [MAX_CHUNKS] pass 1, [MAX_CHUNKS] pass 2 */
    uint32_t *segA = nullptr;         /* PointCloud: per-segment prefixes (allocated on first use, with `signs`) */
    /* batched chunks (isomc_batch_create): `batch` lattices stacked in z (Geo.zper), one implicit tree each */
    uint32_t batch = 0, batch_used = 0;
    bool batch_call = false; /* the extract being enqueued comes from a batch entry point */
    SdfProgram *d_progs = nullptr, *h_progs = nullptr; /* h_progs pinned */
    uint32_t *chunkV = nullptr, *chunkT = nullptr;     /* device: [batch + 1] output slots of the chunks' first vertex / triangle */
    uint32_t *h_chunk = nullptr;                       /* pinned: 2 * (batch + 1) */
    ListBufs L{};
    uint32_t *list_marks = nullptr; /* [c] = list blocks handed out before z-chunk c; [0] = 0 */
    EmitTab *etab = nullptr;
    /* stage pipeline of a device-resident extract (pipe_plan): sign words of z-chunk c+2, counting of c+1 and emission of c run
     * on three streams in CTA-limited grids, so that the HBM-bound stage shares the SMs with the two issue-bound ones */
    cudaStream_t s_sign = nullptr, s_emit = nullptr;
    cudaEvent_t ev_pipe0 = nullptr, ev_pipe1 = nullptr, ev_sign[MAX_CHUNKS] = {}, ev_cnt2[MAX_CHUNKS] = {};
    /* the pipelined sequence (4 launches and 5 event operations per z-chunk) is captured once into a CUDA graph and replayed
     * while everything a kernel takes by value stays the same (pipe_key: pointers, capacities, plan) */
    cudaGraphExec_t pipe_exec = nullptr;
    uint64_t pipe_key[12] = {};
    bool graph_broken = false; /* capture / instantiation failed once on this handle: launch directly */
    uint32_t last_blocks = 0; /* list blocks the previous extract used (sizes the emission grid of small lattices) */
    uint32_t pipe_launches = 0, pipe_chunks = 1, pipe_chunk_l[MAX_CHUNKS + 1] = {};
    /* slab totals exchanged over peer memory (isomc_slab_connect*): own mailbox, the ranks' mailbox addresses, step counter */
    unsigned long long *mailbox = nullptr;
    unsigned long long **d_peers = nullptr;
    std::vector<void *> ipc_opened;
    uint32_t xchg_rank = 0, xchg_n = 0;
    bool seq_exchange = false; /* the launch sequence being built / replayed contains the peer exchange (slab extract in one go) */
    bool step_exchanged = false; /* the id offset of the step in flight came from k_slab_exchange (totals[14] = its time-out flag) */
    /* streamed host-to-host extract (isomc_extract_grid_host_to): copy-in / copy-out streams, per-chunk events */
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[MAX_CHUNKS] = {}, ev_cnt[MAX_CHUNKS] = {}, ev_emit[MAX_CHUNKS] = {};
    uint32_t *chunk_ends = nullptr;   /* device: {vertices, triangles} below the end of z-chunk c */
    uint32_t *h_chunk_ends = nullptr; /* pinned copy */
    bool emit_inline = false;       /* what the extract in flight was enqueued with (re-enqueued after a list grow) */
    int64_t vofs_cached = 0; /* value known to be in *vofs (set to 0 at create); -1 = written by the device */
    McTables *tabs = nullptr;
    unsigned long long *h_totals = nullptr; /* pinned */
    float *stage_grid = nullptr;            /* device copy of a host grid */
    /* results */
    float *xyz = nullptr;
    uint32_t *idx = nullptr;
    uint64_t cap_v = 0, cap_t = 0;
    uint64_t n_v = 0, n_t = 0, n_a = 0;
    bool have_result = false, counted = false, emitted = false, totals_valid = false;
    /* source of the extract in flight (needed to re-run emission after a buffer grow) */
    SrcKind kind = SRC_NONE;
    const float *d_grid = nullptr;
    SdfProgram prog{};
    bool directed = false; /* implicit source sampled as Directed distances (MarchingCubes<Directed>) */
    /* profiling */
    bool profiling = false;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    isomc_stats stats{};
    std::string err;
};

namespace {

int32_t fail(isomc *h, int32_t code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

#define CU(h, call)                                                                                   \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail((h), e_ == cudaErrorMemoryAllocation ? ISOMC_ERR_OOM : ISOMC_ERR_CUDA,       \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

int32_t validate_program(isomc *h, const isomc_sdf_node *prog, uint32_t n, SdfProgram *out) {
    if (!prog || n == 0) return fail(h, ISOMC_ERR_BAD_ARG, "empty SDF program");
    if (n > ISOMC_SDF_MAX_NODES)
        return fail(h, ISOMC_ERR_UNSUPPORTED_SOURCE, "SDF program has %u nodes (max %d)", n, ISOMC_SDF_MAX_NODES);
    int depth = 0, tdepth = 0;
    for (uint32_t i = 0; i < n; ++i) {
        switch (prog[i].op) {
        case ISOMC_SDF_SPHERE: case ISOMC_SDF_TORUS: case ISOMC_SDF_CYLINDER: case ISOMC_SDF_PRISM:
            if (++depth > ISOMC_SDF_MAX_STACK)
                return fail(h, ISOMC_ERR_UNSUPPORTED_SOURCE, "SDF program needs more than %d live values", ISOMC_SDF_MAX_STACK);
            break;
        case ISOMC_SDF_UNION: case ISOMC_SDF_INTERSECTION: case ISOMC_SDF_DIFFERENCE:
            if (depth < 2) return fail(h, ISOMC_ERR_BAD_ARG, "SDF node %u: binary op with < 2 operands", i);
            --depth;
            break;
        case ISOMC_SDF_TRANSLATE_PUSH:
            if (++tdepth > ISOMC_SDF_MAX_TRANSLATE)
                return fail(h, ISOMC_ERR_UNSUPPORTED_SOURCE, "more than %d nested translations", ISOMC_SDF_MAX_TRANSLATE);
            break;
        case ISOMC_SDF_TRANSLATE_POP:
            if (--tdepth < 0) return fail(h, ISOMC_ERR_BAD_ARG, "SDF node %u: TRANSLATE_POP without PUSH", i);
            break;
        default:
            return fail(h, ISOMC_ERR_UNSUPPORTED_SOURCE, "SDF node %u: unknown op %u (arbitrary closures are not a device path)", i, prog[i].op);
        }
    }
    if (depth != 1 || tdepth != 0) return fail(h, ISOMC_ERR_BAD_ARG, "SDF program does not reduce to one value");
    memset(out, 0, sizeof *out);
    memcpy(out->nodes, prog, n * sizeof(isomc_sdf_node));
    out->n = n;
    return ISOMC_OK;
}

int32_t bind_device(isomc *h) {
    CU(h, cudaSetDevice(h->device));
    return ISOMC_OK;
}

int32_t ensure_capacity(isomc *h, uint64_t nv, uint64_t nt) {
    if (nv > h->cap_v) {
        uint64_t want = nv + nv / 16 + 1024;
        if (h->xyz) CU(h, cudaFree(h->xyz));
        h->xyz = nullptr; h->cap_v = 0;
        CU(h, cudaMalloc(&h->xyz, want * 12));
        h->cap_v = want;
    }
    if (nt > h->cap_t) {
        uint64_t want = nt + nt / 16 + 1024;
        if (h->idx) CU(h, cudaFree(h->idx));
        h->idx = nullptr; h->cap_t = 0;
        CU(h, cudaMalloc(&h->idx, want * 12));
        h->cap_t = want;
    }
    return ISOMC_OK;
}

/* list of active cells: grow-only, sized from the previous extract; an overflow is detected from the block counter */
int32_t ensure_list_capacity(isomc *h, uint64_t blocks) {
    if (blocks <= h->L.cap_blocks) return ISOMC_OK;
    if (blocks > 0xFFFFF0ull) return fail(h, ISOMC_ERR_OOM, "active-cell list of %llu blocks exceeds the 32-bit entry positions", (unsigned long long)blocks);
    if (h->L.ent) CU(h, cudaFree(h->L.ent));
    if (h->L.ent_yz) CU(h, cudaFree(h->L.ent_yz));
    if (h->L.blkfill) CU(h, cudaFree(h->L.blkfill));
    h->L.ent = nullptr; h->L.ent_yz = nullptr; h->L.blkfill = nullptr; h->L.cap_blocks = 0;
    CU(h, cudaMalloc(&h->L.ent, blocks * LIST_BLOCK * sizeof(uint2)));
    CU(h, cudaMalloc(&h->L.ent_yz, blocks * LIST_BLOCK * sizeof(uint32_t)));
    CU(h, cudaMalloc(&h->L.blkfill, blocks * sizeof(uint32_t)));
    h->L.cap_blocks = (uint32_t)blocks;
    return ISOMC_OK;
}

/* tile path: entry list and crossing-parameter buffer, in blocks of ENT_BLOCK; grow-only, sized from the previous extract */
int32_t ensure_tile_capacity(isomc *h, uint64_t eb, uint64_t tb) {
    if (eb > 0xFFFFF0ull || tb > 0xFFFFF0ull)
        return fail(h, ISOMC_ERR_OOM, "entry list of %llu / %llu blocks exceeds the 32-bit positions", (unsigned long long)eb, (unsigned long long)tb);
    if (eb > h->TB.cap_eb) {
        if (h->TB.ent) CU(h, cudaFree(h->TB.ent));
        if (h->TB.tq) CU(h, cudaFree(h->TB.tq));
        h->TB.ent = nullptr; h->TB.tq = nullptr; h->TB.cap_eb = 0;
        CU(h, cudaMalloc(&h->TB.ent, eb * ENT_BLOCK * sizeof(uint2)));
        CU(h, cudaMalloc(&h->TB.tq, eb * ENT_BLOCK * 3 * sizeof(float)));
        /* pass 2 fetches all three slots of an entry before it knows which are in use: never read uninitialised memory */
        CU(h, cudaMemsetAsync(h->TB.tq, 0, eb * ENT_BLOCK * 3 * sizeof(float), h->stream));
        h->TB.cap_eb = (uint32_t)eb;
    }
    if (tb > h->TB.cap_tb) {
        if (h->TB.tbuf) CU(h, cudaFree(h->TB.tbuf));
        h->TB.tbuf = nullptr; h->TB.cap_tb = 0;
        CU(h, cudaMalloc(&h->TB.tbuf, tb * ENT_BLOCK * sizeof(float)));
        h->TB.cap_tb = (uint32_t)tb;
    }
    return ISOMC_OK;
}
/* blocks the counting warps may strand: one partly filled block per warp and space */
uint64_t tile_stranded_blocks(const isomc *h) { return (uint64_t)h->sms * 4 * TILE_Y + 16; }

/* sign words + per-segment scratch of the PointCloud path and the cube-index dump (the mesh path needs neither) */
int32_t ensure_signs(isomc *h) {
    const Geo &g = h->g;
    const uint64_t nrows_s = (uint64_t)g.nsl * g.N, nrows_c = (uint64_t)g.ncl * g.ncx;
    if (!h->signs) CU(h, cudaMalloc(&h->signs, (nrows_s * g.nws + 4) * sizeof(uint32_t)));
    if (!h->segA) CU(h, cudaMalloc(&h->segA, (nrows_c * g.nsegx + 4) * sizeof(uint32_t)));
    return ISOMC_OK;
}

void tl_mark(isomc *h, const char *name, uint32_t c, cudaStream_t st) {
    if (!h->timeline) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    h->tl.emplace_back(std::string(name) + "(" + std::to_string(c) + ")", e);
}

int32_t launch_emit_chunk(isomc *h, uint32_t c, cudaStream_t st, int bps = 0) {
    const Geo &g = h->g;
    if (bps == 0 && h->last_blocks) bps = -(int)h->last_blocks;
    const uint32_t l0 = h->chunk_l[c], l1 = h->chunk_l[c + 1];
    if (h->tile_mode) {
        /* pass 2: edge ids, vertex positions and triangles of the chunk's cell layers (warm-up on the layer below) */
        CU(h, isomc_launch_tile_emit(g, h->tg, h->TB, h->etab, h->vofs, h->xyz, h->idx, h->cap_v, h->cap_t, l0, l1,
                                     h->tile_tickets + MAX_CHUNKS + c, h->sms, st));
        h->stats.kernel_launches += 1;
        return ISOMC_OK;
    }
    /* one kernel: edge ids, vertex positions and triangles of the chunk's active cells */
    if (h->kind == SRC_SDF_BATCH)
        CU(h, isomc_launch_emit_list_sdf_batch(g, h->d_progs, h->directed, h->L, h->etab, h->rowV, h->rowT, h->layerTot, h->vofs, h->xyz, h->idx,
                                               h->cap_v, h->cap_t, h->list_marks + c, h->list_marks + c + 1, h->sms, st, bps));
    else if (h->kind == SRC_GRID)
        CU(h, isomc_launch_emit_list_grid(g, h->d_grid, h->L, h->etab, h->rowV, h->rowT, h->layerTot, h->vofs, h->xyz, h->idx,
                                          h->cap_v, h->cap_t, h->list_marks + c, h->list_marks + c + 1, h->sms, st, bps));
    else
        CU(h, isomc_launch_emit_list_sdf(g, h->prog, h->directed, h->L, h->etab, h->rowV, h->rowT, h->layerTot, h->vofs, h->xyz, h->idx,
                                         h->cap_v, h->cap_t, h->list_marks + c, h->list_marks + c + 1, h->sms, st, bps));
    h->stats.kernel_launches += 1;
    return ISOMC_OK;
}

/* CTAs per SM of the three stages while they share the SMs (pipelined extract); 0 = the stage has the GPU to itself */
struct StageGrids { int sign = 0, count = 0, emit = 0; };

/* sign words of the sample rows z-chunk c adds (list path) */
int32_t launch_sign_chunk(isomc *h, uint32_t c, cudaStream_t st, int bps = 0) {
    const Geo &g = h->g;
    const uint32_t l0 = h->chunk_l[c], l1 = h->chunk_l[c + 1];
    const uint32_t row0 = (c == 0 ? 0u : l0 + 1) * g.N, row1 = (l1 + 1) * g.N;
    if (h->kind == SRC_SDF_BATCH) CU(h, isomc_launch_sign_sdf_batch(g, h->d_progs, h->directed, h->signs, row0, row1, h->sms, st));
    else if (h->kind == SRC_GRID) CU(h, isomc_launch_sign_grid(g, h->d_grid, h->signs, row0, row1, h->sms, bps ? bps : 8, st));
    else CU(h, isomc_launch_sign_sdf(g, h->prog, h->directed, h->signs, row0, row1, h->sms, bps ? bps : 8, st));
    h->stats.kernel_launches += 1;
    return ISOMC_OK;
}

/* counting stage of z-chunk c (cell layers [l0, l1)) + the row scan; tile path: one kernel reads the samples once */
int32_t launch_count_chunk(isomc *h, uint32_t c, cudaStream_t st, uint32_t *chunk_end, bool with_sign = true, int bps = 0) {
    const Geo &g = h->g;
    const uint32_t l0 = h->chunk_l[c], l1 = h->chunk_l[c + 1];
    if (h->tile_mode) {
        if (h->kind == SRC_GRID)
            CU(h, isomc_launch_tile_count_grid(g, h->tg, h->d_grid, h->TB, h->etab, l0, l1, h->tile_tickets + c, h->sms, st));
        else
            CU(h, isomc_launch_tile_count_sdf(g, h->tg, h->prog, h->directed, h->TB, h->etab, l0, l1, h->tile_tickets + c, h->sms, st));
        if (h->profiling) { CU(h, cudaEventRecord(h->ev[1], st)); CU(h, cudaEventRecord(h->ev[2], st)); }
        CU(h, isomc_launch_scan(g, h->tg.ppl, h->rowV, h->rowT, h->layerTot, h->totals, h->TB.ctr, nullptr, chunk_end, l0, l1, st));
        if (h->profiling) CU(h, cudaEventRecord(h->ev[3], st));
        h->stats.kernel_launches += 2;
        return ISOMC_OK;
    }
    if (with_sign) {
        int32_t rc = launch_sign_chunk(h, c, st);
        if (rc) return rc;
    }
    if (h->profiling) CU(h, cudaEventRecord(h->ev[1], st));
    CU(h, isomc_launch_count_list(g, h->signs, h->tabs, h->L, h->rowV, h->rowT, h->rowA, h->layerTot, h->ticket + c, l0, l1, h->sms, st, bps));
    if (h->profiling) CU(h, cudaEventRecord(h->ev[2], st));
    CU(h, isomc_launch_scan(g, g.ncx, h->rowV, h->rowT, h->layerTot, h->totals, h->L.ctr, h->list_marks + c + 1, chunk_end, l0, l1, st));
    if (h->batch) { /* per-lattice totals -> output bases, grand totals */
        CU(h, isomc_launch_chunk_bases(h->batch, h->totals, h->L.ctr, h->chunkV, h->chunkT, st));
        h->stats.kernel_launches += 1;
    }
    if (h->profiling) CU(h, cudaEventRecord(h->ev[3], st));
    h->stats.kernel_launches += 2;
    return ISOMC_OK;
}

/*
 * Stage pipeline of a device-resident extract.  The three stages are bound by different things -- k_sign by HBM, k_count_list and
 * k_emit_list by instruction issue and load latency -- and each of them alone leaves the other resource idle.  The lattice is cut
 * into z-chunks (the row scan is causal in z, so a chunk can be emitted as soon as it is counted) and the stages run as
 *
 *   s_sign : sign(0) sign(1) sign(2) ...
 *   stream :         count+scan(0) count+scan(1) ...
 *   s_emit :                       emit(0)       emit(1) ...
 *
 * in grids of `sign` / `count` / `emit` CTAs per SM, chosen so that CTAs of all three are resident on every SM at once.
 * ISOMC_PIPE="chunks,sign,count,emit" overrides the plan ("0" = off).
 */
struct PipePlan { uint32_t chunks = 1; StageGrids sg; };

PipePlan pipe_plan(const isomc *h, bool emit_inline) {
    PipePlan p;
    const Geo &g = h->g;
    if (!emit_inline || h->tile_mode || h->batch || h->profiling || h->timeline || h->seq_exchange || g.ncl < 64) return p;
    /* OFF unless asked for: measured slower than the serial order in every configuration tried (fbm512: serial 0.54 ms; 2 chunks
     * 0.56-0.75 ms, 4 chunks 0.63, 8 chunks 0.66-0.77 with the CUDA graph, 0.97 without).  Each stage needs all the warps an SM
     * holds to hide its own load latency, and per-chunk launches of k_count_list / k_emit_list quantise badly (profiles/r02_history.md) */
    int v[4] = {0, 2, 2, 3};
    const char *e = getenv("ISOMC_PIPE");
    if (!e) return p;
    if (sscanf(e, "%d%*c%d%*c%d%*c%d", &v[0], &v[1], &v[2], &v[3]) < 1 || v[0] < 2) return p; /* "8x2x2x3" (any separator) */
    p.chunks = (uint32_t)v[0] > (uint32_t)MAX_CHUNKS ? (uint32_t)MAX_CHUNKS : (uint32_t)v[0];
    if (p.chunks > g.ncl / 8) p.chunks = g.ncl / 8;
    p.sg.sign = v[1]; p.sg.count = v[2]; p.sg.emit = v[3];
    return p;
}

/* totals of this slab to every rank's mailbox, the other ranks' totals from this rank's, id offset -> *vofs (k_slab_exchange) */
int32_t launch_exchange(isomc *h, cudaStream_t st) {
    static long long timeout = 0;
    if (!timeout) { const char *p = getenv("ISOMC_EXCHANGE_TIMEOUT_MS"); timeout = (long long)(p ? atof(p) : 10000.0) * 2000000ll; }
    CU(h, isomc_launch_slab_exchange(h->d_peers, h->xchg_rank, h->xchg_n, h->g.ghost, h->totals, h->vofs, timeout, st));
    h->vofs_cached = -1;
    h->step_exchanged = true;
    h->totals_valid = false; /* (totals[13..15] are new) */
    h->stats.kernel_launches += 1;
    return ISOMC_OK;
}

int32_t pipelined_launches(isomc *h, const PipePlan &pp);
int32_t serial_launches(isomc *h, bool emit_inline);

/*
 * The launch sequence of an extract (memset + 4 kernels; or the stage pipeline) replayed from a CUDA graph: one cudaGraphLaunch
 * instead of 5 .. 70 API calls, and the dependent launches keep their programmatic edges.  A graph is captured from the very
 * launch code below and stays valid while everything the kernels take BY VALUE stays the same (the key: source, output buffers
 * and capacities, list buffers, stream, plan); anything else re-captures.  Off while profiling (events between the kernels) and
 * for extracts without inline emission (slabs: their sequence is split by the exchange).  ISOMC_GRAPH=0 switches it off.
 */
uint64_t program_hash(const isomc *h) {
    uint64_t x = 1469598103934665603ull; /* FNV-1a over the nodes the kernels receive by value */
    const unsigned char *p = reinterpret_cast<const unsigned char *>(&h->prog);
    for (size_t i = 0; i < sizeof(SdfProgram); ++i) { x ^= p[i]; x *= 1099511628211ull; }
    return x;
}

int32_t enqueue_graphed(isomc *h, const PipePlan &pp, bool emit_inline) {
    static const bool use_graph = !(getenv("ISOMC_GRAPH") && atoi(getenv("ISOMC_GRAPH")) == 0);
    const bool pipelined = pp.chunks > 1;
    if (!use_graph || !emit_inline || h->profiling || h->timeline || h->tile_mode)
        return pipelined ? pipelined_launches(h, pp) : serial_launches(h, emit_inline);
    const uint64_t key[12] = {(uint64_t)(uintptr_t)h->d_grid, (uint64_t)(uintptr_t)h->xyz, (uint64_t)(uintptr_t)h->idx, h->cap_v, h->cap_t,
                              (uint64_t)(uintptr_t)h->L.ent, h->L.cap_blocks, pp.chunks,
                              (uint64_t)pp.sg.sign << 32 | (uint64_t)pp.sg.count << 16 | (uint64_t)pp.sg.emit,
                              (uint64_t)(uintptr_t)h->stream,
                              (uint64_t)h->last_blocks << 16 | (uint64_t)h->kind << 8 | (h->seq_exchange ? 2u : 0u) | (h->directed ? 1u : 0u),
                              h->kind == SRC_SDF ? program_hash(h) : 0ull};
    if (h->pipe_exec && memcmp(key, h->pipe_key, sizeof key) == 0) {
        CU(h, cudaGraphLaunch(h->pipe_exec, h->stream));
        h->stats.kernel_launches = h->pipe_launches;
        h->n_chunks = h->pipe_chunks;
        memcpy(h->chunk_l, h->pipe_chunk_l, sizeof h->chunk_l);
        h->counted = true; h->emitted = true;
        if (h->seq_exchange) { h->vofs_cached = -1; h->step_exchanged = true; h->totals_valid = false; }
        return ISOMC_OK;
    }
    if (h->pipe_exec) { cudaGraphExecDestroy(h->pipe_exec); h->pipe_exec = nullptr; }
    if (pipelined && !h->s_sign) return pipelined_launches(h, pp); /* (its streams and events are created outside a capture) */
    if (h->cap_v == 0 && h->cap_t == 0) return serial_launches(h, emit_inline);
    /* A driver that cannot capture or instantiate this sequence (programmatic edges in a capture need CUDA 12.3+) must not cost
     * the extract: the handle then launches directly from now on. */
    if (h->graph_broken || cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        h->graph_broken = true;
        return pipelined ? pipelined_launches(h, pp) : serial_launches(h, emit_inline);
    }
    int32_t rc = pipelined ? pipelined_launches(h, pp) : serial_launches(h, emit_inline);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
    if (rc == ISOMC_OK && e == cudaSuccess) e = cudaGraphInstantiate(&h->pipe_exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (rc != ISOMC_OK || e != cudaSuccess) { /* nothing has run yet (the launches were only recorded): run them directly */
        cudaGetLastError();
        h->pipe_exec = nullptr;
        h->graph_broken = true;
        h->stats.kernel_launches = 0;
        return pipelined ? pipelined_launches(h, pp) : serial_launches(h, emit_inline);
    }
    memcpy(h->pipe_key, key, sizeof key);
    h->pipe_launches = h->stats.kernel_launches;
    h->pipe_chunks = h->n_chunks;
    memcpy(h->pipe_chunk_l, h->chunk_l, sizeof h->chunk_l);
    CU(h, cudaGraphLaunch(h->pipe_exec, h->stream));
    return ISOMC_OK;
}

int32_t pipelined_launches(isomc *h, const PipePlan &pp) {
    const Geo &g = h->g;
    if (!h->s_sign) {
        CU(h, cudaStreamCreateWithFlags(&h->s_sign, cudaStreamNonBlocking));
        CU(h, cudaStreamCreateWithFlags(&h->s_emit, cudaStreamNonBlocking));
        CU(h, cudaEventCreateWithFlags(&h->ev_pipe0, cudaEventDisableTiming));
        CU(h, cudaEventCreateWithFlags(&h->ev_pipe1, cudaEventDisableTiming));
        for (int c = 0; c < MAX_CHUNKS; ++c) {
            CU(h, cudaEventCreateWithFlags(&h->ev_sign[c], cudaEventDisableTiming));
            CU(h, cudaEventCreateWithFlags(&h->ev_cnt2[c], cudaEventDisableTiming));
        }
    }
    const uint32_t per = (g.ncl + pp.chunks - 1) / pp.chunks;
    const uint32_t n = (g.ncl + per - 1) / per;
    h->n_chunks = n;
    for (uint32_t c = 0; c <= n; ++c) h->chunk_l[c] = c * per < g.ncl ? c * per : g.ncl;
    CU(h, cudaMemsetAsync(h->layerTot, 0, h->zero_bytes, h->stream));
    CU(h, cudaEventRecord(h->ev_pipe0, h->stream));
    CU(h, cudaStreamWaitEvent(h->s_sign, h->ev_pipe0, 0));
    CU(h, cudaStreamWaitEvent(h->s_emit, h->ev_pipe0, 0));
    int32_t rc;
    /* launch order = the order in which the hardware queues may start the kernels: keep the stages interleaved */
    for (uint32_t step = 0; step < n + 2; ++step) {
        if (step < n) {
            rc = launch_sign_chunk(h, step, h->s_sign, pp.sg.sign);
            if (rc) return rc;
            CU(h, cudaEventRecord(h->ev_sign[step], h->s_sign));
        }
        if (step >= 1 && step - 1 < n) {
            const uint32_t c = step - 1;
            CU(h, cudaStreamWaitEvent(h->stream, h->ev_sign[c], 0));
            rc = launch_count_chunk(h, c, h->stream, nullptr, false, pp.sg.count);
            if (rc) return rc;
            CU(h, cudaEventRecord(h->ev_cnt2[c], h->stream));
        }
        if (step >= 2) {
            const uint32_t c = step - 2;
            CU(h, cudaStreamWaitEvent(h->s_emit, h->ev_cnt2[c], 0));
            rc = launch_emit_chunk(h, c, h->s_emit, pp.sg.emit);
            if (rc) return rc;
        }
    }
    CU(h, cudaEventRecord(h->ev_pipe1, h->s_emit));
    CU(h, cudaStreamWaitEvent(h->stream, h->ev_pipe1, 0));
    h->counted = true;
    h->emitted = true;
    return ISOMC_OK;
}

/* phase 1 (+ optionally phase 2 inline): sign bits, counts, scans [, emission] per z-chunk */
int32_t enqueue_count(isomc *h, bool emit_inline) {
    const Geo &g = h->g;
    h->have_result = false; h->counted = false; h->emitted = false; h->totals_valid = false;
    h->stats.kernel_launches = 0; h->stats.emit_reruns = 0;
    h->emit_inline = emit_inline;
    if (g.ncl == 0 || g.ncx == 0) { /* size == 1: the reference visits no cells */
        CU(h, cudaMemsetAsync(h->totals, 0, 15 * sizeof(unsigned long long), h->stream)); /* ([15] is the exchange's step counter) */
        h->counted = true;
        h->emitted = emit_inline;
        return ISOMC_OK;
    }
    return enqueue_graphed(h, pipe_plan(h, emit_inline), emit_inline);
}

/* one chunk, one stream, every stage with the GPU to itself */
int32_t serial_launches(isomc *h, bool emit_inline) {
    const Geo &g = h->g;
    h->n_chunks = 1;
    h->chunk_l[0] = 0; h->chunk_l[1] = g.ncl;
    if (h->profiling) CU(h, cudaEventRecord(h->ev[0], h->stream));
    CU(h, cudaMemsetAsync(h->layerTot, 0, h->zero_bytes, h->stream));
    tl_mark(h, "start", 0, h->stream);
    int32_t rc = launch_count_chunk(h, 0, h->stream, nullptr);
    if (rc) return rc;
    tl_mark(h, "count+scan", 0, h->stream);
    if (h->seq_exchange) {
        rc = launch_exchange(h, h->stream);
        if (rc) return rc;
    }
    if (emit_inline) {
        rc = launch_emit_chunk(h, 0, h->stream);
        if (rc) return rc;
        tl_mark(h, "emit", 0, h->stream);
    }
    h->counted = true;
    h->emitted = emit_inline;
    return ISOMC_OK;
}

/* phase 2 on its own (after a buffer grow, or after the slab all-gather) */
int32_t enqueue_emit(isomc *h) {
    const Geo &g = h->g;
    if (g.ncl == 0 || g.ncx == 0) { h->emitted = true; return ISOMC_OK; }
    CU(h, cudaMemsetAsync(h->tile_tickets + MAX_CHUNKS, 0, MAX_CHUNKS * sizeof(uint32_t), h->stream)); /* emit tickets */
    for (uint32_t c = 0; c < h->n_chunks; ++c) {
        int32_t rc = launch_emit_chunk(h, c, h->stream);
        if (rc) return rc;
    }
    h->emitted = true;
    return ISOMC_OK;
}

int32_t enqueue_count(isomc *h, bool emit_inline);

/* bring the totals to the host (synchronises).  If the active-cell list was too small, grow it to what the count
 * asked for and run the count again (same totals; the first extract of a handle, or a much denser field). */
int32_t fetch_totals(isomc *h) {
    for (int attempt = 0; !h->totals_valid; ++attempt) {
        CU(h, cudaMemcpyAsync(h->h_totals, h->totals, N_TOTALS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        const uint64_t blocks = h->h_totals[7], tblocks = h->h_totals[12];
        if (h->tile_mode ? (blocks <= h->TB.cap_eb && tblocks <= h->TB.cap_tb) : blocks <= h->L.cap_blocks) { h->totals_valid = true; break; }
        if (attempt >= 2) return fail(h, ISOMC_ERR_CUDA, "active-cell list still too small after regrowing (%llu blocks)", (unsigned long long)blocks);
        const uint32_t reruns = h->stats.emit_reruns;
        int32_t rc = h->tile_mode ? ensure_tile_capacity(h, blocks + blocks / 8 + tile_stranded_blocks(h), tblocks + tblocks / 8 + tile_stranded_blocks(h))
                                  : ensure_list_capacity(h, blocks + blocks / 8 + isomc_count_list_max_warps(h->sms));
        if (rc) return rc;
        rc = enqueue_count(h, h->emit_inline);
        if (rc) return rc;
        h->stats.emit_reruns = reruns + 1;
    }
    return ISOMC_OK;
}

int32_t finish_impl(isomc *h) {
    if (!h->counted) return fail(h, ISOMC_ERR_NO_RESULT, "finish() without an extract in flight");
    int32_t rc = fetch_totals(h);
    if (rc) return rc;
    const uint64_t nv = h->h_totals[8], nt = h->h_totals[10];
    if (!h->tile_mode) { /* rounded up to a power of two: the figure is part of the graph key and must not change with every mesh */
        uint32_t b = 16;
        while (b < h->h_totals[7] + 8 && b < 4096) b <<= 1;
        h->last_blocks = b < 4096 ? b : 0u;
    }
    /* ids are u32 and the whole numbering (incl. a slab's ghost layer) must fit */
    if (h->h_totals[0] >= (1ull << 32) || h->h_totals[1] >= (1ull << 32))
        return fail(h, ISOMC_ERR_INDEX_OVERFLOW, "mesh has %llu vertices / %llu triangles: does not fit u32 indices",
                    (unsigned long long)h->h_totals[0], (unsigned long long)h->h_totals[1]);
    if (h->vofs_cached < 0 && h->step_exchanged && h->h_totals[14] != 0)
        return fail(h, ISOMC_ERR_NCCL, "totals exchange over peer memory timed out: rank %llu never published step %llu",
                    (unsigned long long)h->h_totals[14] - 1, (unsigned long long)h->h_totals[15]);
    if (h->vofs_cached < 0 && h->h_totals[13] + h->h_totals[0] >= (1ull << 32)) /* offset derived on the device (slab_emit_gathered) */
        return fail(h, ISOMC_ERR_INDEX_OVERFLOW, "global vertex ids of this slab reach %llu: do not fit u32 indices",
                    (unsigned long long)(h->h_totals[13] + h->h_totals[0]));
    const bool must_rerun = !h->emitted || nv > h->cap_v || nt > h->cap_t;
    if (must_rerun) {
        if (h->emitted) h->stats.emit_reruns = 1;
        rc = ensure_capacity(h, nv, nt);
        if (rc) return rc;
        rc = enqueue_emit(h);
        if (rc) return rc;
    }
    if (h->profiling) CU(h, cudaEventRecord(h->ev[4], h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    if (h->timeline && !h->tl.empty()) {
        for (auto &e : h->tl) {
            float ms = 0;
            cudaEventElapsedTime(&ms, h->tl[0].second, e.second);
            fprintf(stderr, "[isomc timeline] %-16s done at %8.3f ms\n", e.first.c_str(), ms);
            }
        for (auto &e : h->tl) cudaEventDestroy(e.second);
        h->tl.clear();
        fprintf(stderr, "[isomc timeline] %llu list blocks of %u entries for %llu listed cells (%.1f %% of the slots used)\n",
                (unsigned long long)h->h_totals[7], LIST_BLOCK, (unsigned long long)h->h_totals[2],
                h->h_totals[7] ? 100.0 * (double)h->h_totals[2] / ((double)h->h_totals[7] * LIST_BLOCK) : 0.0);
    }
    h->n_v = nv; h->n_t = nt; h->n_a = h->h_totals[11];
    h->have_result = true;
    isomc_stats &s = h->stats;
    s.n_vertices = nv; s.n_triangles = nt; s.n_active_cells = h->n_a;
    s.n_samples = (uint64_t)h->g.N * h->g.N * h->g.nsl;
    s.n_cells = (uint64_t)h->g.ncx * h->g.ncx * (h->g.ncl - h->g.ghost);
    s.algorithmic_bytes = 4 * s.n_samples + 12 * nv + 12 * nt;
    if (h->profiling && h->g.ncl && !must_rerun) {
        cudaEventElapsedTime(&s.ms_sign, h->ev[0], h->ev[1]);
        cudaEventElapsedTime(&s.ms_count, h->ev[1], h->ev[2]);
        cudaEventElapsedTime(&s.ms_scan, h->ev[2], h->ev[3]);
        cudaEventElapsedTime(&s.ms_emit, h->ev[3], h->ev[4]);
        cudaEventElapsedTime(&s.ms_total, h->ev[0], h->ev[4]);
    }
    return ISOMC_OK;
}

int32_t set_vofs(isomc *h, uint32_t v) {
    if (h->vofs_cached == (int64_t)v) return ISOMC_OK;
    CU(h, cudaMemsetAsync(h->vofs, 0, sizeof(uint32_t), h->stream));
    if (v != 0) CU(h, cudaMemcpyAsync(h->vofs, &v, sizeof v, cudaMemcpyHostToDevice, h->stream));
    h->vofs_cached = (int64_t)v;
    return ISOMC_OK;
}

int32_t create_impl(uint32_t size, uint32_t z_begin, uint32_t z_end, int32_t device, isomc_t **out, uint32_t batch = 0) {
    if (!out) return fail(nullptr, ISOMC_ERR_BAD_ARG, "out == NULL");
    *out = nullptr;
    /* size == 0 underflows in the reference (primal_grid.rs:44); size > 8192 overflows the packed row prefixes */
    if (size < 1 || size > 8192) return fail(nullptr, ISOMC_ERR_BAD_ARG, "size %u out of range [1, 8192]", size);
    if (z_begin >= z_end || z_end > size) return fail(nullptr, ISOMC_ERR_BAD_ARG, "bad slab [%u, %u) of %u", z_begin, z_end, size);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, ISOMC_ERR_CUDA, "no CUDA device: %s (libisomc_b200 has no CPU fallback)",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, ISOMC_ERR_BAD_ARG, "device %d out of range (%d devices)", device, ndev);
    isomc *h = new (std::nothrow) isomc();
    if (!h) return fail(nullptr, ISOMC_ERR_OOM, "host allocation failed");
    h->size = size; h->z_begin = z_begin; h->z_end = z_end; h->device = device;
    h->timeline = getenv("ISOMC_TIMELINE") != nullptr;
    Geo &g = h->g;
    g.N = size; g.ncx = size - 1;
    g.nsegx = (g.ncx + 31) / 32; g.nws = (g.nsegx + 2) & ~1u; /* >= nsegx + 1, even: word 2j of a row is 8-byte aligned */
    g.ghost = z_begin > 0 ? 1u : 0u;
    g.gz0 = z_begin - g.ghost;
    g.ncl = (g.ncx == 0) ? 0 : (z_end - z_begin + g.ghost);
    g.nsl = g.ncl + 1;
    g.inv = 1.0f / (float)(size - 1);
    g.row_magic = g.ncx ? ((1ull << 40) + g.ncx - 1) / g.ncx : 0;
    if (batch) { /* `batch` lattices of size^2 x (size+1) samples stacked in z; the cell layer between two of them is dead */
        if (size < 2) {
            delete h;
            return fail(nullptr, ISOMC_ERR_BAD_ARG, "batched chunks need size >= 2");
        }
        g.zper = size + 1;
        g.zmagic = geo_zmagic(g.zper);
        g.ncl = batch * (size + 1) - 1;
        g.nsl = g.ncl + 1;
        if ((uint64_t)g.ncl * g.ncx >= (1ull << 26) || batch > 65535u) {
            delete h;
            return fail(nullptr, ISOMC_ERR_BAD_ARG, "batch of %u chunks of size %u is too large for one handle (%llu cell rows, limit 2^26)", batch,
                        size, (unsigned long long)g.ncl * g.ncx);
        }
        h->batch = batch;
    }
    int32_t rc = ISOMC_OK;
    auto body = [&]() -> int32_t {
        CU(h, cudaSetDevice(device));
        cudaDeviceProp prop;
        CU(h, cudaGetDeviceProperties(&prop, device));
        h->sms = prop.multiProcessorCount;
        CU(h, cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
        h->stream = h->own_stream;
        for (auto &ev : h->ev) CU(h, cudaEventCreate(&ev));
        McTables host_tabs;
        if (isomc_build_tables(&host_tabs)) return fail(h, ISOMC_ERR_BAD_ARG, "case table is inconsistent");
        CU(h, cudaMalloc(&h->tabs, sizeof(McTables)));
        CU(h, cudaMemcpy(h->tabs, &host_tabs, sizeof(McTables), cudaMemcpyHostToDevice));
        const uint64_t nrows_s = (uint64_t)g.nsl * g.N, nrows_c = (uint64_t)g.ncl * g.ncx;
        if (const char *p = getenv("ISOMC_PATH")) h->tile_mode = strcmp(p, "tile") == 0;
        if (h->batch) h->tile_mode = false; /* (stacked lattices are served by the list kernels) */
        h->tg = tile_geo(g);
        EmitTab host_etab;
        isomc_build_emit_tab(host_tabs, &host_etab);
        isomc_tile_fill_eloc(host_etab.eloc);
        CU(h, cudaMalloc(&h->etab, sizeof(EmitTab)));
        CU(h, cudaMemcpy(h->etab, &host_etab, sizeof(EmitTab), cudaMemcpyHostToDevice));
        if (h->tile_mode) {
            /* per row piece: counts -> prefixes (rowV / rowT), entry and t positions, active cells */
            const uint64_t np = (uint64_t)g.ncl * h->tg.ppl;
            CU(h, cudaMalloc(&h->rowV, (np + 4) * sizeof(uint32_t)));
            CU(h, cudaMalloc(&h->rowT, (np + 4) * sizeof(uint32_t)));
            CU(h, cudaMalloc(&h->TB.pE, (np + 4) * sizeof(uint32_t)));
            CU(h, cudaMalloc(&h->TB.pTp, (np + 4) * sizeof(uint32_t)));
            CU(h, cudaMalloc(&h->TB.pA, (np + 4) * sizeof(uint16_t)));
            h->TB.pV = h->rowV; h->TB.pT = h->rowT;
            /* first guess: 1/32 of the cells active, plus the blocks the counting warps may strand */
            const uint64_t guess = nrows_c * g.ncx / 32 / ENT_BLOCK + tile_stranded_blocks(h);
            int32_t lrc = ensure_tile_capacity(h, guess, guess);
            if (lrc) return lrc;
        } else {
            CU(h, cudaMalloc(&h->signs, (nrows_s * g.nws + 4) * sizeof(uint32_t)));
            CU(h, cudaMalloc(&h->L.segrec, (nrows_c * g.nsegx + 4) * sizeof(uint2)));
            CU(h, cudaMalloc(&h->L.segtpre, (nrows_c * g.nsegx + 4) * sizeof(uint32_t)));
            /* first guess: 1/32 of the cells active, plus the block every counting warp may strand */
            int32_t lrc = ensure_list_capacity(h, nrows_c * g.ncx / 32 / LIST_BLOCK + isomc_count_list_max_warps(h->sms) + 16);
            if (lrc) return lrc;
            /* (rowV / rowT / rowA live behind the per-layer totals, see below: one memset per extract zeroes all of it) */
        }
        /* per-layer totals followed by the emit tickets: zeroed by a single memset per extract */
        /* list path: the per-row counts follow in the same allocation -- the counting kernel only writes the rows that have active
         * cells, the others keep the zeros of the memset */
        h->zero_bytes = ((size_t)g.ncl * 3 + 4) * sizeof(unsigned long long) + AUX_WORDS * sizeof(uint32_t);
        const size_t row_words = h->tile_mode ? 0 : (size_t)((nrows_c + 4 + 3) & ~3ull);
        CU(h, cudaMalloc(&h->layerTot, h->zero_bytes + 3 * row_words * sizeof(uint32_t)));
        h->ticket = reinterpret_cast<uint32_t *>(h->layerTot + ((size_t)g.ncl * 3 + 4));
        if (!h->tile_mode) {
            h->rowV = h->ticket + AUX_WORDS;
            h->rowT = h->rowV + row_words;
            h->rowA = h->rowT + row_words;
            h->zero_bytes += 3 * row_words * sizeof(uint32_t);
        }
        h->L.ctr = h->ticket + MAX_CHUNKS;
        h->list_marks = h->ticket + MAX_CHUNKS + 1;
        h->chunk_ends = h->ticket + 2 * MAX_CHUNKS + 2;
        h->TB.ctr = h->ticket + 4 * MAX_CHUNKS + 4;
        h->tile_tickets = h->ticket + 4 * MAX_CHUNKS + 8;
        h->TB.layerTot = h->layerTot;
        CU(h, cudaMalloc(&h->totals, (N_TOTALS + 3 * (size_t)h->batch) * sizeof(unsigned long long)));
        CU(h, cudaMemset(h->totals, 0, (N_TOTALS + 3 * (size_t)h->batch) * sizeof(unsigned long long)));
        if (h->batch) {
            CU(h, cudaMalloc(&h->d_progs, h->batch * sizeof(SdfProgram)));
            CU(h, cudaMallocHost(&h->h_progs, h->batch * sizeof(SdfProgram)));
            CU(h, cudaMalloc(&h->chunkV, (h->batch + 1) * sizeof(uint32_t)));
            CU(h, cudaMalloc(&h->chunkT, (h->batch + 1) * sizeof(uint32_t)));
            CU(h, cudaMallocHost(&h->h_chunk, 2 * (h->batch + 1) * sizeof(uint32_t)));
            h->L.chunkV = h->chunkV; h->L.chunkT = h->chunkT;
        }
        CU(h, cudaMalloc(&h->vofs, sizeof(uint32_t)));
        CU(h, cudaMemset(h->vofs, 0, sizeof(uint32_t)));
        CU(h, cudaMallocHost(&h->h_totals, N_TOTALS * sizeof(unsigned long long)));
        CU(h, cudaMallocHost(&h->h_chunk_ends, 2 * MAX_CHUNKS * sizeof(uint32_t)));
        return ISOMC_OK;
    };
    rc = body();
    if (rc) {
        g_create_error = h->err;
        isomc_destroy(h);
        return rc;
    }
    *out = h;
    return ISOMC_OK;
}

}  // namespace

extern "C" {

const char *isomc_version(void) { return "isomc_b200 0.2.0 (sm_100a)"; }

int32_t isomc_device_count(void) {
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

int32_t isomc_create(uint32_t size, int32_t device, isomc_t **out) { return create_impl(size, 0, size, device, out); }

int32_t isomc_slab_create(uint32_t size, uint32_t z_begin, uint32_t z_end, int32_t device, isomc_t **out) {
    return create_impl(size, z_begin, z_end, device, out);
}

int32_t isomc_destroy(isomc_t *h) {
    if (!h) return ISOMC_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (void *p : h->ipc_opened) cudaIpcCloseMemHandle(p);
    cudaFree(h->mailbox); cudaFree(h->d_peers);
    if (h->pipe_exec) cudaGraphExecDestroy(h->pipe_exec);
    if (h->s_sign) {
        cudaStreamSynchronize(h->s_sign); cudaStreamSynchronize(h->s_emit);
        cudaStreamDestroy(h->s_sign); cudaStreamDestroy(h->s_emit);
        cudaEventDestroy(h->ev_pipe0); cudaEventDestroy(h->ev_pipe1);
        for (int c = 0; c < MAX_CHUNKS; ++c) { cudaEventDestroy(h->ev_sign[c]); cudaEventDestroy(h->ev_cnt2[c]); }
    }
    cudaFree(h->d_progs); cudaFree(h->chunkV); cudaFree(h->chunkT);
    if (h->h_progs) cudaFreeHost(h->h_progs);
    if (h->h_chunk) cudaFreeHost(h->h_chunk);
    cudaFree(h->signs); cudaFree(h->segA);
    if (h->tile_mode) { cudaFree(h->rowV); cudaFree(h->rowT); cudaFree(h->rowA); } /* (list path: part of the layerTot allocation) */
    cudaFree(h->TB.pE); cudaFree(h->TB.pTp); cudaFree(h->TB.pA); cudaFree(h->TB.ent); cudaFree(h->TB.tq); cudaFree(h->TB.tbuf);
    cudaFree(h->layerTot); cudaFree(h->totals); cudaFree(h->vofs); cudaFree(h->tabs);
    cudaFree(h->xyz); cudaFree(h->idx); cudaFree(h->stage_grid);
    cudaFree(h->L.ent); cudaFree(h->L.ent_yz); cudaFree(h->L.segrec); cudaFree(h->L.segtpre); cudaFree(h->L.blkfill); cudaFree(h->etab);
    if (h->h_totals) cudaFreeHost(h->h_totals);
    if (h->h_chunk_ends) cudaFreeHost(h->h_chunk_ends);
    for (int c = 0; c < MAX_CHUNKS; ++c) {
        if (h->ev_in[c]) cudaEventDestroy(h->ev_in[c]);
        if (h->ev_cnt[c]) cudaEventDestroy(h->ev_cnt[c]);
        if (h->ev_emit[c]) cudaEventDestroy(h->ev_emit[c]);
    }
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    for (auto &ev : h->ev) if (ev) cudaEventDestroy(ev);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return ISOMC_OK;
}

const char *isomc_last_error(const isomc_t *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int32_t isomc_get_stream(isomc_t *h, void **cuda_stream) {
    if (!h || !cuda_stream) return ISOMC_ERR_BAD_ARG;
    *cuda_stream = (void *)h->stream;
    return ISOMC_OK;
}

int32_t isomc_set_stream(isomc_t *h, void *cuda_stream) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    int32_t rc = bind_device(h);
    if (rc) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    return ISOMC_OK;
}

int32_t isomc_set_profiling(isomc_t *h, int32_t on) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    h->profiling = on != 0;
    return ISOMC_OK;
}

int32_t isomc_reserve(isomc_t *h, uint64_t n_vertices, uint64_t n_triangles) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    int32_t rc = bind_device(h);
    if (rc) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    return ensure_capacity(h, n_vertices, n_triangles);
}

/* ---- full extracts ------------------------------------------------------------------------ */

static int32_t enqueue_full(isomc_t *h) {
    if (h->batch && !h->batch_call)
        return fail(h, ISOMC_ERR_BAD_ARG, "this handle is a batch of %u chunks: use isomc_extract_sdf_batch / isomc_extract_grid_batch_*", h->batch);
    if (h->z_begin != 0 || h->z_end != h->size)
        return fail(h, ISOMC_ERR_BAD_ARG, "this handle is a slab [%u, %u): use the isomc_slab_* calls", h->z_begin, h->z_end);
    int32_t rc = bind_device(h);
    if (rc) return rc;
    rc = set_vofs(h, 0);
    if (rc) return rc;
    /* optimistic emission into the buffers of the previous extract; finish() re-runs it if they are too small */
    return enqueue_count(h, h->cap_v > 0 || h->cap_t > 0);
}

int32_t isomc_enqueue_grid_device(isomc_t *h, const float *d_grid) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    if (!d_grid) return fail(h, ISOMC_ERR_BAD_ARG, "d_grid == NULL");
    h->kind = SRC_GRID; h->d_grid = d_grid; h->directed = false;
    return enqueue_full(h);
}

int32_t isomc_enqueue_sdf(isomc_t *h, const isomc_sdf_node *prog, uint32_t n_nodes) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    int32_t rc = validate_program(h, prog, n_nodes, &h->prog);
    if (rc) return rc;
    h->kind = SRC_SDF; h->d_grid = nullptr; h->directed = false;
    return enqueue_full(h);
}

/* MarchingCubes::<Directed>::new(size).extract(&Sampler::new(&implicit_tree), ..)  (reference src/distance.rs:72-104) */
int32_t isomc_enqueue_sdf_directed(isomc_t *h, const isomc_sdf_node *prog, uint32_t n_nodes) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    int32_t rc = validate_program(h, prog, n_nodes, &h->prog);
    if (rc) return rc;
    h->kind = SRC_SDF; h->d_grid = nullptr; h->directed = true;
    return enqueue_full(h);
}

int32_t isomc_extract_sdf_directed(isomc_t *h, const isomc_sdf_node *prog, uint32_t n_nodes) {
    int32_t rc = isomc_enqueue_sdf_directed(h, prog, n_nodes);
    return rc ? rc : isomc_finish(h);
}

int32_t isomc_finish(isomc_t *h) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    int32_t rc = bind_device(h);
    if (rc) return rc;
    return finish_impl(h);
}

int32_t isomc_extract_grid_device(isomc_t *h, const float *d_grid) {
    int32_t rc = isomc_enqueue_grid_device(h, d_grid);
    return rc ? rc : isomc_finish(h);
}

int32_t isomc_extract_sdf(isomc_t *h, const isomc_sdf_node *prog, uint32_t n_nodes) {
    int32_t rc = isomc_enqueue_sdf(h, prog, n_nodes);
    return rc ? rc : isomc_finish(h);
}

int32_t isomc_extract_grid_host(isomc_t *h, const float *h_grid) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    if (!h_grid) return fail(h, ISOMC_ERR_BAD_ARG, "h_grid == NULL");
    int32_t rc = bind_device(h);
    if (rc) return rc;
    const size_t bytes = (size_t)h->g.N * h->g.N * h->g.nsl * sizeof(float);
    if (!h->stage_grid) CU(h, cudaMalloc(&h->stage_grid, bytes > 0 ? bytes : 4));
    CU(h, cudaMemcpyAsync(h->stage_grid, h_grid, bytes, cudaMemcpyHostToDevice, h->stream));
    return isomc_extract_grid_device(h, h->stage_grid);
}

/*
 * Host lattice in, host mesh out, pipelined in z-chunks (the row scan is causal in z, so a chunk's part of the mesh
 * is final as soon as the chunk is emitted):
 *
 *   s_in   : H2D(0) H2D(1) H2D(2) ...
 *   stream :        sign/count/scan/emit(0)  sign/count/scan/emit(1) ...
 *   s_out  :                                 D2H(mesh part 0)        D2H(mesh part 1) ...
 *
 * PCIe is full duplex, so the copy-out of the mesh hides behind the copy-in of the lattice; the kernels hide behind
 * both.  The host only waits for one small event per chunk (how many vertices / triangles lie below the chunk's end).
 * Needs output buffers from an earlier extract (optimistic emission); the first extract of a handle takes the plain path.
 */
int32_t isomc_extract_grid_host_to(isomc_t *h, const float *h_grid, float *xyz, uint64_t cap_vertices, uint32_t *idx,
                                   uint64_t cap_triangles) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    if (!h_grid) return fail(h, ISOMC_ERR_BAD_ARG, "h_grid == NULL");
    if (h->z_begin != 0 || h->z_end != h->size)
        return fail(h, ISOMC_ERR_BAD_ARG, "this handle is a slab [%u, %u): use the isomc_slab_* calls", h->z_begin, h->z_end);
    int32_t rc = bind_device(h);
    if (rc) return rc;
    const Geo &g = h->g;
    const size_t bytes = (size_t)g.N * g.N * g.nsl * sizeof(float);
    if (!h->stage_grid) CU(h, cudaMalloc(&h->stage_grid, bytes > 0 ? bytes : 4));
    const bool streamed = g.ncl >= 32 && h->cap_v > 0 && h->cap_t > 0 && !h->profiling;
    bool delivered = false;
    if (!streamed) {
        CU(h, cudaMemcpyAsync(h->stage_grid, h_grid, bytes, cudaMemcpyHostToDevice, h->stream));
        rc = isomc_extract_grid_device(h, h->stage_grid);
        if (rc) return rc;
    } else {
        if (!h->s_in) {
            CU(h, cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
            CU(h, cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
            for (int c = 0; c < MAX_CHUNKS; ++c) {
                CU(h, cudaEventCreateWithFlags(&h->ev_in[c], cudaEventDisableTiming));
                CU(h, cudaEventCreateWithFlags(&h->ev_cnt[c], cudaEventDisableTiming));
                CU(h, cudaEventCreateWithFlags(&h->ev_emit[c], cudaEventDisableTiming));
            }
        }
        rc = set_vofs(h, 0);
        if (rc) return rc;
        h->kind = SRC_GRID; h->d_grid = h->stage_grid; h->directed = false;
        h->have_result = false; h->counted = false; h->emitted = false; h->totals_valid = false;
        h->stats.kernel_launches = 0; h->stats.emit_reruns = 0;
        h->emit_inline = true;
        /* chunk plan: up to MAX_CHUNKS chunks of >= 8 cell layers.  The copy-in of the lattice is what the call waits for, so what
         * is left to do when its last byte has arrived -- the kernels and the copy-out of the LAST chunk -- is the only part
         * that is not hidden: the last chunks are made small (8, 16, ... layers), the ones before share the rest evenly. */
        {
            uint32_t n = g.ncl / 8 < (uint32_t)MAX_CHUNKS ? g.ncl / 8 : (uint32_t)MAX_CHUNKS;
            if (n < 1) n = 1;
            uint32_t tail[MAX_CHUNKS], n_tail = 0, tail_sum = 0;
            static const bool graded = !(getenv("ISOMC_HOST_TAIL") && atoi(getenv("ISOMC_HOST_TAIL")) == 0);
            for (uint32_t t = 8; graded && n_tail + 2 < n && t < g.ncl / n && tail_sum + t < g.ncl / 4; t *= 2) { tail[n_tail++] = t; tail_sum += t; }
            const uint32_t n_head = n - n_tail, head = g.ncl - tail_sum, per = (head + n_head - 1) / n_head;
            uint32_t c = 0, l = 0;
            h->chunk_l[0] = 0;
            while (l < head) { l = l + per < head ? l + per : head; h->chunk_l[++c] = l; }
            for (uint32_t i = n_tail; i-- > 0;) { l += tail[i]; h->chunk_l[++c] = l; }
            h->n_chunks = c;
        }
        const uint64_t cap_v0 = h->cap_v, cap_t0 = h->cap_t;
        CU(h, cudaMemsetAsync(h->layerTot, 0, h->zero_bytes, h->stream));
        for (uint32_t c = 0; c < h->n_chunks; ++c) {
            const uint64_t row0 = (uint64_t)(c == 0 ? 0u : h->chunk_l[c] + 1) * g.N, row1 = (uint64_t)(h->chunk_l[c + 1] + 1) * g.N;
            CU(h, cudaMemcpyAsync(h->stage_grid + row0 * g.N, h_grid + row0 * g.N, (row1 - row0) * g.N * sizeof(float),
                                  cudaMemcpyHostToDevice, h->s_in));
            CU(h, cudaEventRecord(h->ev_in[c], h->s_in));
        }
        for (uint32_t c = 0; c < h->n_chunks; ++c) {
            cudaStream_t st = h->stream;
            CU(h, cudaStreamWaitEvent(st, h->ev_in[c], 0));
            rc = launch_count_chunk(h, c, st, h->chunk_ends + 2 * c);
            if (rc) return rc;
            CU(h, cudaMemcpyAsync(h->h_chunk_ends + 2 * c, h->chunk_ends + 2 * c, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            CU(h, cudaEventRecord(h->ev_cnt[c], st));
            rc = launch_emit_chunk(h, c, st);
            if (rc) return rc;
            CU(h, cudaEventRecord(h->ev_emit[c], st));
        }
        h->counted = true;
        h->emitted = true;
        /* copy-out of each chunk's part of the mesh as soon as its size is known and it is emitted */
        bool fits = true;
        uint64_t v_prev = 0, t_prev = 0;
        for (uint32_t c = 0; c < h->n_chunks && fits; ++c) {
            CU(h, cudaEventSynchronize(h->ev_cnt[c]));
            const uint64_t v_end = h->h_chunk_ends[2 * c], t_end = h->h_chunk_ends[2 * c + 1];
            if (v_end > cap_v0 || t_end > cap_t0 || v_end > cap_vertices || t_end > cap_triangles || !xyz || !idx) { fits = false; break; }
            CU(h, cudaStreamWaitEvent(h->s_out, h->ev_emit[c], 0));
            if (v_end > v_prev) CU(h, cudaMemcpyAsync(xyz + 3 * v_prev, h->xyz + 3 * v_prev, (v_end - v_prev) * 12, cudaMemcpyDeviceToHost, h->s_out));
            if (t_end > t_prev) CU(h, cudaMemcpyAsync(idx + 3 * t_prev, h->idx + 3 * t_prev, (t_end - t_prev) * 12, cudaMemcpyDeviceToHost, h->s_out));
            v_prev = v_end; t_prev = t_end;
        }
        CU(h, cudaStreamSynchronize(h->s_out));
        rc = finish_impl(h); /* synchronises; grows the list / the output buffers and re-runs what is needed */
        if (rc) return rc;
        delivered = fits && h->stats.emit_reruns == 0;
    }
    if (h->n_v > cap_vertices || h->n_t > cap_triangles || (h->n_v && !xyz) || (h->n_t && !idx))
        return fail(h, ISOMC_ERR_BUFFER_TOO_SMALL, "mesh has %llu vertices / %llu triangles, the caller's buffers hold %llu / %llu "
                    "(the result stays on the device: grow and call isomc_copy_out)", (unsigned long long)h->n_v,
                    (unsigned long long)h->n_t, (unsigned long long)cap_vertices, (unsigned long long)cap_triangles);
    if (!delivered) return isomc_copy_out(h, xyz, idx);
    return ISOMC_OK;
}

/* ---- batched chunks: the crate's usage model is many size^3 chunks (src/marching_cubes.rs:44-45, README.md:19),
 * each one `MarchingCubes::new(size).extract(&Sampler::new(&tree_b), ..)` with its own (translated) tree.  A batch handle runs
 * all of them as ONE sign / count / scan / emit sequence over stacked lattices, with one size read-back. */
int32_t isomc_batch_create(uint32_t size, uint32_t n_chunks, int32_t device, isomc_t **out) {
    if (n_chunks < 1) return fail(nullptr, ISOMC_ERR_BAD_ARG, "n_chunks == 0");
    return create_impl(size, 0, size, device, out, n_chunks);
}

static int32_t extract_sdf_batch_impl(isomc_t *h, const isomc_sdf_node *progs, const uint32_t *n_nodes, uint32_t n_chunks, bool directed) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    if (!h->batch) return fail(h, ISOMC_ERR_BAD_ARG, "not a batch handle (isomc_batch_create)");
    if (!progs || !n_nodes || n_chunks < 1 || n_chunks > h->batch)
        return fail(h, ISOMC_ERR_BAD_ARG, "bad batch arguments (%u chunks, the handle holds %u)", n_chunks, h->batch);
    int32_t rc = bind_device(h);
    if (rc) return rc;
    CU(h, cudaStreamSynchronize(h->stream)); /* (the pinned program staging buffer may still be in flight) */
    const isomc_sdf_node *p = progs;
    for (uint32_t b = 0; b < h->batch; ++b) {
        if (b < n_chunks) {
            rc = validate_program(h, p, n_nodes[b], &h->h_progs[b]);
            if (rc) return rc;
            p += n_nodes[b];
        } else { /* unused lattice: a unit sphere far below the origin -- positive everywhere, in every component too */
            static const isomc_sdf_node pad[3] = {{ISOMC_SDF_TRANSLATE_PUSH, -10.0f, -10.0f, -10.0f}, {ISOMC_SDF_SPHERE, 1.0f, 0.0f, 0.0f},
                                                  {ISOMC_SDF_TRANSLATE_POP, 0.0f, 0.0f, 0.0f}};
            rc = validate_program(h, pad, 3, &h->h_progs[b]);
            if (rc) return rc;
        }
    }
    CU(h, cudaMemcpyAsync(h->d_progs, h->h_progs, h->batch * sizeof(SdfProgram), cudaMemcpyHostToDevice, h->stream));
    h->kind = SRC_SDF_BATCH; h->d_grid = nullptr; h->directed = directed;
    h->batch_used = n_chunks;
    h->batch_call = true;
    rc = enqueue_full(h);
    h->batch_call = false;
    return rc ? rc : isomc_finish(h);
}

int32_t isomc_extract_sdf_batch(isomc_t *h, const isomc_sdf_node *progs, const uint32_t *n_nodes, uint32_t n_chunks) {
    return extract_sdf_batch_impl(h, progs, n_nodes, n_chunks, false);
}
/* MarchingCubes<Directed> per chunk (isomc_extract_sdf_directed) */
int32_t isomc_extract_sdf_batch_directed(isomc_t *h, const isomc_sdf_node *progs, const uint32_t *n_nodes, uint32_t n_chunks) {
    return extract_sdf_batch_impl(h, progs, n_nodes, n_chunks, true);
}

/* the same for dense chunks: the handle's `n_chunks` lattices (size * size * (size + 1) f32 each) back to back in device memory --
 * which IS the stacked lattice the kernels walk, so nothing is copied */
int32_t isomc_extract_grid_batch_device(isomc_t *h, const float *d_lattices, uint32_t n_chunks) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    if (!h->batch) return fail(h, ISOMC_ERR_BAD_ARG, "not a batch handle (isomc_batch_create)");
    if (!d_lattices) return fail(h, ISOMC_ERR_BAD_ARG, "d_lattices == NULL");
    if (n_chunks != h->batch)
        return fail(h, ISOMC_ERR_BAD_ARG, "a device-resident batch must fill the handle: %u lattices given, the handle holds %u", n_chunks, h->batch);
    h->kind = SRC_GRID; h->d_grid = d_lattices; h->directed = false;
    h->batch_used = n_chunks;
    h->batch_call = true;
    int32_t rc = enqueue_full(h);
    h->batch_call = false;
    return rc ? rc : isomc_finish(h);
}

/* host lattices: copied into the handle's staging buffer; lattices the call does not fill are padded with a positive value (no surface) */
int32_t isomc_extract_grid_batch_host(isomc_t *h, const float *h_lattices, uint32_t n_chunks) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    if (!h->batch) return fail(h, ISOMC_ERR_BAD_ARG, "not a batch handle (isomc_batch_create)");
    if (!h_lattices || n_chunks < 1 || n_chunks > h->batch)
        return fail(h, ISOMC_ERR_BAD_ARG, "bad batch arguments (%u chunks, the handle holds %u)", n_chunks, h->batch);
    int32_t rc = bind_device(h);
    if (rc) return rc;
    const size_t per = (size_t)h->g.N * h->g.N * (h->g.N + 1) * sizeof(float), bytes = per * h->batch;
    if (!h->stage_grid) CU(h, cudaMalloc(&h->stage_grid, bytes));
    CU(h, cudaMemcpyAsync(h->stage_grid, h_lattices, per * n_chunks, cudaMemcpyHostToDevice, h->stream));
    if (n_chunks < h->batch) /* 0x7F7F7F7F = 3.4e38f: outside everywhere */
        CU(h, cudaMemsetAsync(reinterpret_cast<char *>(h->stage_grid) + per * n_chunks, 0x7F, per * (h->batch - n_chunks), h->stream));
    h->kind = SRC_GRID; h->d_grid = h->stage_grid; h->directed = false;
    h->batch_used = n_chunks;
    h->batch_call = true;
    rc = enqueue_full(h);
    h->batch_call = false;
    return rc ? rc : isomc_finish(h);
}

int32_t isomc_batch_offsets(isomc_t *h, uint64_t *v_offsets, uint64_t *t_offsets) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    if (!h->batch) return fail(h, ISOMC_ERR_BAD_ARG, "not a batch handle (isomc_batch_create)");
    if (!h->have_result) return fail(h, ISOMC_ERR_NO_RESULT, "no extract has completed on this handle");
    int32_t rc = bind_device(h);
    if (rc) return rc;
    const uint32_t n = h->batch + 1;
    CU(h, cudaMemcpyAsync(h->h_chunk, h->chunkV, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(h->h_chunk + n, h->chunkT, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    for (uint32_t b = 0; b <= h->batch_used; ++b) {
        if (v_offsets) v_offsets[b] = h->h_chunk[b];
        if (t_offsets) t_offsets[b] = h->h_chunk[n + b];
    }
    return ISOMC_OK;
}

/* ---- PointCloud::new(size).extract(&source, &mut extractor)  (reference src/point_cloud.rs:50-63) -------- */

static int32_t points_impl(isomc_t *h) {
    if (h->batch) return fail(h, ISOMC_ERR_BAD_ARG, "this handle is a batch of chunks");
    if (h->z_begin != 0 || h->z_end != h->size)
        return fail(h, ISOMC_ERR_BAD_ARG, "point clouds are extracted on whole-lattice handles (this one is a slab [%u, %u))", h->z_begin, h->z_end);
    int32_t rc = bind_device(h);
    if (rc) return rc;
    const Geo &g = h->g;
    h->have_result = false; h->counted = false; h->emitted = false; h->totals_valid = false;
    h->stats.kernel_launches = 0; h->stats.emit_reruns = 0;
    if (g.ncl == 0 || g.ncx == 0) {
        h->n_v = h->n_t = h->n_a = 0;
        h->have_result = true;
        return ISOMC_OK;
    }
    rc = ensure_signs(h);
    if (rc) return rc;
    uint32_t *segA = h->segA;
    CU(h, cudaMemsetAsync(h->layerTot, 0, h->zero_bytes, h->stream));
    if (h->kind == SRC_GRID) CU(h, isomc_launch_sign_grid(g, h->d_grid, h->signs, 0, g.nsl * g.N, h->sms, 8, h->stream));
    else CU(h, isomc_launch_sign_sdf(g, h->prog, h->directed, h->signs, 0, g.nsl * g.N, h->sms, 8, h->stream));
    CU(h, isomc_launch_points_count(g, h->signs, segA, h->rowV, h->rowT, h->layerTot, h->sms, h->stream));
    CU(h, isomc_launch_scan(g, g.ncx, h->rowV, h->rowT, h->layerTot, h->totals, nullptr, nullptr, nullptr, 0, g.ncl, h->stream));
    CU(h, cudaMemcpyAsync(h->h_totals, h->totals, N_TOTALS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    const uint64_t np = h->h_totals[0];
    if (np >= (1ull << 32)) return fail(h, ISOMC_ERR_INDEX_OVERFLOW, "%llu points", (unsigned long long)np);
    rc = ensure_capacity(h, np, 0);
    if (rc) return rc;
    CU(h, isomc_launch_points_emit(g, h->signs, segA, h->rowV, h->xyz, h->cap_v, h->sms, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    h->stats.kernel_launches = 4;
    h->n_v = np; h->n_t = 0; h->n_a = np;
    h->stats.n_vertices = np; h->stats.n_triangles = 0; h->stats.n_active_cells = np;
    h->stats.n_samples = (uint64_t)g.N * g.N * g.nsl;
    h->stats.n_cells = (uint64_t)g.ncx * g.ncx * g.ncl;
    h->stats.algorithmic_bytes = 4 * h->stats.n_samples + 12 * np;
    h->have_result = true;
    return ISOMC_OK;
}

int32_t isomc_points_grid_device(isomc_t *h, const float *d_grid) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    if (!d_grid) return fail(h, ISOMC_ERR_BAD_ARG, "d_grid == NULL");
    h->kind = SRC_GRID; h->d_grid = d_grid; h->directed = false;
    return points_impl(h);
}

int32_t isomc_points_sdf(isomc_t *h, const isomc_sdf_node *prog, uint32_t n_nodes) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    int32_t rc = validate_program(h, prog, n_nodes, &h->prog);
    if (rc) return rc;
    h->kind = SRC_SDF; h->d_grid = nullptr; h->directed = false;
    return points_impl(h);
}

/* PointCloud::<Directed>: the tree sampled as Directed distances (a cell corner is outside iff any component is positive) */
int32_t isomc_points_sdf_directed(isomc_t *h, const isomc_sdf_node *prog, uint32_t n_nodes) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    int32_t rc = validate_program(h, prog, n_nodes, &h->prog);
    if (rc) return rc;
    h->kind = SRC_SDF; h->d_grid = nullptr; h->directed = true;
    return points_impl(h);
}

int32_t isomc_points_grid_host(isomc_t *h, const float *h_grid) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    if (!h_grid) return fail(h, ISOMC_ERR_BAD_ARG, "h_grid == NULL");
    int32_t rc = bind_device(h);
    if (rc) return rc;
    const size_t bytes = (size_t)h->g.N * h->g.N * h->g.nsl * sizeof(float);
    if (!h->stage_grid) CU(h, cudaMalloc(&h->stage_grid, bytes > 0 ? bytes : 4));
    CU(h, cudaMemcpyAsync(h->stage_grid, h_grid, bytes, cudaMemcpyHostToDevice, h->stream));
    return isomc_points_grid_device(h, h->stage_grid);
}

/* ---- results ------------------------------------------------------------------------------ */

int32_t isomc_counts(isomc_t *h, uint64_t *n_vertices, uint64_t *n_triangles, uint64_t *n_active_cells) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    if (!h->have_result) return fail(h, ISOMC_ERR_NO_RESULT, "no extract has completed on this handle");
    if (n_vertices) *n_vertices = h->n_v;
    if (n_triangles) *n_triangles = h->n_t;
    if (n_active_cells) *n_active_cells = h->n_a;
    return ISOMC_OK;
}

/* per cell layer of the last extract: {vertices created, triangles, active cells}; a slab reports its own layers only */
int32_t isomc_layer_counts(isomc_t *h, uint64_t *counts) {
    if (!h || !counts) return ISOMC_ERR_BAD_ARG;
    if (!h->have_result) return fail(h, ISOMC_ERR_NO_RESULT, "no extract has completed on this handle");
    if (h->tile_mode || h->batch) return fail(h, ISOMC_ERR_BAD_ARG, "per-layer counts are kept by whole-lattice and slab handles of the list path");
    int32_t rc = bind_device(h);
    if (rc) return rc;
    const uint32_t n = h->g.ncl - h->g.ghost;
    if (n == 0) return ISOMC_OK;
    CU(h, cudaMemcpyAsync(counts, h->layerTot + 3 * (size_t)h->g.ghost, 3 * (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return ISOMC_OK;
}

int32_t isomc_device_buffers(isomc_t *h, const float **d_xyz, const uint32_t **d_idx) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    if (!h->have_result) return fail(h, ISOMC_ERR_NO_RESULT, "no extract has completed on this handle");
    if (d_xyz) *d_xyz = h->xyz;
    if (d_idx) *d_idx = h->idx;
    return ISOMC_OK;
}

int32_t isomc_copy_out(isomc_t *h, float *xyz, uint32_t *idx) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    if (!h->have_result) return fail(h, ISOMC_ERR_NO_RESULT, "no extract has completed on this handle");
    int32_t rc = bind_device(h);
    if (rc) return rc;
    if (xyz && h->n_v) CU(h, cudaMemcpyAsync(xyz, h->xyz, h->n_v * 12, cudaMemcpyDeviceToHost, h->stream));
    if (idx && h->n_t) CU(h, cudaMemcpyAsync(idx, h->idx, h->n_t * 12, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return ISOMC_OK;
}

/* IndexedInterleavedNormals with a CentralDifference source (reference src/extractor.rs:95-127, src/source.rs:82-94) */
int32_t isomc_copy_out_interleaved_normals(isomc_t *h, const isomc_sdf_node *prog, uint32_t n_nodes, float epsilon, float *xyzn,
                                           uint32_t *idx) {
    return isomc_copy_out_interleaved_normals_at(h, prog, n_nodes, epsilon, 0xFFFFFFFFu, xyzn, idx);
}

int32_t isomc_copy_out_interleaved_normals_at(isomc_t *h, const isomc_sdf_node *prog, uint32_t n_nodes, float epsilon,
                                              uint32_t n_outer_translations, float *xyzn, uint32_t *idx) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    if (!h->have_result) return fail(h, ISOMC_ERR_NO_RESULT, "no extract has completed on this handle");
    if (!(epsilon > 0.0f)) return fail(h, ISOMC_ERR_BAD_ARG, "epsilon must be positive");
    SdfProgram P;
    int32_t rc = validate_program(h, prog, n_nodes, &P);
    if (rc) return rc;
    rc = bind_device(h);
    if (rc) return rc;
    /* translations that enclose the whole program AND lie outside the CentralDifference adaptor are applied to the vertex
     * first (DemoSource around CentralDifference, examples/common/sources.rs:55-60); the differences are taken on what is inside */
    float offsets[ISOMC_SDF_MAX_TRANSLATE][3];
    uint32_t n_off = 0, lo = 0, hi = P.n;
    while (n_off < n_outer_translations && hi - lo >= 3 && P.nodes[lo].op == ISOMC_SDF_TRANSLATE_PUSH && P.nodes[hi - 1].op == ISOMC_SDF_TRANSLATE_POP) {
        int depth = 0;
        bool encloses = true;
        for (uint32_t i = lo; i < hi; ++i) {
            if (P.nodes[i].op == ISOMC_SDF_TRANSLATE_PUSH) ++depth;
            if (P.nodes[i].op == ISOMC_SDF_TRANSLATE_POP && --depth == 0 && i != hi - 1) { encloses = false; break; }
        }
        if (!encloses) break;
        offsets[n_off][0] = P.nodes[lo].a; offsets[n_off][1] = P.nodes[lo].b; offsets[n_off][2] = P.nodes[lo].c;
        ++n_off; ++lo; --hi;
    }
    if (n_outer_translations != 0xFFFFFFFFu && n_off != n_outer_translations)
        return fail(h, ISOMC_ERR_BAD_ARG, "the program is enclosed by %u translations, %u were said to lie outside the CentralDifference", n_off, n_outer_translations);
    SdfProgram inner;
    memset(&inner, 0, sizeof inner);
    inner.n = hi - lo;
    memcpy(inner.nodes, P.nodes + lo, inner.n * sizeof(isomc_sdf_node));
    if (h->n_v) {
        if (!xyzn) return fail(h, ISOMC_ERR_BAD_ARG, "xyzn == NULL");
        float *d_out = nullptr;
        CU(h, cudaMalloc(&d_out, h->n_v * 24));
        cudaError_t e = isomc_launch_normals_cd(inner, offsets, n_off, epsilon, h->xyz, h->n_v, d_out, h->sms, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(xyzn, d_out, h->n_v * 24, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        cudaFree(d_out);
        if (e != cudaSuccess) return fail(h, ISOMC_ERR_CUDA, "normal sampling failed: %s", cudaGetErrorString(e));
    }
    if (idx && h->n_t) {
        CU(h, cudaMemcpyAsync(idx, h->idx, h->n_t * 12, cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
    }
    return ISOMC_OK;
}

int32_t isomc_stats_get(isomc_t *h, isomc_stats *out) {
    if (!h || !out) return ISOMC_ERR_BAD_ARG;
    if (!h->have_result) return fail(h, ISOMC_ERR_NO_RESULT, "no extract has completed on this handle");
    *out = h->stats;
    return ISOMC_OK;
}

/* ---- slabs -------------------------------------------------------------------------------- */

int32_t isomc_slab_count_grid_device(isomc_t *h, const float *d_slab) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    if (!d_slab) return fail(h, ISOMC_ERR_BAD_ARG, "d_slab == NULL");
    int32_t rc = bind_device(h);
    if (rc) return rc;
    h->kind = SRC_GRID; h->d_grid = d_slab; h->directed = false;
    return enqueue_count(h, false);
}

int32_t isomc_slab_count_sdf(isomc_t *h, const isomc_sdf_node *prog, uint32_t n_nodes) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    int32_t rc = bind_device(h);
    if (rc) return rc;
    rc = validate_program(h, prog, n_nodes, &h->prog);
    if (rc) return rc;
    h->kind = SRC_SDF; h->d_grid = nullptr; h->directed = false;
    return enqueue_count(h, false);
}

/* the same with the tree sampled as Directed distances (MarchingCubes<Directed> on a slab) */
int32_t isomc_slab_count_sdf_directed(isomc_t *h, const isomc_sdf_node *prog, uint32_t n_nodes) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    int32_t rc = bind_device(h);
    if (rc) return rc;
    rc = validate_program(h, prog, n_nodes, &h->prog);
    if (rc) return rc;
    h->kind = SRC_SDF; h->d_grid = nullptr; h->directed = true;
    return enqueue_count(h, false);
}

int32_t isomc_slab_totals(isomc_t *h, uint64_t totals[3]) {
    if (!h || !totals) return ISOMC_ERR_BAD_ARG;
    if (!h->counted) return fail(h, ISOMC_ERR_NO_RESULT, "slab_totals before slab_count");
    int32_t rc = bind_device(h);
    if (rc) return rc;
    rc = fetch_totals(h);
    if (rc) return rc;
    totals[0] = h->h_totals[8]; totals[1] = h->h_totals[9]; totals[2] = h->h_totals[10];
    return ISOMC_OK;
}

int32_t isomc_slab_totals_device(isomc_t *h, const uint64_t **d_totals) {
    if (!h || !d_totals) return ISOMC_ERR_BAD_ARG;
    *d_totals = (const uint64_t *)(h->totals + 8);
    return ISOMC_OK;
}

int32_t isomc_slab_emit(isomc_t *h, uint64_t vertex_base, uint64_t boundary_base) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    if (!h->counted) return fail(h, ISOMC_ERR_NO_RESULT, "slab_emit before slab_count");
    int32_t rc = bind_device(h);
    if (rc) return rc;
    const uint64_t ofs = h->g.ghost ? boundary_base : vertex_base;
    if (ofs >= (1ull << 32)) return fail(h, ISOMC_ERR_INDEX_OVERFLOW, "vertex base %llu does not fit u32", (unsigned long long)ofs);
    rc = set_vofs(h, (uint32_t)ofs);
    if (rc) return rc;
    rc = fetch_totals(h);
    if (rc) return rc;
    if (ofs + h->h_totals[0] >= (1ull << 32))
        return fail(h, ISOMC_ERR_INDEX_OVERFLOW, "global vertex ids exceed u32");
    rc = ensure_capacity(h, h->h_totals[8], h->h_totals[10]);
    if (rc) return rc;
    rc = enqueue_emit(h);
    if (rc) return rc;
    return finish_impl(h);
}

int32_t isomc_slab_enqueue_emit_gathered(isomc_t *h, const uint64_t *d_gathered, uint32_t rank, uint32_t n_ranks) {
    if (!h || !d_gathered || rank >= n_ranks) return ISOMC_ERR_BAD_ARG;
    if (!h->counted) return fail(h, ISOMC_ERR_NO_RESULT, "slab_emit before slab_count");
    int32_t rc = bind_device(h);
    if (rc) return rc;
    /* id offset on the device; its 64-bit value goes to totals[13] so that finish() can tell an overflow of the global ids */
    CU(h, isomc_launch_slab_bases((const unsigned long long *)d_gathered, rank, h->g.ghost, h->vofs, h->totals + 13, h->stream));
    h->vofs_cached = -1;
    h->step_exchanged = false;
    h->totals_valid = false; /* (totals[13] is new) */
    h->stats.kernel_launches += 1;
    if (h->cap_v > 0 || h->cap_t > 0) {
        rc = enqueue_emit(h);
        if (rc) return rc;
    }
    return ISOMC_OK;
}

int32_t isomc_slab_emit_gathered(isomc_t *h, const uint64_t *d_gathered, uint32_t rank, uint32_t n_ranks) {
    int32_t rc = isomc_slab_enqueue_emit_gathered(h, d_gathered, rank, n_ranks);
    return rc ? rc : finish_impl(h);
}

/* ---- totals exchange over peer memory (no collective library) ------------------------------ */

static int32_t ensure_mailbox(isomc_t *h) {
    if (h->mailbox) return ISOMC_OK;
    CU(h, cudaMalloc(&h->mailbox, ISOMC_MAILBOX_BYTES));
    CU(h, cudaMemset(h->mailbox, 0, ISOMC_MAILBOX_BYTES));
    CU(h, cudaMalloc(&h->d_peers, ISOMC_MAX_RANKS * sizeof(unsigned long long *)));
    CU(h, cudaMemset(h->d_peers, 0, ISOMC_MAX_RANKS * sizeof(unsigned long long *)));
    return ISOMC_OK;
}

int32_t isomc_slab_mailbox(isomc_t *h, void **d_mailbox) {
    if (!h || !d_mailbox) return ISOMC_ERR_BAD_ARG;
    int32_t rc = bind_device(h);
    if (rc) return rc;
    rc = ensure_mailbox(h);
    if (rc) return rc;
    *d_mailbox = h->mailbox;
    return ISOMC_OK;
}

int32_t isomc_slab_mailbox_ipc(isomc_t *h, void *handle64) {
    if (!h || !handle64) return ISOMC_ERR_BAD_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == ISOMC_IPC_HANDLE_BYTES, "ISOMC_IPC_HANDLE_BYTES");
    int32_t rc = bind_device(h);
    if (rc) return rc;
    rc = ensure_mailbox(h);
    if (rc) return rc;
    cudaIpcMemHandle_t hd;
    CU(h, cudaIpcGetMemHandle(&hd, h->mailbox));
    memcpy(handle64, &hd, sizeof hd);
    return ISOMC_OK;
}

static int32_t connect_impl(isomc_t *h, uint32_t rank, uint32_t n_ranks, std::vector<unsigned long long *> &ptrs) {
    ptrs[rank] = h->mailbox;
    CU(h, cudaStreamSynchronize(h->stream));
    CU(h, cudaMemcpy(h->d_peers, ptrs.data(), n_ranks * sizeof(unsigned long long *), cudaMemcpyHostToDevice));
    h->xchg_rank = rank; h->xchg_n = n_ranks;
    CU(h, cudaMemset(h->mailbox, 0, ISOMC_MAILBOX_BYTES));
    CU(h, cudaMemset(h->totals + 15, 0, sizeof(unsigned long long))); /* step counter of k_slab_exchange */
    if (h->pipe_exec) { cudaGraphExecDestroy(h->pipe_exec); h->pipe_exec = nullptr; }
    return ISOMC_OK;
}

int32_t isomc_slab_connect(isomc_t *h, uint32_t rank, uint32_t n_ranks, void *const *mailboxes) {
    if (!h || !mailboxes || rank >= n_ranks || n_ranks > ISOMC_MAX_RANKS) return h ? fail(h, ISOMC_ERR_BAD_ARG, "bad rank / n_ranks") : ISOMC_ERR_BAD_ARG;
    int32_t rc = bind_device(h);
    if (rc) return rc;
    rc = ensure_mailbox(h);
    if (rc) return rc;
    std::vector<unsigned long long *> ptrs(n_ranks);
    for (uint32_t r = 0; r < n_ranks; ++r) {
        if (r != rank && !mailboxes[r]) return fail(h, ISOMC_ERR_BAD_ARG, "mailbox of rank %u is NULL", r);
        ptrs[r] = (unsigned long long *)mailboxes[r];
        if (r == rank) continue;
        cudaPointerAttributes at;
        CU(h, cudaPointerGetAttributes(&at, mailboxes[r]));
        if (at.device != h->device) { /* another device of this process: peer stores need peer access */
            int can = 0;
            CU(h, cudaDeviceCanAccessPeer(&can, h->device, at.device));
            if (!can) return fail(h, ISOMC_ERR_CUDA, "device %d cannot access device %d as a peer", h->device, at.device);
            cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) return fail(h, ISOMC_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", at.device, cudaGetErrorString(e));
        }
    }
    return connect_impl(h, rank, n_ranks, ptrs);
}

int32_t isomc_slab_connect_ipc(isomc_t *h, uint32_t rank, uint32_t n_ranks, const void *handles) {
    if (!h || !handles || rank >= n_ranks || n_ranks > ISOMC_MAX_RANKS) return h ? fail(h, ISOMC_ERR_BAD_ARG, "bad rank / n_ranks") : ISOMC_ERR_BAD_ARG;
    int32_t rc = bind_device(h);
    if (rc) return rc;
    rc = ensure_mailbox(h);
    if (rc) return rc;
    for (void *p : h->ipc_opened) cudaIpcCloseMemHandle(p);
    h->ipc_opened.clear();
    std::vector<unsigned long long *> ptrs(n_ranks);
    for (uint32_t r = 0; r < n_ranks; ++r) {
        if (r == rank) continue;
        cudaIpcMemHandle_t hd;
        memcpy(&hd, (const char *)handles + (size_t)r * ISOMC_IPC_HANDLE_BYTES, sizeof hd);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(h, ISOMC_ERR_CUDA, "cudaIpcOpenMemHandle (mailbox of rank %u): %s", r, cudaGetErrorString(e));
        h->ipc_opened.push_back(p);
        ptrs[r] = (unsigned long long *)p;
    }
    return connect_impl(h, rank, n_ranks, ptrs);
}

int32_t isomc_slab_enqueue_emit_exchanged(isomc_t *h) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    if (!h->xchg_n) return fail(h, ISOMC_ERR_BAD_ARG, "slab is not connected to its peers (isomc_slab_connect / isomc_slab_connect_ipc)");
    if (!h->counted) return fail(h, ISOMC_ERR_NO_RESULT, "slab_emit before slab_count");
    int32_t rc = bind_device(h);
    if (rc) return rc;
    rc = launch_exchange(h, h->stream);
    if (rc) return rc;
    if (h->cap_v > 0 || h->cap_t > 0) {
        rc = enqueue_emit(h);
        if (rc) return rc;
    }
    return ISOMC_OK;
}

int32_t isomc_slab_emit_exchanged(isomc_t *h) {
    int32_t rc = isomc_slab_enqueue_emit_exchanged(h);
    return rc ? rc : finish_impl(h);
}

/* count + exchange + emit of a slab as ONE launch sequence (replayed from a CUDA graph in steady state) */
int32_t isomc_slab_enqueue_extract_grid_exchanged(isomc_t *h, const float *d_slab) {
    if (!h) return ISOMC_ERR_BAD_ARG;
    if (!d_slab) return fail(h, ISOMC_ERR_BAD_ARG, "d_slab == NULL");
    if (!h->xchg_n) return fail(h, ISOMC_ERR_BAD_ARG, "slab is not connected to its peers (isomc_slab_connect / isomc_slab_connect_ipc)");
    int32_t rc = bind_device(h);
    if (rc) return rc;
    h->kind = SRC_GRID; h->d_grid = d_slab; h->directed = false;
    h->seq_exchange = true;
    rc = enqueue_count(h, h->cap_v > 0 || h->cap_t > 0); /* (first extract: no output buffers yet, finish() sizes them and emits) */
    h->seq_exchange = false;
    return rc;
}

int32_t isomc_slab_extract_grid_exchanged(isomc_t *h, const float *d_slab) {
    int32_t rc = isomc_slab_enqueue_extract_grid_exchanged(h, d_slab);
    return rc ? rc : finish_impl(h);
}

/* ---- debug -------------------------------------------------------------------------------- */

int32_t isomc_debug_cube_indices(isomc_t *h, uint8_t *host_out) {
    if (!h || !host_out) return ISOMC_ERR_BAD_ARG;
    if (!h->counted) return fail(h, ISOMC_ERR_NO_RESULT, "no extract has run on this handle");
    int32_t rc = bind_device(h);
    if (rc) return rc;
    const uint64_t n = (uint64_t)h->g.ncl * h->g.ncx * h->g.ncx;
    if (n == 0) return ISOMC_OK;
    if (h->tile_mode) { /* the tile path keeps no sign words: derive them from the source of the last extract */
        rc = ensure_signs(h);
        if (rc) return rc;
        if (h->kind == SRC_GRID) CU(h, isomc_launch_sign_grid(h->g, h->d_grid, h->signs, 0, h->g.nsl * h->g.N, h->sms, 8, h->stream));
        else CU(h, isomc_launch_sign_sdf(h->g, h->prog, h->directed, h->signs, 0, h->g.nsl * h->g.N, h->sms, 8, h->stream));
    }
    uint8_t *d = nullptr;
    CU(h, cudaMalloc(&d, n));
    cudaError_t e = isomc_launch_cube_indices(h->g, h->signs, h->tabs, d, h->sms, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_out, d, n, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(h, ISOMC_ERR_CUDA, "cube index dump failed: %s", cudaGetErrorString(e));
    return ISOMC_OK;
}

int32_t isomc_debug_sample_sdf(int32_t device, const isomc_sdf_node *prog, uint32_t n_nodes, const float *h_xyz,
                               uint64_t n_points, float *h_out) {
    const int use_chain = (n_nodes & 0x80000000u) != 0;
    n_nodes &= 0x7FFFFFFFu;
    SdfProgram P;
    int32_t rc = validate_program(nullptr, prog, n_nodes, &P);
    if (rc) return rc;
    if (!h_xyz || !h_out) return fail(nullptr, ISOMC_ERR_BAD_ARG, "NULL buffer");
    if (n_points == 0) return ISOMC_OK;
    CU(nullptr, cudaSetDevice(device));
    float *dx = nullptr, *dout = nullptr;
    CU(nullptr, cudaMalloc(&dx, n_points * 12));
    cudaError_t e = cudaMalloc(&dout, n_points * 4);
    if (e == cudaSuccess) e = cudaMemcpy(dx, h_xyz, n_points * 12, cudaMemcpyHostToDevice);
    /* device < 0 is not allowed; the high bit of n_nodes selects the chain evaluator (tests compare both) */
    if (e == cudaSuccess) e = isomc_launch_sample_sdf(P, dx, n_points, dout, use_chain, 0);
    if (e == cudaSuccess) e = cudaMemcpy(h_out, dout, n_points * 4, cudaMemcpyDeviceToHost);
    cudaFree(dx); cudaFree(dout);
    if (e != cudaSuccess) return fail(nullptr, ISOMC_ERR_CUDA, "sdf sampling failed: %s", cudaGetErrorString(e));
    return ISOMC_OK;
}

/* ---- synthetic fields ------------------------------------------------------ */

static uint64_t splitmix64(uint64_t *state) {
    uint64_t z = (*state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static float unit24(uint64_t *state) { return (float)(splitmix64(state) >> 40) * (1.0f / 16777216.0f); }

int32_t isomc_synth_field(int32_t device, int32_t kind, uint32_t size, uint64_t seed, uint32_t z_first,
                          uint32_t n_layers, float *d_out) {
    if (size < 2 || !d_out || n_layers == 0) return fail(nullptr, ISOMC_ERR_BAD_ARG, "bad synth_field arguments");
    SynthParams sp;
    memset(&sp, 0, sizeof sp);
    sp.kind = kind;
    uint64_t st = seed;
    if (kind == ISOMC_FIELD_FBM) {
        /* f(p) = sum_{o<5} 2^-o sum_{k<4} sin(2 pi 4 2^o (d_ok . p) + phi_ok) */
        for (int o = 0; o < 5; ++o)
            for (int k = 0; k < 4; ++k) {
                float dx, dy, dz, len;
                do {
                    dx = 2.0f * unit24(&st) - 1.0f; dy = 2.0f * unit24(&st) - 1.0f; dz = 2.0f * unit24(&st) - 1.0f;
                    len = sqrtf(dx * dx + dy * dy + dz * dz);
                } while (len < 1e-3f);
                const int w = o * 4 + k;
                sp.dx[w] = dx / len; sp.dy[w] = dy / len; sp.dz[w] = dz / len;
                sp.ph[w] = 6.283185307179586f * unit24(&st);
                sp.amp[w] = 1.0f / (float)(1 << o);
                /* 4 * 2^o periods per unit length at 512^3; scaled with the lattice so that the
                 * finest wave stays 8 samples long at other sizes (constant surface density: weak-scaling runs) */
                sp.freq[w] = 6.283185307179586f * 4.0f * (float)(1 << o) * ((float)(size - 1) / 511.0f);
            }
    } else if (kind == ISOMC_FIELD_SPHERE_UNION) {
        for (int s = 0; s < 64; ++s) {
            sp.cx[s] = 0.1f + 0.8f * unit24(&st); sp.cy[s] = 0.1f + 0.8f * unit24(&st); sp.cz[s] = 0.1f + 0.8f * unit24(&st);
            sp.r[s] = 0.05f + 0.1f * unit24(&st);
        }
    } else if (kind != ISOMC_FIELD_GYROID) {
        return fail(nullptr, ISOMC_ERR_BAD_ARG, "unknown field kind %d", kind);
    }
    CU(nullptr, cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(nullptr, cudaGetDeviceProperties(&prop, device));
    CU(nullptr, isomc_launch_synth(sp, size, z_first, n_layers, d_out, prop.multiProcessorCount, 0));
    CU(nullptr, cudaDeviceSynchronize());
    return ISOMC_OK;
}

} /* extern "C" */
