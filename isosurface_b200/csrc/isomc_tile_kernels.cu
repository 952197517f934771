/*
 * isomc_tile_kernels.cu -- the two kernels of the tile path (sm_100a); their bodies are tile_count_item() and
 * tile_emit_item() of isomc_tile.cuh (shared with the host model).
 *
 *   k_tile_count<Src>  pass 1: persistent CTAs take (z-chunk, tile column) items from a ticket counter, chunk-major, so
 *                      that neighbouring columns march the same layers at the same time (their shared halo rows meet
 *                      in L2).  Device grids whose rows are 16-byte aligned are staged by TMA bulk copies
 *                      (cp.async.bulk.shared::cluster.global + mbarrier complete_tx; SASS: UBLKCP) into a three-layer
 *                      ring; other grids and the implicit sources fill a two-layer ring with all threads.
 *                      Replaces k_sign + k_count_list: the samples are read once, nothing is re-read at emission.
 *   k_tile_emit        pass 2: same item scheme; edge-id planes in shared memory, no sample access.
 */
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "isomc_tile.cuh"
#include "isomc_kernels.h"

namespace {

/* ---- sources of pass 1 ---------------------------------------------------------------------------------- */
struct GridBulkSrc { /* TMA-staged device lattice: N % 4 == 0, 16-byte aligned base */
    static constexpr int NS = 3, NC = 1;
    static constexpr bool ASYNC = true;
    const float *p;
    __device__ __forceinline__ const float *base() const { return p; }
    __device__ __forceinline__ void sample(const Geo &g, uint32_t x, uint32_t y, uint32_t lz, float *out) const {
        out[0] = __ldg(p + ((uint64_t)lz * g.N + y) * g.N + x);
    }
};
struct GridBulk2Src { /* the same with a two-layer ring: the copy of a layer is not overlapped within the CTA, but more CTAs fit an SM */
    static constexpr int NS = 2, NC = 1;
    static constexpr bool ASYNC = true;
    const float *p;
    __device__ __forceinline__ const float *base() const { return p; }
    __device__ __forceinline__ void sample(const Geo &g, uint32_t x, uint32_t y, uint32_t lz, float *out) const {
        out[0] = __ldg(p + ((uint64_t)lz * g.N + y) * g.N + x);
    }
};
struct GridPlainSrc { /* any size / alignment: loaded by all threads */
    static constexpr int NS = 2, NC = 1;
    static constexpr bool ASYNC = false;
    const float *p;
    __device__ __forceinline__ const float *base() const { return p; }
    __device__ __forceinline__ void sample(const Geo &g, uint32_t x, uint32_t y, uint32_t lz, float *out) const {
        out[0] = __ldg(p + ((uint64_t)lz * g.N + y) * g.N + x);
    }
};
/* implicit sources: the lattice point is (i as f32) * one_over_size (primal_grid.rs:50,63-67) */
struct SdfTileSrc {
    static constexpr int NS = 2, NC = 1;
    static constexpr bool ASYNC = false;
    SdfProgram prog;
    __device__ __forceinline__ const float *base() const { return nullptr; }
    __device__ __forceinline__ void sample(const Geo &g, uint32_t x, uint32_t y, uint32_t lz, float *out) const {
        out[0] = sdf_eval(prog, __fmul_rn((float)x, g.inv), __fmul_rn((float)y, g.inv), __fmul_rn((float)(g.gz0 + lz), g.inv));
    }
};
struct SdfChainTileSrc {
    static constexpr int NS = 2, NC = 1;
    static constexpr bool ASYNC = false;
    SdfChain chain;
    __device__ __forceinline__ const float *base() const { return nullptr; }
    __device__ __forceinline__ void sample(const Geo &g, uint32_t x, uint32_t y, uint32_t lz, float *out) const {
        out[0] = sdf_chain_eval(chain, __fmul_rn((float)x, g.inv), __fmul_rn((float)y, g.inv), __fmul_rn((float)(g.gz0 + lz), g.inv));
    }
};
struct SdfDirTileSrc { /* MarchingCubes<Directed>: three axis distances per lattice point (distance.rs:72-104) */
    static constexpr int NS = 2, NC = 3;
    static constexpr bool ASYNC = false;
    SdfProgram prog;
    __device__ __forceinline__ const float *base() const { return nullptr; }
    __device__ __forceinline__ void sample(const Geo &g, uint32_t x, uint32_t y, uint32_t lz, float *out) const {
        const Vec3f v = sdf_eval_vec(prog, __fmul_rn((float)x, g.inv), __fmul_rn((float)y, g.inv), __fmul_rn((float)(g.gz0 + lz), g.inv));
        out[0] = v.x; out[1] = v.y; out[2] = v.z;
    }
};

extern __shared__ __align__(16) unsigned char tile_smem_raw[];

template <class Src, int MINB>
__global__ void __launch_bounds__(TILE_NT, MINB) k_tile_count(Src src, Geo g, TileGeo tg, TileBufs B, const EmitTab *__restrict__ tabg,
                                                         uint32_t lz0, uint32_t lz1, uint32_t zc, uint32_t *ticket) {
    using Smem = CountSmem<Src::NS, Src::NC>;
    Smem &S = *reinterpret_cast<Smem *>(tile_smem_raw);
    __shared__ uint32_t s_item;
    if (Src::ASYNC && threadIdx.x == 0) {
        for (int s = 0; s < Src::NS; ++s) tile_mbar_init(&S.mbar[s], 1);
        tile_mbar_fence_init();
    }
    __syncthreads();
    Cta c;
    c.tid = threadIdx.x;
    c.w.lane = threadIdx.x & 31u;
    c.w.emu = nullptr;
    c.bemu = nullptr;
    CountCtx X;
    X.curE.pos = X.curE.end = X.curT.pos = X.curT.end = 0;
    X.phase = 0;
    const uint32_t nchunks = (lz1 - lz0 + zc - 1) / zc, nitems = nchunks * tg.ncols;
    for (;;) {
        if (threadIdx.x == 0) s_item = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t item = s_item;
        __syncthreads();
        if (item >= nitems) break;
        const uint32_t chunk = item / tg.ncols, col = item - chunk * tg.ncols;
        const uint32_t l0 = lz0 + chunk * zc, l1 = min(lz1, l0 + zc);
        tile_count_item(c, g, tg, src, S, B, tabg, col, l0, l1, X);
    }
}

struct EmitArgsDev {
    const uint32_t *pV, *pT, *pE, *pTp;
    const uint16_t *pA;
    const uint2 *ent;
    const float *tq, *tbuf;
    const uint32_t *vofs_ptr;
    const unsigned long long *layerTot;
    const uint32_t *ctr;
    uint32_t cap_eb, cap_tb;
    unsigned long long cap_v, cap_t;
    float *xyz;
    uint32_t *idx;
};

__global__ void __launch_bounds__(EMIT_NT, 7) k_tile_emit(Geo g, TileGeo tg, EmitArgsDev A, const EmitTab *__restrict__ tabg, uint32_t lz0,
                                                     uint32_t lz1, uint32_t *ticket) {
    EmitSmem &S = *reinterpret_cast<EmitSmem *>(tile_smem_raw);
    __shared__ uint32_t s_item;
    /* entry list / t buffer overflow: what was counted is incomplete; the host grows the buffers and re-runs */
    if (A.ctr[0] > A.cap_eb || A.ctr[1] > A.cap_tb) return;
    for (uint32_t i = threadIdx.x; i < 256; i += EMIT_NT) {
        S.tri[i] = (tabg->tri[i] & 0x0FFFFFFFFFFFFFFFull) | (unsigned long long)tabg->ntri[i] << 60;
        S.emask[i] = tabg->emask[i];
        S.rank3[i] = tabg->rank3[i];
    }
    if (threadIdx.x < 24) {
        const uint32_t loc = tabg->eloc[threadIdx.x / 12][threadIdx.x % 12];
        S.etab[threadIdx.x / 12][threadIdx.x % 12] = loc;
        S.eofs[threadIdx.x / 12][threadIdx.x % 12] = make_uint2((loc & 255u) * 4u, (loc >> 12) * 2u);
    }
    EmitParams P;
    P.pV = A.pV; P.pT = A.pT; P.pE = A.pE; P.pTp = A.pTp; P.pA = A.pA;
    P.ent = A.ent; P.tq = A.tq; P.tbuf = A.tbuf;
    P.vofs = *A.vofs_ptr;
    P.ghostV = g.ghost ? (uint32_t)A.layerTot[0] : 0u;
    P.ghostT = g.ghost ? (uint32_t)A.layerTot[1] : 0u;
    P.first_own_layer = g.ghost;
    P.cap_v = A.cap_v; P.cap_t = A.cap_t;
    P.xyz = A.xyz; P.idx = A.idx;
    Cta c;
    c.tid = threadIdx.x;
    c.w.lane = threadIdx.x & 31u;
    c.w.emu = nullptr;
    c.bemu = nullptr;
    const uint32_t nchunks = (lz1 - lz0 + EMIT_ZC - 1) / EMIT_ZC, nitems = nchunks * tg.ncols_emit;
    for (;;) {
        if (threadIdx.x == 0) s_item = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t item = s_item;
        __syncthreads();
        if (item >= nitems) break;
        const uint32_t chunk = item / tg.ncols_emit, col = item - chunk * tg.ncols_emit;
        const uint32_t l0 = lz0 + chunk * EMIT_ZC, l1 = min(lz1, l0 + EMIT_ZC);
        tile_emit_item(c, g, tg, S, P, tabg, col, l0, l1);
    }
}

/* cell layers per counting item: about four items per resident CTA, 8..64 layers (every item re-reads one sample layer) */
uint32_t count_chunk_layers(uint32_t nl, uint32_t ncols, uint32_t slots) {
    uint64_t zc = ((uint64_t)nl * ncols + 4ull * slots - 1) / (4ull * slots);
    if (const char *p = getenv("ISOMC_TILE_ZC")) zc = (uint64_t)atoi(p);
    if (zc < 8) zc = 8;
    if (zc > 64) zc = 64;
    return (uint32_t)zc;
}

template <class Src, int MINB>
cudaError_t launch_count(const Src &src, const Geo &g, const TileGeo &tg, const TileBufs &B, const EmitTab *tab, uint32_t lz0, uint32_t lz1,
                         uint32_t *ticket, int sms, cudaStream_t st) {
    using Smem = CountSmem<Src::NS, Src::NC>;
    static bool attr_set = false; /* per instantiation */
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_tile_count<Src, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const uint32_t slots = (uint32_t)sms * MINB;
    const uint32_t zc = count_chunk_layers(lz1 - lz0, tg.ncols, slots);
    const uint64_t nitems = (uint64_t)((lz1 - lz0 + zc - 1) / zc) * tg.ncols;
    const uint32_t grid = (uint32_t)(nitems < slots ? nitems : slots);
    k_tile_count<Src, MINB><<<grid, TILE_NT, sizeof(Smem), st>>>(src, g, tg, B, tab, lz0, lz1, zc, ticket);
    return cudaGetLastError();
}

} /* namespace */

size_t isomc_tile_emit_smem_bytes() { return sizeof(EmitSmem); }

void isomc_tile_fill_eloc(uint32_t eloc[2][12]) {
    for (uint32_t par = 0; par < 2; ++par)
        for (uint32_t e = 0; e < 12; ++e) eloc[par][e] = tile_edge_loc(par, e);
}

cudaError_t isomc_launch_tile_count_grid(const Geo &g, const TileGeo &tg, const float *d_grid, const TileBufs &B, const EmitTab *tab,
                                         uint32_t lz0, uint32_t lz1, uint32_t *ticket, int sms, cudaStream_t st) {
    static int plain = -1;
    if (plain < 0) {
        const char *p = getenv("ISOMC_FILL");
        plain = (p && strcmp(p, "plain") == 0) ? 1 : 0;
    }
    static int ring = -1;
    if (ring < 0) {
        const char *p = getenv("ISOMC_RING");
        ring = p ? atoi(p) : 3;
    }
    const bool aligned = (g.N % 4u) == 0 && (reinterpret_cast<uintptr_t>(d_grid) & 15u) == 0;
    if (aligned && !plain && ring == 2) return launch_count<GridBulk2Src, 5>(GridBulk2Src{d_grid}, g, tg, B, tab, lz0, lz1, ticket, sms, st);
    if (aligned && !plain) return launch_count<GridBulkSrc, 4>(GridBulkSrc{d_grid}, g, tg, B, tab, lz0, lz1, ticket, sms, st);
    return launch_count<GridPlainSrc, 4>(GridPlainSrc{d_grid}, g, tg, B, tab, lz0, lz1, ticket, sms, st);
}

cudaError_t isomc_launch_tile_count_sdf(const Geo &g, const TileGeo &tg, const SdfProgram &prog, bool directed, const TileBufs &B,
                                        const EmitTab *tab, uint32_t lz0, uint32_t lz1, uint32_t *ticket, int sms, cudaStream_t st) {
    if (directed) return launch_count<SdfDirTileSrc, 1>(SdfDirTileSrc{prog}, g, tg, B, tab, lz0, lz1, ticket, sms, st);
    SdfChainTileSrc csrc;
    if (sdf_to_chain(prog, &csrc.chain)) return launch_count<SdfChainTileSrc, 4>(csrc, g, tg, B, tab, lz0, lz1, ticket, sms, st);
    return launch_count<SdfTileSrc, 4>(SdfTileSrc{prog}, g, tg, B, tab, lz0, lz1, ticket, sms, st);
}

cudaError_t isomc_launch_tile_emit(const Geo &g, const TileGeo &tg, const TileBufs &B, const EmitTab *tab, const uint32_t *vofs,
                                   float *xyz, uint32_t *idx, uint64_t cap_v, uint64_t cap_t, uint32_t lz0, uint32_t lz1, uint32_t *ticket,
                                   int sms, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_tile_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EmitSmem));
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    EmitArgsDev A;
    A.pV = B.pV; A.pT = B.pT; A.pE = B.pE; A.pTp = B.pTp; A.pA = B.pA;
    A.ent = B.ent; A.tq = B.tq; A.tbuf = B.tbuf;
    A.vofs_ptr = vofs;
    A.layerTot = B.layerTot;
    A.ctr = B.ctr;
    A.cap_eb = B.cap_eb; A.cap_tb = B.cap_tb;
    A.cap_v = cap_v; A.cap_t = cap_t;
    A.xyz = xyz; A.idx = idx;
    const uint64_t nitems = (uint64_t)((lz1 - lz0 + EMIT_ZC - 1) / EMIT_ZC) * tg.ncols_emit;
    const uint32_t slots = (uint32_t)sms * 7u;
    const uint32_t grid = (uint32_t)(nitems < slots ? nitems : slots);
    k_tile_emit<<<grid, EMIT_NT, sizeof(EmitSmem), st>>>(g, tg, A, tab, lz0, lz1, ticket);
    return cudaGetLastError();
}
