/*
 * isomc_list_kernels.cu -- the active-cell-list form of output sizing and emission (sm_100a).
 *
 *   K2L k_count_list   replaces k_count: classifies 32-cell segments bit-parallel (phase A, lane per
 *                      segment), then walks the warp's active cells with one lane per CELL (phase B): cube
 *                      index, triangles, in-row vertex / triangle prefixes -> one 12-byte list entry per
 *                      active cell, one record per active segment (isomc_cell.cuh).  List space is handed
 *                      out in blocks of LIST_BLOCK entries by one atomic per block, not per warp pass.
 *   K4L k_emit_list    replaces k_emit + k_vertex: one lane per list entry runs emit_cell(): edge ids from
 *                      the entries of the cells that created them (edge ownership replaces the reference's
 *                      HashMap index cache, index_cache.rs / mesh.rs:240-251), the cell's own vertices
 *                      (distance.rs:64-69) and its triangles (marching_cubes_impl.rs:102-117) go straight
 *                      to their final slots.  No bricks, no halo recomputation, no descriptors.
 *
 * Every lane of both hot loops works on an ACTIVE cell, which is what the brick kernels could not offer
 * (profiles/r01_history.md: 22 of 32 lanes, 1.6 active cells per segment against a warp maximum of 19).
 */
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#include "isomc_cell.cuh"
#include "isomc_kernels.h"
#include "isomc_launch.cuh"

namespace {

/* cell rows [row0, row1); the body is count_list_warp() of isomc_cell.cuh (shared with the host model) */
template <bool WIDE, int MINB>
__global__ void __launch_bounds__(256, MINB) k_count_list(Geo g, const uint32_t *__restrict__ signs, const uint8_t *__restrict__ ntri_g,
                                                    ListBufs L, CountOut out, uint32_t gshift, uint32_t row0, uint32_t row1,
                                                    uint32_t *ticket, uint32_t task_passes) {
    __shared__ uint8_t s_ntri[256];
    __shared__ SegQueue s_q[8];
    __shared__ uint8_t s_nth8[256 * 8];
    isomc_pdl_trigger();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { s_ntri[i] = ntri_g[i]; nth8_fill(s_nth8, (uint32_t)i); } /* (static tables) */
    __syncthreads();
    isomc_pdl_wait(); /* the sign words come from the kernel before */
    const uint32_t warp = threadIdx.x >> 5;
    const Warp w{threadIdx.x & 31u, nullptr};
    count_list_warp<WIDE>(w, g, signs, s_ntri, s_nth8, L, out, gshift, row0, row1, ticket, s_q[warp], gridDim.x * (blockDim.x >> 5), task_passes);
}

/* list blocks [*blk_first, *blk_end): a CTA per block (strided over the CTAs), one lane per entry.  (Warps drawing 32-entry pieces
 * from a ticket counter instead was measured: emit 0.273 -> 0.325 ms at 512^3, 1.29 -> 2.68 ms at 2048^3 -- the eight warps of a
 * CTA working on ONE block is what keeps the neighbour look-ups in L1.) */
template <class Src, int MINB>
__global__ void __launch_bounds__(LIST_BLOCK, MINB) k_emit_list(Src src, Geo g, ListBufs L, const EmitTab *__restrict__ tab_g,
                                                          const uint32_t *__restrict__ rowPV, const uint32_t *__restrict__ rowPT,
                                                          const unsigned long long *__restrict__ layerTot,
                                                          const uint32_t *__restrict__ vofs_ptr, float *__restrict__ xyz,
                                                          uint32_t *__restrict__ idx, unsigned long long cap_v,
                                                          unsigned long long cap_t, const uint32_t *__restrict__ blk_first,
                                                          const uint32_t *__restrict__ blk_end) {
    __shared__ EmitTab T;
    __shared__ uint32_t s_eid[12 * LIST_BLOCK];
    isomc_pdl_trigger();
    { /* (the table is written once at create: safe to read before the kernels earlier in the stream are done) */
        const uint32_t *srcw = reinterpret_cast<const uint32_t *>(tab_g);
        uint32_t *dstw = reinterpret_cast<uint32_t *>(&T);
        for (uint32_t i = threadIdx.x; i < sizeof(EmitTab) / 4; i += LIST_BLOCK) dstw[i] = srcw[i];
    }
    __syncthreads();
    isomc_pdl_wait(); /* list, prefixes, marks: the counting kernel and the row scan */
    if (*L.ctr > L.cap_blocks) return; /* list overflow: entries are incomplete; the host grows the list and re-runs */
    const uint32_t b0 = blk_first ? *blk_first : 0u, b1 = *blk_end;
    EmitArgs A;
    A.rowPV = rowPV; A.rowPT = rowPT;
    A.vofs = *vofs_ptr;
    A.ghostV = g.ghost ? (uint32_t)layerTot[0] : 0u;
    A.ghostT = g.ghost ? (uint32_t)layerTot[1] : 0u;
    A.first_own_layer = g.ghost;
    A.cap_v = cap_v; A.cap_t = cap_t;
    A.xyz = xyz; A.idx = idx;
    /* the entries of the next piece of work are requested before the current one is worked on: the first load of an iteration
     * (a DRAM miss) was the largest single stall of the kernel */
    /* (the fill of a block is fetched two blocks ahead, its entries one block ahead and only below the fill: no load ever
     * touches list space that was not written) */
    uint32_t b = b0 + blockIdx.x, bn = b + gridDim.x;
    uint32_t fill = b < b1 ? L.blkfill[b] : 0u, fill_n = bn < b1 ? L.blkfill[bn] : 0u, yz = 0;
    uint2 ea = make_uint2(0u, 0u);
    if (threadIdx.x < fill) {
        const uint64_t k = (uint64_t)b * LIST_BLOCK + threadIdx.x;
        ea = L.ent[k]; yz = L.ent_yz[k];
    }
    while (b < b1) {
        const uint32_t bnn = bn + gridDim.x;
        const uint32_t fill_nn = bnn < b1 ? L.blkfill[bnn] : 0u;
        uint32_t yz_n = 0;
        uint2 ea_n = make_uint2(0u, 0u);
        if (threadIdx.x < fill_n) {
            const uint64_t kn = (uint64_t)bn * LIST_BLOCK + threadIdx.x;
            ea_n = L.ent[kn]; yz_n = L.ent_yz[kn];
        }
        if (threadIdx.x < fill) emit_cell(g, src, T, L, A, (uint64_t)b * LIST_BLOCK + threadIdx.x, ea, yz, s_eid + threadIdx.x, LIST_BLOCK);
        b = bn; bn = bnn; fill = fill_n; fill_n = fill_nn; ea = ea_n; yz = yz_n;
    }
}

inline uint32_t grid_for(uint64_t warps_needed, int sms, int warps_per_block, int blocks_per_sm) {
    uint64_t blocks = (warps_needed + warps_per_block - 1) / warps_per_block;
    const uint64_t cap = (uint64_t)sms * (blocks_per_sm < 1 ? 1 : blocks_per_sm);
    if (blocks > cap) blocks = cap;
    return (uint32_t)(blocks < 1 ? 1 : blocks);
}

/* CTAs per SM of k_emit_list, i.e. the register budget vs. the warps in flight (ISOMC_EMIT_MINB = 4, 5 or 6).  Default: 5
 * (48 registers); 4 (64 registers) on lattices wider than 1024, where the kernel is bound by the sample gathers and more warps
 * in flight only thrash (2048^3: 1.39 ms with 5, 1.25 ms with 4; 1024^3: 1.27 with 5, 1.30 with 4).  CTAs of 128 threads (twice
 * as many) were measured too: slower at 512^3 and 1024^3 (profiles/r02_history.md). */
static int emit_list_minb(const Geo &g) {
    static int v = -1;
    if (v < 0) {
        const char *p = getenv("ISOMC_EMIT_MINB");
        v = p ? atoi(p) : 0;
        if (v != 4 && v != 5 && v != 6) v = 0;
    }
    return v ? v : (g.N > 1024 ? 4 : 5);
}

template <class Src>
cudaError_t launch_emit_list(const Src &src, const Geo &g, const ListBufs &L, const EmitTab *tab, const uint32_t *rowPV,
                             const uint32_t *rowPT, const unsigned long long *layerTot, const uint32_t *vofs, float *xyz,
                             uint32_t *idx, uint64_t cap_v, uint64_t cap_t, const uint32_t *blk_first, const uint32_t *blk_end,
                             int sms, cudaStream_t st, int grid_bps) {
    const int minb = emit_list_minb(g);
    uint32_t grid = (uint32_t)(sms * (grid_bps > 0 && grid_bps < minb ? grid_bps : minb));
    /* grid_bps < 0: -(list blocks the previous extract of the handle used): a small lattice does not need 740 CTAs that each
     * copy the tables and find nothing to do (any grid is correct: the CTAs stride over the blocks) */
    if (grid_bps < 0 && (uint32_t)(-grid_bps) < grid) grid = (uint32_t)(-grid_bps);
#define ISOMC_EMIT_LAUNCH(M) isomc_launch(k_emit_list<Src, M>, grid, LIST_BLOCK, st, isomc_pdl_for((unsigned long long)g.N * g.N * g.nsl), src, g, L, tab, rowPV, rowPT, layerTot, vofs, xyz, idx, (unsigned long long)cap_v, (unsigned long long)cap_t, blk_first, blk_end)
    switch (minb) {
    case 4: return ISOMC_EMIT_LAUNCH(4);
    case 6: return ISOMC_EMIT_LAUNCH(6);
    default: return ISOMC_EMIT_LAUNCH(5);
    }
#undef ISOMC_EMIT_LAUNCH
}

} /* namespace */

/* CTAs per SM of k_count_list: 4 = 64 registers, 5 = 48, 6 = 40 with a few spills (ISOMC_LIST_MINB, default 4: measured best on the dense field) */
static int count_list_minb() {
    static int v = 0;
    if (v == 0) {
        const char *p = getenv("ISOMC_LIST_MINB");
        v = p ? atoi(p) : 4;
        if (v != 4 && v != 5 && v != 6) v = 4;
    }
    return v;
}

/* warps k_count_list may run: each can strand one partly filled block */
uint32_t isomc_count_list_max_warps(int sms) { return (uint32_t)(sms * 6 * 8); }

template <bool WIDE>
static void launch_count_list_v(uint32_t grid, cudaStream_t st, const Geo &g, const uint32_t *signs, const uint8_t *ntri,
                                const ListBufs &L, const CountOut &out, uint32_t gshift, uint32_t row0, uint32_t row1, uint32_t *ticket) {
    static int tp = -1; /* ISOMC_COUNT_TASK: passes per ticket (0 = count_task_passes()) */
    if (tp < 0) { const char *p = getenv("ISOMC_COUNT_TASK"); tp = p ? atoi(p) : 0; if (tp < 0) tp = 0; }
    switch (count_list_minb()) {
    default: isomc_launch(k_count_list<WIDE, 4>, grid, 256, st, isomc_pdl_for((unsigned long long)g.N * g.N * g.nsl), g, signs, ntri, L, out, gshift, row0, row1, ticket, (uint32_t)tp); break;
    case 6: isomc_launch(k_count_list<WIDE, 6>, grid, 256, st, isomc_pdl_for((unsigned long long)g.N * g.N * g.nsl), g, signs, ntri, L, out, gshift, row0, row1, ticket, (uint32_t)tp); break;
    case 5: isomc_launch(k_count_list<WIDE, 5>, grid, 256, st, isomc_pdl_for((unsigned long long)g.N * g.N * g.nsl), g, signs, ntri, L, out, gshift, row0, row1, ticket, (uint32_t)tp); break;
    }
}

cudaError_t isomc_launch_count_list(const Geo &g, const uint32_t *signs, const McTables *tabs, const ListBufs &L, uint32_t *rowV,
                                    uint32_t *rowT, uint32_t *rowA, unsigned long long *layerTot, uint32_t *ticket, uint32_t lz0,
                                    uint32_t lz1, int sms, cudaStream_t st, int grid_bps) {
    const uint32_t npair = (g.nsegx + 1) / 2; /* lanes per row: every lane scans two neighbouring segments */
    uint32_t gshift = 0;
    while ((1u << gshift) < npair && gshift < 5) ++gshift;
    const uint32_t row0 = lz0 * g.ncx, row1 = lz1 * g.ncx;
    const uint8_t *ntri = reinterpret_cast<const uint8_t *>(tabs) + offsetof(McTables, ntri);
    const CountOut out{rowV, rowT, rowA, layerTot};
    const int per_sm = grid_bps > 0 && grid_bps < count_list_minb() ? grid_bps : count_list_minb();
    if (npair <= 32) {
        const uint32_t rpw = 32u >> gshift;
        const uint64_t warps = ((uint64_t)(row1 - row0) + rpw - 1) / rpw;
        launch_count_list_v<false>(grid_for(warps, sms, 8, per_sm), st, g, signs, ntri, L, out, gshift, row0, row1, ticket);
    } else {
        launch_count_list_v<true>(grid_for(row1 - row0, sms, 8, per_sm), st, g, signs, ntri, L, out, gshift, row0, row1, ticket);
    }
    return cudaGetLastError();
}

cudaError_t isomc_launch_emit_list_grid(const Geo &g, const float *d_grid, const ListBufs &L, const EmitTab *tab,
                                        const uint32_t *rowPV, const uint32_t *rowPT, const unsigned long long *layerTot,
                                        const uint32_t *vofs, float *xyz, uint32_t *idx, uint64_t cap_v, uint64_t cap_t,
                                        const uint32_t *blk_first, const uint32_t *blk_end, int sms, cudaStream_t st, int grid_bps) {
    return launch_emit_list(GridSrc{d_grid}, g, L, tab, rowPV, rowPT, layerTot, vofs, xyz, idx, cap_v, cap_t, blk_first, blk_end,
                            sms, st, grid_bps);
}

cudaError_t isomc_launch_emit_list_sdf_batch(const Geo &g, const SdfProgram *d_progs, bool directed, const ListBufs &L, const EmitTab *tab,
                                             const uint32_t *rowPV, const uint32_t *rowPT, const unsigned long long *layerTot,
                                             const uint32_t *vofs, float *xyz, uint32_t *idx, uint64_t cap_v, uint64_t cap_t,
                                             const uint32_t *blk_first, const uint32_t *blk_end, int sms, cudaStream_t st, int grid_bps) {
    if (directed)
        return launch_emit_list(SdfBatchDirSrc{d_progs}, g, L, tab, rowPV, rowPT, layerTot, vofs, xyz, idx, cap_v, cap_t, blk_first, blk_end, sms, st, grid_bps);
    return launch_emit_list(SdfBatchSrc{d_progs}, g, L, tab, rowPV, rowPT, layerTot, vofs, xyz, idx, cap_v, cap_t, blk_first, blk_end, sms, st, grid_bps);
}

cudaError_t isomc_launch_emit_list_sdf(const Geo &g, const SdfProgram &prog, bool directed, const ListBufs &L, const EmitTab *tab,
                                       const uint32_t *rowPV, const uint32_t *rowPT, const unsigned long long *layerTot,
                                       const uint32_t *vofs, float *xyz, uint32_t *idx, uint64_t cap_v, uint64_t cap_t,
                                       const uint32_t *blk_first, const uint32_t *blk_end, int sms, cudaStream_t st, int grid_bps) {
    if (directed)
        return launch_emit_list(SdfDirSrc{prog}, g, L, tab, rowPV, rowPT, layerTot, vofs, xyz, idx, cap_v, cap_t, blk_first, blk_end, sms, st, grid_bps);
    SdfChainSrc csrc;
    if (sdf_to_chain(prog, &csrc.chain))
        return launch_emit_list(csrc, g, L, tab, rowPV, rowPT, layerTot, vofs, xyz, idx, cap_v, cap_t, blk_first, blk_end, sms, st, grid_bps);
    return launch_emit_list(SdfSrc{prog}, g, L, tab, rowPV, rowPT, layerTot, vofs, xyz, idx, cap_v, cap_t, blk_first, blk_end, sms,
                            st, grid_bps);
}
