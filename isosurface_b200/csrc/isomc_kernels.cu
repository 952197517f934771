/*
 * isomc_kernels.cu -- the sm_100a kernels of the MarchingCubes extract path.
 *
 * Pipeline (one stream, no host round trip in steady state):
 *
 *   K1 k_sign<Src>   sample -> inside bit.  One bit per lattice point (`!(v > 0)`,
 *                    marching_cubes_impl.rs:32 / distance.rs:52-54), packed 32 per word with a
 *                    warp ballot.  Grid sources stream every f32 exactly once (HBM bound);
 *                    implicit sources evaluate the SDF program instead of loading.
 *   K2 k_count       per 32-cell segment: bit-parallel classification.  Crossed-edge masks are
 *                    XORs of sign words, the "edges this cell creates" count is a bit-sliced sum
 *                    of the owned masks, triangle counts come from ntri[ci'] for active cells
 *                    only.  Writes within-row exclusive prefixes per segment and row totals.
 *   K3 k_scan_rows   exclusive scan over cell rows in (z, y) order (+ totals).
 *   K4 k_emit<Src>   bricks of 128x8x4 cells: compacts active cells, interpolates each owned
 *                    edge once (edge ownership replaces the reference's HashMap index cache,
 *                    index_cache.rs / mesh.rs:240-251) and writes u32 indices in reference order.
 *
 * Vertex numbering = reference numbering: id(cell, e) = (# vertices created by earlier cells in
 * (z,y,x) order) + (# edges the cell creates that precede e in first-appearance order of its
 * triangle list).  See SURVEY.md 3.1-9 and DESIGN.md.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "isomc_device.cuh"
#include "isomc_kernels.h"
#include "isomc_tables.h"

/* ------------------------------------------------------------------------------------------ */
/* small helpers                                                                                */
/* ------------------------------------------------------------------------------------------ */

__device__ __forceinline__ uint32_t lo32(uint64_t v) { return (uint32_t)v; }

/* bit-sliced add of a 1-bit-per-cell mask into a 4-bit-per-cell counter (c0 = LSB plane) */
__device__ __forceinline__ void bs_add(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t m) {
    uint32_t k0 = c0 & m; c0 ^= m;
    uint32_t k1 = c1 & k0; c1 ^= k0;
    uint32_t k2 = c2 & k1; c2 ^= k1;
    c3 ^= k2;
}

/*
 * Per-cell count of the vertices a cell creates, for 32 cells at once, as 4 bit planes.
 * Inputs are the inside bits of the 8 corner rows aligned so that bit j of a0/b0/c0/d0 is the
 * corner at the cell's own x and bit j of an/bn/cn/dn the corner at x+1:
 *   a: (y, z)   b: (y+1, z)   c: (y, z+1)   d: (y+1, z+1)
 * Ownership: every cell creates e5, e6, e10; cells with global z == 0 also
 * e1, e2 (and e0 if y == 0, e3 if x == 0); cells with y == 0 also e4, e9 (e8 if x == 0); cells
 * with x == 0 also e7, e11.  x0m has the bit of the x == 0 cell set (or is 0).
 */
__device__ __forceinline__ uint4 owned_planes(uint32_t a0, uint32_t an, uint32_t b0, uint32_t bn,
                                              uint32_t c0, uint32_t cn, uint32_t d0, uint32_t dn,
                                              bool Z0, bool Y0, uint32_t x0m, uint32_t vm) {
    uint32_t p0 = 0, p1 = 0, p2 = 0, p3 = 0;
    bs_add(p0, p1, p2, p3, (cn ^ dn) & vm);          /* e5: corners 5-6 */
    bs_add(p0, p1, p2, p3, (dn ^ d0) & vm);          /* e6: corners 6-7 */
    bs_add(p0, p1, p2, p3, (bn ^ dn) & vm);          /* e10: corners 2-6 */
    if (Z0) {
        bs_add(p0, p1, p2, p3, (an ^ bn) & vm);      /* e1: corners 1-2 */
        bs_add(p0, p1, p2, p3, (bn ^ b0) & vm);      /* e2: corners 2-3 */
        bs_add(p0, p1, p2, p3, (b0 ^ a0) & vm & x0m);/* e3: corners 3-0 */
        if (Y0) bs_add(p0, p1, p2, p3, (a0 ^ an) & vm); /* e0: corners 0-1 */
    }
    if (Y0) {
        bs_add(p0, p1, p2, p3, (c0 ^ cn) & vm);      /* e4: corners 4-5 */
        bs_add(p0, p1, p2, p3, (an ^ cn) & vm);      /* e9: corners 1-5 */
        bs_add(p0, p1, p2, p3, (a0 ^ c0) & vm & x0m);/* e8: corners 0-4 */
    }
    if (x0m) {
        bs_add(p0, p1, p2, p3, (d0 ^ c0) & vm & x0m);/* e7: corners 7-4 */
        bs_add(p0, p1, p2, p3, (b0 ^ d0) & vm & x0m);/* e11: corners 3-7 */
    }
    return make_uint4(p0, p1, p2, p3);
}

__device__ __forceinline__ uint32_t planes_count(uint4 p, uint32_t m) {
    return __popc(p.x & m) + 2 * __popc(p.y & m) + 4 * __popc(p.z & m) + 8 * __popc(p.w & m);
}

/* cells whose 8 corners are neither all inside nor all outside (point_cloud.rs:58) */
__device__ __forceinline__ uint32_t active_mask(uint32_t a0, uint32_t an, uint32_t b0, uint32_t bn,
                                                uint32_t c0, uint32_t cn, uint32_t d0, uint32_t dn, uint32_t vm) {
    uint32_t all_in = a0 & an & b0 & bn & c0 & cn & d0 & dn;
    uint32_t any_in = a0 | an | b0 | bn | c0 | cn | d0 | dn;
    return any_in & ~all_in & vm;
}

__device__ __forceinline__ uint32_t valid_mask(uint32_t n) { /* low n bits, n in [0, 32] */
    return n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u);
}

/* ------------------------------------------------------------------------------------------ */
/* K1: sample -> inside bits                                                                    */
/* ------------------------------------------------------------------------------------------ */

template <class Src>
__global__ void __launch_bounds__(256) k_sign(Src src, Geo g, uint32_t *__restrict__ signs) {
    constexpr int U = 8;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    const uint32_t nrows = g.nsl * g.N;
    for (uint32_t row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < nrows; row += nwarps) {
        const uint32_t lz = row / g.N, y = row - lz * g.N;
        uint32_t *out = signs + (uint64_t)row * g.nws;
        for (uint32_t w0 = 0; w0 < g.nws; w0 += U) {
            float v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                uint32_t x = (w0 + u) * 32 + lane;
                v[u] = (x < g.N) ? src.at(g, x, y, lz) : 1.0f;
            }
            uint32_t mine = 0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                uint32_t b = __ballot_sync(0xFFFFFFFFu, !(v[u] > 0.0f));
                if (lane == (uint32_t)u) mine = b;
            }
            if (lane < U && w0 + lane < g.nws) out[w0 + lane] = mine;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* K2: per-segment counts, within-row prefixes, row totals                                      */
/* ------------------------------------------------------------------------------------------ */

__global__ void __launch_bounds__(256) k_count(Geo g, const uint32_t *__restrict__ signs,
                                               const McTables *__restrict__ tabs, uint32_t *__restrict__ segpre,
                                               uint32_t *__restrict__ rowV, uint32_t *__restrict__ rowT,
                                               uint32_t *__restrict__ rowA, unsigned long long *__restrict__ layerTot) {
    __shared__ uint8_t s_ntri[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_ntri[i] = tabs->ntri[i];
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31;
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    const uint32_t nrows = g.ncl * g.ncx;
    const uint64_t layer_stride = (uint64_t)g.N * g.nws;
    for (uint32_t row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < nrows; row += nwarps) {
        const uint32_t lz = row / g.ncx, y = row - lz * g.ncx;
        const bool Z0 = (g.gz0 + lz) == 0, Y0 = (y == 0);
        const uint32_t *r00 = signs + ((uint64_t)lz * g.N + y) * g.nws;
        const uint32_t *r01 = r00 + g.nws, *r10 = r00 + layer_stride, *r11 = r10 + g.nws;
        uint32_t carryV = 0, carryT = 0, carryA = 0;
        for (uint32_t s0 = 0; s0 < g.nsegx; s0 += 32) {
            const uint32_t s = s0 + lane;
            uint32_t nv = 0, nt = 0, na = 0;
            if (s < g.nsegx) {
                uint32_t a0 = __ldg(r00 + s), a1 = __ldg(r00 + s + 1);
                uint32_t b0 = __ldg(r01 + s), b1 = __ldg(r01 + s + 1);
                uint32_t c0 = __ldg(r10 + s), c1 = __ldg(r10 + s + 1);
                uint32_t d0 = __ldg(r11 + s), d1 = __ldg(r11 + s + 1);
                uint32_t all_or = a0 | b0 | c0 | d0 | (a1 & 1u) | (b1 & 1u) | (c1 & 1u) | (d1 & 1u);
                uint32_t all_and = a0 & b0 & c0 & d0;
                bool uniform = (all_or == 0u) || (all_and == 0xFFFFFFFFu && (a1 & b1 & c1 & d1 & 1u));
                if (!uniform) {
                    uint32_t an = __funnelshift_r(a0, a1, 1), bn = __funnelshift_r(b0, b1, 1);
                    uint32_t cn = __funnelshift_r(c0, c1, 1), dn = __funnelshift_r(d0, d1, 1);
                    uint32_t vm = valid_mask(g.ncx - s * 32);
                    uint4 pl = owned_planes(a0, an, b0, bn, c0, cn, d0, dn, Z0, Y0, s == 0 ? 1u : 0u, vm);
                    nv = planes_count(pl, 0xFFFFFFFFu);
                    uint32_t act = active_mask(a0, an, b0, bn, c0, cn, d0, dn, vm);
                    na = __popc(act);
                    while (act) {
                        uint32_t i = __ffs(act) - 1;
                        act &= act - 1;
                        uint32_t ci = (__funnelshift_r(a0, a1, i) & 3u) | (__funnelshift_r(b0, b1, i) & 3u) << 2 |
                                      (__funnelshift_r(c0, c1, i) & 3u) << 4 | (__funnelshift_r(d0, d1, i) & 3u) << 6;
                        nt += s_ntri[ci];
                    }
                }
            }
            /* warp inclusive scan of (nv, nt) packed as 2 x 16 bit is not safe in general (a row of
             * boundary cells can exceed 16 bits only for N > 5461, rejected at create) */
            uint32_t pk = nv | nt << 16, inc = pk;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= (uint32_t)d) inc += o;
            }
            uint32_t exc = inc - pk;
            if (s < g.nsegx)
                segpre[(uint64_t)row * g.nsegx + s] = ((carryV + (exc & 0xFFFFu)) & 0xFFFFu) | (carryT + (exc >> 16)) << 16;
            uint32_t tot = __shfl_sync(0xFFFFFFFFu, inc, 31);
            carryV += tot & 0xFFFFu;
            carryT += tot >> 16;
            uint32_t sa = na;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) sa += __shfl_xor_sync(0xFFFFFFFFu, sa, d);
            carryA += sa;
        }
        if (lane == 0) {
            rowV[row] = carryV;
            rowT[row] = carryT;
            rowA[row] = carryA;
            if (carryV | carryT) {
                atomicAdd(&layerTot[3 * lz + 0], (unsigned long long)carryV);
                atomicAdd(&layerTot[3 * lz + 1], (unsigned long long)carryT);
                atomicAdd(&layerTot[3 * lz + 2], (unsigned long long)carryA);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* K3: exclusive scan over cell rows (one CTA per cell layer) + totals                          */
/* ------------------------------------------------------------------------------------------ */

template <typename T>
__device__ __forceinline__ T block_excl_scan_256(T v, T *s_warp, T &total) {
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= (uint32_t)d) inc += o;
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    T wbase = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        T x = s_warp[i];
        if ((uint32_t)i < w) wbase += x;
        tot += x;
    }
    __syncthreads();
    total = tot;
    return wbase + inc - v;
}

/* in: rowV/rowT hold per-row counts; out: exclusive prefixes over rows in (lz, y) order, with a
 * sentinel entry [nrows] = grand total.  totals (u64):
 *   [0] V incl. ghost layer  [1] T incl. ghost  [2] active cells incl. ghost
 *   [3] V prefix at the start of the last cell layer  [4..6] V, T, active of the ghost layer
 *   [8] vertices owned  [9] owned vertices created before the last cell layer  [10] triangles owned */
__global__ void __launch_bounds__(256) k_scan_rows(Geo g, uint32_t *__restrict__ rowV, uint32_t *__restrict__ rowT,
                                                   const unsigned long long *__restrict__ layerTot,
                                                   unsigned long long *__restrict__ totals) {
    __shared__ unsigned long long s_w[8];
    __shared__ unsigned long long s_base[2];
    const uint32_t lz = blockIdx.x;
    /* base = sum of the totals of the layers below (<= 4096 values) */
    unsigned long long bv = 0, bt = 0, ba = 0;
    for (uint32_t l = threadIdx.x; l < lz; l += blockDim.x) {
        bv += layerTot[3 * l];
        bt += layerTot[3 * l + 1];
        ba += layerTot[3 * l + 2];
    }
    unsigned long long tv, tt, ta;
    block_excl_scan_256<unsigned long long>(bv, s_w, tv);
    block_excl_scan_256<unsigned long long>(bt, s_w, tt);
    block_excl_scan_256<unsigned long long>(ba, s_w, ta);
    if (threadIdx.x == 0) { s_base[0] = tv; s_base[1] = tt; }
    __syncthreads();
    const uint32_t per = (g.ncx + 255) / 256;
    const uint32_t y_begin = threadIdx.x * per, y_end = min(g.ncx, y_begin + per);
    uint32_t sv = 0, st = 0;
    for (uint32_t y = y_begin; y < y_end; ++y) {
        sv += rowV[lz * g.ncx + y];
        st += rowT[lz * g.ncx + y];
    }
    unsigned long long tot;
    unsigned long long ev = block_excl_scan_256<unsigned long long>((unsigned long long)sv, s_w, tot);
    unsigned long long et = block_excl_scan_256<unsigned long long>((unsigned long long)st, s_w, tot);
    uint32_t pv = (uint32_t)(s_base[0] + ev), pt = (uint32_t)(s_base[1] + et);
    for (uint32_t y = y_begin; y < y_end; ++y) {
        uint32_t cv = rowV[lz * g.ncx + y], ct = rowT[lz * g.ncx + y];
        rowV[lz * g.ncx + y] = pv;
        rowT[lz * g.ncx + y] = pt;
        pv += cv;
        pt += ct;
    }
    if (lz == g.ncl - 1 && threadIdx.x == 0) {
        unsigned long long V = tv + layerTot[3 * lz], T = tt + layerTot[3 * lz + 1], A = ta + layerTot[3 * lz + 2];
        unsigned long long gV = g.ghost ? layerTot[0] : 0, gT = g.ghost ? layerTot[1] : 0, gA = g.ghost ? layerTot[2] : 0;
        totals[0] = V; totals[1] = T; totals[2] = A; totals[3] = tv;
        totals[4] = gV; totals[5] = gT; totals[6] = gA; totals[7] = 0;
        totals[8] = V - gV; totals[9] = tv - gV; totals[10] = T - gT; totals[11] = A - gA;
        rowV[g.ncl * g.ncx] = (uint32_t)V;
        rowT[g.ncl * g.ncx] = (uint32_t)T;
    }
}

/* vertex-id offset of a slab from the all-gathered per-rank totals {V, V_before_last, T} */
__global__ void k_slab_bases(const unsigned long long *__restrict__ gathered, uint32_t rank, uint32_t ghost,
                             uint32_t *__restrict__ vofs) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long vbase = 0;
        for (uint32_t h = 0; h < rank; ++h) vbase += gathered[3 * h];
        unsigned long long ofs = vbase;
        if (ghost && rank > 0) ofs = vbase - gathered[3 * (rank - 1)] + gathered[3 * (rank - 1) + 1];
        *vofs = (uint32_t)ofs;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* K4: emission                                                                                 */
/* ------------------------------------------------------------------------------------------ */

constexpr int BX = 4;  /* brick: segments of 32 cells in x */
constexpr int BY = 8;  /* rows */
constexpr int BZ = 4;  /* layers */
constexpr int EMIT_THREADS = 256;
constexpr int LIST_CAP = BX * 32 * BY * BZ;
constexpr int NWIN = (BZ + 2) * (BY + 2) * (BX + 1);
constexpr int NDESC = (BZ + 1) * (BY + 1) * (BX + 1);

struct EmitShared {
    uint64_t tri[256];
    uint64_t order[256];
    uint64_t win[NWIN];
    uint4 planes[NDESC];
    uint2 list[LIST_CAP];
    uint32_t dbase[NDESC];
    uint16_t before[256][12];
    uint16_t emask[256];
    uint16_t ownmask[8];
    uint8_t ntri[256];
    uint8_t owner[8][12];
    uint8_t ends[12];
    uint32_t list_n;
    uint32_t work;
    uint32_t ticket;
};

size_t isomc_emit_smem_bytes(uint32_t nws) {
    return sizeof(EmitShared) + (size_t)(BZ + 2) * (BY + 2) * (nws + 2) * sizeof(uint32_t);
}

__device__ __forceinline__ int win_index(int li, int ri, int si) { return (li * (BY + 2) + ri) * (BX + 1) + si; }
__device__ __forceinline__ int desc_index(int li, int ri, int si) { return (li * (BY + 1) + ri) * (BX + 1) + si; }

/* vertices created by cells before the one at descriptor bit p (p = 0: last cell of the previous
 * segment, p = i+1: cell i of this segment) */
__device__ __forceinline__ uint32_t vertex_prefix(const EmitShared &S, int di, uint32_t p) {
    const uint4 c = S.planes[di];
    const uint32_t base = S.dbase[di];
    if (p == 0) return base - ((c.x & 1u) + 2u * (c.y & 1u) + 4u * (c.z & 1u) + 8u * (c.w & 1u));
    const uint32_t lt = (uint32_t)((1ull << p) - 1ull) & ~1u;
    return base + planes_count(c, lt);
}

/* natural cube index of the cell at descriptor bit p whose low corner row is window (li, ri, si) */
__device__ __forceinline__ uint32_t cube_index_at(const EmitShared &S, int li, int ri, int si, uint32_t p) {
    const uint64_t wa = S.win[win_index(li, ri, si)], wb = S.win[win_index(li, ri + 1, si)];
    const uint64_t wc = S.win[win_index(li + 1, ri, si)], wd = S.win[win_index(li + 1, ri + 1, si)];
    return (lo32(wa >> p) & 3u) | (lo32(wb >> p) & 3u) << 2 | (lo32(wc >> p) & 3u) << 4 | (lo32(wd >> p) & 3u) << 6;
}

template <class Src>
__global__ void __launch_bounds__(EMIT_THREADS) k_emit(Src src, Geo g, const uint32_t *__restrict__ signs,
                                                      const uint32_t *__restrict__ segpre,
                                                      const uint32_t *__restrict__ rowPV,
                                                      const uint32_t *__restrict__ rowPT,
                                                      const McTables *__restrict__ tabs,
                                                      const unsigned long long *__restrict__ totals,
                                                      const uint32_t *__restrict__ vofs_ptr, uint32_t *__restrict__ ticket,
                                                      float *__restrict__ xyz, uint32_t *__restrict__ idx,
                                                      unsigned long long cap_v, unsigned long long cap_t) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EmitShared &S = *reinterpret_cast<EmitShared *>(smem_raw);
    uint32_t *s_words = reinterpret_cast<uint32_t *>(smem_raw + sizeof(EmitShared));
    const uint32_t WS = g.nws + 2; /* [0] = pad for segment -1, [1 + w] = word w, [nws + 1] = pad */

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 256; i += EMIT_THREADS) {
        S.tri[i] = tabs->tri[i];
        S.order[i] = tabs->order[i];
        S.emask[i] = tabs->emask[i];
        S.ntri[i] = tabs->ntri[i];
    }
    for (int i = tid; i < 256 * 12; i += EMIT_THREADS) (&S.before[0][0])[i] = (&tabs->before[0][0])[i];
    if (tid < 8) S.ownmask[tid] = tabs->ownmask[tid];
    if (tid < 96) (&S.owner[0][0])[tid] = (&tabs->owner[0][0])[tid];
    if (tid < 12) S.ends[tid] = tabs->ends[tid];

    const uint32_t vofs = *vofs_ptr;
    const uint32_t ghostV = (uint32_t)totals[4], ghostT = (uint32_t)totals[5];
    const uint32_t first_own_layer = g.ghost ? 1u : 0u;
    const uint32_t nby = (g.ncx + BY - 1) / BY, nbz = (g.ncl + BZ - 1) / BZ, nbx = (g.nsegx + BX - 1) / BX;
    const uint32_t n_brick_rows = nby * nbz;

    for (;;) {
        __syncthreads();
        if (tid == 0) S.ticket = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t brow = S.ticket;
        if (brow >= n_brick_rows) break;
        const uint32_t bz = brow / nby, by = brow - bz * nby;
        const int lz0 = (int)(bz * BZ), y0 = (int)(by * BY);

        /* any triangles in these rows? (prefix differences; rows are consecutive in (lz, y) order) */
        if (tid == 0) {
            uint32_t t = 0;
            for (int l = lz0; l < lz0 + BZ && l < (int)g.ncl; ++l) {
                if ((uint32_t)l < first_own_layer) continue;
                uint32_t rb = (uint32_t)l * g.ncx + (uint32_t)y0;
                uint32_t re = (uint32_t)l * g.ncx + min((uint32_t)(y0 + BY), g.ncx);
                t += rowPT[re] - rowPT[rb];
            }
            S.work = t;
        }
        __syncthreads();
        if (S.work == 0) continue;

        /* stage the sign words of the brick row (+1 halo row/layer on each side) */
        for (uint32_t rr = warp; rr < (BZ + 2) * (BY + 2); rr += EMIT_THREADS / 32) {
            const int li = (int)(rr / (BY + 2)), ri = (int)(rr % (BY + 2));
            const int l = lz0 - 1 + li, r = y0 - 1 + ri;
            uint32_t *dst = s_words + (size_t)rr * WS;
            const bool ok = l >= 0 && l < (int)g.nsl && r >= 0 && r < (int)g.N;
            const uint32_t *srcw = signs + ((uint64_t)(ok ? l : 0) * g.N + (ok ? r : 0)) * g.nws;
            for (uint32_t w = lane; w < WS; w += 32) dst[w] = (ok && w >= 1 && w <= g.nws) ? __ldg(srcw + w - 1) : 0u;
        }
        __syncthreads();

        for (uint32_t bx = 0; bx < nbx; ++bx) {
            const int sx0 = (int)(bx * BX);
            /* windows: bit j of win(l, r, s) = inside bit of sample x = 32 s - 1 + j */
            for (int w = tid; w < NWIN; w += EMIT_THREADS) {
                const int si = w % (BX + 1), rr = w / (BX + 1);
                const int s = sx0 - 1 + si;
                uint64_t v = 0;
                if (s >= 0 && s < (int)g.nsegx) {
                    const uint32_t *rw = s_words + (size_t)rr * WS + 1 + s; /* rw[-1] is valid storage */
                    const uint32_t wm = rw[-1], w0 = rw[0], w1 = rw[1];
                    v = (uint64_t)__funnelshift_r(wm, w0, 31) | (uint64_t)__funnelshift_r(w0, w1, 31) << 32;
                }
                S.win[w] = v;
            }
            if (tid == 0) S.list_n = 0;
            __syncthreads();

            /* descriptors: bit planes of "vertices created" per cell + absolute prefix at segment start */
            for (int d = tid; d < NDESC; d += EMIT_THREADS) {
                const int si = d % (BX + 1), ri = (d / (BX + 1)) % (BY + 1), li = d / ((BX + 1) * (BY + 1));
                const int l = lz0 - 1 + li, r = y0 - 1 + ri, s = sx0 - 1 + si;
                uint4 pl = make_uint4(0, 0, 0, 0);
                uint32_t base = 0;
                if (l >= 0 && l < (int)g.ncl && r >= 0 && r < (int)g.ncx && s >= 0 && s < (int)g.nsegx) {
                    const uint64_t wa = S.win[win_index(li, ri, si)], wb = S.win[win_index(li, ri + 1, si)];
                    const uint64_t wc = S.win[win_index(li + 1, ri, si)], wd = S.win[win_index(li + 1, ri + 1, si)];
                    /* bit j <-> cell 32 s - 1 + j; valid cells are [0, ncx) */
                    const int jhi = (int)g.ncx - 32 * s; /* last valid bit */
                    uint32_t vm = valid_mask((uint32_t)min(32, jhi + 1));
                    if (s == 0) vm &= ~1u;
                    pl = owned_planes(lo32(wa), lo32(wa >> 1), lo32(wb), lo32(wb >> 1), lo32(wc), lo32(wc >> 1),
                                      lo32(wd), lo32(wd >> 1), (g.gz0 + (uint32_t)l) == 0, r == 0, s == 0 ? 2u : 0u, vm);
                    const uint32_t row = (uint32_t)l * g.ncx + (uint32_t)r;
                    base = rowPV[row] + (__ldg(segpre + (uint64_t)row * g.nsegx + s) & 0xFFFFu);
                }
                S.planes[d] = pl;
                S.dbase[d] = base;
            }

            /* compaction of the brick's active cells; one warp per 32-cell segment, lane = cell */
            for (int sg = warp; sg < BX * BY * BZ; sg += EMIT_THREADS / 32) {
                const int sl = sg % BX, rl = (sg / BX) % BY, ll = sg / (BX * BY);
                const int l = lz0 + ll, r = y0 + rl, s = sx0 + sl;
                if (l >= (int)g.ncl || (uint32_t)l < first_own_layer || r >= (int)g.ncx || s >= (int)g.nsegx) continue;
                const uint32_t x = (uint32_t)s * 32 + lane;
                const uint32_t ci = cube_index_at(S, ll + 1, rl + 1, sl + 1, lane + 1);
                const bool active = x < g.ncx && ci != 0 && ci != 255;
                const uint32_t nt = active ? S.ntri[ci] : 0;
                const uint32_t am = __ballot_sync(0xFFFFFFFFu, active);
                if (am == 0) continue;
                uint32_t inc = nt;
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
                    uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, dd);
                    if (lane >= (uint32_t)dd) inc += o;
                }
                const uint32_t row = (uint32_t)l * g.ncx + (uint32_t)r;
                const uint32_t tseg = rowPT[row] + (__ldg(segpre + (uint64_t)row * g.nsegx + s) >> 16);
                uint32_t pos0 = 0;
                if (lane == 0) pos0 = atomicAdd(&S.list_n, (uint32_t)__popc(am));
                pos0 = __shfl_sync(0xFFFFFFFFu, pos0, 0);
                if (active) {
                    const uint32_t pos = pos0 + __popc(am & ((1u << lane) - 1u));
                    S.list[pos] = make_uint2(lane | (uint32_t)sl << 5 | (uint32_t)rl << 8 | (uint32_t)ll << 12 | ci << 16,
                                             tseg + inc - nt);
                }
            }
            __syncthreads();

            /* one thread per active cell: create the owned vertices, write the triangles */
            const uint32_t n_act = S.list_n;
            for (uint32_t k = tid; k < n_act; k += EMIT_THREADS) {
                const uint2 ent = S.list[k];
                const uint32_t i = ent.x & 31u, ci = (ent.x >> 16) & 255u;
                const int sl = (int)((ent.x >> 5) & 7u), rl = (int)((ent.x >> 8) & 15u), ll = (int)((ent.x >> 12) & 15u);
                const uint32_t x = (uint32_t)(sx0 + sl) * 32 + i, y = (uint32_t)(y0 + rl), lz = (uint32_t)(lz0 + ll);
                const uint32_t gz = g.gz0 + lz;
                const uint32_t bflags = (x == 0 ? 1u : 0u) | (y == 0 ? 2u : 0u) | (gz == 0 ? 4u : 0u);
                const uint32_t em = S.emask[ci];

                /* ---- vertices this cell creates (mesh.rs:240-251 cache-miss path), in first-appearance order */
                const uint32_t owned = em & S.ownmask[bflags];
                if (owned) {
                    uint32_t slot = vertex_prefix(S, desc_index(ll + 1, rl + 1, sl + 1), i + 1) - ghostV;
                    uint64_t ord = S.order[ci];
                    for (uint32_t rem = em; rem; rem &= rem - 1, ord >>= 4) {
                        const uint32_t e = (uint32_t)ord & 15u;
                        if (!(owned >> e & 1u)) continue;
                        const uint32_t en = S.ends[e];
                        const uint32_t ux = x + (en & 1u), uy = y + (en >> 1 & 1u), uz = en >> 2 & 1u;
                        const uint32_t vx = x + (en >> 4 & 1u), vy = y + (en >> 5 & 1u), vz = en >> 6 & 1u;
                        const float a = src.at(g, ux, uy, lz + uz), b = src.at(g, vx, vy, lz + vz);
                        /* distance.rs:64-69 */
                        const float delta = __fsub_rn(b, a);
                        const float t = (delta == 0.0f) ? 0.5f : __fdiv_rn(-a, delta);
                        const float omt = __fsub_rn(1.0f, t);
                        const float pax = __fmul_rn((float)ux, g.inv), pay = __fmul_rn((float)uy, g.inv);
                        const float paz = __fmul_rn((float)(gz + uz), g.inv);
                        const float pbx = __fmul_rn((float)vx, g.inv), pby = __fmul_rn((float)vy, g.inv);
                        const float pbz = __fmul_rn((float)(gz + vz), g.inv);
                        if (slot < cap_v) {
                            float *o = xyz + (uint64_t)slot * 3;
                            o[0] = __fadd_rn(__fmul_rn(pax, omt), __fmul_rn(pbx, t));
                            o[1] = __fadd_rn(__fmul_rn(pay, omt), __fmul_rn(pby, t));
                            o[2] = __fadd_rn(__fmul_rn(paz, omt), __fmul_rn(pbz, t));
                        }
                        ++slot;
                    }
                }

                /* ---- triangles (march_cube, marching_cubes_impl.rs:102-117), ids by edge ownership */
                uint64_t tl = S.tri[ci];
                uint32_t tslot = ent.y - ghostT;
                const uint32_t nt = S.ntri[ci];
                for (uint32_t t = 0; t < nt; ++t, ++tslot) {
                    uint32_t ids[3];
#pragma unroll
                    for (int q = 0; q < 3; ++q, tl >>= 4) {
                        const uint32_t e = (uint32_t)tl & 15u;
                        const uint32_t ow = S.owner[bflags][e];
                        const int dx = ow & 1, dy = ow >> 1 & 1, dz = ow >> 2 & 1;
                        const uint32_t e2 = ow >> 4;
                        const uint32_t p = i + 1 - dx;
                        const uint32_t oci = cube_index_at(S, ll + 1 - dz, rl + 1 - dy, sl + 1, p);
                        const uint32_t ob = ((x - dx) == 0 ? 1u : 0u) | ((y - dy) == 0 ? 2u : 0u) | ((gz - dz) == 0 ? 4u : 0u);
                        const uint32_t rank = __popc((uint32_t)S.before[oci][e2] & (uint32_t)S.ownmask[ob]);
                        ids[q] = vofs + vertex_prefix(S, desc_index(ll + 1 - dz, rl + 1 - dy, sl + 1), p) + rank;
                    }
                    if (tslot < cap_t) {
                        uint32_t *o = idx + (uint64_t)tslot * 3;
                        o[0] = ids[0]; o[1] = ids[1]; o[2] = ids[2];
                    }
                }
            }
            __syncthreads();
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* debug / parity kernels                                                                       */
/* ------------------------------------------------------------------------------------------ */

__global__ void k_cube_indices(Geo g, const uint32_t *__restrict__ signs, const McTables *__restrict__ tabs,
                               uint8_t *__restrict__ out) {
    const uint64_t ncell = (uint64_t)g.ncl * g.ncx * g.ncx;
    for (uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; c < ncell; c += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t x = (uint32_t)(c % g.ncx), y = (uint32_t)((c / g.ncx) % g.ncx), lz = (uint32_t)(c / ((uint64_t)g.ncx * g.ncx));
        uint32_t ci = 0;
        for (int n = 0; n < 8; ++n) {
            const uint32_t sx = x + (n & 1), sy = y + (n >> 1 & 1), sz = lz + (n >> 2 & 1);
            const uint32_t w = signs[((uint64_t)sz * g.N + sy) * g.nws + (sx >> 5)];
            ci |= (w >> (sx & 31) & 1u) << n;
        }
        out[c] = tabs->ref_of_nat[ci];
    }
}

__global__ void k_sample_sdf(SdfProgram prog, const float *__restrict__ xyz, uint64_t n, float *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = sdf_eval(prog, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}

/* ------------------------------------------------------------------------------------------ */
/* synthetic fields (bench / tests only)                                          */
/* ------------------------------------------------------------------------------------------ */

__global__ void k_synth(SynthParams sp, uint32_t size, float inv, uint32_t z_first, uint32_t n_layers, float *__restrict__ out) {
    const uint64_t n = (uint64_t)n_layers * size * size;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t x = (uint32_t)(i % size), y = (uint32_t)((i / size) % size), z = z_first + (uint32_t)(i / ((uint64_t)size * size));
        const float px = (float)x * inv, py = (float)y * inv, pz = (float)z * inv;
        float f = 0.0f;
        if (sp.kind == ISOMC_FIELD_FBM) {
            for (int w = 0; w < 20; ++w)
                f += sp.amp[w] * sinf(sp.freq[w] * (sp.dx[w] * px + sp.dy[w] * py + sp.dz[w] * pz) + sp.ph[w]);
        } else if (sp.kind == ISOMC_FIELD_GYROID) {
            const float k = 6.283185307179586f * 8.0f;
            const float X = k * px, Y = k * py, Z = k * pz;
            f = sinf(X) * cosf(Y) + sinf(Y) * cosf(Z) + sinf(Z) * cosf(X);
        } else {
            f = 1e30f;
            for (int s = 0; s < 64; ++s) {
                const float ddx = px - sp.cx[s], ddy = py - sp.cy[s], ddz = pz - sp.cz[s];
                f = fminf(f, sqrtf(ddx * ddx + ddy * ddy + ddz * ddz) - sp.r[s]);
            }
        }
        out[i] = f;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* launchers                                                                                    */
/* ------------------------------------------------------------------------------------------ */

static inline uint32_t grid_for(uint64_t warps_needed, int sms, int warps_per_block, int blocks_per_sm) {
    uint64_t blocks = (warps_needed + warps_per_block - 1) / warps_per_block;
    uint64_t cap = (uint64_t)sms * blocks_per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (uint32_t)blocks;
}

cudaError_t isomc_launch_sign_grid(const Geo &g, const float *d_grid, uint32_t *signs, int sms, cudaStream_t st) {
    GridSrc src{d_grid};
    k_sign<GridSrc><<<grid_for((uint64_t)g.nsl * g.N, sms, 8, 8), 256, 0, st>>>(src, g, signs);
    return cudaGetLastError();
}
cudaError_t isomc_launch_sign_sdf(const Geo &g, const SdfProgram &prog, uint32_t *signs, int sms, cudaStream_t st) {
    SdfSrc src{prog};
    k_sign<SdfSrc><<<grid_for((uint64_t)g.nsl * g.N, sms, 8, 8), 256, 0, st>>>(src, g, signs);
    return cudaGetLastError();
}
cudaError_t isomc_launch_count(const Geo &g, const uint32_t *signs, const McTables *tabs, uint32_t *segpre,
                               uint32_t *rowV, uint32_t *rowT, uint32_t *rowA, unsigned long long *layerTot,
                               int sms, cudaStream_t st) {
    k_count<<<grid_for((uint64_t)g.ncl * g.ncx, sms, 8, 8), 256, 0, st>>>(g, signs, tabs, segpre, rowV, rowT, rowA, layerTot);
    return cudaGetLastError();
}
cudaError_t isomc_launch_scan(const Geo &g, uint32_t *rowV, uint32_t *rowT, const unsigned long long *layerTot,
                              unsigned long long *totals, cudaStream_t st) {
    k_scan_rows<<<g.ncl, 256, 0, st>>>(g, rowV, rowT, layerTot, totals);
    return cudaGetLastError();
}
cudaError_t isomc_launch_slab_bases(const unsigned long long *gathered, uint32_t rank, uint32_t ghost, uint32_t *vofs,
                                    cudaStream_t st) {
    k_slab_bases<<<1, 32, 0, st>>>(gathered, rank, ghost, vofs);
    return cudaGetLastError();
}

template <class Src>
static cudaError_t launch_emit(Src src, const Geo &g, const uint32_t *signs, const uint32_t *segpre,
                               const uint32_t *rowPV, const uint32_t *rowPT, const McTables *tabs,
                               const unsigned long long *totals, const uint32_t *vofs, uint32_t *ticket, float *xyz,
                               uint32_t *idx, uint64_t cap_v, uint64_t cap_t, int sms, cudaStream_t st) {
    const size_t smem = isomc_emit_smem_bytes(g.nws);
    static size_t configured = 0; /* per template instantiation */
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(k_emit<Src>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    int per_sm = 1;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_emit<Src>, EMIT_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const uint32_t nby = (g.ncx + BY - 1) / BY, nbz = (g.ncl + BZ - 1) / BZ;
    uint64_t blocks = (uint64_t)nby * nbz;
    if (blocks > (uint64_t)sms * per_sm) blocks = (uint64_t)sms * per_sm;
    k_emit<Src><<<(uint32_t)blocks, EMIT_THREADS, smem, st>>>(src, g, signs, segpre, rowPV, rowPT, tabs, totals, vofs,
                                                              ticket, xyz, idx, cap_v, cap_t);
    return cudaGetLastError();
}

cudaError_t isomc_launch_emit_grid(const Geo &g, const float *d_grid, const uint32_t *signs, const uint32_t *segpre,
                                   const uint32_t *rowPV, const uint32_t *rowPT, const McTables *tabs,
                                   const unsigned long long *totals, const uint32_t *vofs, uint32_t *ticket, float *xyz,
                                   uint32_t *idx, uint64_t cap_v, uint64_t cap_t, int sms, cudaStream_t st) {
    return launch_emit(GridSrc{d_grid}, g, signs, segpre, rowPV, rowPT, tabs, totals, vofs, ticket, xyz, idx, cap_v, cap_t, sms, st);
}
cudaError_t isomc_launch_emit_sdf(const Geo &g, const SdfProgram &prog, const uint32_t *signs, const uint32_t *segpre,
                                  const uint32_t *rowPV, const uint32_t *rowPT, const McTables *tabs,
                                  const unsigned long long *totals, const uint32_t *vofs, uint32_t *ticket, float *xyz,
                                  uint32_t *idx, uint64_t cap_v, uint64_t cap_t, int sms, cudaStream_t st) {
    return launch_emit(SdfSrc{prog}, g, signs, segpre, rowPV, rowPT, tabs, totals, vofs, ticket, xyz, idx, cap_v, cap_t, sms, st);
}

cudaError_t isomc_launch_cube_indices(const Geo &g, const uint32_t *signs, const McTables *tabs, uint8_t *out, int sms,
                                      cudaStream_t st) {
    k_cube_indices<<<sms * 8, 256, 0, st>>>(g, signs, tabs, out);
    return cudaGetLastError();
}
cudaError_t isomc_launch_sample_sdf(const SdfProgram &prog, const float *xyz, uint64_t n, float *out, cudaStream_t st) {
    k_sample_sdf<<<(uint32_t)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256), 256, 0, st>>>(prog, xyz, n, out);
    return cudaGetLastError();
}
cudaError_t isomc_launch_synth(const SynthParams &sp, uint32_t size, uint32_t z_first, uint32_t n_layers, float *out,
                               int sms, cudaStream_t st) {
    const float inv = 1.0f / (float)(size - 1);
    k_synth<<<sms * 8, 256, 0, st>>>(sp, size, inv, z_first, n_layers, out);
    return cudaGetLastError();
}
