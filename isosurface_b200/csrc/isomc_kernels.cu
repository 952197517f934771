/*
 * isomc_kernels.cu -- sm_100a kernels of the MarchingCubes extract path: the two kernels every path shares, and the
 * older "brick" form of output sizing and emission (ISOMC_EMIT=brick; the default is the active-cell-list form in
 * isomc_list_kernels.cu, which replaced it: profiles/r01_history.md).
 *
 * Shared (one stream, no host round trip in steady state):
 *
 *   K1 k_sign_vec4 / k_sign<Src>
 *                    sample -> inside bit.  One bit per lattice point (`!(v > 0)`,
 *                    marching_cubes_impl.rs:32 / distance.rs:52-54), 32 per word.  Grid sources
 *                    stream every f32 exactly once as float4s (HBM bound); implicit sources evaluate
 *                    the SDF program instead of loading (Directed: outside iff any component > 0, distance.rs:77-80).
 *   K3 k_scan_rows   exclusive scan over cell rows in (z, y) order (+ totals, list marks); causal in z.
 *
 * Brick path:
 *
 *   K2 k_count       warp-autonomous, lane per 32-cell segment: bit-parallel classification.
 *                    Crossed-edge masks are XORs of sign words, the "edges this cell creates" count
 *                    is a bit-sliced sum of the owned masks, triangle counts come from ntri[ci'] for
 *                    active cells only.  Writes within-row exclusive prefixes and row totals.
 *   K4 k_emit        warp-autonomous bricks of 32x4x4 cells: flat cell and triangle lists, 16-bit
 *                    id planes in shared memory (edge ownership replaces the reference's HashMap
 *                    index cache, index_cache.rs / mesh.rs:240-251); writes u32 indices in reference
 *                    order and one 12-byte descriptor per created vertex.
 *   K5 k_vertex<Src> descriptor -> position (distance.rs:64-69), in place.
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

/* generic path: any source, any size/alignment; one lane per sample, ballot per 32 samples */
template <class Src>
__global__ void __launch_bounds__(256) k_sign(Src src, Geo g, uint32_t *__restrict__ signs, uint32_t row0, uint32_t row1) {
    constexpr int U = 8;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t row = row0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < row1; row += nwarps) {
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

/* dense-grid fast path (N % 4 == 0, 16-byte aligned base): every lane streams float4s (512
 * contiguous bytes per warp instruction), turns them into a 4-bit nibble and the nibbles of 8
 * neighbouring lanes are OR-combined with 3 shuffles into one 32-sample word. */
template <int U>
__global__ void __launch_bounds__(256) k_sign_vec4(const float4 *__restrict__ grid4, Geo g, uint32_t *__restrict__ signs,
                                                   uint32_t row0, uint32_t row1) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    const uint32_t n4 = g.N >> 2;            /* float4 per sample row */
    const uint32_t steps = (n4 + 31) >> 5;   /* 128-sample chunks per row */
    const uint32_t sh = (lane & 7u) * 4u;
    for (uint32_t row = row0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < row1; row += nwarps) {
        const float4 *rp = grid4 + (uint64_t)row * n4;
        uint32_t *out = signs + (uint64_t)row * g.nws;
        for (uint32_t c0 = 0; c0 < steps; c0 += U) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t i4 = (c0 + u) * 32 + lane;
                v[u] = (i4 < n4) ? __ldg(rp + i4) : make_float4(1.0f, 1.0f, 1.0f, 1.0f);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                uint32_t nib = (!(v[u].x > 0.0f) ? 1u : 0u) | (!(v[u].y > 0.0f) ? 2u : 0u) |
                               (!(v[u].z > 0.0f) ? 4u : 0u) | (!(v[u].w > 0.0f) ? 8u : 0u);
                uint32_t w = nib << sh;
                w |= __shfl_xor_sync(0xFFFFFFFFu, w, 1);
                w |= __shfl_xor_sync(0xFFFFFFFFu, w, 2);
                w |= __shfl_xor_sync(0xFFFFFFFFu, w, 4);
                const uint32_t widx = (c0 + u) * 4 + (lane >> 3);
                if ((lane & 7u) == 0 && widx < g.nws) out[widx] = w;
            }
        }
        /* padding words past the last 128-sample chunk */
        const uint32_t done = ((steps + U - 1) / U) * U * 4;
        if (done + lane < g.nws) out[done + lane] = 0u;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* K2: per-segment counts, within-row prefixes, row totals                                      */
/* ------------------------------------------------------------------------------------------ */

/* Warp-autonomous (no block barriers in the loop): a lane owns one 32-cell segment; a warp covers
 * 32/G consecutive cell rows of G = pow2 >= nsegx segments each (one row in 32-segment chunks when
 * nsegx > 32) and scans them with width-G shuffles.  Vertex counts are pure bit-parallel work on the
 * sign words; triangle counts walk the (few) active cells of the segment. */
struct SegCounts { uint32_t nv, nt, na; };

__device__ __forceinline__ SegCounts count_segment(const Geo &g, const uint32_t *__restrict__ signs, const uint8_t *s_ntri,
                                                   uint32_t row, uint32_t lz, uint32_t s) {
    SegCounts c = {0u, 0u, 0u};
    const uint32_t y = row - lz * g.ncx;
    const uint32_t *r00 = signs + (uint64_t)(row + lz) * g.nws + s; /* sample row lz*N + y = row + lz */
    const uint32_t *r01 = r00 + g.nws, *r10 = r00 + (uint64_t)g.N * g.nws, *r11 = r10 + g.nws;
    const uint32_t a0 = __ldg(r00), a1 = __ldg(r00 + 1), b0 = __ldg(r01), b1 = __ldg(r01 + 1);
    const uint32_t c0 = __ldg(r10), c1 = __ldg(r10 + 1), d0 = __ldg(r11), d1 = __ldg(r11 + 1);
    const uint32_t all_or = a0 | b0 | c0 | d0 | ((a1 | b1 | c1 | d1) & 1u);
    const uint32_t all_and = a0 & b0 & c0 & d0;
    if ((all_or == 0u) || (all_and == 0xFFFFFFFFu && (a1 & b1 & c1 & d1 & 1u))) return c;
    const uint32_t an = __funnelshift_r(a0, a1, 1), bn = __funnelshift_r(b0, b1, 1);
    const uint32_t cn = __funnelshift_r(c0, c1, 1), dn = __funnelshift_r(d0, d1, 1);
    const uint32_t vm = valid_mask(g.ncx - s * 32);
    uint32_t act = active_mask(a0, an, b0, bn, c0, cn, d0, dn, vm);
    if (act == 0) return c;
    const uint4 pl = owned_planes(a0, an, b0, bn, c0, cn, d0, dn, (g.gz0 + lz) == 0, y == 0, s == 0 ? 1u : 0u, vm);
    c.nv = planes_count(pl, 0xFFFFFFFFu);
    c.na = __popc(act);
    while (act) {
        const uint32_t i = __ffs(act) - 1;
        act &= act - 1;
        const uint32_t ci = (__funnelshift_r(a0, a1, i) & 3u) | (__funnelshift_r(b0, b1, i) & 3u) << 2 |
                            (__funnelshift_r(c0, c1, i) & 3u) << 4 | (__funnelshift_r(d0, d1, i) & 3u) << 6;
        c.nt += s_ntri[ci];
    }
    return c;
}

__global__ void __launch_bounds__(256) k_count(Geo g, const uint32_t *__restrict__ signs, const McTables *__restrict__ tabs,
                                               uint32_t *__restrict__ segpre, uint32_t *__restrict__ rowV,
                                               uint32_t *__restrict__ rowT, uint32_t *__restrict__ rowA,
                                               unsigned long long *__restrict__ layerTot, uint32_t gshift, uint32_t row0,
                                               uint32_t row1) {
    __shared__ uint8_t s_ntri[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_ntri[i] = tabs->ntri[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwarps = gridDim.x * (blockDim.x >> 5);
    const uint32_t nrows = row1;
    if (g.nsegx <= 32) {
        const uint32_t G = 1u << gshift, rpw = 32u >> gshift, sub = lane >> gshift, s = lane & (G - 1);
        const uint32_t niter = (row1 - row0 + rpw - 1) / rpw;
        for (uint32_t it = gwarp; it < niter; it += nwarps) {
            const uint32_t row = row0 + it * rpw + sub;
            const bool valid = row < nrows && s < g.nsegx;
            const uint32_t lz = (row < nrows ? row : row0) / g.ncx;
            SegCounts c = {0u, 0u, 0u};
            if (valid) c = count_segment(g, signs, s_ntri, row, lz, s);
            const uint32_t pk = c.nv | c.nt << 16; /* 16-bit fields: row totals < 65536 for size <= 8192 */
            uint32_t inc = pk, acta = c.na;
            for (uint32_t d = 1; d < G; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d, G);
                if (s >= d) inc += o;
                acta += __shfl_xor_sync(0xFFFFFFFFu, acta, d, G);
            }
            if (valid) segpre[(uint64_t)row * g.nsegx + s] = inc - pk;
            const uint32_t tot = __shfl_sync(0xFFFFFFFFu, inc, G - 1, G);
            if (s == 0 && row < nrows) {
                const uint32_t tv = tot & 0xFFFFu, tt = tot >> 16;
                rowV[row] = tv; rowT[row] = tt; rowA[row] = acta;
                if (tot) {
                    atomicAdd(&layerTot[3 * lz + 0], (unsigned long long)tv);
                    atomicAdd(&layerTot[3 * lz + 1], (unsigned long long)tt);
                    atomicAdd(&layerTot[3 * lz + 2], (unsigned long long)acta);
                }
            }
        }
    } else {
        for (uint32_t row = row0 + gwarp; row < nrows; row += nwarps) {
            const uint32_t lz = row / g.ncx;
            uint32_t carry = 0, acta = 0;
            for (uint32_t s0 = 0; s0 < g.nsegx; s0 += 32) {
                const uint32_t s = s0 + lane;
                SegCounts c = {0u, 0u, 0u};
                if (s < g.nsegx) c = count_segment(g, signs, s_ntri, row, lz, s);
                const uint32_t pk = c.nv | c.nt << 16;
                uint32_t inc = pk;
                acta += c.na;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                    if (lane >= (uint32_t)d) inc += o;
                }
                if (s < g.nsegx) segpre[(uint64_t)row * g.nsegx + s] = carry + inc - pk;
                carry += __shfl_sync(0xFFFFFFFFu, inc, 31);
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) acta += __shfl_xor_sync(0xFFFFFFFFu, acta, d);
            if (lane == 0) {
                const uint32_t tv = carry & 0xFFFFu, tt = carry >> 16;
                rowV[row] = tv; rowT[row] = tt; rowA[row] = acta;
                if (carry) {
                    atomicAdd(&layerTot[3 * lz + 0], (unsigned long long)tv);
                    atomicAdd(&layerTot[3 * lz + 1], (unsigned long long)tt);
                    atomicAdd(&layerTot[3 * lz + 2], (unsigned long long)acta);
                }
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
 *   [3] V prefix at the start of the last cell layer  [4..6] V, T, active of the ghost layer  [7] list blocks asked for
 *   [8] vertices owned  [9] owned vertices created before the last cell layer  [10] triangles owned */
__global__ void __launch_bounds__(256) k_scan_rows(Geo g, uint32_t *__restrict__ rowV, uint32_t *__restrict__ rowT,
                                                   const unsigned long long *__restrict__ layerTot,
                                                   unsigned long long *__restrict__ totals,
                                                   const uint32_t *__restrict__ list_ctr, uint32_t *__restrict__ list_mark,
                                                   uint32_t *__restrict__ chunk_end, uint32_t lz_first) {
    __shared__ unsigned long long s_w[8];
    if (list_mark && blockIdx.x == 0 && threadIdx.x == 0) *list_mark = *list_ctr; /* list blocks handed out up to this z-chunk */
    __shared__ unsigned long long s_base[2];
    const uint32_t lz = lz_first + blockIdx.x;
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
    if (chunk_end && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) { /* ids / slots below the end of this z-chunk */
        chunk_end[0] = (uint32_t)(tv + layerTot[3 * lz]);
        chunk_end[1] = (uint32_t)(tt + layerTot[3 * lz + 1]);
    }
    if (lz == g.ncl - 1 && threadIdx.x == 0) {
        unsigned long long V = tv + layerTot[3 * lz], T = tt + layerTot[3 * lz + 1], A = ta + layerTot[3 * lz + 2];
        unsigned long long gV = g.ghost ? layerTot[0] : 0, gT = g.ghost ? layerTot[1] : 0, gA = g.ghost ? layerTot[2] : 0;
        totals[0] = V; totals[1] = T; totals[2] = A; totals[3] = tv;
        totals[4] = gV; totals[5] = gT; totals[6] = gA;
        totals[7] = list_ctr ? *list_ctr : 0u; /* list blocks the count asked for (active-cell-list path) */
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
/*
 * Warp-autonomous: every warp pulls "strips" (BY rows x BZ layers x all x) from a ticket counter and
 * walks the strip's bricks of 32 x BY x BZ cells that contain triangles.  No block barriers; each
 * warp owns a private slice of shared memory.  A brick plus its one-cell halo on the low side of
 * every axis (the cells that created the vertices the brick's triangles refer to) is the "region".
 * Everything variable-length is flattened before it is processed:
 *
 *   P1   one lane per region row (= one 32-cell segment): crossed-edge masks and bit-sliced
 *        "vertices created" counts from the sign words; list positions by warp scans; the active
 *        mask is expanded into a flat cell list (cube index + triangles of the earlier cells).
 *   P2   one lane per active cell of the region: id of the first vertex it creates and the ids of
 *        all edges it creates -> shared-memory id planes (one per edge axis, indexed by the cell that
 *        would create the edge in an unbounded grid; 16 bit, relative to the creating row).  Own
 *        cells also write one 12-byte *vertex descriptor* (creator cell + edge) into the slot of each
 *        vertex they create (k_vertex turns descriptors into positions) and list their triangles.
 *   B    one lane per triangle: three id-plane lookups, one 12-byte store (u32 x 3).
 */

constexpr int BY = 4;  /* brick rows */
constexpr int BZ = 4;  /* brick layers */
constexpr int EMIT_WARPS = 6;
constexpr int EMIT_THREADS = EMIT_WARPS * 32;
constexpr int RX = 33, RY = BY + 1, RZ = BZ + 1; /* region extents in cells */
constexpr int NREGION = RX * RY * RZ;
constexpr int NTASK = RZ * RY;                   /* region rows = P1 tasks, one per lane */
constexpr int NROWS_OWN = BY * BZ;
constexpr int TRI_CAP = 1024;                    /* triangles listed per pass; one row (32 * 5) always fits */
constexpr int CELL_CAP = 512;                    /* active cells per pass; one region layer (5 * 33) always fits */
static_assert(NTASK <= 32 && NREGION <= 1024 && NROWS_OWN <= 32, "one task per lane; list entry bit fields");

struct __align__(16) SegDesc {
    uint32_t p0, p1, p2, p3; /* bit planes of "vertices created" per cell */
    uint32_t vbase;          /* id (before vofs) of the first vertex created in this segment */
    uint32_t tseg;           /* slot of the first triangle of this segment */
    uint32_t cpos_tch;       /* cell-list start | triangle-list start << 16 */
    uint32_t info;           /* y==0 | z==0 << 1 | listed << 2 | (s == 0) << 3 | inside bits of sample 32s-1 in the 4 rows << 4 |
                                inside bit of sample 32s in the 4 rows << 8 */
};

struct WarpShared {
    SegDesc seg[NTASK];
    uint32_t cellmap[CELL_CAP];   /* task | i << 5 | x-halo << 10 | ci' << 11 | triangles of earlier cells of the segment << 19 */
    uint32_t trilist[TRI_CAP];    /* region pos (10) | ci' << 10 | t << 18 | task << 21 */
    uint32_t rowbase[NTASK];      /* id (before vofs) at the start of each region row; virtual rows: clamped row */
    uint16_t plane[3 * NREGION];  /* id of the x / y / z edge created by (virtual) cell, relative to rowbase of its row */
    uint16_t pad[3];
};

struct EmitShared {
    uint64_t tri[256];
    WarpShared w[EMIT_WARPS];
    uint16_t emask[256];
    uint16_t ownmask[8];
    int16_t offs[12];   /* plane index of edge e seen from a cell at region pos cp: cp + offs[e] */
    int32_t look[12];   /* (offs[e] + 4096) | rowback[e] << 16: one load per lookup in phase B */
    uint8_t rowback[12];/* region rows between a cell and the (virtual) creator of its edge e */
    uint8_t bstep[12];  /* dx | dy << 1 | dz << 2 of that step */
    uint8_t ntri[256];
    uint8_t rank3[256];
};

size_t isomc_emit_smem_bytes(uint32_t) { return sizeof(EmitShared); }

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t lane, uint32_t &total) {
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= (uint32_t)d) inc += o;
    }
    total = __shfl_sync(0xFFFFFFFFu, inc, 31);
    return inc - v;
}

__global__ void __launch_bounds__(EMIT_THREADS, 3) k_emit(Geo g, const uint32_t *__restrict__ signs,
                                                      const uint32_t *__restrict__ segpre,
                                                      const uint32_t *__restrict__ rowPV,
                                                      const uint32_t *__restrict__ rowPT,
                                                      const McTables *__restrict__ tabs,
                                                      const unsigned long long *__restrict__ layerTot,
                                                      const uint32_t *__restrict__ vofs_ptr, uint32_t *__restrict__ ticket,
                                                      uint32_t *__restrict__ vdesc, uint32_t *__restrict__ idx,
                                                      unsigned long long cap_v, unsigned long long cap_t, uint32_t strip0,
                                                      uint32_t strip1) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EmitShared &S = *reinterpret_cast<EmitShared *>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 256; i += EMIT_THREADS) {
        S.tri[i] = tabs->tri[i];
        S.emask[i] = tabs->emask[i];
        S.ntri[i] = tabs->ntri[i];
        S.rank3[i] = tabs->rank3[i];
    }
    if (tid < 8) S.ownmask[tid] = tabs->ownmask[tid];
    if (tid < 12) {
        const uint32_t ow = tabs->owner[0][tid]; /* unbounded grid: every step back is allowed */
        const uint32_t e2 = ow >> 4;             /* always one of 5 (y edge), 6 (x edge), 10 (z edge) */
        const int back = (int)((ow & 1) + (ow >> 1 & 1) * RX + (ow >> 2 & 1) * RX * RY);
        const int axis = e2 == 6 ? 0 : e2 == 5 ? 1 : 2;
        S.bstep[tid] = (uint8_t)(ow & 7);
        S.rowback[tid] = (uint8_t)((ow >> 1 & 1) + (ow >> 2 & 1) * RY);
        S.offs[tid] = (int16_t)(axis * NREGION - back);
        S.look[tid] = (int32_t)((axis * NREGION - back + 4096) | ((ow >> 1 & 1) + (ow >> 2 & 1) * RY) << 16);
    }
    __syncthreads(); /* the only block barrier: tables ready */
    WarpShared &W = S.w[warp];

    const uint32_t vofs = *vofs_ptr;
    const uint32_t ghostV = g.ghost ? (uint32_t)layerTot[0] : 0u, ghostT = g.ghost ? (uint32_t)layerTot[1] : 0u;
    const int first_own_layer = g.ghost ? 1 : 0;
    const uint32_t nby = (g.ncx + BY - 1) / BY;
    /* P1 task of this lane: region row (t_rz, t_ry) */
    const bool has_task = lane < NTASK;
    const int t_rz = (int)(lane / RY), t_ry = (int)(lane % RY);
    const bool t_own_pos = has_task && t_rz >= 1 && t_ry >= 1;
    const int t_q = (t_rz - 1) * BY + (t_ry - 1); /* own-row ordinal (valid when t_own_pos) */

    for (;;) {
        uint32_t strip = 0;
        if (lane == 0) strip = strip0 + atomicAdd(ticket, 1u);
        strip = __shfl_sync(0xFFFFFFFFu, strip, 0);
        if (strip >= strip1) break;
        const uint32_t bz = strip / nby, by = strip - bz * nby;
        const int lz0 = (int)(bz * BZ), y0 = (int)(by * BY);

        /* region rows of this lane: prefixes at the row start; any triangles in the own rows? */
        const int l = lz0 - 1 + t_rz, r = y0 - 1 + t_ry;
        const bool row_ok = has_task && l >= 0 && l < (int)g.ncl && r >= 0 && r < (int)g.ncx;
        const uint32_t row = row_ok ? (uint32_t)l * g.ncx + (uint32_t)r : 0u;
        uint32_t pv = 0, pt = 0, ptn = 0;
        if (row_ok) { pv = rowPV[row]; pt = rowPT[row]; ptn = rowPT[row + 1]; }
        const bool own_row = row_ok && t_own_pos && l >= first_own_layer;
        if (__ballot_sync(0xFFFFFFFFu, own_row && ptn != pt) == 0) continue;
        /* id base of every region row; rows below the grid (virtual creators of boundary edges) use the clamped row */
        {
            const int cz = max(t_rz, (lz0 == 0) ? 1 : 0), cy = max(t_ry, (y0 == 0) ? 1 : 0);
            const uint32_t b = __shfl_sync(0xFFFFFFFFu, pv, has_task ? cz * RY + cy : 0);
            if (has_task) W.rowbase[lane] = b + vofs; /* global id base of the row */
        }
        const uint32_t *my_signs = signs + ((uint64_t)(row_ok ? l : 0) * g.N + (uint32_t)(row_ok ? r : 0)) * g.nws;
        const uint32_t *my_sp = segpre + (uint64_t)row * g.nsegx;
        const uint32_t gz = g.gz0 + (uint32_t)max(l, 0);

        for (uint32_t s0 = 0; s0 < g.nsegx; s0 += 32) {
            /* bricks (= segments s0 + lane) with triangles: differences of the within-row prefixes of the own rows */
            uint32_t has = 0;
            {
                const uint32_t sb = s0 + lane;
#pragma unroll 4
                for (int q = 0; q < NROWS_OWN; ++q) {
                    const int ql = lz0 + q / BY, qr = y0 + q % BY;
                    const bool ok = ql >= first_own_layer && ql < (int)g.ncl && qr < (int)g.ncx && sb < g.nsegx;
                    const uint32_t qrow = ok ? (uint32_t)ql * g.ncx + (uint32_t)qr : 0u;
                    const uint32_t *qp = segpre + (uint64_t)qrow * g.nsegx + (ok ? sb : 0u);
                    const uint32_t t0 = __ldg(qp) >> 16;
                    const uint32_t t1 = (ok && sb + 1 < g.nsegx) ? __ldg(qp + 1) >> 16 : rowPT[qrow + 1] - rowPT[qrow];
                    if (ok) has |= t1 ^ t0;
                }
            }
            uint32_t todo = __ballot_sync(0xFFFFFFFFu, has != 0);
            if (todo == 0) continue;
            /* software pipeline over the bricks with work: the sign words / prefixes of the NEXT brick are loaded
             * while the current one is processed, so their L2 latency is hidden behind a whole brick of work */
            uint32_t n_s = s0 + (uint32_t)__ffs(todo) - 1;
            todo &= todo - 1;
            uint32_t na0 = 0, na1 = 0, nb0 = 0, nb1 = 0, nc0 = 0, nc1 = 0, nd0 = 0, nd1 = 0, nsp = 0, nspn = 0, npw = 0;
#define ISOMC_BRICK_LOADS(S_)                                                                                              \
    if (row_ok) {                                                                                                          \
        const uint32_t *wa = my_signs + (S_), *wb = wa + g.nws, *wc = wa + (size_t)g.N * g.nws, *wd = wc + g.nws;          \
        na0 = __ldg(wa); na1 = __ldg(wa + 1); nb0 = __ldg(wb); nb1 = __ldg(wb + 1);                                        \
        nc0 = __ldg(wc); nc1 = __ldg(wc + 1); nd0 = __ldg(wd); nd1 = __ldg(wd + 1);                                        \
        nsp = __ldg(my_sp + (S_));                                                                                         \
        nspn = (S_) + 1 < g.nsegx ? __ldg(my_sp + (S_) + 1) >> 16 : ptn - pt;                                              \
        npw = 0;                                                                                                           \
        if ((S_) > 0) npw = (__ldg(wa - 1) >> 31) | (__ldg(wb - 1) >> 31) << 1 | (__ldg(wc - 1) >> 31) << 2 | (__ldg(wd - 1) >> 31) << 3; \
    }
            ISOMC_BRICK_LOADS(n_s)
            for (bool more = true; more;) {
                const uint32_t s = n_s;
                const uint32_t a0 = na0, a1 = na1, b0 = nb0, b1 = nb1, c0 = nc0, c1 = nc1, d0 = nd0, d1 = nd1;
                const uint32_t sp = nsp, spn = nspn, prevbits = npw;
                more = todo != 0;
                if (more) {
                    n_s = s0 + (uint32_t)__ffs(todo) - 1;
                    todo &= todo - 1;
                    ISOMC_BRICK_LOADS(n_s)
                }

                uint32_t act = 0;
                uint4 pl = make_uint4(0, 0, 0, 0);
                {
                    const uint32_t all_or = a0 | b0 | c0 | d0 | ((a1 | b1 | c1 | d1) & 1u);
                    const uint32_t all_and = a0 & b0 & c0 & d0;
                    const bool uniform = (all_or == 0u) || (all_and == 0xFFFFFFFFu && (a1 & b1 & c1 & d1 & 1u));
                    if (row_ok && !uniform) {
                        const uint32_t an = __funnelshift_r(a0, a1, 1), bn = __funnelshift_r(b0, b1, 1);
                        const uint32_t cn = __funnelshift_r(c0, c1, 1), dn = __funnelshift_r(d0, d1, 1);
                        const uint32_t vm = valid_mask(g.ncx - s * 32);
                        act = active_mask(a0, an, b0, bn, c0, cn, d0, dn, vm);
                        if (act) pl = owned_planes(a0, an, b0, bn, c0, cn, d0, dn, gz == 0, r == 0, s == 0 ? 1u : 0u, vm);
                    }
                }
                const uint32_t nextbits = (a0 & 1u) | (b0 & 1u) << 1 | (c0 & 1u) << 2 | (d0 & 1u) << 3;
                const uint32_t both = prevbits | nextbits << 4;
                const uint32_t hx = (row_ok && s > 0 && both != 0 && both != 255) ? 1u : 0u; /* x-halo cell active */
                const uint32_t nt_seg = spn - (sp >> 16);

                /* pass plan.  Usually everything fits one pass: fill the id planes for the whole region and list all
                 * own rows.  A dense brick is done in pieces: one fill pass per region layer, then list passes over the
                 * own rows (a layer at a time if it fits the lists, else a row at a time; one row always fits). */
                const uint32_t my_cells = (uint32_t)__popc(act);
                const uint32_t my_tris = (own_row && act) ? nt_seg : 0u;
                const bool cells_fit = __reduce_add_sync(0xFFFFFFFFu, my_cells + hx) <= CELL_CAP;
                const bool simple = cells_fit && __reduce_add_sync(0xFFFFFFFFu, my_tris) <= TRI_CAP;
                const int n_fill = cells_fit ? 1 : RZ; /* fill passes of the piecewise plan */
                int pass = 0, lo = 0;
                for (;;) {
                    bool part, listed, with_hx;
                    int hi = lo;
                    if (simple) {
                        part = (act | hx) != 0; with_hx = true; listed = own_row && act != 0;
                    } else if (pass < n_fill) {
                        part = has_task && (cells_fit || t_rz == pass) && (act | hx) != 0; with_hx = true; listed = false;
                    } else {
                        const bool g4 = own_row && act != 0 && t_q >= lo && t_q < lo + BY;
                        const bool fits = __reduce_add_sync(0xFFFFFFFFu, g4 ? my_cells : 0u) <= CELL_CAP &&
                                          __reduce_add_sync(0xFFFFFFFFu, g4 ? my_tris : 0u) <= TRI_CAP;
                        hi = fits ? lo + BY : lo + 1;
                        listed = own_row && act != 0 && t_q >= lo && t_q < hi;
                        part = listed; with_hx = false;
                    }
                    uint32_t n_cells, n_tri;
                    const uint32_t cpos = warp_excl_scan(part ? my_cells + (with_hx ? hx : 0u) : 0u, lane, n_cells);
                    const uint32_t tch = warp_excl_scan(listed ? nt_seg : 0u, lane, n_tri);
                    __syncwarp();
                    if (part) {
                        SegDesc &D = W.seg[lane];
                        D.p0 = pl.x; D.p1 = pl.y; D.p2 = pl.z; D.p3 = pl.w;
                        D.vbase = pv + (sp & 0xFFFFu);
                        D.tseg = pt + (sp >> 16);
                        D.cpos_tch = cpos | tch << 16;
                        D.info = (r == 0 ? 1u : 0u) | (gz == 0 ? 2u : 0u) | (listed ? 4u : 0u) | (s == 0 ? 8u : 0u) | both << 4;
                        uint32_t k = cpos, tpre = 0, m = act;
                        while (m) { /* expansion: one store per active cell (cube index + triangles of the earlier cells) */
                            const uint32_t i = __ffs(m) - 1;
                            m &= m - 1;
                            const uint32_t ci = (__funnelshift_r(a0, a1, i) & 3u) | (__funnelshift_r(b0, b1, i) & 3u) << 2 |
                                                (__funnelshift_r(c0, c1, i) & 3u) << 4 | (__funnelshift_r(d0, d1, i) & 3u) << 6;
                            W.cellmap[k++] = lane | i << 5 | ci << 11 | tpre << 19;
                            if (listed) tpre += S.ntri[ci];
                        }
                        if (with_hx && hx) W.cellmap[k] = lane | 1u << 10;
                    }
                    __syncwarp();

                    /* ---------------- P2: one lane per active cell of the region ---------------- */
                    for (uint32_t base = 0; base < n_cells; base += 32) {
                        const uint32_t k = base + lane;
                        if (k < n_cells) {
                            const uint32_t cm = W.cellmap[k];
                            const uint32_t task = cm & 31u, i = (cm >> 5) & 31u;
                            const SegDesc &D = W.seg[task];
                            const uint32_t info = D.info;
                            bool lst = (info >> 2 & 1u) != 0;
                            uint32_t ci, vid, bfl;
                            int cp;
                            const int rq = (int)task;
                            if (cm >> 10 & 1u) { /* x-halo cell: corners from samples 32s-1 and 32s */
                                const uint32_t bb = info >> 4;
                                ci = (bb & 1u) | (bb >> 4 & 1u) << 1 | (bb >> 1 & 1u) << 2 | (bb >> 5 & 1u) << 3 | (bb >> 2 & 1u) << 4 |
                                     (bb >> 6 & 1u) << 5 | (bb >> 3 & 1u) << 6 | (bb >> 7 & 1u) << 7;
                                bfl = (info & 1u) << 1 | (info >> 1 & 1u) << 2;
                                vid = D.vbase - __popc((uint32_t)S.emask[ci] & (uint32_t)S.ownmask[bfl]);
                                cp = rq * RX;
                                lst = false; /* belongs to the brick on the left */
                            } else {
                                ci = (cm >> 11) & 255u;
                                vid = D.vbase + planes_count(make_uint4(D.p0, D.p1, D.p2, D.p3), (1u << i) - 1u);
                                bfl = ((info >> 3 & 1u) && i == 0 ? 1u : 0u) | (info & 1u) << 1 | (info >> 1 & 1u) << 2;
                                cp = rq * RX + 1 + (int)i;
                            }
                            const uint32_t em = S.emask[ci];
                            const int rz = rq / RY, ry = rq - rz * RY;
                            const uint32_t rel = vid + vofs - W.rowbase[rq]; /* < 65536: row totals are 16 bit */
                            /* creator cell for the vertex descriptors: x | y << 16, local layer | e << 16 */
                            const uint32_t dxy = (s * 32 + i) | (uint32_t)(y0 + ry - 1) << 16, dlz = (uint32_t)(lz0 + rz - 1);
                            if (bfl == 0) { /* interior: creates exactly its crossed e5, e6, e10, ranks from rank3 */
                                const uint32_t r3 = S.rank3[ci];
                                if (em >> 6 & 1u) {
                                    const uint32_t rk = r3 >> 2 & 3u, slot = vid + rk - ghostV;
                                    W.plane[cp] = (uint16_t)(rel + rk);
                                    if (lst && slot < cap_v) { vdesc[3 * (uint64_t)slot] = dxy; vdesc[3 * (uint64_t)slot + 1] = dlz | 6u << 16; }
                                }
                                if (em >> 5 & 1u) {
                                    const uint32_t rk = r3 & 3u, slot = vid + rk - ghostV;
                                    W.plane[NREGION + cp] = (uint16_t)(rel + rk);
                                    if (lst && slot < cap_v) { vdesc[3 * (uint64_t)slot] = dxy; vdesc[3 * (uint64_t)slot + 1] = dlz | 5u << 16; }
                                }
                                if (em >> 10 & 1u) {
                                    const uint32_t rk = r3 >> 4 & 3u, slot = vid + rk - ghostV;
                                    W.plane[2 * NREGION + cp] = (uint16_t)(rel + rk);
                                    if (lst && slot < cap_v) { vdesc[3 * (uint64_t)slot] = dxy; vdesc[3 * (uint64_t)slot + 1] = dlz | 10u << 16; }
                                }
                            } else { /* on a low boundary face: more edges, first-appearance order decides the ranks */
                                const uint32_t owned = em & S.ownmask[bfl];
                                const int rx = cp - rq * RX;
                                uint64_t ord = tabs->order[ci];
                                uint32_t rk = 0;
                                for (uint32_t rem = em; rem; rem &= rem - 1, ord >>= 4) {
                                    const uint32_t e = (uint32_t)ord & 15u;
                                    if (!(owned >> e & 1u)) continue;
                                    const uint32_t st = S.bstep[e], slot = vid + rk - ghostV;
                                    /* the (virtual) creator's row has the same id base (clamped rows) */
                                    if ((int)(st & 1u) <= rx && (int)(st >> 1 & 1u) <= ry && (int)(st >> 2 & 1u) <= rz)
                                        W.plane[cp + S.offs[e]] = (uint16_t)(vid + vofs + rk - W.rowbase[rq - S.rowback[e]]);
                                    if (lst && slot < cap_v) { vdesc[3 * (uint64_t)slot] = dxy; vdesc[3 * (uint64_t)slot + 1] = dlz | e << 16; }
                                    ++rk;
                                }
                            }
                            if (lst) { /* triangle list: position = segment start + triangles of the earlier cells (from P1) */
                                const uint32_t tp = (D.cpos_tch >> 16) + (cm >> 19);
                                const uint32_t ent = (uint32_t)cp | ci << 10 | task << 21;
                                const uint32_t nt = S.ntri[ci];
                                uint32_t *tl = W.trilist + tp; /* nt is 1..5 */
                                tl[0] = ent;
                                if (nt > 1) tl[1] = ent | 1u << 18;
                                if (nt > 2) tl[2] = ent | 2u << 18;
                                if (nt > 3) tl[3] = ent | 3u << 18;
                                if (nt > 4) tl[4] = ent | 4u << 18;
                            }
                        }
                    }
                    __syncwarp();

                    /* ---------------- B: one lane per triangle ---------------- */
                    for (uint32_t base = 0; base < n_tri; base += 32) {
                        const uint32_t j = base + lane;
                        if (j < n_tri) {
                            const uint32_t ent = W.trilist[j];
                            const int cp = (int)(ent & 1023u);
                            const uint32_t ci = (ent >> 10) & 255u, t = (ent >> 18) & 7u, task = ent >> 21;
                            const uint32_t edges = (uint32_t)(S.tri[ci] >> (12 * t));
                            const uint32_t k0 = (uint32_t)S.look[edges & 15u], k1 = (uint32_t)S.look[(edges >> 4) & 15u], k2 = (uint32_t)S.look[(edges >> 8) & 15u];
                            const int cq = cp - 4096;
                            const uint32_t i0 = W.rowbase[task - (k0 >> 16)] + W.plane[cq + (int)(k0 & 0xFFFFu)];
                            const uint32_t i1 = W.rowbase[task - (k1 >> 16)] + W.plane[cq + (int)(k1 & 0xFFFFu)];
                            const uint32_t i2 = W.rowbase[task - (k2 >> 16)] + W.plane[cq + (int)(k2 & 0xFFFFu)];
                            const SegDesc &D = W.seg[task];
                            const uint32_t tslot = D.tseg + (j - (D.cpos_tch >> 16)) - ghostT;
                            if (tslot < cap_t) {
                                uint32_t *o = idx + (uint64_t)tslot * 3;
                                o[0] = i0; o[1] = i1; o[2] = i2;
                            }
                        }
                    }
                    __syncwarp();
                    if (simple) break;
                    if (pass >= n_fill) {
                        lo = hi;
                        if (lo >= NROWS_OWN) break;
                    }
                    ++pass;
                }
                __syncwarp();
            }
#undef ISOMC_BRICK_LOADS
        }
    }
}

/* vertex descriptor (written by k_emit into the vertex's own 12-byte slot) -> position, in place.
 * descriptor: [0] = x | y << 16 of the creating cell, [1] = local layer | edge << 16.
 * Interpolation = Signed::find_crossing_point (distance.rs:64-69) between the edge's ends in
 * EDGE_CONNECTION direction of the creating cell, corner coordinates = (i as f32) * inv. */
template <class Src>
__global__ void __launch_bounds__(256) k_vertex(Src src, Geo g, const McTables *__restrict__ tabs,
                                                const unsigned long long *__restrict__ layerTot,
                                                const uint32_t *__restrict__ rowPV, float *__restrict__ xyz,
                                                unsigned long long cap_v, uint32_t lz_begin, uint32_t lz_end) {
    __shared__ uint8_t s_ends[12];
    if (threadIdx.x < 12) s_ends[threadIdx.x] = tabs->ends[threadIdx.x];
    __syncthreads();
    /* vertices created by cell layers [lz_begin, lz_end): slots [first, n) */
    const uint32_t ghostV = g.ghost ? (uint32_t)layerTot[0] : 0u;
    const unsigned long long first = rowPV[(uint64_t)lz_begin * g.ncx] - ghostV;
    unsigned long long n = (unsigned long long)rowPV[(uint64_t)(lz_end - 1) * g.ncx] + layerTot[3 * (lz_end - 1)] - ghostV;
    if (n > cap_v) n = cap_v;
    const uint32_t *desc = reinterpret_cast<const uint32_t *>(xyz);
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long v0 = first + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; v0 < n; v0 += 4 * stride) {
        /* four independent vertices per iteration: all sample loads are issued before the first use */
        uint32_t dd0[4], dd1[4];
        float sa[4], sb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned long long v = v0 + u * stride;
            dd0[u] = v < n ? desc[3 * v] : 0u;
            dd1[u] = v < n ? desc[3 * v + 1] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t x = dd0[u] & 0xFFFFu, y = dd0[u] >> 16, lz = dd1[u] & 0xFFFFu, en = s_ends[(dd1[u] >> 16) & 15u];
            sa[u] = src.at(g, x + (en & 1u), y + (en >> 1 & 1u), lz + (en >> 2 & 1u));
            sb[u] = src.at(g, x + (en >> 4 & 1u), y + (en >> 5 & 1u), lz + (en >> 6 & 1u));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned long long v = v0 + u * stride;
            if (v >= n) break;
            const uint32_t x = dd0[u] & 0xFFFFu, y = dd0[u] >> 16, lz = dd1[u] & 0xFFFFu, en = s_ends[(dd1[u] >> 16) & 15u];
            const uint32_t gz = g.gz0 + lz;
            const uint32_t ux = x + (en & 1u), uy = y + (en >> 1 & 1u), uz = en >> 2 & 1u;
            const uint32_t vx = x + (en >> 4 & 1u), vy = y + (en >> 5 & 1u), vz = en >> 6 & 1u;
            const float a = sa[u], b = sb[u];
            const float delta = __fsub_rn(b, a);
            const float t = (delta == 0.0f) ? 0.5f : __fdiv_rn(-a, delta);
            const float omt = __fsub_rn(1.0f, t);
            const float pax = __fmul_rn((float)ux, g.inv), pay = __fmul_rn((float)uy, g.inv), paz = __fmul_rn((float)(gz + uz), g.inv);
            const float pbx = __fmul_rn((float)vx, g.inv), pby = __fmul_rn((float)vy, g.inv), pbz = __fmul_rn((float)(gz + vz), g.inv);
            xyz[3 * v] = __fadd_rn(__fmul_rn(pax, omt), __fmul_rn(pbx, t));
            xyz[3 * v + 1] = __fadd_rn(__fmul_rn(pay, omt), __fmul_rn(pby, t));
            xyz[3 * v + 2] = __fadd_rn(__fmul_rn(paz, omt), __fmul_rn(pbz, t));
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
/* same points through the chain evaluator (when the program has a chain form): out2 must equal out bit for bit */
__global__ void k_sample_sdf_chain(SdfChain chain, const float *__restrict__ xyz, uint64_t n, float *__restrict__ out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = sdf_chain_eval(chain, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
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
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    uint64_t blocks = (warps_needed + warps_per_block - 1) / warps_per_block;
    uint64_t cap = (uint64_t)sms * blocks_per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (uint32_t)blocks;
}

/* sample rows [row0, row1) of the handle's lattice (row = local layer * N + y) */
cudaError_t isomc_launch_sign_grid(const Geo &g, const float *d_grid, uint32_t *signs, uint32_t row0, uint32_t row1, int sms,
                                   int ctas_per_sm, cudaStream_t st) {
    const uint32_t grid = grid_for((uint64_t)(row1 - row0), sms, 8, ctas_per_sm);
    if ((g.N & 3u) == 0 && (reinterpret_cast<uintptr_t>(d_grid) & 15u) == 0) {
        const float4 *g4 = reinterpret_cast<const float4 *>(d_grid);
        if (g.N >= 1024) k_sign_vec4<8><<<grid, 256, 0, st>>>(g4, g, signs, row0, row1);
        else k_sign_vec4<4><<<grid, 256, 0, st>>>(g4, g, signs, row0, row1);
    } else {
        GridSrc src{d_grid};
        k_sign<GridSrc><<<grid, 256, 0, st>>>(src, g, signs, row0, row1);
    }
    return cudaGetLastError();
}
cudaError_t isomc_launch_sign_sdf(const Geo &g, const SdfProgram &prog, bool directed, uint32_t *signs, uint32_t row0, uint32_t row1,
                                  int sms, int ctas_per_sm, cudaStream_t st) {
    if (directed) {
        k_sign<SdfDirSrc><<<grid_for((uint64_t)(row1 - row0), sms, 8, ctas_per_sm), 256, 0, st>>>(SdfDirSrc{prog}, g, signs, row0, row1);
        return cudaGetLastError();
    }
    SdfChainSrc csrc;
    if (sdf_to_chain(prog, &csrc.chain)) {
        k_sign<SdfChainSrc><<<grid_for((uint64_t)(row1 - row0), sms, 8, ctas_per_sm), 256, 0, st>>>(csrc, g, signs, row0, row1);
    } else {
        SdfSrc src{prog};
        k_sign<SdfSrc><<<grid_for((uint64_t)(row1 - row0), sms, 8, ctas_per_sm), 256, 0, st>>>(src, g, signs, row0, row1);
    }
    return cudaGetLastError();
}
/* cell layers [lz0, lz1) */
cudaError_t isomc_launch_count(const Geo &g, const uint32_t *signs, const McTables *tabs, uint32_t *segpre,
                               uint32_t *rowV, uint32_t *rowT, uint32_t *rowA, unsigned long long *layerTot,
                               uint32_t lz0, uint32_t lz1, int sms, int ctas_per_sm, cudaStream_t st) {
    uint32_t gshift = 0;
    while ((1u << gshift) < g.nsegx && gshift < 5) ++gshift;
    const uint32_t rpw = g.nsegx <= 32 ? (32u >> gshift) : 1u;
    const uint32_t row0 = lz0 * g.ncx, row1 = lz1 * g.ncx;
    const uint64_t warps = ((uint64_t)(row1 - row0) + rpw - 1) / rpw;
    k_count<<<grid_for(warps, sms, 8, ctas_per_sm), 256, 0, st>>>(g, signs, tabs, segpre, rowV, rowT, rowA, layerTot, gshift, row0, row1);
    return cudaGetLastError();
}
cudaError_t isomc_launch_scan(const Geo &g, uint32_t *rowV, uint32_t *rowT, const unsigned long long *layerTot,
                              unsigned long long *totals, const uint32_t *list_ctr, uint32_t *list_mark, uint32_t *chunk_end,
                              uint32_t lz0, uint32_t lz1, cudaStream_t st) {
    k_scan_rows<<<lz1 - lz0, 256, 0, st>>>(g, rowV, rowT, layerTot, totals, list_ctr, list_mark, chunk_end, lz0);
    return cudaGetLastError();
}
cudaError_t isomc_launch_slab_bases(const unsigned long long *gathered, uint32_t rank, uint32_t ghost, uint32_t *vofs,
                                    cudaStream_t st) {
    k_slab_bases<<<1, 32, 0, st>>>(gathered, rank, ghost, vofs);
    return cudaGetLastError();
}

int isomc_emit_layers_per_brick() { return BZ; }

/* cell layers [lz0, lz1), lz0 a multiple of BZ */
static cudaError_t launch_emit(const Geo &g, const uint32_t *signs, const uint32_t *segpre, const uint32_t *rowPV,
                               const uint32_t *rowPT, const McTables *tabs, const unsigned long long *layerTot,
                               const uint32_t *vofs, uint32_t *ticket, float *xyz, uint32_t *idx, uint64_t cap_v,
                               uint64_t cap_t, uint32_t lz0, uint32_t lz1, int sms, cudaStream_t st) {
    const size_t smem = isomc_emit_smem_bytes(g.nws);
    static int per_sm = 0;
    if (per_sm == 0) {
        cudaError_t e = cudaFuncSetAttribute(k_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int n = 1;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_emit, EMIT_THREADS, smem);
        if (e != cudaSuccess) return e;
        per_sm = n < 1 ? 1 : n;
    }
    const uint32_t nby = (g.ncx + BY - 1) / BY;
    const uint32_t strip0 = (lz0 / BZ) * nby, strip1 = ((lz1 + BZ - 1) / BZ) * nby;
    uint64_t blocks = ((uint64_t)(strip1 - strip0) + EMIT_WARPS - 1) / EMIT_WARPS;
    if (blocks > (uint64_t)sms * per_sm) blocks = (uint64_t)sms * per_sm;
    k_emit<<<(uint32_t)blocks, EMIT_THREADS, smem, st>>>(g, signs, segpre, rowPV, rowPT, tabs, layerTot, vofs, ticket,
                                                         reinterpret_cast<uint32_t *>(xyz), idx, cap_v, cap_t, strip0, strip1);
    return cudaGetLastError();
}

cudaError_t isomc_launch_emit(const Geo &g, const uint32_t *signs, const uint32_t *segpre, const uint32_t *rowPV,
                              const uint32_t *rowPT, const McTables *tabs, const unsigned long long *layerTot,
                              const uint32_t *vofs, uint32_t *ticket, float *xyz, uint32_t *idx, uint64_t cap_v, uint64_t cap_t,
                              uint32_t lz0, uint32_t lz1, int sms, cudaStream_t st) {
    return launch_emit(g, signs, segpre, rowPV, rowPT, tabs, layerTot, vofs, ticket, xyz, idx, cap_v, cap_t, lz0, lz1, sms, st);
}
cudaError_t isomc_launch_vertex_grid(const Geo &g, const float *d_grid, const McTables *tabs, const unsigned long long *layerTot,
                                     const uint32_t *rowPV, float *xyz, uint64_t cap_v, uint32_t lz0, uint32_t lz1, int sms,
                                     int ctas_per_sm, cudaStream_t st) {
    k_vertex<GridSrc><<<sms * ctas_per_sm, 256, 0, st>>>(GridSrc{d_grid}, g, tabs, layerTot, rowPV, xyz, cap_v, lz0, lz1);
    return cudaGetLastError();
}
cudaError_t isomc_launch_vertex_sdf(const Geo &g, const SdfProgram &prog, const McTables *tabs, const unsigned long long *layerTot,
                                    const uint32_t *rowPV, float *xyz, uint64_t cap_v, uint32_t lz0, uint32_t lz1, int sms,
                                    int ctas_per_sm, cudaStream_t st) {
    SdfChainSrc csrc;
    if (sdf_to_chain(prog, &csrc.chain))
        k_vertex<SdfChainSrc><<<sms * ctas_per_sm, 256, 0, st>>>(csrc, g, tabs, layerTot, rowPV, xyz, cap_v, lz0, lz1);
    else
        k_vertex<SdfSrc><<<sms * ctas_per_sm, 256, 0, st>>>(SdfSrc{prog}, g, tabs, layerTot, rowPV, xyz, cap_v, lz0, lz1);
    return cudaGetLastError();
}

cudaError_t isomc_launch_cube_indices(const Geo &g, const uint32_t *signs, const McTables *tabs, uint8_t *out, int sms,
                                      cudaStream_t st) {
    k_cube_indices<<<sms * 8, 256, 0, st>>>(g, signs, tabs, out);
    return cudaGetLastError();
}
/* use_chain != 0: evaluate through the chain form (returns cudaErrorInvalidValue if the program has none) */
cudaError_t isomc_launch_sample_sdf(const SdfProgram &prog, const float *xyz, uint64_t n, float *out, int use_chain, cudaStream_t st) {
    const uint32_t grid = (uint32_t)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256);
    if (use_chain) {
        SdfChain chain;
        if (!sdf_to_chain(prog, &chain)) return cudaErrorInvalidValue;
        k_sample_sdf_chain<<<grid, 256, 0, st>>>(chain, xyz, n, out);
    } else {
        k_sample_sdf<<<grid, 256, 0, st>>>(prog, xyz, n, out);
    }
    return cudaGetLastError();
}
cudaError_t isomc_launch_synth(const SynthParams &sp, uint32_t size, uint32_t z_first, uint32_t n_layers, float *out,
                               int sms, cudaStream_t st) {
    const float inv = 1.0f / (float)(size - 1);
    k_synth<<<sms * 8, 256, 0, st>>>(sp, size, inv, z_first, n_layers, out);
    return cudaGetLastError();
}
