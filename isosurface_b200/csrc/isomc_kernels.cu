/*
 * isomc_kernels.cu -- sm_100a kernels shared by the extract paths, plus debug and synthesis kernels.
 *
 *   k_sign_vec4 / k_sign<Src>
 *                    sample -> inside bit.  One bit per lattice point (`!(v > 0)`,
 *                    marching_cubes_impl.rs:32 / distance.rs:52-54), 32 per word.  Used by the PointCloud
 *                    path and the cube-index dump; the mesh path produces its sign bits inside pass 1
 *                    of the tile path (isomc_tile_kernels.cu) and never stores them.
 *   k_scan_rows      exclusive scan over row pieces in (z, y, x-tile) order (+ totals); causal in z.
 *   k_slab_bases     id offset of a z-slab from the all-gathered per-rank totals.
 *
 * Vertex numbering = reference numbering: id(cell, e) = (# vertices created by earlier cells in
 * (z,y,x) order) + (# edges the cell creates that precede e in first-appearance order of its
 * triangle list).  See SURVEY.md 3.1-9 and DESIGN.md.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "isomc_device.cuh"
#include "isomc_kernels.h"
#include "isomc_launch.cuh"
#include "isomc_tables.h"

/* ------------------------------------------------------------------------------------------ */
/* K1: sample -> inside bits                                                                    */
/* ------------------------------------------------------------------------------------------ */

/* generic path: any source, any size/alignment; one lane per sample, ballot per 32 samples */
template <class Src>
__global__ void __launch_bounds__(256) k_sign(Src src, Geo g, uint32_t *__restrict__ signs, uint32_t row0, uint32_t row1) {
    constexpr int U = 8;
    isomc_pdl_trigger();
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
    isomc_pdl_trigger();
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
/* exclusive scan over row pieces (one CTA per cell layer) + totals                             */
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

/* in: rowV/rowT hold per-piece counts (ppl pieces per cell layer: a piece is a cell row, or one x-tile of it on the
 * tile path); out: exclusive prefixes over pieces in (lz, y, x-tile) order, with a sentinel entry [ncl * ppl] = grand
 * total.  totals (u64):
 *   [0] V incl. ghost layer  [1] T incl. ghost  [2] active cells incl. ghost
 *   [3] V prefix at the start of the last cell layer  [4..6] V, T, active of the ghost layer  [7] list blocks asked for
 *   [8] vertices owned  [9] owned vertices created before the last cell layer  [10] triangles owned
 *   [11] active cells owned  [12] t blocks asked for (tile path) */
__global__ void __launch_bounds__(256) k_scan_rows(Geo g, uint32_t ppl, uint32_t *__restrict__ rowV, uint32_t *__restrict__ rowT,
                                                   const unsigned long long *__restrict__ layerTot,
                                                   unsigned long long *__restrict__ totals,
                                                   const uint32_t *__restrict__ list_ctr, uint32_t *__restrict__ list_mark,
                                                   uint32_t *__restrict__ chunk_end, uint32_t lz_first) {
    __shared__ unsigned long long s_w[8];
    isomc_pdl_trigger();
    isomc_pdl_wait(); /* row counts, layer totals and the list counter come from the counting kernel */
    if (list_mark && blockIdx.x == 0 && threadIdx.x == 0) *list_mark = *list_ctr; /* list blocks handed out up to this z-chunk */
    __shared__ unsigned long long s_base[2];
    const uint32_t lz = lz_first + blockIdx.x;
    /* base = sum of the totals of the layers below (batched chunks: of the layers below in the same lattice, ids are chunk-local) */
    const uint32_t l_first = geo_chunk(g, lz) * g.zper; /* (0 unless the handle is a batch) */
    unsigned long long bv = 0, bt = 0, ba = 0;
    for (uint32_t l = l_first + threadIdx.x; l < lz; l += blockDim.x) {
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
    const uint32_t per = (ppl + 255) / 256;
    const uint32_t y_begin = min(ppl, threadIdx.x * per), y_end = min(ppl, y_begin + per);
    const uint64_t lbase = (uint64_t)lz * ppl;
    uint32_t sv = 0, st = 0;
    for (uint32_t y = y_begin; y < y_end; ++y) {
        sv += rowV[lbase + y];
        st += rowT[lbase + y];
    }
    unsigned long long tot;
    unsigned long long ev = block_excl_scan_256<unsigned long long>((unsigned long long)sv, s_w, tot);
    unsigned long long et = block_excl_scan_256<unsigned long long>((unsigned long long)st, s_w, tot);
    uint32_t pv = (uint32_t)(s_base[0] + ev), pt = (uint32_t)(s_base[1] + et);
    for (uint32_t y = y_begin; y < y_end; ++y) {
        uint32_t cv = rowV[lbase + y], ct = rowT[lbase + y];
        rowV[lbase + y] = pv;
        rowT[lbase + y] = pt;
        pv += cv;
        pt += ct;
    }
    if (chunk_end && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) { /* ids / slots below the end of this z-chunk */
        chunk_end[0] = (uint32_t)(tv + layerTot[3 * lz]);
        chunk_end[1] = (uint32_t)(tt + layerTot[3 * lz + 1]);
    }
    if (g.zper && lz - geo_chunk(g, lz) * g.zper == g.zper - 2 && threadIdx.x == 0) { /* last cell layer of a lattice of the batch: its totals */
        unsigned long long *ct = totals + 16 + 3 * (size_t)geo_chunk(g, lz);
        ct[0] = tv + layerTot[3 * lz]; ct[1] = tt + layerTot[3 * lz + 1]; ct[2] = ta + layerTot[3 * lz + 2];
    }
    if (lz == g.ncl - 1 && threadIdx.x == 0 && !g.zper) {
        unsigned long long V = tv + layerTot[3 * lz], T = tt + layerTot[3 * lz + 1], A = ta + layerTot[3 * lz + 2];
        unsigned long long gV = g.ghost ? layerTot[0] : 0, gT = g.ghost ? layerTot[1] : 0, gA = g.ghost ? layerTot[2] : 0;
        totals[0] = V; totals[1] = T; totals[2] = A; totals[3] = tv;
        totals[4] = gV; totals[5] = gT; totals[6] = gA;
        totals[7] = list_ctr ? list_ctr[0] : 0u; /* list blocks the count asked for */
        totals[8] = V - gV; totals[9] = tv - gV; totals[10] = T - gT; totals[11] = A - gA;
        totals[12] = (list_ctr && !list_mark) ? list_ctr[1] : 0u; /* tile path: blocks of crossing parameters asked for */
        rowV[(uint64_t)g.ncl * ppl] = (uint32_t)V;
        rowT[(uint64_t)g.ncl * ppl] = (uint32_t)T;
    }
}

/* batched chunks: totals[16 + 3b ..] = {V, T, A} of chunk b -> output slot of each chunk's first vertex / triangle (exclusive
 * prefixes, n + 1 entries each) and the grand totals in the slots finish() reads */
__global__ void __launch_bounds__(256) k_chunk_bases(uint32_t n, unsigned long long *__restrict__ totals, const uint32_t *__restrict__ list_ctr,
                                                     uint32_t *__restrict__ chunkV, uint32_t *__restrict__ chunkT) {
    __shared__ unsigned long long s_w[8];
    __shared__ unsigned long long s_carry[3];
    if (threadIdx.x < 3) s_carry[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t b0 = 0; b0 < n; b0 += 256) {
        const uint32_t b = b0 + threadIdx.x;
        const unsigned long long v = b < n ? totals[16 + 3 * (size_t)b] : 0, t = b < n ? totals[16 + 3 * (size_t)b + 1] : 0,
                                 a = b < n ? totals[16 + 3 * (size_t)b + 2] : 0;
        unsigned long long sv, st, sa;
        const unsigned long long ev = block_excl_scan_256<unsigned long long>(v, s_w, sv);
        const unsigned long long et = block_excl_scan_256<unsigned long long>(t, s_w, st);
        block_excl_scan_256<unsigned long long>(a, s_w, sa);
        if (b < n) { chunkV[b] = (uint32_t)(s_carry[0] + ev); chunkT[b] = (uint32_t)(s_carry[1] + et); }
        __syncthreads();
        if (threadIdx.x == 0) { s_carry[0] += sv; s_carry[1] += st; s_carry[2] += sa; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const unsigned long long V = s_carry[0], T = s_carry[1], A = s_carry[2];
        chunkV[n] = (uint32_t)V; chunkT[n] = (uint32_t)T;
        totals[0] = V; totals[1] = T; totals[2] = A; totals[3] = 0;
        totals[4] = 0; totals[5] = 0; totals[6] = 0;
        totals[7] = list_ctr ? list_ctr[0] : 0u;
        totals[8] = V; totals[9] = 0; totals[10] = T; totals[11] = A; totals[12] = 0;
    }
}

/* vertex-id offset of a slab from the all-gathered per-rank totals {V, V_before_last, T} */
__global__ void k_slab_bases(const unsigned long long *__restrict__ gathered, uint32_t rank, uint32_t ghost,
                             uint32_t *__restrict__ vofs, unsigned long long *__restrict__ ofs64) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long vbase = 0;
        for (uint32_t h = 0; h < rank; ++h) vbase += gathered[3 * h];
        unsigned long long ofs = vbase;
        if (ghost && rank > 0) ofs = vbase - gathered[3 * (rank - 1)] + gathered[3 * (rank - 1) + 1];
        *vofs = (uint32_t)ofs;
        if (ofs64) *ofs64 = ofs; /* untruncated: the host checks that ofs + local ids fit u32 */
    }
}

/*
 * The same exchange WITHOUT a collective library: the slab totals travel as peer stores over NVLink / NVSwitch.
 * Every rank owns a mailbox (2 parities x ISOMC_MAX_RANKS slots of {V, V_before_last, T, seq}) in its device memory and holds the
 * addresses of all ranks' mailboxes (same process: plain device pointers with peer access enabled; other processes: CUDA IPC
 * mappings).  Thread r of this one-CTA kernel writes this rank's totals into slot [parity][rank] of rank r's mailbox, fences,
 * releases the slot's sequence number, then waits for rank r's slot in the rank's OWN mailbox and takes its totals; thread 0
 * derives the id offset as k_slab_bases does.  Two parities are enough: a rank publishes step s+2 only after it has seen every
 * rank's step s+1, and a rank publishes s+1 after its own wait of step s (stream order).  The step number is kept in totals[15]
 * and advanced here.  A wait that lasts longer than
 * `timeout_cycles` gives up and flags the step (totals[14] = 1 + the silent rank) instead of hanging the device.
 */
__global__ void __launch_bounds__(ISOMC_MAX_RANKS) k_slab_exchange(unsigned long long *const *__restrict__ peers, uint32_t rank,
                                                                  uint32_t n_ranks, uint32_t ghost,
                                                                  unsigned long long *__restrict__ totals, uint32_t *__restrict__ vofs,
                                                                  long long timeout_cycles) {
    __shared__ unsigned long long s_tot[ISOMC_MAX_RANKS][3];
    __shared__ unsigned long long s_seq;
    __shared__ uint32_t s_bad;
    isomc_pdl_trigger();
    isomc_pdl_wait(); /* totals[8..10] come from the row scan */
    const uint32_t r = threadIdx.x;
    if (r == 0) { /* the step number lives on the device (totals[15]): the launch carries nothing that changes from step to step, so
                     the whole slab extract can be replayed from a CUDA graph */
        s_bad = 0;
        s_seq = totals[15] + 1ull;
        totals[15] = s_seq;
    }
    __syncthreads();
    const unsigned long long seq = s_seq;
    const uint32_t par = (uint32_t)(seq & 1ull);
    if (r < n_ranks) {
        volatile unsigned long long *dst = peers[r] + ((size_t)par * ISOMC_MAX_RANKS + rank) * 4;
        dst[0] = totals[8]; dst[1] = totals[9]; dst[2] = totals[10];
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst + 3), "l"(seq) : "memory");
        const unsigned long long *src = peers[rank] + ((size_t)par * ISOMC_MAX_RANKS + r) * 4;
        const long long t0 = clock64();
        unsigned long long got = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(src + 3) : "memory");
            if (got == seq) break;
            if (clock64() - t0 > timeout_cycles) { atomicMax(&s_bad, r + 1); break; }
            __nanosleep(100);
        }
        const volatile unsigned long long *vs = src;
        s_tot[r][0] = vs[0]; s_tot[r][1] = vs[1]; s_tot[r][2] = vs[2];
    }
    __syncthreads();
    if (r == 0) {
        unsigned long long vbase = 0;
        for (uint32_t h = 0; h < rank; ++h) vbase += s_tot[h][0];
        unsigned long long ofs = vbase;
        if (ghost && rank > 0) ofs = vbase - s_tot[rank - 1][0] + s_tot[rank - 1][1];
        *vofs = (uint32_t)ofs;
        totals[13] = ofs; /* untruncated: the host checks that ofs + local ids fit u32 */
        totals[14] = s_bad;
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
cudaError_t isomc_launch_scan(const Geo &g, uint32_t ppl, uint32_t *rowV, uint32_t *rowT, const unsigned long long *layerTot,
                              unsigned long long *totals, const uint32_t *list_ctr, uint32_t *list_mark, uint32_t *chunk_end,
                              uint32_t lz0, uint32_t lz1, cudaStream_t st) {
    return isomc_launch(k_scan_rows, lz1 - lz0, 256, st, isomc_pdl_for((unsigned long long)g.N * g.N * g.nsl), g, ppl, rowV, rowT, layerTot, totals, list_ctr, list_mark, chunk_end, lz0);
}
cudaError_t isomc_launch_chunk_bases(uint32_t n, unsigned long long *totals, const uint32_t *list_ctr, uint32_t *chunkV, uint32_t *chunkT,
                                     cudaStream_t st) {
    k_chunk_bases<<<1, 256, 0, st>>>(n, totals, list_ctr, chunkV, chunkT);
    return cudaGetLastError();
}
cudaError_t isomc_launch_sign_sdf_batch(const Geo &g, const SdfProgram *d_progs, bool directed, uint32_t *signs, uint32_t row0,
                                        uint32_t row1, int sms, cudaStream_t st) {
    const uint32_t grid = grid_for((uint64_t)(row1 - row0), sms, 8, 8);
    if (directed) k_sign<SdfBatchDirSrc><<<grid, 256, 0, st>>>(SdfBatchDirSrc{d_progs}, g, signs, row0, row1);
    else k_sign<SdfBatchSrc><<<grid, 256, 0, st>>>(SdfBatchSrc{d_progs}, g, signs, row0, row1);
    return cudaGetLastError();
}
cudaError_t isomc_launch_slab_bases(const unsigned long long *gathered, uint32_t rank, uint32_t ghost, uint32_t *vofs,
                                    unsigned long long *ofs64, cudaStream_t st) {
    k_slab_bases<<<1, 32, 0, st>>>(gathered, rank, ghost, vofs, ofs64);
    return cudaGetLastError();
}

cudaError_t isomc_launch_slab_exchange(unsigned long long *const *d_peers, uint32_t rank, uint32_t n_ranks, uint32_t ghost,
                                       unsigned long long *totals, uint32_t *vofs, long long timeout_cycles, cudaStream_t st) {
    return isomc_launch(k_slab_exchange, 1, ISOMC_MAX_RANKS, st, true, d_peers, rank, n_ranks, ghost, totals, vofs, timeout_cycles);
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
