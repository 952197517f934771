/*
 * isomc_points.cu -- the crate's other sinks of the same traversal: PointCloud extraction and central-difference normals.
 *
 * PointCloud (reference src/point_cloud.rs:50-63) on the sign words:
 * every active cell (cube index neither 0 nor 255) emits the midpoint of its corners 0 and 6,
 * `corners[0].lerp(corners[6], 0.5)` (src/math/vector.rs:325-333), in (z, y, x) cell order.
 *
 *   k_sign      (shared with MarchingCubes)  sample -> inside bit
 *   k_pc_count  lane per 32-cell segment: active mask, in-row prefix of the active-cell count, row totals
 *   k_scan_rows (shared)                     exclusive prefix over cell rows
 *   k_pc_emit   lane per segment: one 12-byte store per active cell at its final position
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "isomc_device.cuh"
#include "isomc_kernels.h"

namespace {

/* active cells of segment s of cell row (lz, y): corners neither all inside nor all outside */
__device__ __forceinline__ uint32_t seg_active(const Geo &g, const uint32_t *__restrict__ signs, uint32_t row, uint32_t lz, uint32_t s) {
    const uint32_t *r00 = signs + (uint64_t)(row + lz) * g.nws + s; /* sample row lz*N + y = row + lz */
    const uint32_t *r01 = r00 + g.nws, *r10 = r00 + (uint64_t)g.N * g.nws, *r11 = r10 + g.nws;
    const uint32_t a0 = __ldg(r00), a1 = __ldg(r00 + 1), b0 = __ldg(r01), b1 = __ldg(r01 + 1);
    const uint32_t c0 = __ldg(r10), c1 = __ldg(r10 + 1), d0 = __ldg(r11), d1 = __ldg(r11 + 1);
    const uint32_t an = __funnelshift_r(a0, a1, 1), bn = __funnelshift_r(b0, b1, 1);
    const uint32_t cn = __funnelshift_r(c0, c1, 1), dn = __funnelshift_r(d0, d1, 1);
    const uint32_t ncell = g.ncx - s * 32;
    const uint32_t vm = ncell >= 32 ? 0xFFFFFFFFu : ((1u << ncell) - 1u);
    const uint32_t all_in = a0 & an & b0 & bn & c0 & cn & d0 & dn;
    const uint32_t any_in = a0 | an | b0 | bn | c0 | cn | d0 | dn;
    return any_in & ~all_in & vm;
}

__global__ void __launch_bounds__(256) k_pc_count(Geo g, const uint32_t *__restrict__ signs, uint32_t *__restrict__ segA,
                                                  uint32_t *__restrict__ rowV, uint32_t *__restrict__ rowT,
                                                  unsigned long long *__restrict__ layerTot, uint32_t row0, uint32_t row1) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t row = row0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < row1; row += nwarps) {
        const uint32_t lz = row / g.ncx;
        uint32_t carry = 0;
        for (uint32_t s0 = 0; s0 < g.nsegx; s0 += 32) {
            const uint32_t s = s0 + lane;
            const uint32_t na = s < g.nsegx ? (uint32_t)__popc(seg_active(g, signs, row, lz, s)) : 0u;
            uint32_t inc = na;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= (uint32_t)d) inc += o;
            }
            if (s < g.nsegx) segA[(uint64_t)row * g.nsegx + s] = carry + inc - na;
            carry += __shfl_sync(0xFFFFFFFFu, inc, 31);
        }
        if (lane == 0) {
            rowV[row] = carry; /* the row scan's "vertex" channel carries the points */
            rowT[row] = 0;
            if (carry) {
                atomicAdd(&layerTot[3 * lz + 0], (unsigned long long)carry);
                atomicAdd(&layerTot[3 * lz + 2], (unsigned long long)carry);
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_pc_emit(Geo g, const uint32_t *__restrict__ signs, const uint32_t *__restrict__ segA,
                                                 const uint32_t *__restrict__ rowPV, float *__restrict__ xyz,
                                                 unsigned long long cap_v, uint32_t row0, uint32_t row1) {
    const uint64_t nseg = (uint64_t)(row1 - row0) * g.nsegx;
    for (uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; q < nseg; q += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t row = row0 + (uint32_t)(q / g.nsegx), s = (uint32_t)(q % g.nsegx);
        const uint32_t lz = row / g.ncx, y = row - lz * g.ncx;
        uint32_t act = seg_active(g, signs, row, lz, s);
        if (!act) continue;
        unsigned long long slot = (unsigned long long)rowPV[row] + segA[(uint64_t)row * g.nsegx + s];
        /* lerp(corners[0], corners[6], 0.5): of = 1.0 - 0.5; of * c0 + 0.5 * c6, corner coordinate = (i as f32) * inv */
        const float py = __fadd_rn(__fmul_rn(0.5f, __fmul_rn((float)y, g.inv)), __fmul_rn(0.5f, __fmul_rn((float)(y + 1), g.inv)));
        const float pz = __fadd_rn(__fmul_rn(0.5f, __fmul_rn((float)(g.gz0 + lz), g.inv)),
                                   __fmul_rn(0.5f, __fmul_rn((float)(g.gz0 + lz + 1), g.inv)));
        for (; act; act &= act - 1, ++slot) {
            if (slot >= cap_v) break;
            const uint32_t x = s * 32 + (uint32_t)__ffs(act) - 1u;
            float *o = xyz + 3 * slot;
            o[0] = __fadd_rn(__fmul_rn(0.5f, __fmul_rn((float)x, g.inv)), __fmul_rn(0.5f, __fmul_rn((float)(x + 1), g.inv)));
            o[1] = py;
            o[2] = pz;
        }
    }
}

}  // namespace

cudaError_t isomc_launch_points_count(const Geo &g, const uint32_t *signs, uint32_t *segA, uint32_t *rowV, uint32_t *rowT,
                                      unsigned long long *layerTot, int sms, cudaStream_t st) {
    const uint32_t rows = g.ncl * g.ncx;
    uint32_t grid = (rows + 7) / 8;
    if (grid > (uint32_t)sms * 8) grid = (uint32_t)sms * 8;
    k_pc_count<<<grid < 1 ? 1 : grid, 256, 0, st>>>(g, signs, segA, rowV, rowT, layerTot, 0u, rows);
    return cudaGetLastError();
}

cudaError_t isomc_launch_points_emit(const Geo &g, const uint32_t *signs, const uint32_t *segA, const uint32_t *rowPV, float *xyz,
                                     uint64_t cap_v, int sms, cudaStream_t st) {
    k_pc_emit<<<sms * 8, 256, 0, st>>>(g, signs, segA, rowPV, xyz, cap_v, 0u, g.ncl * g.ncx);
    return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------------------------
 * Normals for extractor::IndexedInterleavedNormals (reference src/extractor.rs:95-127): every vertex v of the
 * last extract is followed by `source.sample_normal(v)`, with `source` = a CentralDifference adaptor around an
 * implicit tree (src/source.rs:82-94), possibly inside translations (examples/common/sources.rs:55-60:
 * q = p - offset, then the wrapped source's normal at q):
 *
 *     n = ( f(q + dx) - f(q - dx),  f(q + dy) - f(q - dy),  f(q + dz) - f(q - dz) ) / (2 * epsilon)
 *
 * with dx = (epsilon, 0, 0) etc. added as whole vectors (the other two components get + 0.0 / - 0.0), every
 * operation rounded to binary32 in the reference's order.  One lane per vertex, six evaluations of the program.
 * ------------------------------------------------------------------------------------------------------------ */
struct NormalOffsets {
    float off[ISOMC_TR_DEPTH][3];
    uint32_t n;
};

__global__ void __launch_bounds__(256) k_normals_cd(SdfProgram prog, NormalOffsets tr, float eps, const float *__restrict__ xyz,
                                                    uint64_t n_vertices, float *__restrict__ out /* 6 floats per vertex */) {
    const float two_eps = __fmul_rn(2.0f, eps);
    for (uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; v < n_vertices; v += (uint64_t)gridDim.x * blockDim.x) {
        const float px = xyz[3 * v], py = xyz[3 * v + 1], pz = xyz[3 * v + 2];
        float qx = px, qy = py, qz = pz;
        for (uint32_t k = 0; k < tr.n; ++k) {
            qx = __fsub_rn(qx, tr.off[k][0]); qy = __fsub_rn(qy, tr.off[k][1]); qz = __fsub_rn(qz, tr.off[k][2]);
        }
        const float z = 0.0f;
        const float vx = __fsub_rn(sdf_eval(prog, __fadd_rn(qx, eps), __fadd_rn(qy, z), __fadd_rn(qz, z)),
                                   sdf_eval(prog, __fsub_rn(qx, eps), __fsub_rn(qy, z), __fsub_rn(qz, z)));
        const float vy = __fsub_rn(sdf_eval(prog, __fadd_rn(qx, z), __fadd_rn(qy, eps), __fadd_rn(qz, z)),
                                   sdf_eval(prog, __fsub_rn(qx, z), __fsub_rn(qy, eps), __fsub_rn(qz, z)));
        const float vz = __fsub_rn(sdf_eval(prog, __fadd_rn(qx, z), __fadd_rn(qy, z), __fadd_rn(qz, eps)),
                                   sdf_eval(prog, __fsub_rn(qx, z), __fsub_rn(qy, z), __fsub_rn(qz, eps)));
        float *o = out + 6 * v;
        o[0] = px; o[1] = py; o[2] = pz;
        o[3] = __fdiv_rn(vx, two_eps); o[4] = __fdiv_rn(vy, two_eps); o[5] = __fdiv_rn(vz, two_eps);
    }
}

/* inner = program without the translations that enclose all of it (applied to the vertex first, outermost first) */
cudaError_t isomc_launch_normals_cd(const SdfProgram &inner, const float (*offsets)[3], uint32_t n_offsets, float eps,
                                    const float *xyz, uint64_t n_vertices, float *out, int sms, cudaStream_t st) {
    NormalOffsets tr;
    tr.n = n_offsets;
    for (uint32_t k = 0; k < ISOMC_TR_DEPTH; ++k)
        for (int j = 0; j < 3; ++j) tr.off[k][j] = k < n_offsets ? offsets[k][j] : 0.0f;
    k_normals_cd<<<sms * 4, 256, 0, st>>>(inner, tr, eps, xyz, n_vertices, out);
    return cudaGetLastError();
}
