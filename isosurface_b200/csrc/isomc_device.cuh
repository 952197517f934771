/*
 * isomc_device.cuh -- shared device-side definitions: geometry, sources, SDF evaluator.
 *
 * Arithmetic contract: everything that feeds a sign test or a vertex position is IEEE
 * binary32 with explicit round-to-nearest intrinsics (__fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn/
 * __fsqrt_rn), which nvcc never contracts into FMA -- rustc does not contract either, so the
 * device reproduces the reference's bits.
 */
#ifndef ISOMC_DEVICE_CUH
#define ISOMC_DEVICE_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/isomc.h"

#define ISOMC_VAL_DEPTH 8   /* value stack held in registers */
#define ISOMC_TR_DEPTH 4    /* nested translations */

struct Geo {
    uint32_t N;      /* lattice points per x/y axis (= size) */
    uint32_t ncx;    /* cells per x/y axis = N-1 */
    uint32_t nsegx;  /* 32-cell segments per cell row = ceil(ncx/32) */
    uint32_t nws;    /* sign words per sample row = nsegx+1 (last word may be all padding) */
    uint32_t ncl;    /* cell layers processed by this handle (incl. the ghost layer of a slab) */
    uint32_t nsl;    /* sample layers = ncl+1 */
    uint32_t gz0;    /* global z of local layer 0 */
    uint32_t ghost;  /* 1: local cell layer 0 belongs to the previous slab (counted, not emitted) */
    float inv;       /* 1.0f / (float)(N-1)   (primal_grid.rs:44-45) */
};

struct SdfProgram {
    isomc_sdf_node nodes[ISOMC_SDF_MAX_NODES];
    uint32_t n;
};

/* One implicit-source sample: reference src/implicit/{sphere,torus,cylinder,rectangular_prism,csg}.rs
 * scalar impls and the translate of examples/common/sources.rs:38-43, same operation order. */
__device__ __forceinline__ float sdf_eval(const SdfProgram &P, float px, float py, float pz) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f, s5 = 0.f, s6 = 0.f, s7 = 0.f;
    float tx0 = 0.f, ty0 = 0.f, tz0 = 0.f, tx1 = 0.f, ty1 = 0.f, tz1 = 0.f;
    float tx2 = 0.f, ty2 = 0.f, tz2 = 0.f, tx3 = 0.f, ty3 = 0.f, tz3 = 0.f;
#define ISOMC_PUSH(v) do { s7 = s6; s6 = s5; s5 = s4; s4 = s3; s3 = s2; s2 = s1; s1 = s0; s0 = (v); } while (0)
#define ISOMC_POP1() do { s1 = s2; s2 = s3; s3 = s4; s4 = s5; s5 = s6; s6 = s7; } while (0)
    for (uint32_t i = 0; i < P.n; ++i) {
        const uint32_t op = P.nodes[i].op;
        const float a = P.nodes[i].a, b = P.nodes[i].b, c = P.nodes[i].c;
        switch (op) {
        case ISOMC_SDF_SPHERE: {
            float l2 = __fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz));
            ISOMC_PUSH(__fsub_rn(__fsqrt_rn(l2), a));
        } break;
        case ISOMC_SDF_TORUS: {
            float qx = __fsub_rn(fabsf(__fsqrt_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)))), a);
            float len = __fsqrt_rn(__fadd_rn(__fmul_rn(qx, qx), __fmul_rn(pz, pz)));
            ISOMC_PUSH(__fsub_rn(len, b));
        } break;
        case ISOMC_SDF_CYLINDER: {
            float qx = __fsub_rn(fabsf(__fsqrt_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)))), a);
            float qz = __fsub_rn(fabsf(pz), b);
            float dx = fmaxf(qx, 0.0f), dy = fmaxf(qz, 0.0f);
            float dl = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(0.0f, 0.0f)));
            ISOMC_PUSH(__fadd_rn(fminf(fmaxf(qx, qz), 0.0f), dl));
        } break;
        case ISOMC_SDF_PRISM: {
            float qx = __fsub_rn(fabsf(px), a), qy = __fsub_rn(fabsf(py), b), qz = __fsub_rn(fabsf(pz), c);
            float mx = fmaxf(qx, 0.0f), my = fmaxf(qy, 0.0f), mz = fmaxf(qz, 0.0f);
            float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(mx, mx), __fmul_rn(my, my)), __fmul_rn(mz, mz)));
            float mc = fmaxf(qx, fmaxf(qy, qz));
            ISOMC_PUSH(__fadd_rn(len, fminf(mc, 0.0f)));
        } break;
        case ISOMC_SDF_UNION: { float r = fminf(s1, s0); ISOMC_POP1(); s0 = r; } break;
        case ISOMC_SDF_INTERSECTION: { float r = fmaxf(s1, s0); ISOMC_POP1(); s0 = r; } break;
        case ISOMC_SDF_DIFFERENCE: { float r = fmaxf(s0, -s1); ISOMC_POP1(); s0 = r; } break;
        case ISOMC_SDF_TRANSLATE_PUSH:
            tx3 = tx2; ty3 = ty2; tz3 = tz2; tx2 = tx1; ty2 = ty1; tz2 = tz1;
            tx1 = tx0; ty1 = ty0; tz1 = tz0; tx0 = px; ty0 = py; tz0 = pz;
            px = __fsub_rn(px, a); py = __fsub_rn(py, b); pz = __fsub_rn(pz, c);
            break;
        case ISOMC_SDF_TRANSLATE_POP:
            px = tx0; py = ty0; pz = tz0;
            tx0 = tx1; ty0 = ty1; tz0 = tz1; tx1 = tx2; ty1 = ty2; tz1 = tz2;
            tx2 = tx3; ty2 = ty3; tz2 = tz3;
            break;
        default: break;
        }
    }
#undef ISOMC_PUSH
#undef ISOMC_POP1
    return s0;
}

/* Sources as seen by the kernels: value at lattice point (x, y, local layer lz / global gz). */
struct GridSrc {
    const float *__restrict__ p; /* first sample layer of the handle's slab */
    __device__ __forceinline__ float at(const Geo &g, uint32_t x, uint32_t y, uint32_t lz) const {
        return __ldg(p + ((uint64_t)lz * g.N + y) * g.N + x);
    }
};
struct SdfSrc {
    SdfProgram prog;
    __device__ __forceinline__ float at(const Geo &g, uint32_t x, uint32_t y, uint32_t lz) const {
        /* primal_grid.rs:50,63-67: (i as f32) * one_over_size */
        return sdf_eval(prog, __fmul_rn((float)x, g.inv), __fmul_rn((float)y, g.inv),
                        __fmul_rn((float)(g.gz0 + lz), g.inv));
    }
};

#endif
