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
#include <math.h>
#include <stdint.h>

#include "../../include/isomc.h"

#define ISOMC_VAL_DEPTH 8   /* value stack held in registers */
#define ISOMC_TR_DEPTH 4    /* nested translations */

#define ISOMC_HD __host__ __device__ __forceinline__

ISOMC_HD uint32_t hd_popc(uint32_t v) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__popc(v);
#else
    return (uint32_t)__builtin_popcount(v);
#endif
}
ISOMC_HD uint32_t hd_ffs0(uint32_t v) { /* index of the lowest set bit, v != 0 */
#ifdef __CUDA_ARCH__
    return (uint32_t)__ffs((int)v) - 1u;
#else
    return (uint32_t)__builtin_ctz(v);
#endif
}
ISOMC_HD float hd_mul(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    return a * b; /* host model is compiled with -ffp-contract=off */
#endif
}
ISOMC_HD float hd_add(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
ISOMC_HD float hd_sub(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
ISOMC_HD float hd_div(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}

ISOMC_HD float hd_sqrt(float a) {
#ifdef __CUDA_ARCH__
    return __fsqrt_rn(a);
#else
    return sqrtf(a);
#endif
}

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
    uint64_t row_magic; /* ceil(2^40 / ncx): row / ncx == (row * row_magic) >> 40 for row < 2^26, ncx < 2^13 */
    uint32_t zper;   /* 0, or (batched chunks) sample layers per chunk = N+1: the handle's layers are B lattices stacked in z;
                        layer l belongs to chunk l / zper at z = l % zper, and cell layer z = zper-1 (between two chunks) is dead */
    uint32_t zmagic; /* 0, or ceil(2^32 / zper): l / zper == (l * zmagic) >> 32 for l * (zper + 1) < 2^32 (no division in the kernels) */
};

/* z of a (cell or sample) layer within its lattice: what the reference's loop variable is (primal_grid.rs:59-67) */
/* chunk of a layer of a batch handle (0 otherwise): multiply-high, no division (an integer division per list entry cost the
 * emission kernel 9 % of its instructions) */
ISOMC_HD uint32_t geo_chunk(const Geo &g, uint32_t lz) {
#ifdef __CUDA_ARCH__
    return __umulhi(lz, g.zmagic);
#else
    return (uint32_t)(((uint64_t)lz * g.zmagic) >> 32);
#endif
}
ISOMC_HD uint32_t geo_z(const Geo &g, uint32_t lz) { return g.gz0 + lz - geo_chunk(g, lz) * g.zper; } /* (a batch handle has gz0 == 0) */
ISOMC_HD bool geo_dead(const Geo &g, uint32_t lz) { return g.zper != 0 && lz - geo_chunk(g, lz) * g.zper == g.zper - 1; }
static inline uint32_t geo_zmagic(uint32_t zper) { return zper ? (uint32_t)(((1ull << 32) + zper - 1) / zper) : 0u; }

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

/*
 * Chain form of an SDF program: ((L0 op0 L1) op1 L2) ... -- every tree the reference's demos build
 * (examples/sampler.rs:70-95) is left-deep.  No stacks, no per-node dispatch beyond one switch per leaf;
 * same arithmetic, same operation order as sdf_eval.  Built on the host by sdf_to_chain(); programs that
 * are not left-deep (or nest more than two translations around a leaf) keep the generic interpreter.
 */
#define ISOMC_CHAIN_MAX_LEAVES 8
struct SdfLeaf {
    uint32_t type;      /* ISOMC_SDF_SPHERE .. PRISM */
    float a, b, c;
    uint32_t n_off;     /* translations around this leaf, outermost first (applied in that order) */
    float off[2][3];
    uint32_t op;        /* how this leaf combines with the chain so far (unused for leaf 0) */
};
struct SdfChain {
    SdfLeaf leaf[ISOMC_CHAIN_MAX_LEAVES];
    uint32_t n;
};

__device__ __forceinline__ float sdf_leaf(const SdfLeaf &L, float px, float py, float pz) {
    if (L.n_off > 0) { px = __fsub_rn(px, L.off[0][0]); py = __fsub_rn(py, L.off[0][1]); pz = __fsub_rn(pz, L.off[0][2]); }
    if (L.n_off > 1) { px = __fsub_rn(px, L.off[1][0]); py = __fsub_rn(py, L.off[1][1]); pz = __fsub_rn(pz, L.off[1][2]); }
    switch (L.type) {
    case ISOMC_SDF_SPHERE:
        return __fsub_rn(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz))), L.a);
    case ISOMC_SDF_TORUS: {
        const float qx = __fsub_rn(fabsf(__fsqrt_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)))), L.a);
        return __fsub_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(qx, qx), __fmul_rn(pz, pz))), L.b);
    }
    case ISOMC_SDF_CYLINDER: {
        const float qx = __fsub_rn(fabsf(__fsqrt_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)))), L.a);
        const float qz = __fsub_rn(fabsf(pz), L.b);
        const float dx = fmaxf(qx, 0.0f), dy = fmaxf(qz, 0.0f);
        const float dl = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(0.0f, 0.0f)));
        return __fadd_rn(fminf(fmaxf(qx, qz), 0.0f), dl);
    }
    default: { /* ISOMC_SDF_PRISM */
        const float qx = __fsub_rn(fabsf(px), L.a), qy = __fsub_rn(fabsf(py), L.b), qz = __fsub_rn(fabsf(pz), L.c);
        const float mx = fmaxf(qx, 0.0f), my = fmaxf(qy, 0.0f), mz = fmaxf(qz, 0.0f);
        const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(mx, mx), __fmul_rn(my, my)), __fmul_rn(mz, mz)));
        return __fadd_rn(len, fminf(fmaxf(qx, fmaxf(qy, qz)), 0.0f));
    }
    }
}

__device__ __forceinline__ float sdf_chain_eval(const SdfChain &C, float px, float py, float pz) {
    float acc = sdf_leaf(C.leaf[0], px, py, pz);
    for (uint32_t i = 1; i < C.n; ++i) {
        const float v = sdf_leaf(C.leaf[i], px, py, pz);
        const uint32_t op = C.leaf[i].op;
        acc = op == ISOMC_SDF_UNION ? fminf(acc, v) : op == ISOMC_SDF_INTERSECTION ? fmaxf(acc, v) : fmaxf(v, -acc);
    }
    return acc;
}

/* host: postfix program -> chain; false if the program is not a left-deep chain */
static inline bool sdf_to_chain(const SdfProgram &P, SdfChain *out) {
    float offs[ISOMC_TR_DEPTH][3];
    int n_off = 0, depth = 0; /* depth: values on the stack (1 = chain so far, 2 = chain + pending leaf) */
    SdfChain C;
    C.n = 0;
    for (uint32_t i = 0; i < P.n; ++i) {
        const isomc_sdf_node &nd = P.nodes[i];
        switch (nd.op) {
        case ISOMC_SDF_SPHERE: case ISOMC_SDF_TORUS: case ISOMC_SDF_CYLINDER: case ISOMC_SDF_PRISM: {
            if (depth >= 2 || C.n >= ISOMC_CHAIN_MAX_LEAVES || n_off > 2) return false;
            if (depth == 1 && C.n == 0) return false;
            SdfLeaf &L = C.leaf[C.n++];
            L.type = nd.op; L.a = nd.a; L.b = nd.b; L.c = nd.c; L.op = 0;
            L.n_off = (uint32_t)n_off;
            for (int k = 0; k < 2; ++k)
                for (int j = 0; j < 3; ++j) L.off[k][j] = k < n_off ? offs[k][j] : 0.0f;
            ++depth;
        } break;
        case ISOMC_SDF_UNION: case ISOMC_SDF_INTERSECTION: case ISOMC_SDF_DIFFERENCE:
            if (depth != 2 || C.n < 2) return false;
            C.leaf[C.n - 1].op = nd.op; /* (chain, leaf): first-pushed = chain = field `a` */
            depth = 1;
            break;
        case ISOMC_SDF_TRANSLATE_PUSH:
            /* a translation opened while a chain value is pending would have to apply to later leaves only: fine,
             * but one opened around an already-combined value cannot be expressed -> only leaves inherit offsets */
            if (n_off >= ISOMC_TR_DEPTH) return false;
            offs[n_off][0] = nd.a; offs[n_off][1] = nd.b; offs[n_off][2] = nd.c;
            ++n_off;
            break;
        case ISOMC_SDF_TRANSLATE_POP:
            if (n_off <= 0) return false;
            --n_off;
            break;
        default: return false;
        }
    }
    if (depth != 1 || C.n < 1) return false;
    *out = C;
    return true;
}

/*
 * Directed distances: VectorSource::sample_vector of the implicit shapes (reference src/implicit/sphere.rs:41-57,
 * torus.rs:47-97, cylinder.rs:50-72, rectangular_prism.rs:42-74, csg.rs:41-45,74-78,102-106; the translation of
 * examples/common/sources.rs:46-51), same operation order; f32::min / f32::max ignore a NaN operand like fminf / fmaxf.
 * Host + device: tests/list_model.cu runs the same function on the CPU.
 */
#define ISOMC_F32_MAX 3.40282347e+38f
struct Vec3f { float x, y, z; };

ISOMC_HD Vec3f sdf_eval_vec(const SdfProgram &P, float px, float py, float pz) {
    Vec3f vs[ISOMC_VAL_DEPTH];
    float ts[ISOMC_TR_DEPTH][3];
    int nv = 0, nt = 0;
    for (uint32_t i = 0; i < P.n; ++i) {
        const uint32_t op = P.nodes[i].op;
        const float pa = P.nodes[i].a, pb = P.nodes[i].b, pc = P.nodes[i].c;
        const float ax = fabsf(px), ay = fabsf(py), az = fabsf(pz); /* the point flipped into the positive octant */
        Vec3f r;
        r.x = r.y = r.z = 0.0f;
        if (op == ISOMC_SDF_SPHERE) {
            const float r2 = hd_mul(pa, pa);
            const float l_yz = hd_sub(r2, hd_add(hd_mul(ay, ay), hd_mul(az, az)));
            const float l_xz = hd_sub(r2, hd_add(hd_mul(ax, ax), hd_mul(az, az)));
            const float l_xy = hd_sub(r2, hd_add(hd_mul(ax, ax), hd_mul(ay, ay)));
            r.x = l_yz < 0.0f ? ISOMC_F32_MAX : hd_sub(ax, hd_sqrt(l_yz));
            r.y = l_xz < 0.0f ? ISOMC_F32_MAX : hd_sub(ay, hd_sqrt(l_xz));
            r.z = l_xy < 0.0f ? ISOMC_F32_MAX : hd_sub(az, hd_sqrt(l_xy));
        } else if (op == ISOMC_SDF_TORUS) {
            const float R = pa, tr = pb;
            const float l_xy = hd_sub(hd_sqrt(hd_add(hd_mul(ax, ax), hd_mul(ay, ay))), R);
            const float tz = hd_sqrt(hd_sub(hd_mul(tr, tr), hd_mul(az, az))); /* NaN beyond the tube: comparisons are false then */
            const float rx = hd_add(R, tz), ry = hd_sub(R, tz);
            if (az > tr || ay > hd_add(R, tz)) r.x = ISOMC_F32_MAX;
            else if (ax == 0.0f) r.x = hd_sub(fabsf(hd_sub(ay, R)), tr);
            else r.x = fmaxf(hd_sub(ax, hd_sqrt(hd_sub(hd_mul(rx, rx), hd_mul(ay, ay)))), hd_sub(hd_sqrt(hd_sub(hd_mul(ry, ry), hd_mul(ay, ay))), ax));
            if (az > tr || ax > hd_add(R, tz)) r.y = ISOMC_F32_MAX;
            else if (ay == 0.0f) r.y = hd_sub(fabsf(hd_sub(ax, R)), tr);
            else r.y = fmaxf(hd_sub(ay, hd_sqrt(hd_sub(hd_mul(rx, rx), hd_mul(ax, ax)))), hd_sub(hd_sqrt(hd_sub(hd_mul(ry, ry), hd_mul(ax, ax))), ay));
            if (fabsf(l_xy) > tr) r.z = ISOMC_F32_MAX;
            else r.z = hd_sub(az, hd_sqrt(hd_sub(hd_mul(tr, tr), hd_mul(l_xy, l_xy))));
        } else if (op == ISOMC_SDF_CYLINDER) {
            const float R = pa, h = pb, R2 = hd_mul(R, R);
            r.x = (az > h || ay > R) ? ISOMC_F32_MAX : hd_sub(ax, hd_sqrt(hd_sub(R2, hd_mul(ay, ay))));
            r.y = (az > h || ax > R) ? ISOMC_F32_MAX : hd_sub(ay, hd_sqrt(hd_sub(R2, hd_mul(ax, ax))));
            r.z = (hd_add(hd_mul(ax, ax), hd_mul(ay, ay)) > R2) ? ISOMC_F32_MAX : hd_sub(az, h);
        } else if (op == ISOMC_SDF_PRISM) {
            const bool ox = hd_sub(ax, pa) > 0.0f, oy = hd_sub(ay, pb) > 0.0f, oz = hd_sub(az, pc) > 0.0f;
            const float mx = hd_mul((oy || oz) ? 1.0f : -1.0f, ISOMC_F32_MAX), my = hd_mul((ox || oz) ? 1.0f : -1.0f, ISOMC_F32_MAX);
            const float mz = hd_mul((ox || oy) ? 1.0f : -1.0f, ISOMC_F32_MAX);
            const bool inside = ax < pa && ay < pb && az < pc;
            const float cx = inside ? fmaxf(ax, pa) : fminf(ax, pa), cy = inside ? fmaxf(ay, pb) : fminf(ay, pb);
            const float cz = inside ? fmaxf(az, pc) : fminf(az, pc);
            r.x = fmaxf(hd_sub(ax, cx), mx); r.y = fmaxf(hd_sub(ay, cy), my); r.z = fmaxf(hd_sub(az, cz), mz);
        } else if (op == ISOMC_SDF_UNION || op == ISOMC_SDF_INTERSECTION || op == ISOMC_SDF_DIFFERENCE) {
            const Vec3f A = vs[nv - 2], B = vs[nv - 1]; /* first-pushed = field `a` */
            if (op == ISOMC_SDF_UNION) { r.x = fminf(A.x, B.x); r.y = fminf(A.y, B.y); r.z = fminf(A.z, B.z); }
            else if (op == ISOMC_SDF_INTERSECTION) { r.x = fmaxf(A.x, B.x); r.y = fmaxf(A.y, B.y); r.z = fmaxf(A.z, B.z); }
            else { r.x = fmaxf(B.x, -A.x); r.y = fmaxf(B.y, -A.y); r.z = fmaxf(B.z, -A.z); }
            vs[nv - 2] = r;
            --nv;
            continue;
        } else if (op == ISOMC_SDF_TRANSLATE_PUSH) {
            ts[nt][0] = px; ts[nt][1] = py; ts[nt][2] = pz;
            ++nt;
            px = hd_sub(px, pa); py = hd_sub(py, pb); pz = hd_sub(pz, pc);
            continue;
        } else { /* ISOMC_SDF_TRANSLATE_POP (programs are validated on the host) */
            --nt;
            px = ts[nt][0]; py = ts[nt][1]; pz = ts[nt][2];
            continue;
        }
        vs[nv++] = r;
    }
    return vs[0];
}

/* Sources as seen by the kernels: value at lattice point (x, y, local layer lz / global gz). */
struct GridSrc {
    const float *__restrict__ p; /* first sample layer of the handle's slab */
    __device__ __forceinline__ float at(const Geo &g, uint32_t x, uint32_t y, uint32_t lz) const {
        return __ldg(p + ((uint64_t)lz * g.N + y) * g.N + x);
    }
    /* the two ends of a lattice edge along `axis` (a scalar field has one value whatever the axis) */
    __device__ __forceinline__ void pair(const Geo &g, uint32_t ux, uint32_t uy, uint32_t uz, uint32_t vx, uint32_t vy, uint32_t vz,
                                         uint32_t, float &a, float &b) const {
        a = at(g, ux, uy, uz);
        b = at(g, vx, vy, vz);
    }
    /* the (a, b) ends of the three edges a cell creates, all meeting at corner 6 = (x+1, y+1, lz+1):
     * e5 = corners 5 -> 6 (y), e6 = corners 6 -> 7 (x), e10 = corners 2 -> 6 (z); only the requested ones are loaded */
    __device__ __forceinline__ void corner6(const Geo &g, uint32_t x, uint32_t y, uint32_t lz, bool n5, bool n6, bool n10, float &a5,
                                            float &b5, float &a6, float &b6, float &a10, float &b10) const {
        const float *q = p + ((uint64_t)(lz + 1) * g.N + (y + 1)) * g.N + (x + 1);
        const float s6 = __ldg(q);
        a5 = n5 ? __ldg(q - g.N) : 0.0f; b5 = s6;
        a6 = s6; b6 = n6 ? __ldg(q - 1) : 0.0f;
        a10 = n10 ? __ldg(q - (uint64_t)g.N * g.N) : 0.0f; b10 = s6;
    }
};
/* pair() / corner6() for scalar sources that evaluate instead of loading */
#define ISOMC_SCALAR_EDGE_SAMPLES                                                                                                  \
    __device__ __forceinline__ void pair(const Geo &g, uint32_t ux, uint32_t uy, uint32_t uz, uint32_t vx, uint32_t vy, uint32_t vz, \
                                         uint32_t, float &a, float &b) const {                                                     \
        a = at(g, ux, uy, uz);                                                                                                      \
        b = at(g, vx, vy, vz);                                                                                                      \
    }                                                                                                                               \
    __device__ __forceinline__ void corner6(const Geo &g, uint32_t x, uint32_t y, uint32_t lz, bool n5, bool n6, bool n10, float &a5, \
                                            float &b5, float &a6, float &b6, float &a10, float &b10) const {                       \
        const float s6 = at(g, x + 1, y + 1, lz + 1);                                                                               \
        a5 = n5 ? at(g, x + 1, y, lz + 1) : 0.0f; b5 = s6;                                                                          \
        a6 = s6; b6 = n6 ? at(g, x, y + 1, lz + 1) : 0.0f;                                                                          \
        a10 = n10 ? at(g, x + 1, y + 1, lz) : 0.0f; b10 = s6;                                                                       \
    }
struct SdfSrc {
    SdfProgram prog;
    __device__ __forceinline__ float at(const Geo &g, uint32_t x, uint32_t y, uint32_t lz) const {
        /* primal_grid.rs:50,63-67: (i as f32) * one_over_size */
        return sdf_eval(prog, __fmul_rn((float)x, g.inv), __fmul_rn((float)y, g.inv),
                        __fmul_rn((float)geo_z(g, lz), g.inv));
    }
    ISOMC_SCALAR_EDGE_SAMPLES
};
struct SdfChainSrc {
    SdfChain chain;
    __device__ __forceinline__ float at(const Geo &g, uint32_t x, uint32_t y, uint32_t lz) const {
        return sdf_chain_eval(chain, __fmul_rn((float)x, g.inv), __fmul_rn((float)y, g.inv),
                              __fmul_rn((float)geo_z(g, lz), g.inv));
    }
    ISOMC_SCALAR_EDGE_SAMPLES
};

/* batched chunks: B implicit trees, one per stacked lattice (the programs live in global memory; a warp's row is one chunk's) */
struct SdfBatchSrc {
    const SdfProgram *progs;
    __device__ __forceinline__ float at(const Geo &g, uint32_t x, uint32_t y, uint32_t lz) const {
        const uint32_t b = geo_chunk(g, lz);
        return sdf_eval(progs[b], __fmul_rn((float)x, g.inv), __fmul_rn((float)y, g.inv), __fmul_rn((float)(lz - b * g.zper), g.inv));
    }
    ISOMC_SCALAR_EDGE_SAMPLES
};

/* MarchingCubes<Directed> over an implicit tree (reference src/distance.rs:72-104): outside iff any component > 0,
 * crossings interpolated from the component along the edge's own axis.  The sampling members of a source with a vec() */
#define ISOMC_VECTOR_EDGE_SAMPLES                                                                                                   \
    /* value for the sign test: positive iff some component is positive (NaN components never are) */                              \
    __device__ __forceinline__ float at(const Geo &g, uint32_t x, uint32_t y, uint32_t lz) const {                                  \
        const Vec3f v = vec(g, x, y, lz);                                                                                           \
        return (v.x > 0.0f || v.y > 0.0f || v.z > 0.0f) ? 1.0f : -1.0f;                                                            \
    }                                                                                                                               \
    static __device__ __forceinline__ float comp(const Vec3f &v, uint32_t axis) { return axis == 0 ? v.x : axis == 1 ? v.y : v.z; } \
    __device__ __forceinline__ void pair(const Geo &g, uint32_t ux, uint32_t uy, uint32_t uz, uint32_t vx, uint32_t vy, uint32_t vz, \
                                         uint32_t axis, float &a, float &b) const {                                                 \
        a = comp(vec(g, ux, uy, uz), axis);                                                                                         \
        b = comp(vec(g, vx, vy, vz), axis);                                                                                         \
    }                                                                                                                               \
    __device__ __forceinline__ void corner6(const Geo &g, uint32_t x, uint32_t y, uint32_t lz, bool n5, bool n6, bool n10, float &a5, \
                                            float &b5, float &a6, float &b6, float &a10, float &b10) const {                       \
        const Vec3f s6 = vec(g, x + 1, y + 1, lz + 1);                                                                              \
        a5 = n5 ? vec(g, x + 1, y, lz + 1).y : 0.0f; b5 = s6.y;                                                                     \
        a6 = s6.x; b6 = n6 ? vec(g, x, y + 1, lz + 1).x : 0.0f;                                                                     \
        a10 = n10 ? vec(g, x + 1, y + 1, lz).z : 0.0f; b10 = s6.z;                                                                  \
    }

struct SdfDirSrc {
    SdfProgram prog;
    __device__ __forceinline__ Vec3f vec(const Geo &g, uint32_t x, uint32_t y, uint32_t lz) const {
        return sdf_eval_vec(prog, __fmul_rn((float)x, g.inv), __fmul_rn((float)y, g.inv), __fmul_rn((float)geo_z(g, lz), g.inv));
    }
    ISOMC_VECTOR_EDGE_SAMPLES
};

/* ... and over the B trees of a batch */
struct SdfBatchDirSrc {
    const SdfProgram *progs;
    __device__ __forceinline__ Vec3f vec(const Geo &g, uint32_t x, uint32_t y, uint32_t lz) const {
        const uint32_t b = geo_chunk(g, lz);
        return sdf_eval_vec(progs[b], __fmul_rn((float)x, g.inv), __fmul_rn((float)y, g.inv), __fmul_rn((float)(lz - b * g.zper), g.inv));
    }
    ISOMC_VECTOR_EDGE_SAMPLES
};

#endif
