/*
 * isomc_cell.cuh -- the per-item arithmetic of the active-cell-list path, written once as
 * `__host__ __device__` code: the kernels (isomc_list_kernels.cu) wrap it with the warp-level
 * plumbing, and tests/list_model.cu runs the very same functions item by item on the host against the
 * CPU restatement of the reference (the image this is developed in has no GPU; the model pins the entry formats, the
 * neighbour lookups and the table use before a kernel ever runs).
 *
 * Active-cell list.  k_count_list appends one entry per active cell (a cell whose 8 corners are not all
 * on one side, reference src/marching_cubes_impl.rs:26-37 + marching_cubes_tables.rs:49-70):
 *
 *     ent[k]    = { vrel | tseg << 16,  x | ci' << 16 }     ent_yz[k] = y | local layer << 16
 *
 *   vrel  vertices created by the earlier cells of the same cell row      (id = rowPV[row] + vrel)
 *   tseg  triangles of the earlier cells of the same 32-cell segment
 *   ci'   natural cube index (isomc_tables.h)
 *
 * and, per 32-cell segment that has active cells,
 *
 *     segrec[row * nsegx + s] = { list position of the segment's first active cell, active mask }
 *     segtpre[row * nsegx + s] = triangles of the earlier segments of the same cell row
 *                                                          (triangle slot = rowPT[row] + segtpre + tseg)
 *
 * so that "the entry of cell (x, row)" is segrec.x + popc(segrec.y & below(x & 31)).  Entries of one segment
 * are contiguous and in x order; beyond that the list order is arbitrary (blocks of LIST_BLOCK entries are
 * handed out by an atomic counter), which is fine: every output slot is computed, never appended.
 *
 * emit_cell() is the whole of `march_cube` + the index-cache lookups (reference
 * src/marching_cubes_impl.rs:102-117, src/index_cache.rs:38-60, src/mesh.rs:240-251) for one cell:
 * the ids of its crossed edges come from the entries of the (at most 7) earlier cells that created them,
 * its own created vertices are interpolated (src/distance.rs:64-69) and written to their final slots, its
 * triangles are written to theirs.
 */
#ifndef ISOMC_CELL_CUH
#define ISOMC_CELL_CUH

#include <stdint.h>

#include "isomc_device.cuh"
#include "isomc_tables.h"

constexpr uint32_t LIST_BLOCK = 256; /* entries per list block = threads per emit CTA */

/* tables emit_cell needs (copied into shared memory by the kernel); derived from McTables on the host */
struct EmitTab {
    uint64_t tri[256];        /* the case's edges, 4 bits each, table order (marching_cubes_tables.rs:74-331) */
    uint16_t before[256][12]; /* edges that appear before e in the case's first-appearance order */
    uint16_t emask[256];      /* crossed edges */
    uint16_t ownmask[8];      /* edges a cell with boundary flags b creates */
    uint16_t stepedges[8][8]; /* [b][d]: edges of a cell with flags b that were created by the cell at -d (d = dx|dy<<1|dz<<2) */
    uint8_t owner[8][12];     /* [b][e]: d | e' << 4 */
    uint8_t ends[12];
    uint8_t ntri[256];
    uint8_t rank3[256];       /* interior cells: rank of e5 | e6 << 2 | e10 << 4 among the three edges such a cell creates */
    uint8_t pad[4];
    uint32_t eloc[2][12];     /* tile path: plane location of edge e on a cell layer of parity p (isomc_tile.cuh: tile_edge_loc) */
};

static inline void isomc_build_emit_tab(const McTables &m, EmitTab *t) {
    memset(t, 0, sizeof *t);
    memcpy(t->tri, m.tri, sizeof t->tri);
    memcpy(t->before, m.before, sizeof t->before);
    memcpy(t->emask, m.emask, sizeof t->emask);
    memcpy(t->ownmask, m.ownmask, sizeof t->ownmask);
    memcpy(t->owner, m.owner, sizeof t->owner);
    memcpy(t->ends, m.ends, sizeof t->ends);
    memcpy(t->ntri, m.ntri, sizeof t->ntri);
    memcpy(t->rank3, m.rank3, sizeof t->rank3);
    for (int b = 0; b < 8; ++b)
        for (int e = 0; e < 12; ++e) t->stepedges[b][m.owner[b][e] & 7] |= (uint16_t)(1u << e);
}

struct ListBufs {
    uint2 *ent;
    uint32_t *ent_yz;
    uint2 *segrec;
    uint32_t *segtpre;
    uint32_t *blkfill;  /* valid entries of each handed-out block */
    uint32_t *ctr;      /* [0] blocks handed out so far (may exceed cap_blocks: the host then grows the list and re-runs) */
    uint32_t cap_blocks;
    const uint32_t *chunkV, *chunkT; /* batched chunks: output slot of a chunk's first vertex / triangle (ids stay chunk-local); else NULL */
};

ISOMC_HD uint32_t cell_flags(const Geo &g, uint32_t x, uint32_t y, uint32_t lz) {
    return (x == 0 ? 1u : 0u) | (y == 0 ? 2u : 0u) | (geo_z(g, lz) == 0 ? 4u : 0u);
}

struct EmitArgs {
    const uint32_t *rowPV, *rowPT; /* exclusive prefixes over cell rows (local numbering incl. a slab's ghost layer) */
    uint32_t vofs;                 /* local id -> global id */
    uint32_t ghostV, ghostT;       /* vertices / triangles of the ghost layer: local id/slot -> output slot */
    uint32_t first_own_layer;
    uint64_t cap_v, cap_t;
    float *xyz;
    uint32_t *idx;
};

/* entry of the active cell (x2, row2): position from the segment record, then { id of its first vertex, cube index } */
ISOMC_HD void creator_lookup(const Geo &g, const ListBufs &L, const uint32_t *rowPV, uint32_t x2, uint32_t row2, uint32_t &vid2,
                             uint32_t &ci2) {
    const uint2 rec = L.segrec[(uint64_t)row2 * g.nsegx + (x2 >> 5)];
    const uint32_t k2 = rec.x + hd_popc(rec.y & ((1u << (x2 & 31u)) - 1u));
    vid2 = 0; ci2 = 0;
    if (k2 < L.cap_blocks * LIST_BLOCK) { /* (out of range only after a list overflow; the host re-runs then) */
        const uint2 e2 = L.ent[k2];
        vid2 = rowPV[row2] + (e2.x & 0xFFFFu);
        ci2 = e2.y >> 16 & 255u;
    }
}

/*
 * One active cell: list entry k = (ea, yz).  eid = 12 words of scratch (stride eid_stride) for the ids of the cell's
 * crossed edges.  Src::at(g, x, y, lz) is the sample at a lattice point of the handle's slab.
 */
template <class Src>
ISOMC_HD void emit_cell(const Geo &g, const Src &src, const EmitTab &T, const ListBufs &L, const EmitArgs &A, uint64_t k,
                        uint2 ea, uint32_t yz, uint32_t *eid, uint32_t eid_stride) {
    const uint32_t x = ea.y & 0xFFFFu, ci = ea.y >> 16 & 255u, y = yz & 0xFFFFu, lz = yz >> 16;
    if (lz < A.first_own_layer) return; /* ghost layer of a slab: looked up by our first layer, emitted by the previous rank */
    const uint32_t row = lz * g.ncx + y;
    const uint32_t em = T.emask[ci];
    const uint32_t vid = A.rowPV[row] + (ea.x & 0xFFFFu);
    const uint32_t gz = geo_z(g, lz);
    /* id -> output slot: a slab drops its ghost layer's vertices; batched chunks keep chunk-local ids and add the chunk's base */
    uint32_t vbase = 0u - A.ghostV, tbase = 0u - A.ghostT;
    if (L.chunkV) { const uint32_t b = geo_chunk(g, lz); vbase = L.chunkV[b]; tbase = L.chunkT[b]; }
    /* slot of the cell's first triangle: requested here, with the first wave of loads (the compiler will not move these loads
     * up across the vertex stores below by itself) */
    const uint32_t tslot32 = A.rowPT[row] + L.segtpre[(uint64_t)row * g.nsegx + (x >> 5)] + (ea.x >> 16) + tbase;
    uint32_t owned;

    if (x >= 2 && y >= 2 && gz >= 2) {
        /* Interior cell whose six possible creators are interior too: every cell involved creates exactly its crossed
         * e5 (y edge), e6 (x edge), e10 (z edge), numbered by rank3[] of its own case.  Which earlier cell created
         * which of my edges (isomc_tables.h owner[0][]):  -x: e7 as e5, e11 as e10;  -y: e4 as e6, e9 as e10;
         * -x-y: e8 as e10;  -z: e1 as e5, e2 as e6;  -x-z: e3 as e5;  -y-z: e0 as e6.  Straight-line, predicated. */
        owned = em & (1u << 5 | 1u << 6 | 1u << 10);
        /* samples for the vertices this cell creates: requested first, used last */
        const bool o5 = (em >> 5 & 1u) != 0, o6 = (em >> 6 & 1u) != 0, o10 = (em >> 10 & 1u) != 0;
        float a5 = 0.0f, b5 = 0.0f, a6 = 0.0f, b6 = 0.0f, a10 = 0.0f, b10 = 0.0f;
        if (owned) src.corner6(g, x, y, lz, o5, o6, o10, a5, b5, a6, b6, a10, b10);
        const uint32_t r3 = T.rank3[ci];
        eid[5 * eid_stride] = vid + (r3 & 3u);
        eid[6 * eid_stride] = vid + (r3 >> 2 & 3u);
        eid[10 * eid_stride] = vid + (r3 >> 4 & 3u);
        /* all segment records first, then all entries, then the ranks: the loads of one stage are independent, so their
         * latencies overlap (the kernel is latency-bound, not bandwidth-bound) */
        const bool n1 = (em & (1u << 7 | 1u << 11)) != 0, n2 = (em & (1u << 4 | 1u << 9)) != 0, n3 = (em & (1u << 8)) != 0;
        const bool n4 = (em & (1u << 1 | 1u << 2)) != 0, n5 = (em & (1u << 3)) != 0, n6 = (em & 1u) != 0;
        const bool n1g = n1 && (x & 31u) == 0; /* -x neighbour in the previous segment; otherwise it is list entry k - 1 */
        const uint32_t rowy = row - 1, rowz = row - g.ncx, rowyz = row - g.ncx - 1;
        const uint32_t sx = x >> 5, sxm = (x - 1) >> 5, bx = (1u << (x & 31u)) - 1u, bxm = (1u << ((x - 1) & 31u)) - 1u;
        const uint2 z2 = make_uint2(0u, 0u);
        const uint2 rA = n1g ? L.segrec[(uint64_t)row * g.nsegx + sxm] : z2;
        const uint2 rB = n2 ? L.segrec[(uint64_t)rowy * g.nsegx + sx] : z2;
        const uint2 rC = n3 ? L.segrec[(uint64_t)rowy * g.nsegx + sxm] : z2;
        const uint2 rD = n4 ? L.segrec[(uint64_t)rowz * g.nsegx + sx] : z2;
        const uint2 rE = n5 ? L.segrec[(uint64_t)rowz * g.nsegx + sxm] : z2;
        const uint2 rF = n6 ? L.segrec[(uint64_t)rowyz * g.nsegx + sx] : z2;
        const uint32_t pvy = (n2 || n3) ? A.rowPV[rowy] : 0u, pvz = (n4 || n5) ? A.rowPV[rowz] : 0u, pvyz = n6 ? A.rowPV[rowyz] : 0u;
        const uint32_t cap = L.cap_blocks * LIST_BLOCK; /* (positions beyond it only after a list overflow; the host re-runs then) */
        const uint32_t kA = n1g ? rA.x + hd_popc(rA.y & bxm) : (uint32_t)k - 1u;
        const uint32_t kB = rB.x + hd_popc(rB.y & bx), kC = rC.x + hd_popc(rC.y & bxm);
        const uint32_t kD = rD.x + hd_popc(rD.y & bx), kE = rE.x + hd_popc(rE.y & bxm), kF = rF.x + hd_popc(rF.y & bx);
        const uint2 eA = (n1 && kA < cap) ? L.ent[kA] : z2, eB = (n2 && kB < cap) ? L.ent[kB] : z2;
        const uint2 eC = (n3 && kC < cap) ? L.ent[kC] : z2, eD = (n4 && kD < cap) ? L.ent[kD] : z2;
        const uint2 eE = (n5 && kE < cap) ? L.ent[kE] : z2, eF = (n6 && kF < cap) ? L.ent[kF] : z2;
        const uint32_t pv0 = vid - (ea.x & 0xFFFFu); /* rowPV[row] */
        uint32_t q;
        q = T.rank3[eA.y >> 16 & 255u];
        eid[7 * eid_stride] = pv0 + (eA.x & 0xFFFFu) + (q & 3u);
        eid[11 * eid_stride] = pv0 + (eA.x & 0xFFFFu) + (q >> 4 & 3u);
        q = T.rank3[eB.y >> 16 & 255u];
        eid[4 * eid_stride] = pvy + (eB.x & 0xFFFFu) + (q >> 2 & 3u);
        eid[9 * eid_stride] = pvy + (eB.x & 0xFFFFu) + (q >> 4 & 3u);
        eid[8 * eid_stride] = pvy + (eC.x & 0xFFFFu) + (T.rank3[eC.y >> 16 & 255u] >> 4 & 3u);
        q = T.rank3[eD.y >> 16 & 255u];
        eid[1 * eid_stride] = pvz + (eD.x & 0xFFFFu) + (q & 3u);
        eid[2 * eid_stride] = pvz + (eD.x & 0xFFFFu) + (q >> 2 & 3u);
        eid[3 * eid_stride] = pvz + (eE.x & 0xFFFFu) + (T.rank3[eE.y >> 16 & 255u] & 3u);
        eid[0] = pvyz + (eF.x & 0xFFFFu) + (T.rank3[eF.y >> 16 & 255u] >> 2 & 3u);
        /* the (at most three) vertices this cell creates all end at corner 6 = (x+1, y+1, z+1):
         *   e5 = corners 5 -> 6 (y edge), e6 = corners 6 -> 7 (x edge), e10 = corners 2 -> 6 (z edge) */
        if (owned) {
            const float fx0 = hd_mul((float)x, g.inv), fx1 = hd_mul((float)(x + 1), g.inv);
            const float fy0 = hd_mul((float)y, g.inv), fy1 = hd_mul((float)(y + 1), g.inv);
            const float fz0 = hd_mul((float)gz, g.inv), fz1 = hd_mul((float)(gz + 1), g.inv);
            const uint64_t s0 = (uint64_t)(vid + vbase);
            if (o5 && s0 + (r3 & 3u) < A.cap_v) {
                const float delta = hd_sub(b5, a5), t = (delta == 0.0f) ? 0.5f : hd_div(-a5, delta), omt = hd_sub(1.0f, t);
                float *o = A.xyz + 3 * (s0 + (r3 & 3u));
                o[0] = hd_add(hd_mul(fx1, omt), hd_mul(fx1, t));
                o[1] = hd_add(hd_mul(fy0, omt), hd_mul(fy1, t));
                o[2] = hd_add(hd_mul(fz1, omt), hd_mul(fz1, t));
            }
            if (o6 && s0 + (r3 >> 2 & 3u) < A.cap_v) {
                const float delta = hd_sub(b6, a6), t = (delta == 0.0f) ? 0.5f : hd_div(-a6, delta), omt = hd_sub(1.0f, t);
                float *o = A.xyz + 3 * (s0 + (r3 >> 2 & 3u));
                o[0] = hd_add(hd_mul(fx1, omt), hd_mul(fx0, t));
                o[1] = hd_add(hd_mul(fy1, omt), hd_mul(fy1, t));
                o[2] = hd_add(hd_mul(fz1, omt), hd_mul(fz1, t));
            }
            if (o10 && s0 + (r3 >> 4 & 3u) < A.cap_v) {
                const float delta = hd_sub(b10, a10), t = (delta == 0.0f) ? 0.5f : hd_div(-a10, delta), omt = hd_sub(1.0f, t);
                float *o = A.xyz + 3 * (s0 + (r3 >> 4 & 3u));
                o[0] = hd_add(hd_mul(fx1, omt), hd_mul(fx1, t));
                o[1] = hd_add(hd_mul(fy1, omt), hd_mul(fy1, t));
                o[2] = hd_add(hd_mul(fz0, omt), hd_mul(fz1, t));
            }
            owned = 0; /* done here */
        }
    } else {
        /* on or next to a low boundary face: creators and ranks through the general tables */
        const uint32_t bfl = cell_flags(g, x, y, lz);
        owned = em & T.ownmask[bfl];
        for (uint32_t m = owned; m; m &= m - 1) {
            const uint32_t e = hd_ffs0(m);
            eid[e * eid_stride] = vid + hd_popc(T.before[ci][e] & owned);
        }
        for (uint32_t d = 1; d < 8; ++d) {
            uint32_t es = em & T.stepedges[bfl][d];
            if (!es) continue;
            const uint32_t x2 = x - (d & 1u), y2 = y - (d >> 1 & 1u), lz2 = lz - (d >> 2 & 1u);
            uint32_t vid2, ci2;
            creator_lookup(g, L, A.rowPV, x2, lz2 * g.ncx + y2, vid2, ci2);
            const uint32_t own2 = T.ownmask[cell_flags(g, x2, y2, lz2)];
            for (; es; es &= es - 1) {
                const uint32_t e = hd_ffs0(es);
                eid[e * eid_stride] = vid2 + hd_popc(T.before[ci2][T.owner[bfl][e] >> 4] & own2);
            }
        }
    }

    /* vertices this cell creates: Signed::find_crossing_point (distance.rs:64-69) between the edge's ends in
     * EDGE_CONNECTION direction, corner coordinates = (i as f32) * inv (primal_grid.rs:50,63-67) */
    for (uint32_t m = owned; m; m &= m - 1) {
        const uint32_t e = hd_ffs0(m);
        const uint64_t slot = (uint64_t)(eid[e * eid_stride] + vbase);
        if (slot >= A.cap_v) continue;
        const uint32_t en = T.ends[e];
        const uint32_t ux = x + (en & 1u), uy = y + (en >> 1 & 1u), uz = lz + (en >> 2 & 1u);
        const uint32_t vx = x + (en >> 4 & 1u), vy = y + (en >> 5 & 1u), vz = lz + (en >> 6 & 1u);
        float a, b;
        src.pair(g, ux, uy, uz, vx, vy, vz, ((en ^ en >> 4) & 7u) >> 1, a, b); /* axis of the edge: 0 x, 1 y, 2 z */
        const float delta = hd_sub(b, a);
        const float t = (delta == 0.0f) ? 0.5f : hd_div(-a, delta);
        const float omt = hd_sub(1.0f, t);
        const float pax = hd_mul((float)ux, g.inv), pay = hd_mul((float)uy, g.inv), paz = hd_mul((float)(gz + (uz - lz)), g.inv);
        const float pbx = hd_mul((float)vx, g.inv), pby = hd_mul((float)vy, g.inv), pbz = hd_mul((float)(gz + (vz - lz)), g.inv);
        float *o = A.xyz + 3 * slot;
        o[0] = hd_add(hd_mul(pax, omt), hd_mul(pbx, t));
        o[1] = hd_add(hd_mul(pay, omt), hd_mul(pby, t));
        o[2] = hd_add(hd_mul(paz, omt), hd_mul(pbz, t));
    }

    /* triangles in table order (march_cube, marching_cubes_impl.rs:106-116) */
    const uint64_t tslot = (uint64_t)tslot32;
    uint32_t nt = T.ntri[ci];
    if (tslot >= A.cap_t) nt = 0;
    else if (tslot + nt > A.cap_t) nt = (uint32_t)(A.cap_t - tslot);
    const uint64_t tri = T.tri[ci];
    uint32_t *o = A.idx + 3 * tslot;
    /* at most five triangles per case: unrolled, so that the shifts and the store offsets are immediates and nothing is carried
     * from one triangle to the next (the rolled loop cost 31 instructions per triangle, most of them bookkeeping) */
#pragma unroll
    for (uint32_t t = 0; t < 5; ++t) {
        if (t < nt) {
            const uint32_t c = (uint32_t)(tri >> (12 * t)) & 0xFFFu;
            o[3 * t + 0] = eid[(c & 15u) * eid_stride] + A.vofs;
            o[3 * t + 1] = eid[(c >> 4 & 15u) * eid_stride] + A.vofs;
            o[3 * t + 2] = eid[(c >> 8) * eid_stride] + A.vofs;
        }
    }
}

/* ---- classification of one 32-cell segment from the inside bits ------------------------------ */

ISOMC_HD uint32_t hd_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { /* sh in [0, 31] */
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}

ISOMC_HD void hd_bs_add(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t m) {
    uint32_t k0 = c0 & m; c0 ^= m;
    uint32_t k1 = c1 & k0; c1 ^= k0;
    uint32_t k2 = c2 & k1; c2 ^= k1;
    c3 ^= k2;
}

struct SegClass {
    uint32_t act;            /* active cells */
    uint32_t p0, p1, p2, p3; /* bit planes of "vertices this cell creates" */
    uint32_t a0, b0, c0, d0; /* inside bits of the 4 corner rows at the cells' own x: (y,z) (y+1,z) (y,z+1) (y+1,z+1) */
    uint32_t nb;             /* bit 0 of the next word of each row: a | b << 1 | c << 2 | d << 3 */
};

/* w[8] = a0 a1 b0 b1 c0 c1 d0 d1 (word s and s+1 of the four sample rows); returns false if no cell is active.
 * Ownership: SURVEY.md 3.1-9 (every cell creates e5, e6, e10; cells on the low faces also the edges in those faces). */
ISOMC_HD bool classify_segment(const Geo &g, const uint32_t w[8], uint32_t s, uint32_t y, uint32_t lz, SegClass &o) {
    const uint32_t a0 = w[0], a1 = w[1], b0 = w[2], b1 = w[3], c0 = w[4], c1 = w[5], d0 = w[6], d1 = w[7];
    o.act = 0; o.p0 = o.p1 = o.p2 = o.p3 = 0;
    o.a0 = a0; o.b0 = b0; o.c0 = c0; o.d0 = d0;
    o.nb = (a1 & 1u) | (b1 & 1u) << 1 | (c1 & 1u) << 2 | (d1 & 1u) << 3;
    const uint32_t all_or = a0 | b0 | c0 | d0 | ((a1 | b1 | c1 | d1) & 1u);
    const uint32_t all_and = a0 & b0 & c0 & d0;
    if ((all_or == 0u) || (all_and == 0xFFFFFFFFu && (a1 & b1 & c1 & d1 & 1u))) return false;
    const uint32_t an = hd_funnel_r(a0, a1, 1), bn = hd_funnel_r(b0, b1, 1);
    const uint32_t cn = hd_funnel_r(c0, c1, 1), dn = hd_funnel_r(d0, d1, 1);
    const uint32_t ncell = g.ncx - s * 32;
    const uint32_t vm = ncell >= 32 ? 0xFFFFFFFFu : ((1u << ncell) - 1u);
    const uint32_t all_in = a0 & an & b0 & bn & c0 & cn & d0 & dn;
    const uint32_t any_in = a0 | an | b0 | bn | c0 | cn | d0 | dn;
    o.act = any_in & ~all_in & vm;
    if (o.act == 0) return false;
    const bool Z0 = geo_z(g, lz) == 0, Y0 = y == 0;
    const uint32_t x0m = s == 0 ? 1u : 0u;
    uint32_t p0 = 0, p1 = 0, p2 = 0, p3 = 0;
    hd_bs_add(p0, p1, p2, p3, (cn ^ dn) & vm);           /* e5: corners 5-6 */
    hd_bs_add(p0, p1, p2, p3, (dn ^ d0) & vm);           /* e6: corners 6-7 */
    hd_bs_add(p0, p1, p2, p3, (bn ^ dn) & vm);           /* e10: corners 2-6 */
    if (Z0) {
        hd_bs_add(p0, p1, p2, p3, (an ^ bn) & vm);       /* e1 */
        hd_bs_add(p0, p1, p2, p3, (bn ^ b0) & vm);       /* e2 */
        hd_bs_add(p0, p1, p2, p3, (b0 ^ a0) & vm & x0m); /* e3 */
        if (Y0) hd_bs_add(p0, p1, p2, p3, (a0 ^ an) & vm); /* e0 */
    }
    if (Y0) {
        hd_bs_add(p0, p1, p2, p3, (c0 ^ cn) & vm);       /* e4 */
        hd_bs_add(p0, p1, p2, p3, (an ^ cn) & vm);       /* e9 */
        hd_bs_add(p0, p1, p2, p3, (a0 ^ c0) & vm & x0m); /* e8 */
    }
    if (x0m) {
        hd_bs_add(p0, p1, p2, p3, (d0 ^ c0) & vm & x0m); /* e7 */
        hd_bs_add(p0, p1, p2, p3, (b0 ^ d0) & vm & x0m); /* e11 */
    }
    o.p0 = p0; o.p1 = p1; o.p2 = p2; o.p3 = p3;
    return true;
}

ISOMC_HD void seg_clear(SegClass &C) {
    C.act = 0; C.p0 = C.p1 = C.p2 = C.p3 = 0; C.a0 = C.b0 = C.c0 = C.d0 = 0; C.nb = 0;
}

ISOMC_HD uint32_t seg_planes_count(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3, uint32_t m) {
    return hd_popc(p0 & m) + 2 * hd_popc(p1 & m) + 4 * hd_popc(p2 & m) + 8 * hd_popc(p3 & m);
}

/* natural cube index of cell i of a segment: bits i, i+1 of the four rows (bit 32 = bit 0 of the next word, in nb) */
ISOMC_HD uint32_t seg_cube_index(uint32_t a0, uint32_t b0, uint32_t c0, uint32_t d0, uint32_t nb, uint32_t i) {
    return (hd_funnel_r(a0, nb, i) & 3u) | (hd_funnel_r(b0, nb >> 1, i) & 3u) << 2 | (hd_funnel_r(c0, nb >> 2, i) & 3u) << 4 |
           (hd_funnel_r(d0, nb >> 3, i) & 3u) << 6;
}

/* nth8[m * 8 + j] = position of the j-th set bit of the byte m (256 x 8 bytes, built by nth8_fill) */
ISOMC_HD void nth8_fill(uint8_t *nth8, uint32_t m) {
    uint32_t j = 0;
    for (uint32_t b = 0; b < 8; ++b)
        if (m >> b & 1u) nth8[m * 8 + j++] = (uint8_t)b;
    for (; j < 8; ++j) nth8[m * 8 + j] = 0;
}

/* position of the j-th (0-based) set bit of m; j < popc(m) */
ISOMC_HD uint32_t nth_set_bit(const uint8_t *nth8, uint32_t m, uint32_t j) {
    const uint32_t c0 = hd_popc(m & 0xFFu), c1 = hd_popc(m & 0xFFFFu), c2 = hd_popc(m & 0xFFFFFFu);
    const uint32_t byte = (j >= c0 ? 1u : 0u) + (j >= c1 ? 1u : 0u) + (j >= c2 ? 1u : 0u);
    const uint32_t before = j >= c2 ? c2 : (j >= c1 ? c1 : (j >= c0 ? c0 : 0u));
    return byte * 8 + nth8[((m >> (8 * byte)) & 0xFFu) * 8 + (j - before)];
}


/* ---- warp-cooperative part of k_count_list ---------------------------------------------------
 * Written against a tiny warp interface so that tests/list_model.cu can run the SAME source on the host
 * (32 threads per emulated warp, a barrier per shuffle; ISOMC_HOST_MODEL).  On the device every call is
 * the intrinsic it names. */
struct Warp {
    uint32_t lane;
    void *emu; /* host model only */
};
#if !defined(__CUDA_ARCH__) && defined(ISOMC_HOST_MODEL)
uint32_t isomc_emu_shfl(void *emu, uint32_t lane, uint32_t v, uint32_t src);
void isomc_emu_sync(void *emu, uint32_t lane);
uint32_t isomc_emu_atomic_add_u32(uint32_t *p, uint32_t v);
uint32_t isomc_emu_next_task(uint32_t *ticket); /* the model deals the tasks out to its emulated warps itself */
void isomc_emu_atomic_add_u64(unsigned long long *p, unsigned long long v);
#endif

ISOMC_HD uint32_t w_shfl(const Warp &w, uint32_t v, uint32_t src) {
#if defined(__CUDA_ARCH__)
    return __shfl_sync(0xFFFFFFFFu, v, src);
#elif defined(ISOMC_HOST_MODEL)
    return isomc_emu_shfl(w.emu, w.lane, v, src & 31u);
#else
    return v;
#endif
}
/* value of lane - d (own value for the first d lanes) */
ISOMC_HD uint32_t w_shfl_up(const Warp &w, uint32_t v, uint32_t d) {
#if defined(__CUDA_ARCH__)
    return __shfl_up_sync(0xFFFFFFFFu, v, d);
#elif defined(ISOMC_HOST_MODEL)
    return isomc_emu_shfl(w.emu, w.lane, v, w.lane >= d ? w.lane - d : w.lane);
#else
    return v;
#endif
}
ISOMC_HD uint32_t w_ballot(const Warp &w, bool pred) {
#if defined(__CUDA_ARCH__)
    return __ballot_sync(0xFFFFFFFFu, pred);
#elif defined(ISOMC_HOST_MODEL)
    uint32_t v = pred ? 1u << w.lane : 0u;
    for (uint32_t d = 1; d < 32; d <<= 1) v |= isomc_emu_shfl(w.emu, w.lane, v, w.lane ^ d);
    return v;
#else
    return pred ? 1u : 0u;
#endif
}
ISOMC_HD uint32_t hd_clz(uint32_t v) { /* v != 0 */
#ifdef __CUDA_ARCH__
    return (uint32_t)__clz((int)v);
#else
    return (uint32_t)__builtin_clz(v);
#endif
}

ISOMC_HD void w_sync(const Warp &w) {
#if defined(__CUDA_ARCH__)
    __syncwarp();
#elif defined(ISOMC_HOST_MODEL)
    isomc_emu_sync(w.emu, w.lane);
#endif
}
ISOMC_HD uint32_t hd_atomic_add(uint32_t *p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#elif defined(ISOMC_HOST_MODEL)
    return isomc_emu_atomic_add_u32(p, v);
#else
    return 0;
#endif
}
ISOMC_HD void hd_atomic_add64(unsigned long long *p, unsigned long long v) {
#if defined(__CUDA_ARCH__)
    atomicAdd(p, v);
#elif defined(ISOMC_HOST_MODEL)
    isomc_emu_atomic_add_u64(p, v);
#endif
}
ISOMC_HD uint32_t hd_ldg(const uint32_t *p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

/* exclusive scan over the warp; total = sum over all lanes */
ISOMC_HD uint32_t w_excl_scan(const Warp &w, uint32_t v, uint32_t &total) {
    uint32_t inc = v;
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1) {
        const uint32_t o = w_shfl_up(w, inc, d);
        if (w.lane >= d) inc += o;
    }
    total = w_shfl(w, inc, 31);
    return inc - v;
}

/* warp-private cursor into the list: [pos, end) is what is left of the blocks [b0, b0 + m) this warp holds */
struct ListCursor {
    uint32_t pos, end, b0, m;
};

ISOMC_HD void list_close(const ListBufs &L, ListCursor &c, uint32_t lane) {
    if (c.m && c.b0 + c.m <= L.cap_blocks && lane < c.m) {
        const uint32_t b = c.b0 + lane, lo = b * LIST_BLOCK;
        L.blkfill[b] = c.pos <= lo ? 0u : (c.pos - lo >= LIST_BLOCK ? LIST_BLOCK : c.pos - lo);
    }
    c.m = 0;
}

/* List space for the n (1..1024) cells of a flush window.  The cells of one SEGMENT must be contiguous, the window need
 * not be: the leading segments that still fit the warp's current block (cend <= room; cend = running cell count, one per
 * lane, non-decreasing) stay there, the rest goes to freshly handed-out blocks.  Cell k of the window lives at
 * (k < n_old ? pos_old + k : pos_new + k - n_old); ok = false after a list overflow (nothing may be written then). */
struct ListSpan {
    uint32_t n_old, pos_old, pos_new;
};
ISOMC_HD ListSpan list_alloc(const Warp &w, const ListBufs &L, ListCursor &c, uint32_t n, uint32_t cend, bool &ok) {
    ListSpan sp;
    const uint32_t room = c.end - c.pos;
    const uint32_t fit = hd_popc(w_ballot(w, cend <= room)); /* leading lanes whose cells all fit */
    sp.n_old = fit ? w_shfl(w, cend, fit - 1) : 0u;
    sp.pos_old = c.pos;
    sp.pos_new = 0;
    bool ok_old = c.m == 0 || c.b0 + c.m <= L.cap_blocks;
    c.pos += sp.n_old;
    ok = ok_old;
    if (sp.n_old < n) {
        list_close(L, c, w.lane);
        const uint32_t need = n - sp.n_old, m = (need + LIST_BLOCK - 1) / LIST_BLOCK;
        uint32_t b0 = 0;
        if (w.lane == 0) b0 = hd_atomic_add(L.ctr, m);
        b0 = w_shfl(w, b0, 0);
        c.b0 = b0; c.m = m;
        c.pos = b0 * LIST_BLOCK;
        c.end = (b0 + m) * LIST_BLOCK;
        sp.pos_new = c.pos;
        c.pos += need;
        ok = ok_old && b0 + m <= L.cap_blocks;
    }
    return sp;
}

struct CountOut {
    uint32_t *rowV, *rowT, *rowA;     /* per cell row: vertices created, triangles, active cells */
    unsigned long long *layerTot;     /* per cell layer: the same three, summed */
};

/*
 * k_count_list, one warp.  Two stages with a queue in between, so that both run with full warps whatever the
 * density of the field:
 *
 *   scan     lane per 32-cell segment, a pass = 32 segments in (row, x) order: load the 8 sign words, keep the
 *            segments that are not uniform (they are the ones with active cells) in a ring of SEGQ_CAP raw segments.
 *            Rows without any are finished on the spot (three zeros).
 *   flush    whenever 32 segments wait (and at the end): lane per queued segment: classify (active mask, bit-sliced
 *            "vertices created" planes), list space for the window's active cells; then prefixes within
 *            the cell row of vertices and active cells by a segmented scan (a row's segments are consecutive in the
 *            queue; the row still open at the end of a window is carried in registers); then lane per CELL (list_cells):
 *            cube index, triangles -> list entries; then back to lane per segment: the same scan for the triangles ->
 *            segment records, row totals.
 */
constexpr uint32_t SEGQ_CAP = 128; /* up to 31 waiting + 64 from one pass */
constexpr uint32_t SEGQ_FIRST = 1u << 16, SEGQ_LAST = 1u << 17; /* meta = s | flags: first / last queued segment of its row */

struct SegQueue {             /* one per warp, shared memory */
    uint32_t w[8][SEGQ_CAP];  /* a0 a1 b0 b1 c0 c1 d0 d1 */
    uint32_t row[SEGQ_CAP];
    uint32_t meta[SEGQ_CAP];
    uint32_t segT[32];        /* triangles of each segment of the window being flushed */
};

struct CountState {           /* warp-uniform */
    uint32_t enq, deq;        /* segments ever queued / flushed */
    ListCursor cur;
    uint32_t p_va, p_t;       /* the row left open by the last window: vertices | active cells << 16, triangles so far */
};

/* lane per cell over the n_cells active cells of a flush window; lane l holds segment l (C, cpos/cend = range of its
 * cells within the window, yznb = y | layer << 13 | next-word bits << 26, s = segment index in its row) */
ISOMC_HD void list_cells(const Warp &w, const ListBufs &L, const uint8_t *s_ntri, const uint8_t *nth8, const SegClass &C,
                         uint32_t cpos, uint32_t cend, uint32_t vpre, uint32_t yznb, uint32_t s, uint32_t n_cells, ListSpan sp,
                         bool ok, uint32_t *segT) {
    const uint32_t lane = w.lane;
    uint32_t carry_t = 0; /* triangles so far of the segment that straddles the 32-cell step boundary */
    for (uint32_t kb = 0; kb < n_cells; kb += 32) {
        const uint32_t k = kb + lane;
        const bool live = k < n_cells;
        /* segment of cell k: the first lane whose range ends beyond k (cend is non-decreasing) */
        uint32_t seg = 0;
#pragma unroll
        for (uint32_t step = 16; step; step >>= 1) {
            const uint32_t t = w_shfl(w, cend, seg + step - 1);
            if (t <= k) seg += step;
        }
        const uint32_t sc = w_shfl(w, cpos, seg), ec = w_shfl(w, cend, seg), am = w_shfl(w, C.act, seg);
        const uint32_t a0 = w_shfl(w, C.a0, seg), b0 = w_shfl(w, C.b0, seg);
        const uint32_t c0 = w_shfl(w, C.c0, seg), d0 = w_shfl(w, C.d0, seg);
        const uint32_t p0 = w_shfl(w, C.p0, seg), p1 = w_shfl(w, C.p1, seg);
        const uint32_t p2 = w_shfl(w, C.p2, seg), p3 = w_shfl(w, C.p3, seg);
        const uint32_t yz = w_shfl(w, yznb, seg), sk = w_shfl(w, s, seg), vp = w_shfl(w, vpre, seg);
        const uint32_t i = live ? nth_set_bit(nth8, am, k - sc) : 0u;
        const uint32_t ci = seg_cube_index(a0, b0, c0, d0, yz >> 26, i);
        const uint32_t nt = live ? (uint32_t)s_ntri[ci] : 0u;
        /* inclusive scan of nt within the segment: lane - d belongs to my segment iff d <= (cells of it before me in this step) */
        const uint32_t dist = live ? (k - sc < lane ? k - sc : lane) : 0u;
        uint32_t incl = nt;
#pragma unroll
        for (uint32_t d = 1; d < 32; d <<= 1) {
            const uint32_t o = w_shfl_up(w, incl, d);
            if (d <= dist) incl += o;
        }
        if (live && sc < kb) incl += carry_t; /* my segment began in the previous step */
        carry_t = w_shfl(w, incl, 31);        /* (only used if lane 31's segment continues) */
        if (live) {
            if (k == ec - 1) segT[seg] = incl; /* last cell of the segment */
            if (ok) {
                const uint32_t pos = k < sp.n_old ? sp.pos_old + k : sp.pos_new + (k - sp.n_old);
                L.ent[pos] = make_uint2((vp + seg_planes_count(p0, p1, p2, p3, (1u << i) - 1u)) | (incl - nt) << 16, (sk * 32 + i) | ci << 16);
                L.ent_yz[pos] = (yz & 0x1FFFu) | (yz >> 13 & 0x1FFFu) << 16;
            }
        }
    }
}

/* flush the n (1..32) oldest queued segments */
ISOMC_HD void count_flush(const Warp &w, const Geo &g, const uint8_t *s_ntri, const uint8_t *nth8, const ListBufs &L,
                          const CountOut &out, SegQueue &Q, CountState &S, uint32_t n) {
    const uint32_t lane = w.lane;
    const bool valid = lane < n;
    const uint32_t slot = (S.deq + lane) & (SEGQ_CAP - 1);
    uint32_t row = 0, meta = 0, lz = 0, y = 0, nv = 0;
    SegClass C;
    seg_clear(C);
    if (valid) {
        uint32_t wd[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) wd[j] = Q.w[j][slot];
        row = Q.row[slot];
        meta = Q.meta[slot];
        lz = (uint32_t)(((uint64_t)row * g.row_magic) >> 40); /* row / ncx */
        y = row - lz * g.ncx;
        if (classify_segment(g, wd, meta & 0xFFFFu, y, lz, C)) nv = seg_planes_count(C.p0, C.p1, C.p2, C.p3, 0xFFFFFFFFu);
    }
    const uint32_t na = hd_popc(C.act);
    uint32_t n_cells;
    const uint32_t cpos = w_excl_scan(w, na, n_cells);
    /* prefixes within the cell row: segmented inclusive scans over the window, heads = first queued segment of a row.
     * Vertices and active cells now, triangles after the cell pass. */
    const uint32_t head = (valid && (meta & SEGQ_FIRST)) ? 1u : 0u;
    uint32_t v1 = nv | na << 16, f = head;
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1) {
        const uint32_t o1 = w_shfl_up(w, v1, d), of = w_shfl_up(w, f, d);
        if (lane >= d) {
            if (!f) v1 += o1;
            f |= of;
        }
    }
    const bool cont = f == 0; /* no head at or before me: my row was left open by the previous window */
    if (cont) v1 += S.p_va;
    Q.segT[lane] = 0;
    w_sync(w);
    bool ok = true;
    ListSpan sp;
    sp.n_old = sp.pos_old = sp.pos_new = 0;
    if (n_cells) {
        sp = list_alloc(w, L, S.cur, n_cells, cpos + na, ok);
        list_cells(w, L, s_ntri, nth8, C, cpos, cpos + na, (v1 & 0xFFFFu) - nv, y | lz << 13 | C.nb << 26, meta & 0xFFFFu, n_cells, sp,
                   ok, Q.segT);
    }
    w_sync(w);
    const uint32_t nt = Q.segT[lane];
    uint32_t v2 = nt;
    f = head;
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1) {
        const uint32_t o2 = w_shfl_up(w, v2, d), of = w_shfl_up(w, f, d);
        if (lane >= d) {
            if (!f) v2 += o2;
            f |= of;
        }
    }
    if (cont) v2 += S.p_t;
    if (valid && na && ok) {
        const uint64_t q = (uint64_t)row * g.nsegx + (meta & 0xFFFFu);
        L.segrec[q] = make_uint2(cpos < sp.n_old ? sp.pos_old + cpos : sp.pos_new + (cpos - sp.n_old), C.act);
        L.segtpre[q] = v2 - nt;
    }
    if (valid && (meta & SEGQ_LAST)) { /* the row is complete */
        const uint32_t tv = v1 & 0xFFFFu, ta = v1 >> 16;
        out.rowV[row] = tv; out.rowT[row] = v2; out.rowA[row] = ta;
        if (tv) hd_atomic_add64(&out.layerTot[3 * lz + 0], (unsigned long long)tv);
        if (v2) hd_atomic_add64(&out.layerTot[3 * lz + 1], (unsigned long long)v2);
        if (ta) hd_atomic_add64(&out.layerTot[3 * lz + 2], (unsigned long long)ta);
    }
    const uint32_t lmeta = w_shfl(w, meta, n - 1), lv1 = w_shfl(w, v1, n - 1), lv2 = w_shfl(w, v2, n - 1);
    if (lmeta & SEGQ_LAST) { S.p_va = 0; S.p_t = 0; } else { S.p_va = lv1; S.p_t = lv2; }
    S.deq += n;
}

ISOMC_HD bool seg_uniform(const uint32_t wd[8]) { /* all 4 x 33 samples on one side: no active cell */
    const uint32_t all_or = wd[0] | wd[2] | wd[4] | wd[6] | ((wd[1] | wd[3] | wd[5] | wd[7]) & 1u);
    const uint32_t all_and = wd[0] & wd[2] & wd[4] & wd[6];
    return (all_or == 0u) || (all_and == 0xFFFFFFFFu && (wd[1] & wd[3] & wd[5] & wd[7] & 1u));
}

/* the sign words of segments 2j and 2j + 1 of a cell row: one 8-byte and one 4-byte load per sample row (nws is even, so
 * word 2j of every row is 8-byte aligned); wa = words of segment 2j, wb = of segment 2j + 1 (valid_b: it exists) */
ISOMC_HD void load_pair(const Geo &g, const uint32_t *signs, uint32_t row, uint32_t lz, uint32_t j, bool valid_b, uint32_t wa[8],
                        uint32_t wb[8]) {
    const uint32_t *r = signs + (uint64_t)(row + lz) * g.nws + 2 * j; /* sample row lz*N + y = row + lz */
    const uint64_t dz = (uint64_t)g.N * g.nws;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t *q = r + (k & 1 ? g.nws : 0u) + (k & 2 ? dz : 0u);
#if defined(__CUDA_ARCH__)
        const uint2 p = __ldg(reinterpret_cast<const uint2 *>(q));
#else
        const uint2 p = *reinterpret_cast<const uint2 *>(q);
#endif
        const uint32_t third = valid_b ? hd_ldg(q + 2) : 0u;
        wa[2 * k] = p.x; wa[2 * k + 1] = p.y;
        wb[2 * k] = p.y; wb[2 * k + 1] = third;
    }
}

/* The scan stage's test on the raw loads: px / py / pt = words 2j, 2j + 1, 2j + 2 of the four sample rows.  A segment is uniform
 * (no active cell) iff its 4 x 33 samples are all outside or all inside.  Same answers as seg_uniform() on the assembled words;
 * this is the form the hot loop uses (most passes of a sparse field keep nothing and end here). */
ISOMC_HD void pair_nonuniform(const uint32_t px[4], const uint32_t py[4], const uint32_t pt[4], bool valid_b, bool &ka, bool &kb) {
    const uint32_t ox = px[0] | px[1] | px[2] | px[3], ax = px[0] & px[1] & px[2] & px[3];
    const uint32_t oy = py[0] | py[1] | py[2] | py[3], ay = py[0] & py[1] & py[2] & py[3];
    const uint32_t ot = (pt[0] | pt[1] | pt[2] | pt[3]) & 1u, at = pt[0] & pt[1] & pt[2] & pt[3] & 1u;
    const bool ua = ((ox | (oy & 1u)) == 0u) || (ax == 0xFFFFFFFFu && (ay & 1u) != 0u);
    const bool ub = ((oy | ot) == 0u) || (ay == 0xFFFFFFFFu && at != 0u);
    ka = !ua;
    kb = valid_b && !ub;
}

/* the same loads as load_pair(), left as loaded */
ISOMC_HD void load_pair_raw(const Geo &g, const uint32_t *signs, uint32_t row, uint32_t lz, uint32_t j, bool valid_b, uint32_t px[4],
                            uint32_t py[4], uint32_t pt[4]) {
    const uint32_t *r = signs + (uint64_t)(row + lz) * g.nws + 2 * j; /* sample row lz*N + y = row + lz */
    const uint64_t dz = (uint64_t)g.N * g.nws;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t *q = r + (k & 1 ? g.nws : 0u) + (k & 2 ? dz : 0u);
#if defined(__CUDA_ARCH__)
        const uint2 p = __ldg(reinterpret_cast<const uint2 *>(q));
#else
        const uint2 p = *reinterpret_cast<const uint2 *>(q);
#endif
        px[k] = p.x; py[k] = p.y;
        pt[k] = valid_b ? hd_ldg(q + 2) : 0u;
    }
}

/* queue the segments a pass keeps: lane order, and within a lane segment a before segment b */
ISOMC_HD void pair_enqueue(const Warp &w, SegQueue &Q, CountState &S, uint32_t mask_a, uint32_t mask_b, bool keep_a, bool keep_b,
                           const uint32_t wa[8], const uint32_t wb[8], uint32_t row, uint32_t meta_a, uint32_t meta_b) {
    const uint32_t below = (1u << w.lane) - 1u;
    const uint32_t pos = S.enq + hd_popc(mask_a & below) + hd_popc(mask_b & below);
    if (keep_a) {
        const uint32_t slot = pos & (SEGQ_CAP - 1);
#pragma unroll
        for (int k = 0; k < 8; ++k) Q.w[k][slot] = wa[k];
        Q.row[slot] = row;
        Q.meta[slot] = meta_a;
    }
    if (keep_b) {
        const uint32_t slot = (pos + (keep_a ? 1u : 0u)) & (SEGQ_CAP - 1);
#pragma unroll
        for (int k = 0; k < 8; ++k) Q.w[k][slot] = wb[k];
        Q.row[slot] = row;
        Q.meta[slot] = meta_b;
    }
    S.enq += hd_popc(mask_a) + hd_popc(mask_b);
    w_sync(w);
}

/* consecutive passes a warp takes per ticket: neighbouring rows end up next to each other in the list, so that the
 * emission's look-ups of the -y neighbours (and the scan's own sign-word loads) hit lines that are already close */
#ifndef ISOMC_COUNT_LONG_TASK_AT
#define ISOMC_COUNT_LONG_TASK_AT (1u << 20) /* (the host model is also built with a small value to cover the long-task branch) */
#endif
/* Passes per ticket, from what was measured on a B200 (profiles/r02_history.md: count-task sweeps on whole lattices and on the
 * slabs of an 8-way split): a task should cover about 16 consecutive cell rows while a pass holds several rows (512^3: 0.120 ms
 * with 3-4 passes of 4 rows, 0.134 with 8; 1024^3: 0.616 ms with 8 passes of 2 rows, 0.643 with 27), and 64 rows once a row fills
 * the warp (2048^3: 1.37 ms with 64, 1.89 with 27, 2.73 with 4; one rank's eighth of it: 0.246 / 0.307 / 0.427 ms with 64 / 32 / 8)
 * -- but never so many that the warps in flight get fewer than two tasks each. */
ISOMC_HD uint32_t count_task_passes(uint32_t n_passes, uint32_t n_warps, uint32_t rows_per_pass) {
    if (n_passes > ISOMC_COUNT_LONG_TASK_AT) return 64u;
    const uint32_t want = rows_per_pass <= 1u ? 64u : (16u + rows_per_pass - 1u) / rows_per_pass;
    const uint32_t fair = n_passes / (2u * (n_warps ? n_warps : 1u));
    const uint32_t p = want < fair ? want : fair;
    return p < 1u ? 1u : p;
}

ISOMC_HD uint32_t next_task(const Warp &w, uint32_t *ticket) {
    uint32_t t = 0;
#if !defined(__CUDA_ARCH__) && defined(ISOMC_HOST_MODEL)
    if (w.lane == 0) t = isomc_emu_next_task(ticket);
#else
    if (w.lane == 0) t = hd_atomic_add(ticket, 1u);
#endif
    return w_shfl(w, t, 0);
}

/* One warp's share of cell rows [row0, row1): tasks of count_task_passes() consecutive passes, handed out by a ticket
 * counter.  Every lane scans TWO neighbouring segments.
 * WIDE = rows of more than 64 segments (a pass is a 64-segment chunk of one row); else a pass covers 32 >> gshift whole
 * rows of 1 << gshift lanes (segment pairs) each. */
template <bool WIDE>
ISOMC_HD void count_list_warp(const Warp &w, const Geo &g, const uint32_t *signs, const uint8_t *s_ntri, const uint8_t *nth8,
                              const ListBufs &L, const CountOut &out, uint32_t gshift, uint32_t row0, uint32_t row1, uint32_t *ticket,
                              SegQueue &Q, uint32_t n_warps, uint32_t task_passes = 0 /* 0: count_task_passes() */) {
    const uint32_t lane = w.lane;
    CountState S;
    S.enq = S.deq = 0;
    S.cur.pos = S.cur.end = S.cur.b0 = S.cur.m = 0;
    S.p_va = S.p_t = 0;
    if (!WIDE) {
        const uint32_t G = 1u << gshift, rpw = 32u >> gshift, sub = lane >> gshift, j = lane & (G - 1);
        const uint32_t gmask = (G >= 32 ? 0xFFFFFFFFu : ((1u << G) - 1u)) << (sub << gshift);
        const bool va_lane = 2 * j < g.nsegx, vb_lane = 2 * j + 1 < g.nsegx;
        const uint32_t niter = (row1 - row0 + rpw - 1) / rpw, P = task_passes ? task_passes : count_task_passes(niter, n_warps, rpw);
        for (uint32_t task = next_task(w, ticket); task * P < niter; task = next_task(w, ticket))
        for (uint32_t it = task * P; it < niter && it < (task + 1) * P; ++it) {
            /* (rows without active cells are not written at all: rowV / rowT / rowA are zeroed before the launch) */
            const uint32_t row = row0 + it * rpw + sub;
            uint32_t px[4], py[4], pt[4];
            bool ka = false, kb = false;
            if (row < row1 && va_lane) {
                const uint32_t lz = (uint32_t)(((uint64_t)row * g.row_magic) >> 40);
                load_pair_raw(g, signs, row, lz, j, vb_lane, px, py, pt);
                pair_nonuniform(px, py, pt, vb_lane, ka, kb);
                if (g.zper && geo_dead(g, lz)) ka = kb = false; /* (batched chunks: the layer between two lattices has no cells) */
            }
            const uint32_t ma = w_ballot(w, ka), mb = w_ballot(w, kb);
            if ((ma | mb) == 0) continue;
            const uint32_t rany = (ma | mb) & gmask;
            uint32_t wa[8], wb[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) { wa[2 * k] = px[k]; wa[2 * k + 1] = py[k]; wb[2 * k] = py[k]; wb[2 * k + 1] = pt[k]; }
            uint32_t meta_a = 2 * j, meta_b = 2 * j + 1;
            if (ka || kb) {
                if (hd_ffs0(rany) == lane) { if (ka) meta_a |= SEGQ_FIRST; else meta_b |= SEGQ_FIRST; }
                if (31u - hd_clz(rany) == lane) { if (kb) meta_b |= SEGQ_LAST; else meta_a |= SEGQ_LAST; }
            }
            pair_enqueue(w, Q, S, ma, mb, ka, kb, wa, wb, row, meta_a, meta_b);
            while (S.enq - S.deq >= 32) count_flush(w, g, s_ntri, nth8, L, out, Q, S, 32);
        }
    } else {
        const uint32_t nrow = row1 - row0, P = task_passes ? task_passes : count_task_passes(nrow, n_warps, 1u);
        for (uint32_t task = next_task(w, ticket); task * P < nrow; task = next_task(w, ticket))
        for (uint32_t row = row0 + task * P; row < row1 && row < row0 + (task + 1) * P; ++row) {
            const uint32_t lz = (uint32_t)(((uint64_t)row * g.row_magic) >> 40);
            const bool dead = geo_dead(g, lz);
            bool row_has = false;
            uint32_t last_seq = 0;
            for (uint32_t s0 = 0; s0 < g.nsegx; s0 += 64) {
                const uint32_t j = (s0 >> 1) + lane;
                const bool va = 2 * j < g.nsegx, vb = 2 * j + 1 < g.nsegx;
                uint32_t px[4], py[4], pt[4];
                bool ka = false, kb = false;
                if (va && !dead) {
                    load_pair_raw(g, signs, row, lz, j, vb, px, py, pt);
                    pair_nonuniform(px, py, pt, vb, ka, kb);
                }
                const uint32_t ma = w_ballot(w, ka), mb = w_ballot(w, kb);
                if ((ma | mb) == 0) continue;
                uint32_t wa[8], wb[8];
#pragma unroll
                for (int k = 0; k < 4; ++k) { wa[2 * k] = px[k]; wa[2 * k + 1] = py[k]; wb[2 * k] = py[k]; wb[2 * k + 1] = pt[k]; }
                uint32_t meta_a = 2 * j, meta_b = 2 * j + 1;
                if (!row_has && (ka || kb) && hd_ffs0(ma | mb) == lane) { if (ka) meta_a |= SEGQ_FIRST; else meta_b |= SEGQ_FIRST; }
                pair_enqueue(w, Q, S, ma, mb, ka, kb, wa, wb, row, meta_a, meta_b);
                row_has = true;
                last_seq = S.enq - 1;
                while (S.enq - S.deq >= 32) count_flush(w, g, s_ntri, nth8, L, out, Q, S, 32);
            }
            /* end of the row: close it (a row without active cells keeps the zeros it was given before the launch) */
            if (!row_has) {
            } else if ((int32_t)(last_seq - S.deq) >= 0) { /* its last segment still waits: mark it */
                if (lane == 0) Q.meta[last_seq & (SEGQ_CAP - 1)] |= SEGQ_LAST;
                w_sync(w);
            } else { /* all of it is flushed: it is the row the last window left open */
                if (lane == 0) {
                    const uint32_t tv = S.p_va & 0xFFFFu, ta = S.p_va >> 16;
                    out.rowV[row] = tv; out.rowT[row] = S.p_t; out.rowA[row] = ta;
                    if (tv) hd_atomic_add64(&out.layerTot[3 * lz + 0], (unsigned long long)tv);
                    if (S.p_t) hd_atomic_add64(&out.layerTot[3 * lz + 1], (unsigned long long)S.p_t);
                    if (ta) hd_atomic_add64(&out.layerTot[3 * lz + 2], (unsigned long long)ta);
                }
                S.p_va = 0; S.p_t = 0;
            }
        }
    }
    while (S.enq != S.deq) {
        const uint32_t n = S.enq - S.deq;
        count_flush(w, g, s_ntri, nth8, L, out, Q, S, n < 32 ? n : 32);
    }
    list_close(L, S.cur, lane);
}

#endif /* ISOMC_CELL_CUH */
