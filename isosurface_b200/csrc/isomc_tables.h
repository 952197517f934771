/*
 * isomc_tables.h -- device lookup tables, derived on the host from the 256-case triangle table.
 *
 * Everything here is indexed by the *natural* cube index ci' whose bit n = dx | dy<<1 | dz<<2
 * is the inside flag of the corner at offset (dx,dy,dz).  The reference numbers corners in the
 * order of its CORNERS table (src/marching_cubes_tables.rs:16-25); corner i of that order maps
 * to natural bit nat_of_ref[i] = {0,1,3,2,4,5,7,6}.  Edge numbers 0..11 are the reference's
 * (EDGE_CONNECTION, marching_cubes_tables.rs:32-45).
 *
 * Derived tables:
 *   ntri[ci']        triangles of the case (march_cube loop bound, marching_cubes_impl.rs:106-109)
 *   emask[ci']       12-bit set of crossed edges (== EDGE_CROSSING_MASK, tables.rs:49-70)
 *   tri[ci']         the case's edges, 4 bits each, 3 per triangle, table order
 *   order[ci']       crossed edges in order of first appearance in tri[ci'] -- the order in which
 *                    MeshTopologyBuilder::add_vertex (mesh.rs:240-251) first sees them
 *   before[ci'][e]   set of edges that appear before e in order[ci']
 *   ownmask[b]       edges a cell *creates* given its boundary flags b = (x==0)|(y==0)<<1|(z==0)<<2:
 *                    those no earlier cell in (z,y,x) order contains
 *   owner[b][e]      for a cell with flags b, which earlier cell created edge e and under which
 *                    local edge number: dx | dy<<1 | dz<<2 | e'<<4
 *   rank3[ci']       for interior cells (which create exactly e5, e6, e10): rank of e5 | e6<<2 | e10<<4
 *                    among those three in order[ci']
 *   ends[e]          corner offsets of the edge's two ends in EDGE_CONNECTION direction:
 *                    (ux|uy<<1|uz<<2) | (vx|vy<<1|vz<<2)<<4
 */
#ifndef ISOMC_TABLES_H
#define ISOMC_TABLES_H

#include <stdint.h>
#include <string.h>

#include "isomc_case_table.h"

struct McTables {
    uint64_t tri[256];
    uint64_t order[256];
    uint16_t before[256][12];
    uint16_t emask[256];
    uint8_t ntri[256];
    uint16_t ownmask[8];
    uint8_t owner[8][12];
    uint8_t ends[12];
    uint8_t ref_of_nat[256]; /* natural cube index -> reference cube index (debug/parity) */
    uint8_t rank3[256];
    uint8_t pad[4];
};

static inline int isomc_build_tables(McTables *t) {
    static const int corner_off[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0},
                                         {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
    static const int edge_ends[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6},
                                         {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
    memset(t, 0, sizeof *t);
    int nat_of_ref[8];
    for (int i = 0; i < 8; ++i) nat_of_ref[i] = corner_off[i][0] | corner_off[i][1] << 1 | corner_off[i][2] << 2;

    /* edge geometry: axis and the two perpendicular local coordinates */
    int e_axis[12], e_perp[12][2]; /* perp coords in increasing axis order */
    int edge_by_geo[3][2][2];
    for (int e = 0; e < 12; ++e) {
        const int *u = corner_off[edge_ends[e][0]], *v = corner_off[edge_ends[e][1]];
        int axis = (u[0] != v[0]) ? 0 : (u[1] != v[1]) ? 1 : 2;
        int k = 0;
        for (int a = 0; a < 3; ++a)
            if (a != axis) e_perp[e][k++] = u[a];
        e_axis[e] = axis;
        edge_by_geo[axis][e_perp[e][0]][e_perp[e][1]] = e;
        int un = u[0] | u[1] << 1 | u[2] << 2, vn = v[0] | v[1] << 1 | v[2] << 2;
        t->ends[e] = (uint8_t)(un | vn << 4);
    }
    for (int b = 0; b < 8; ++b) {
        uint16_t own = 0;
        for (int e = 0; e < 12; ++e) {
            int axis = e_axis[e], step[3] = {0, 0, 0}, np[2], k = 0;
            for (int a = 0; a < 3; ++a) {
                if (a == axis) continue;
                int on_low_boundary = (b >> a) & 1;
                int l = e_perp[e][k];
                step[a] = (l == 0 && !on_low_boundary);
                np[k] = step[a] ? 1 : l;
                ++k;
            }
            if (!step[0] && !step[1] && !step[2]) own |= (uint16_t)(1u << e);
            int e2 = edge_by_geo[axis][np[0]][np[1]];
            t->owner[b][e] = (uint8_t)(step[0] | step[1] << 1 | step[2] << 2 | e2 << 4);
        }
        t->ownmask[b] = own;
    }

    for (int cref = 0; cref < 256; ++cref) {
        int cnat = 0;
        for (int i = 0; i < 8; ++i)
            if (cref >> i & 1) cnat |= 1 << nat_of_ref[i];
        t->ref_of_nat[cnat] = (uint8_t)cref;
        const char *s = ISOMC_TRI_HEX[cref];
        int n = (int)strlen(s);
        if (n % 3 != 0 || n > 15) return -1;
        uint64_t tri = 0, order = 0;
        uint16_t seen = 0;
        int nseen = 0;
        for (int k = 0; k < n; ++k) {
            char c = s[k];
            int e = (c >= '0' && c <= '9') ? c - '0' : c - 'a' + 10;
            if (e < 0 || e > 11) return -1;
            tri |= (uint64_t)e << (4 * k);
            if (!(seen >> e & 1)) {
                t->before[cnat][e] = seen;
                order |= (uint64_t)e << (4 * nseen++);
                seen |= (uint16_t)(1u << e);
            }
        }
        tri |= (uint64_t)0xF << (4 * n);
        uint16_t from_signs = 0;
        for (int e = 0; e < 12; ++e)
            if ((cref >> edge_ends[e][0] & 1) != (cref >> edge_ends[e][1] & 1)) from_signs |= (uint16_t)(1u << e);
        if (from_signs != seen) return -1; /* table must use exactly the crossed edges */
        t->tri[cnat] = tri;
        t->order[cnat] = order;
        t->emask[cnat] = seen;
        t->ntri[cnat] = (uint8_t)(n / 3);
        const uint16_t interior = (uint16_t)(1u << 5 | 1u << 6 | 1u << 10);
        t->rank3[cnat] = (uint8_t)(__builtin_popcount(t->before[cnat][5] & interior) |
                                   __builtin_popcount(t->before[cnat][6] & interior) << 2 |
                                   __builtin_popcount(t->before[cnat][10] & interior) << 4);
    }
    return 0;
}
#endif
