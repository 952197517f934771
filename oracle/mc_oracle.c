/*
 * mc_oracle.c -- CPU ORACLE for the MarchingCubes extract path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is a plain-C restatement of the algorithm in swiftcoder/isosurface
 * (reference tree, read-only, paths relative to /root/reference):
 *
 *   src/marching_cubes.rs:59-82            wiring: traverse -> classify -> crossings -> faces
 *   src/traversal/primal_grid.rs:39-89     lattice, two-layer z sweep, the `0..size` z quirk
 *   src/marching_cubes_impl.rs:26-37       classify_corners   (bit i set iff !(v_i > 0))
 *   src/marching_cubes_impl.rs:39-56       find_edge_crossings
 *   src/marching_cubes_impl.rs:102-117     march_cube
 *   src/distance.rs:52-54,64-69            Signed::is_positive / find_crossing_point
 *   src/index_cache.rs:49-60               GridKey (ordered lattice-point pair)
 *   src/mesh.rs:64-100,240-255             add_vertex / add_face / extract_indices
 *   src/math/vector.rs:45-52,125-132       len_sq fold order, componentwise ops
 *   src/implicit/{sphere,torus,cylinder,rectangular_prism,csg}.rs   sample_scalar impls
 *   examples/common/sources.rs:38-43       the (0.5,0.5,0.5) translation
 *
 * It exists so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs can check and time the reference algorithm.  Nothing in the
 * product path (isosurface_b200/, include/) may call into it.
 *
 * PARITY PINNING STATUS: the reference cannot be compiled here (no cargo/rustc), and its
 * own test-suite pins only the SDF sample values (sphere.rs:73-76, torus.rs:118-122,
 * cylinder.rs:96-99, rectangular_prism.rs:90-93, csg.rs:121-124,152-155) -- those known
 * answers are checked in tests/test_oracle.py.  For the extraction stage (classification ->
 * mesh) the reference holds no golden vectors: "parity unpinned" by reference-run data; it
 * is anchored instead on the SURVEY.md 8(c) known answers (an independent numpy restatement
 * by the surveyor: counts + SHA-256 of vertex and index streams) and on mesh invariants.
 *
 * Build:  gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (see oracle/Makefile).
 * All arithmetic on the hot path is IEEE binary32 with no contraction, as rustc emits.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "mc_case_table.h"

/* ------------------------------------------------------------------------------------ */
/* Tables (reference src/marching_cubes_tables.rs:16-25,32-45,49-70,74-331)              */
/* ------------------------------------------------------------------------------------ */

static const int CORNER_OFF[8][3] = { /* tables.rs:16-25 */
    {0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
static const int EDGE_ENDS[12][2] = { /* tables.rs:32-45 */
    {0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};

static int8_t TRI[256][16];
static uint16_t EDGE_MASK[256];
static int tables_ready = 0;

static int hexval(char c) { return (c >= '0' && c <= '9') ? c - '0' : c - 'a' + 10; }

static int tables_init(void) {
    if (tables_ready) return 0;
    for (int c = 0; c < 256; ++c) {
        const char *s = ORACLE_TRI_HEX[c];
        int n = (int)strlen(s);
        uint16_t from_rows = 0;
        for (int k = 0; k < 16; ++k) TRI[c][k] = (k < n) ? (int8_t)hexval(s[k]) : -1;
        for (int k = 0; k < n; ++k) from_rows |= (uint16_t)(1u << TRI[c][k]);
        /* The reference's EDGE_CROSSING_MASK (tables.rs:49-70) equals the set of edges whose
         * two corners differ in sign; re-derive it and insist both derivations agree. */
        uint16_t from_signs = 0;
        for (int e = 0; e < 12; ++e) {
            int u = (c >> EDGE_ENDS[e][0]) & 1, v = (c >> EDGE_ENDS[e][1]) & 1;
            if (u != v) from_signs |= (uint16_t)(1u << e);
        }
        if (from_rows != from_signs) return -1;
        EDGE_MASK[c] = from_signs;
    }
    tables_ready = 1;
    return 0;
}

/* Export the decoded tables so tests can hash them against the SURVEY pins. */
int oracle_tables(int8_t *tri_256x16, uint16_t *edge_mask_256, int32_t *corners_8x3, int32_t *edges_12x2) {
    if (tables_init()) return -1;
    memcpy(tri_256x16, TRI, sizeof TRI);
    memcpy(edge_mask_256, EDGE_MASK, sizeof EDGE_MASK);
    for (int i = 0; i < 8; ++i)
        for (int k = 0; k < 3; ++k) corners_8x3[3 * i + k] = CORNER_OFF[i][k];
    for (int i = 0; i < 12; ++i)
        for (int k = 0; k < 2; ++k) edges_12x2[2 * i + k] = EDGE_ENDS[i][k];
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* Implicit sources (reference src/implicit/ *.rs, scalar side only)                     */
/* ------------------------------------------------------------------------------------ */

typedef struct { float x, y, z; } v3;

/* Same binary layout as isomc_sdf_node in include/isomc.h (kept separate on purpose). */
typedef struct { uint32_t op; float a, b, c; } osdf_node;
enum {
    O_SPHERE = 1,       /* a = radius                         sphere.rs:35-39  */
    O_TORUS = 2,        /* a = radius, b = tube_radius        torus.rs:40-46   */
    O_CYLINDER = 3,     /* a = radius, b = half_length        cylinder.rs:41-48 */
    O_PRISM = 4,        /* a,b,c = half_extent                rectangular_prism.rs:36-41 */
    O_UNION = 16,       /* min(a, b)                          csg.rs:35-39 */
    O_INTERSECTION = 17,/* max(a, b)                          csg.rs:68-72 */
    O_DIFFERENCE = 18,  /* max(b, -a)                         csg.rs:96-100 */
    O_TRANSLATE_PUSH = 32, /* q = p - (a,b,c) for the nodes up to the matching POP
                              (examples/common/sources.rs:38-43) */
    O_TRANSLATE_POP = 33
};
#define OSDF_MAX_STACK 16

static float sdf_eval(const osdf_node *prog, uint32_t n, v3 p, int *err) {
    float vs[OSDF_MAX_STACK];
    v3 ps[OSDF_MAX_STACK];
    int nv = 0, np = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const osdf_node *nd = &prog[i];
        switch (nd->op) {
        case O_SPHERE: {
            if (nv >= OSDF_MAX_STACK) { *err = 1; return 0; }
            float l2 = (p.x * p.x + p.y * p.y) + p.z * p.z; /* vector.rs fold!: ((x*x + y*y) + z*z) */
            vs[nv++] = sqrtf(l2) - nd->a;
        } break;
        case O_TORUS: {
            if (nv >= OSDF_MAX_STACK) { *err = 1; return 0; }
            float qx = fabsf(sqrtf(p.x * p.x + p.y * p.y)) - nd->a;
            float len = sqrtf(qx * qx + p.z * p.z);
            vs[nv++] = len - nd->b;
        } break;
        case O_CYLINDER: {
            if (nv >= OSDF_MAX_STACK) { *err = 1; return 0; }
            float qx = fabsf(sqrtf(p.x * p.x + p.y * p.y)) - nd->a;
            float qz = fabsf(p.z) - nd->b;
            float dx = fmaxf(qx, 0.0f), dy = fmaxf(qz, 0.0f), dz = 0.0f;
            float dl = sqrtf((dx * dx + dy * dy) + dz * dz);
            vs[nv++] = fminf(fmaxf(qx, qz), 0.0f) + dl;
        } break;
        case O_PRISM: {
            if (nv >= OSDF_MAX_STACK) { *err = 1; return 0; }
            float qx = fabsf(p.x) - nd->a, qy = fabsf(p.y) - nd->b, qz = fabsf(p.z) - nd->c;
            float mx = fmaxf(qx, 0.0f), my = fmaxf(qy, 0.0f), mz = fmaxf(qz, 0.0f);
            float len = sqrtf((mx * mx + my * my) + mz * mz);
            float mc = fmaxf(qx, fmaxf(qy, qz)); /* vector.rs:245-247 max_component */
            vs[nv++] = len + fminf(mc, 0.0f);
        } break;
        case O_UNION:
            if (nv < 2) { *err = 1; return 0; }
            vs[nv - 2] = fminf(vs[nv - 2], vs[nv - 1]);
            --nv;
            break;
        case O_INTERSECTION:
            if (nv < 2) { *err = 1; return 0; }
            vs[nv - 2] = fmaxf(vs[nv - 2], vs[nv - 1]);
            --nv;
            break;
        case O_DIFFERENCE: /* Difference{a,b}: max(b, -a) with a pushed first */
            if (nv < 2) { *err = 1; return 0; }
            vs[nv - 2] = fmaxf(vs[nv - 1], -vs[nv - 2]);
            --nv;
            break;
        case O_TRANSLATE_PUSH:
            if (np >= OSDF_MAX_STACK) { *err = 1; return 0; }
            ps[np++] = p;
            p.x = p.x - nd->a; p.y = p.y - nd->b; p.z = p.z - nd->c;
            break;
        case O_TRANSLATE_POP:
            if (np < 1) { *err = 1; return 0; }
            p = ps[--np];
            break;
        default:
            *err = 1;
            return 0;
        }
    }
    if (nv != 1 || np != 0) { *err = 1; return 0; }
    return vs[0];
}

int oracle_sample_sdf(const osdf_node *prog, uint32_t n, const float *xyz, uint64_t npts, float *out) {
    int err = 0;
    for (uint64_t i = 0; i < npts; ++i) {
        v3 p = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        out[i] = sdf_eval(prog, n, p, &err);
        if (err) return -2;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* Sources as seen by the traversal: value at lattice point (x, y, z)                    */
/* ------------------------------------------------------------------------------------ */

typedef struct {
    const osdf_node *prog; uint32_t nprog;  /* procedural */
    const float *grid;                      /* or dense N*N*(zcells+1) f32, x fastest */
    uint32_t size;
    int err;
} source_t;

static inline float source_at(source_t *s, v3 corner, uint32_t x, uint32_t y, uint32_t z) {
    if (s->grid) return s->grid[((uint64_t)z * s->size + y) * s->size + x];
    return sdf_eval(s->prog, s->nprog, corner, &s->err);
}

/* Fill a dense lattice exactly as PrimalGrid::traverse would have sampled it
 * (primal_grid.rs:44-53,61-70): coordinate = (i as f32) * (1.0 / (size-1) as f32). */
int oracle_fill_grid_sdf(uint32_t size, const osdf_node *prog, uint32_t n, uint32_t z_layers, float *grid) {
    if (size < 2) return -1;
    float inv = 1.0f / (float)(size - 1);
    int err = 0;
    for (uint32_t z = 0; z < z_layers; ++z)
        for (uint32_t y = 0; y < size; ++y)
            for (uint32_t x = 0; x < size; ++x) {
                v3 p = {(float)x * inv, (float)y * inv, (float)z * inv};
                grid[((uint64_t)z * size + y) * size + x] = sdf_eval(prog, n, p, &err);
            }
    return err ? -2 : 0;
}

/* ------------------------------------------------------------------------------------ */
/* Hash containers.  "faithful" mode mirrors the reference's data-structure cost:        */
/*   IndexCache = HashMap<GridKey(48 B), VertexHandle>   (index_cache.rs:18,25-46)       */
/*   MeshTopology.edges = HashSet<Edge(16 B)>, edge_to_face = HashMap<Edge, Vec<Face>>   */
/*   (mesh.rs:44-49,71-88) -- maintained but never read on the MC path.                  */
/* Rust's default hasher is SipHash-1-3; restated here from the published algorithm.     */
/* "lean" mode keeps the same dedup semantics with a 64-bit packed key and no edge maps. */
/* ------------------------------------------------------------------------------------ */

#define ROTL(x, b) (uint64_t)(((x) << (b)) | ((x) >> (64 - (b))))
#define SIPROUND do { v0 += v1; v1 = ROTL(v1, 13); v1 ^= v0; v0 = ROTL(v0, 32); v2 += v3; v3 = ROTL(v3, 16); v3 ^= v2; \
    v0 += v3; v3 = ROTL(v3, 21); v3 ^= v0; v2 += v1; v1 = ROTL(v1, 17); v1 ^= v2; v2 = ROTL(v2, 32); } while (0)

static uint64_t siphash13_words(const uint64_t *w, int nwords) {
    const uint64_t k0 = 0x0706050403020100ull, k1 = 0x0f0e0d0c0b0a0908ull;
    uint64_t v0 = 0x736f6d6570736575ull ^ k0, v1 = 0x646f72616e646f6dull ^ k1;
    uint64_t v2 = 0x6c7967656e657261ull ^ k0, v3 = 0x7465646279746573ull ^ k1;
    for (int i = 0; i < nwords; ++i) { v3 ^= w[i]; SIPROUND; v0 ^= w[i]; }
    uint64_t b = (uint64_t)(nwords * 8) << 56;
    v3 ^= b; SIPROUND; v0 ^= b;
    v2 ^= 0xff; SIPROUND; SIPROUND; SIPROUND;
    return v0 ^ v1 ^ v2 ^ v3;
}

typedef struct { uint64_t k[6]; uint64_t val; uint64_t hash; uint8_t used; } gk_slot;
typedef struct { gk_slot *s; uint64_t cap, len; } gk_map;

static int gk_grow(gk_map *m) {
    uint64_t ncap = m->cap ? m->cap * 2 : 16;
    gk_slot *ns = (gk_slot *)calloc(ncap, sizeof(gk_slot));
    if (!ns) return -1;
    for (uint64_t i = 0; i < m->cap; ++i)
        if (m->s[i].used) {
            uint64_t j = m->s[i].hash & (ncap - 1);
            while (ns[j].used) j = (j + 1) & (ncap - 1);
            ns[j] = m->s[i];
        }
    free(m->s);
    m->s = ns; m->cap = ncap;
    return 0;
}
/* returns slot (existing or fresh); *found says which */
static gk_slot *gk_find_or_insert(gk_map *m, const uint64_t key[6], int nwords_hash, int *found) {
    if ((m->len + 1) * 8 > m->cap * 7) if (gk_grow(m)) return NULL;
    uint64_t h = siphash13_words(key, nwords_hash);
    uint64_t j = h & (m->cap - 1);
    while (m->s[j].used) {
        if (m->s[j].hash == h && !memcmp(m->s[j].k, key, 48)) { *found = 1; return &m->s[j]; }
        j = (j + 1) & (m->cap - 1);
    }
    *found = 0;
    m->s[j].used = 1; m->s[j].hash = h; memcpy(m->s[j].k, key, 48);
    m->len++;
    return &m->s[j];
}

/* lean map: 64-bit key -> u32 */
typedef struct { uint64_t *k; uint32_t *v; uint64_t cap, len; } lean_map;
static inline uint64_t mix64(uint64_t z) {
    z ^= z >> 33; z *= 0xff51afd7ed558ccdull; z ^= z >> 33; z *= 0xc4ceb9fe1a85ec53ull; z ^= z >> 33;
    return z;
}
static int lean_grow(lean_map *m) {
    uint64_t ncap = m->cap ? m->cap * 2 : 1024;
    uint64_t *nk = (uint64_t *)malloc(ncap * 8);
    uint32_t *nv = (uint32_t *)malloc(ncap * 4);
    if (!nk || !nv) { free(nk); free(nv); return -1; }
    memset(nk, 0xff, ncap * 8);
    for (uint64_t i = 0; i < m->cap; ++i)
        if (m->k[i] != UINT64_MAX) {
            uint64_t j = mix64(m->k[i]) & (ncap - 1);
            while (nk[j] != UINT64_MAX) j = (j + 1) & (ncap - 1);
            nk[j] = m->k[i]; nv[j] = m->v[i];
        }
    free(m->k); free(m->v);
    m->k = nk; m->v = nv; m->cap = ncap;
    return 0;
}

typedef struct { uint64_t *p; uint32_t len, cap; } fvec; /* Vec<FaceHandle> */

/* ------------------------------------------------------------------------------------ */
/* Output sink == extractor::IndexedVertices (extractor.rs:72-93)                        */
/* ------------------------------------------------------------------------------------ */

typedef struct {
    float *xyz; uint64_t n_vertices, cap_v;
    uint32_t *idx; uint64_t n_triangles, cap_i;   /* idx holds 3*n_triangles entries */
    uint64_t n_active_cells;
} oracle_mesh;

static int push_vertex(oracle_mesh *m, v3 p) {
    if (m->n_vertices == m->cap_v) {
        uint64_t nc = m->cap_v ? m->cap_v * 2 : 4; /* Vec growth: amortised doubling */
        float *nx = (float *)realloc(m->xyz, nc * 12);
        if (!nx) return -1;
        m->xyz = nx; m->cap_v = nc;
    }
    float *d = m->xyz + 3 * m->n_vertices++;
    d[0] = p.x; d[1] = p.y; d[2] = p.z;
    return 0;
}
static int push_indices(oracle_mesh *m, uint64_t a, uint64_t b, uint64_t c) {
    if (m->n_triangles == m->cap_i) {
        uint64_t nc = m->cap_i ? m->cap_i * 2 : 4;
        uint32_t *ni = (uint32_t *)realloc(m->idx, nc * 12);
        if (!ni) return -1;
        m->idx = ni; m->cap_i = nc;
    }
    uint32_t *d = m->idx + 3 * m->n_triangles++;
    d[0] = (uint32_t)a; d[1] = (uint32_t)b; d[2] = (uint32_t)c; /* `index as u32`, extractor.rs:90-92 */
    return 0;
}

void oracle_mesh_free(oracle_mesh *m) {
    free(m->xyz); free(m->idx);
    memset(m, 0, sizeof *m);
}

/* ------------------------------------------------------------------------------------ */
/* The extraction itself                                                                 */
/* ------------------------------------------------------------------------------------ */

typedef struct { v3 corner; float value; } lattice_entry; /* (Vec3, Signed) = 16 B, primal_grid.rs:20 */

enum { MODE_FAITHFUL = 0, MODE_LEAN = 1 };

static int extract_impl(source_t *src, uint32_t z_cells, int mode, oracle_mesh *out, uint8_t *cube_index_out) {
    const uint32_t size = src->size;
    memset(out, 0, sizeof *out);
    if (tables_init()) return -3;
    if (size < 1) return -1;
    int rc = 0;
    lattice_entry *layers[2] = {NULL, NULL};
    gk_map vmap = {0}, edge_set = {0}, edge_to_face = {0};
    fvec *fvecs = NULL; uint64_t n_fvecs = 0, cap_fvecs = 0;
    lean_map lmap = {0};
    /* faces are recorded during traversal and emitted afterwards (mesh.rs:91-100) */
    uint64_t *faces = NULL, n_faces = 0, cap_faces = 0;

    layers[0] = (lattice_entry *)malloc((size_t)size * size * sizeof(lattice_entry));
    layers[1] = (lattice_entry *)malloc((size_t)size * size * sizeof(lattice_entry));
    if (!layers[0] || !layers[1]) { rc = -4; goto done; }

    const uint32_t size_minus_one = size - 1;
    const float one_over_size = 1.0f / (float)size_minus_one;

    for (uint32_t y = 0; y < size; ++y)
        for (uint32_t x = 0; x < size; ++x) {
            v3 c = {(float)x * one_over_size, (float)y * one_over_size, 0.0f};
            lattice_entry e = {c, source_at(src, c, x, y, 0)};
            layers[0][(size_t)y * size + x] = e;
        }

    for (uint32_t z = 0; z < z_cells; ++z) {
        for (uint32_t y = 0; y < size; ++y)
            for (uint32_t x = 0; x < size; ++x) {
                v3 c = {(float)x * one_over_size, (float)y * one_over_size, (float)(z + 1) * one_over_size};
                lattice_entry e = {c, source_at(src, c, x, y, z + 1)};
                layers[1][(size_t)y * size + x] = e;
            }
        if (src->err) { rc = -2; goto done; }

        for (uint32_t y = 0; y < size_minus_one; ++y)
            for (uint32_t x = 0; x < size_minus_one; ++x) {
                uint64_t keys[8][3];
                v3 corners[8];
                float values[8];
                for (int i = 0; i < 8; ++i) {
                    keys[i][0] = x + CORNER_OFF[i][0];
                    keys[i][1] = y + CORNER_OFF[i][1];
                    keys[i][2] = z + CORNER_OFF[i][2];
                    lattice_entry e =
                        layers[CORNER_OFF[i][2]][(size_t)(y + CORNER_OFF[i][1]) * size + x + CORNER_OFF[i][0]];
                    corners[i] = e.corner;
                    values[i] = e.value;
                }
                /* classify_corners */
                unsigned cube_index = 0;
                for (int i = 0; i < 8; ++i)
                    if (!(values[i] > 0.0f)) cube_index |= 1u << i;
                if (cube_index_out)
                    cube_index_out[((uint64_t)z * size_minus_one + y) * size_minus_one + x] = (uint8_t)cube_index;
                if (cube_index != 0 && cube_index != 255) out->n_active_cells++; /* point_cloud.rs:58 */

                /* find_edge_crossings: every masked edge, also ones a previous cell created */
                v3 vertices[12];
                unsigned edges = EDGE_MASK[cube_index];
                for (int i = 0; i < 12; ++i)
                    if (edges & (1u << i)) {
                        int u = EDGE_ENDS[i][0], v = EDGE_ENDS[i][1];
                        float a = values[u], b = values[v];
                        float delta = b - a;
                        float t = (delta == 0.0f) ? 0.5f : -a / delta;
                        float omt = 1.0f - t;
                        v3 pa = corners[u], pb = corners[v], r;
                        r.x = pa.x * omt + pb.x * t;
                        r.y = pa.y * omt + pb.y * t;
                        r.z = pa.z * omt + pb.z * t;
                        vertices[i] = r;
                    }

                /* march_cube */
                for (int i = 0; i < 5; ++i) {
                    if (TRI[cube_index][3 * i] < 0) break;
                    uint64_t h[3];
                    for (int k = 0; k < 3; ++k) {
                        int e = TRI[cube_index][3 * i + k];
                        int u = EDGE_ENDS[e][0], v = EDGE_ENDS[e][1];
                        const uint64_t *a = keys[u], *b = keys[v];
                        /* GridKey::new: tuple compare, smaller lattice point first */
                        int a_gt_b = (a[0] != b[0]) ? (a[0] > b[0]) : (a[1] != b[1]) ? (a[1] > b[1]) : (a[2] > b[2]);
                        if (a_gt_b) { const uint64_t *t = a; a = b; b = t; }
                        if (mode == MODE_FAITHFUL) {
                            uint64_t key[6] = {a[0], a[1], a[2], b[0], b[1], b[2]};
                            int found;
                            gk_slot *s = gk_find_or_insert(&vmap, key, 6, &found);
                            if (!s) { rc = -4; goto done; }
                            if (!found) {
                                s->val = out->n_vertices; /* MeshTopology::add_vertex: running counter */
                                if (push_vertex(out, vertices[e])) { rc = -4; goto done; }
                            }
                            h[k] = s->val;
                        } else {
                            /* lattice point of the smaller end + axis identifies the grid edge */
                            uint64_t axis = (b[0] != a[0]) ? 0 : (b[1] != a[1]) ? 1 : 2;
                            uint64_t key = (a[0] | (a[1] << 20) | (a[2] << 40)) | (axis << 60);
                            if ((lmap.len + 1) * 8 > lmap.cap * 5) if (lean_grow(&lmap)) { rc = -4; goto done; }
                            uint64_t j = mix64(key) & (lmap.cap - 1);
                            while (lmap.k[j] != UINT64_MAX && lmap.k[j] != key) j = (j + 1) & (lmap.cap - 1);
                            if (lmap.k[j] == UINT64_MAX) {
                                lmap.k[j] = key; lmap.v[j] = (uint32_t)out->n_vertices; lmap.len++;
                                if (push_vertex(out, vertices[e])) { rc = -4; goto done; }
                            }
                            h[k] = lmap.v[j];
                        }
                    }
                    /* MeshTopology::add_face */
                    uint64_t face = n_faces;
                    if (n_faces == cap_faces) {
                        uint64_t nc = cap_faces ? cap_faces * 2 : 4;
                        uint64_t *nf = (uint64_t *)realloc(faces, nc * 24);
                        if (!nf) { rc = -4; goto done; }
                        faces = nf; cap_faces = nc;
                    }
                    faces[3 * n_faces] = h[0]; faces[3 * n_faces + 1] = h[1]; faces[3 * n_faces + 2] = h[2];
                    n_faces++;
                    if (mode == MODE_FAITHFUL) {
                        for (int k = 0; k < 3; ++k) {
                            uint64_t p = h[k], q = h[(k + 1) % 3];
                            uint64_t ek[6] = {p < q ? p : q, p < q ? q : p, 0, 0, 0, 0};
                            int found;
                            if (!gk_find_or_insert(&edge_set, ek, 2, &found)) { rc = -4; goto done; }
                            gk_slot *s = gk_find_or_insert(&edge_to_face, ek, 2, &found);
                            if (!s) { rc = -4; goto done; }
                            if (!found) {
                                if (n_fvecs == cap_fvecs) {
                                    uint64_t nc = cap_fvecs ? cap_fvecs * 2 : 1024;
                                    fvec *nf = (fvec *)realloc(fvecs, nc * sizeof(fvec));
                                    if (!nf) { rc = -4; goto done; }
                                    fvecs = nf; cap_fvecs = nc;
                                }
                                fvecs[n_fvecs].p = NULL; fvecs[n_fvecs].len = fvecs[n_fvecs].cap = 0;
                                s->val = n_fvecs++;
                            }
                            fvec *fv = &fvecs[s->val];
                            if (fv->len == fv->cap) {
                                uint32_t nc = fv->cap ? fv->cap * 2 : 4;
                                uint64_t *np_ = (uint64_t *)realloc(fv->p, (size_t)nc * 8);
                                if (!np_) { rc = -4; goto done; }
                                fv->p = np_; fv->cap = nc;
                            }
                            fv->p[fv->len++] = face;
                        }
                    }
                }
            }
        { lattice_entry *t = layers[0]; layers[0] = layers[1]; layers[1] = t; }
    }
    if (src->err) { rc = -2; goto done; }

    /* extract_indices: after the traversal, three per face, face order */
    for (uint64_t f = 0; f < n_faces; ++f)
        if (push_indices(out, faces[3 * f], faces[3 * f + 1], faces[3 * f + 2])) { rc = -4; goto done; }

done:
    free(layers[0]); free(layers[1]);
    free(vmap.s); free(edge_set.s); free(edge_to_face.s);
    for (uint64_t i = 0; i < n_fvecs; ++i) free(fvecs[i].p);
    free(fvecs); free(lmap.k); free(lmap.v); free(faces);
    if (rc) oracle_mesh_free(out);
    return rc;
}

/* MarchingCubes::<Signed>::new(size).extract(&Sampler::new(&implicit_tree), &mut IndexedVertices) */
int oracle_extract_sdf(uint32_t size, const osdf_node *prog, uint32_t n, int mode, oracle_mesh *out) {
    source_t s = {prog, n, NULL, size, 0};
    return extract_impl(&s, size, mode, out, NULL);
}

/* Same with a dense lattice source: grid is size*size*(z_cells+1) f32, x fastest.  z_cells == size
 * is the full extract; a smaller z_cells traverses only the first z_cells cell layers (used to
 * bound CPU-baseline timing runs; the result is the exact prefix of the full mesh's cell order). */
int oracle_extract_grid(uint32_t size, const float *grid, uint32_t z_cells, int mode, oracle_mesh *out) {
    source_t s = {NULL, 0, grid, size, 0};
    return extract_impl(&s, z_cells, mode, out, NULL);
}

/* Per-cell cube_index dump ((size-1)^2 * z_cells bytes, x fastest) for active-cell-set parity. */
int oracle_cube_indices(uint32_t size, const float *grid, uint32_t z_cells, uint8_t *cube_index) {
    source_t s = {NULL, 0, grid, size, 0};
    oracle_mesh tmp;
    int rc = extract_impl(&s, z_cells, MODE_LEAN, &tmp, cube_index);
    if (!rc) oracle_mesh_free(&tmp);
    return rc;
}

/* ------------------------------------------------------------------------------------ */
/* PointCloud::<Signed>::new(size).extract(&source, &mut extractor)                      */
/*   reference src/point_cloud.rs:50-63: the same PrimalGrid traversal; every cell whose */
/*   cube_index is neither 0 nor 255 emits corners[0].lerp(corners[6], 0.5)              */
/*   (src/math/vector.rs:325-333: of = 1.0 - f; of * self.c + f * other.c per component) */
/*   in cell order.  No face data.                                                       */
/* ------------------------------------------------------------------------------------ */
static int point_cloud_impl(source_t *src, uint32_t z_cells, oracle_mesh *out) {
    const uint32_t size = src->size;
    memset(out, 0, sizeof *out);
    if (size < 1) return -1;
    int rc = 0;
    lattice_entry *layers[2];
    layers[0] = (lattice_entry *)malloc((size_t)size * size * sizeof(lattice_entry));
    layers[1] = (lattice_entry *)malloc((size_t)size * size * sizeof(lattice_entry));
    if (!layers[0] || !layers[1]) { rc = -4; goto done; }
    const uint32_t size_minus_one = size - 1;
    const float one_over_size = 1.0f / (float)size_minus_one;
    for (uint32_t y = 0; y < size; ++y)
        for (uint32_t x = 0; x < size; ++x) {
            v3 c = {(float)x * one_over_size, (float)y * one_over_size, 0.0f};
            lattice_entry e = {c, source_at(src, c, x, y, 0)};
            layers[0][(size_t)y * size + x] = e;
        }
    for (uint32_t z = 0; z < z_cells; ++z) {
        for (uint32_t y = 0; y < size; ++y)
            for (uint32_t x = 0; x < size; ++x) {
                v3 c = {(float)x * one_over_size, (float)y * one_over_size, (float)(z + 1) * one_over_size};
                lattice_entry e = {c, source_at(src, c, x, y, z + 1)};
                layers[1][(size_t)y * size + x] = e;
            }
        if (src->err) { rc = -2; goto done; }
        for (uint32_t y = 0; y < size_minus_one; ++y)
            for (uint32_t x = 0; x < size_minus_one; ++x) {
                unsigned cube_index = 0;
                for (int i = 0; i < 8; ++i) {
                    float v = layers[CORNER_OFF[i][2]][(size_t)(y + CORNER_OFF[i][1]) * size + x + CORNER_OFF[i][0]].value;
                    if (!(v > 0.0f)) cube_index |= 1u << i;
                }
                if (cube_index != 0 && cube_index != 255) {
                    v3 a = layers[0][(size_t)y * size + x].corner;             /* corners[0] */
                    v3 b = layers[1][(size_t)(y + 1) * size + x + 1].corner;   /* corners[6] */
                    const float f = 0.5f, of = 1.0f - f;
                    v3 p = {of * a.x + f * b.x, of * a.y + f * b.y, of * a.z + f * b.z};
                    if (push_vertex(out, p)) { rc = -4; goto done; }
                    out->n_active_cells++;
                }
            }
        lattice_entry *t = layers[0]; layers[0] = layers[1]; layers[1] = t;
    }
done:
    free(layers[0]); free(layers[1]);
    if (rc) oracle_mesh_free(out);
    return rc;
}

int oracle_point_cloud_sdf(uint32_t size, const osdf_node *prog, uint32_t n, oracle_mesh *out) {
    source_t s = {prog, n, NULL, size, 0};
    return point_cloud_impl(&s, size, out);
}

int oracle_point_cloud_grid(uint32_t size, const float *grid, uint32_t z_cells, oracle_mesh *out) {
    source_t s = {NULL, 0, grid, size, 0};
    return point_cloud_impl(&s, z_cells, out);
}

/* ------------------------------------------------------------------------------------ */
/* IndexedInterleavedNormals (reference src/extractor.rs:113-122) over a                 */
/* CentralDifference source (src/source.rs:82-94), optionally inside DemoSource-style    */
/* translations (examples/common/sources.rs:55-60: q = p - Vec3::from_scalar(0.5)).      */
/* For each vertex v: out = [v.x v.y v.z n.x n.y n.z].                                    */
/* ------------------------------------------------------------------------------------ */
int oracle_interleaved_normals_cd(const osdf_node *inner, uint32_t n, float epsilon, const float *offsets, uint32_t n_offsets,
                                  const float *xyz, uint64_t n_vertices, float *out) {
    int err = 0;
    for (uint64_t i = 0; i < n_vertices; ++i) {
        v3 p = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, q = p;
        for (uint32_t k = 0; k < n_offsets; ++k) {
            q.x = q.x - offsets[3 * k]; q.y = q.y - offsets[3 * k + 1]; q.z = q.z - offsets[3 * k + 2];
        }
        const v3 d[3] = {{epsilon, 0.0f, 0.0f}, {0.0f, epsilon, 0.0f}, {0.0f, 0.0f, epsilon}};
        float nn[3];
        for (int a = 0; a < 3; ++a) {
            v3 plus = {q.x + d[a].x, q.y + d[a].y, q.z + d[a].z}, minus = {q.x - d[a].x, q.y - d[a].y, q.z - d[a].z};
            nn[a] = sdf_eval(inner, n, plus, &err) - sdf_eval(inner, n, minus, &err);
        }
        const float two_eps = 2.0f * epsilon;
        float *o = out + 6 * i;
        o[0] = p.x; o[1] = p.y; o[2] = p.z;
        o[3] = nn[0] / two_eps; o[4] = nn[1] / two_eps; o[5] = nn[2] / two_eps;
    }
    return err ? -2 : 0;
}

/* ------------------------------------------------------------------------------------ */
/* MarchingCubes::<Directed>: the same traversal with a signed distance along each of    */
/* the three cardinal axes (reference src/distance.rs:43-45,72-104).                      */
/*   sample_vector:  src/implicit/sphere.rs:41-57, torus.rs:47-97, cylinder.rs:50-72,     */
/*                   rectangular_prism.rs:42-74, csg.rs:41-45,74-78,102-106,              */
/*                   examples/common/sources.rs:46-51 (q = p - 0.5)                       */
/*   is_positive:    any component > 0                                  distance.rs:77-80 */
/*   crossing point: t from the component along the edge's axis          distance.rs:90-103 */
/* f32::min / f32::max ignore a NaN operand, as fminf / fmaxf do.                        */
/* ------------------------------------------------------------------------------------ */
#define OSDF_FMAX 3.40282347e+38f /* std::f32::MAX */

static v3 sdf_eval_vec(const osdf_node *prog, uint32_t n, v3 p, int *err) {
    v3 vs[OSDF_MAX_STACK], ps[OSDF_MAX_STACK];
    const v3 zero = {0.0f, 0.0f, 0.0f};
    int nv = 0, np = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const osdf_node *nd = &prog[i];
        const v3 a = {fabsf(p.x), fabsf(p.y), fabsf(p.z)}; /* "flip the point into the positive quadrant" */
        v3 r = zero;
        switch (nd->op) {
        case O_SPHERE: {
            const float r2 = nd->a * nd->a;
            const float l_yz = r2 - (a.y * a.y + a.z * a.z), l_xz = r2 - (a.x * a.x + a.z * a.z), l_xy = r2 - (a.x * a.x + a.y * a.y);
            r.x = l_yz < 0.0f ? OSDF_FMAX : a.x - sqrtf(l_yz);
            r.y = l_xz < 0.0f ? OSDF_FMAX : a.y - sqrtf(l_xz);
            r.z = l_xy < 0.0f ? OSDF_FMAX : a.z - sqrtf(l_xy);
        } break;
        case O_TORUS: {
            const float R = nd->a, tr = nd->b;
            const float l = sqrtf(a.x * a.x + a.y * a.y), l_xy = l - R;
            const float tz = sqrtf(tr * tr - a.z * a.z); /* tube_radius_at_z: NaN beyond the tube, comparisons are then false */
            const float rx = R + tz, ry = R - tz;
            if (a.z > tr || a.y > R + tz) r.x = OSDF_FMAX;
            else if (a.x == 0.0f) r.x = fabsf(a.y - R) - tr;
            else r.x = fmaxf(a.x - sqrtf(rx * rx - a.y * a.y), sqrtf(ry * ry - a.y * a.y) - a.x);
            if (a.z > tr || a.x > R + tz) r.y = OSDF_FMAX;
            else if (a.y == 0.0f) r.y = fabsf(a.x - R) - tr;
            else r.y = fmaxf(a.y - sqrtf(rx * rx - a.x * a.x), sqrtf(ry * ry - a.x * a.x) - a.y);
            if (fabsf(l_xy) > tr) r.z = OSDF_FMAX;
            else r.z = a.z - sqrtf(tr * tr - l_xy * l_xy);
        } break;
        case O_CYLINDER: {
            const float R = nd->a, h = nd->b;
            r.x = (a.z > h || a.y > R) ? OSDF_FMAX : a.x - sqrtf(R * R - a.y * a.y);
            r.y = (a.z > h || a.x > R) ? OSDF_FMAX : a.y - sqrtf(R * R - a.x * a.x);
            r.z = (a.x * a.x + a.y * a.y > R * R) ? OSDF_FMAX : a.z - h;
        } break;
        case O_PRISM: {
            const v3 he = {nd->a, nd->b, nd->c};
            v3 mask;
            mask.x = (((a.y - he.y) > 0.0f || (a.z - he.z) > 0.0f) ? 1.0f : -1.0f) * OSDF_FMAX;
            mask.y = (((a.x - he.x) > 0.0f || (a.z - he.z) > 0.0f) ? 1.0f : -1.0f) * OSDF_FMAX;
            mask.z = (((a.x - he.x) > 0.0f || (a.y - he.y) > 0.0f) ? 1.0f : -1.0f) * OSDF_FMAX;
            v3 c;
            if (a.x < he.x && a.y < he.y && a.z < he.z) { c.x = fmaxf(a.x, he.x); c.y = fmaxf(a.y, he.y); c.z = fmaxf(a.z, he.z); }
            else { c.x = fminf(a.x, he.x); c.y = fminf(a.y, he.y); c.z = fminf(a.z, he.z); }
            r.x = fmaxf(a.x - c.x, mask.x); r.y = fmaxf(a.y - c.y, mask.y); r.z = fmaxf(a.z - c.z, mask.z);
        } break;
        case O_UNION: case O_INTERSECTION: case O_DIFFERENCE: {
            if (nv < 2) { *err = 1; return zero; }
            const v3 A = vs[nv - 2], B = vs[nv - 1];
            if (nd->op == O_UNION) { r.x = fminf(A.x, B.x); r.y = fminf(A.y, B.y); r.z = fminf(A.z, B.z); }
            else if (nd->op == O_INTERSECTION) { r.x = fmaxf(A.x, B.x); r.y = fmaxf(A.y, B.y); r.z = fmaxf(A.z, B.z); }
            else { r.x = fmaxf(B.x, -A.x); r.y = fmaxf(B.y, -A.y); r.z = fmaxf(B.z, -A.z); } /* b.max(-a) */
            vs[nv - 2] = r;
            --nv;
            continue;
        }
        case O_TRANSLATE_PUSH:
            if (np >= OSDF_MAX_STACK) { *err = 1; return zero; }
            ps[np++] = p;
            p.x = p.x - nd->a; p.y = p.y - nd->b; p.z = p.z - nd->c;
            continue;
        case O_TRANSLATE_POP:
            if (np < 1) { *err = 1; return zero; }
            p = ps[--np];
            continue;
        default:
            *err = 1;
            return zero;
        }
        if (nv >= OSDF_MAX_STACK) { *err = 1; return zero; }
        vs[nv++] = r;
    }
    if (nv != 1 || np != 0) { *err = 1; return zero; }
    return vs[0];
}

int oracle_sample_sdf_vector(const osdf_node *prog, uint32_t n, const float *xyz, uint64_t npts, float *out) {
    int err = 0;
    for (uint64_t i = 0; i < npts; ++i) {
        v3 p = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        v3 r = sdf_eval_vec(prog, n, p, &err);
        if (err) return -2;
        out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
    }
    return 0;
}

/* MarchingCubes::<Directed>::new(size).extract(&Sampler::new(&implicit_tree), &mut IndexedVertices); lean dedup
 * (same semantics as the faithful maps: first sight of a lattice edge creates the vertex). */
int oracle_extract_sdf_directed(uint32_t size, const osdf_node *prog, uint32_t n, oracle_mesh *out) {
    memset(out, 0, sizeof *out);
    if (tables_init()) return -3;
    if (size < 1) return -1;
    int rc = 0, err = 0;
    typedef struct { v3 corner; v3 value; } dentry;
    dentry *layers[2];
    lean_map lmap = {0};
    uint64_t *faces = NULL, n_faces = 0, cap_faces = 0;
    layers[0] = (dentry *)malloc((size_t)size * size * sizeof(dentry));
    layers[1] = (dentry *)malloc((size_t)size * size * sizeof(dentry));
    if (!layers[0] || !layers[1]) { rc = -4; goto done; }
    const uint32_t sm1 = size - 1;
    const float inv = 1.0f / (float)sm1;
    for (uint32_t y = 0; y < size; ++y)
        for (uint32_t x = 0; x < size; ++x) {
            v3 c = {(float)x * inv, (float)y * inv, 0.0f};
            dentry e = {c, sdf_eval_vec(prog, n, c, &err)};
            layers[0][(size_t)y * size + x] = e;
        }
    for (uint32_t z = 0; z < size; ++z) {
        for (uint32_t y = 0; y < size; ++y)
            for (uint32_t x = 0; x < size; ++x) {
                v3 c = {(float)x * inv, (float)y * inv, (float)(z + 1) * inv};
                dentry e = {c, sdf_eval_vec(prog, n, c, &err)};
                layers[1][(size_t)y * size + x] = e;
            }
        if (err) { rc = -2; goto done; }
        for (uint32_t y = 0; y < sm1; ++y)
            for (uint32_t x = 0; x < sm1; ++x) {
                uint64_t keys[8][3];
                v3 corners[8], values[8];
                unsigned cube_index = 0;
                for (int i = 0; i < 8; ++i) {
                    keys[i][0] = x + CORNER_OFF[i][0]; keys[i][1] = y + CORNER_OFF[i][1]; keys[i][2] = z + CORNER_OFF[i][2];
                    dentry e = layers[CORNER_OFF[i][2]][(size_t)(y + CORNER_OFF[i][1]) * size + x + CORNER_OFF[i][0]];
                    corners[i] = e.corner; values[i] = e.value;
                    const int positive = e.value.x > 0.0f || e.value.y > 0.0f || e.value.z > 0.0f; /* Directed::is_positive */
                    if (!positive) cube_index |= 1u << i;
                }
                if (cube_index != 0 && cube_index != 255) out->n_active_cells++;
                v3 vertices[12];
                const unsigned edges = EDGE_MASK[cube_index];
                for (int i = 0; i < 12; ++i)
                    if (edges & (1u << i)) {
                        const int u = EDGE_ENDS[i][0], v = EDGE_ENDS[i][1];
                        const v3 pa = corners[u], pb = corners[v];
                        /* axis = (p_a - p_b).abs().max_component_index()   (vector.rs:255-263) */
                        const float dx = fabsf(pa.x - pb.x), dy = fabsf(pa.y - pb.y), dz = fabsf(pa.z - pb.z);
                        const int axis = (dx > dy && dx > dz) ? 0 : (dy > dz) ? 1 : 2;
                        const float a = axis == 0 ? values[u].x : axis == 1 ? values[u].y : values[u].z;
                        const float b = axis == 0 ? values[v].x : axis == 1 ? values[v].y : values[v].z;
                        const float delta = b - a;
                        const float t = (delta == 0.0f) ? 0.5f : -a / delta;
                        const float omt = 1.0f - t;
                        v3 r = {pa.x * omt + pb.x * t, pa.y * omt + pb.y * t, pa.z * omt + pb.z * t};
                        vertices[i] = r;
                    }
                for (int i = 0; i < 5; ++i) {
                    if (TRI[cube_index][3 * i] < 0) break;
                    uint64_t h[3];
                    for (int k = 0; k < 3; ++k) {
                        const int e = TRI[cube_index][3 * i + k];
                        const int u = EDGE_ENDS[e][0], v = EDGE_ENDS[e][1];
                        const uint64_t *a = keys[u], *b = keys[v];
                        const int a_gt_b = (a[0] != b[0]) ? (a[0] > b[0]) : (a[1] != b[1]) ? (a[1] > b[1]) : (a[2] > b[2]);
                        if (a_gt_b) { const uint64_t *t = a; a = b; b = t; }
                        /* lattice edge -> handle: the edge is (lower point, axis) */
                        const uint64_t ax = (a[0] != b[0]) ? 0 : (a[1] != b[1]) ? 1 : 2;
                        const uint64_t key = (((a[2] * size + a[1]) * size + a[0]) * 3 + ax) + 1;
                        if ((lmap.len + 1) * 8 > lmap.cap * 5) if (lean_grow(&lmap)) { rc = -4; goto done; }
                        uint64_t j = mix64(key) & (lmap.cap - 1);
                        while (lmap.k[j] != UINT64_MAX && lmap.k[j] != key) j = (j + 1) & (lmap.cap - 1);
                        if (lmap.k[j] == UINT64_MAX) {
                            lmap.k[j] = key; lmap.v[j] = (uint32_t)out->n_vertices; lmap.len++;
                            if (push_vertex(out, vertices[e])) { rc = -4; goto done; }
                        }
                        h[k] = lmap.v[j];
                    }
                    if (n_faces + 3 > cap_faces) {
                        cap_faces = cap_faces ? cap_faces * 2 : 4096;
                        uint64_t *nf = (uint64_t *)realloc(faces, cap_faces * sizeof(uint64_t));
                        if (!nf) { rc = -4; goto done; }
                        faces = nf;
                    }
                    faces[n_faces] = h[0]; faces[n_faces + 1] = h[1]; faces[n_faces + 2] = h[2];
                    n_faces += 3;
                }
            }
        { dentry *t = layers[0]; layers[0] = layers[1]; layers[1] = t; }
    }
    for (uint64_t i = 0; i < n_faces; i += 3)
        if (push_indices(out, faces[i], faces[i + 1], faces[i + 2])) { rc = -4; goto done; }
done:
    free(layers[0]); free(layers[1]); free(lmap.k); free(lmap.v); free(faces);
    if (rc) oracle_mesh_free(out);
    return rc;
}

/* PointCloud::<Directed>::new(size).extract(&Sampler::new(&implicit_tree), ..): reference src/point_cloud.rs:50-63 with
 * D = Directed -- the same traversal, classify_corners through Directed::is_positive (src/distance.rs:77-80: outside iff any
 * component is positive), one point corners[0].lerp(corners[6], 0.5) per cell whose cube index is neither 0 nor 255. */
int oracle_point_cloud_sdf_directed(uint32_t size, const osdf_node *prog, uint32_t n, oracle_mesh *out) {
    memset(out, 0, sizeof *out);
    if (size < 1) return -1;
    int rc = 0, err = 0;
    unsigned char *inside[2];
    inside[0] = (unsigned char *)malloc((size_t)size * size);
    inside[1] = (unsigned char *)malloc((size_t)size * size);
    if (!inside[0] || !inside[1]) { rc = -4; goto done; }
    const uint32_t sm1 = size - 1;
    const float inv = 1.0f / (float)sm1;
    for (uint32_t z = 0; z <= size; ++z) { /* sample layer z; cells of layer z - 1 once it is there */
        for (uint32_t y = 0; y < size; ++y)
            for (uint32_t x = 0; x < size; ++x) {
                const v3 c = {(float)x * inv, (float)y * inv, (float)z * inv};
                const v3 v = sdf_eval_vec(prog, n, c, &err);
                inside[1][(size_t)y * size + x] = !(v.x > 0.0f || v.y > 0.0f || v.z > 0.0f);
            }
        if (err) { rc = -2; goto done; }
        if (z > 0)
            for (uint32_t y = 0; y < sm1; ++y)
                for (uint32_t x = 0; x < sm1; ++x) {
                    unsigned cube_index = 0;
                    for (int i = 0; i < 8; ++i)
                        if (inside[CORNER_OFF[i][2]][(size_t)(y + CORNER_OFF[i][1]) * size + x + CORNER_OFF[i][0]]) cube_index |= 1u << i;
                    if (cube_index != 0 && cube_index != 255) {
                        const v3 a = {(float)x * inv, (float)y * inv, (float)(z - 1) * inv};           /* corners[0] */
                        const v3 b = {(float)(x + 1) * inv, (float)(y + 1) * inv, (float)z * inv};     /* corners[6] */
                        const float f = 0.5f, of = 1.0f - f;
                        const v3 p = {of * a.x + f * b.x, of * a.y + f * b.y, of * a.z + f * b.z};
                        if (push_vertex(out, p)) { rc = -4; goto done; }
                        out->n_active_cells++;
                    }
                }
        { unsigned char *t = inside[0]; inside[0] = inside[1]; inside[1] = t; }
    }
done:
    free(inside[0]); free(inside[1]);
    if (rc) oracle_mesh_free(out);
    return rc;
}
