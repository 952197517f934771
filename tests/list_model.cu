/*
 * list_model.cu -- HOST model of the active-cell-list path.  TEST INFRASTRUCTURE ONLY (never linked into
 * libisomc_b200.so; built by tests/test_list_model.py with `nvcc -DISOMC_HOST_MODEL`).
 *
 * It runs the very source the kernels run -- count_list_warp() and emit_cell() of
 * isosurface_b200/csrc/isomc_cell.cuh -- on the CPU: each emulated warp is 32 coroutines, a shuffle is
 * "post my value, wait for the other 31, read the source lane's".  Warps are executed in a shuffled order
 * so that the list blocks are handed out in an order that has nothing to do with the cell order, as on
 * the GPU.  The row scan (k_scan_rows) is restated as plain prefix sums.  The result is compared with the
 * oracle by the Python test; nothing here is a product path.
 */
#ifndef ISOMC_HOST_MODEL
#define ISOMC_HOST_MODEL
#endif
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <vector>

#include "../isosurface_b200/csrc/isomc_cell.cuh"

namespace {

struct Emu {
    ucontext_t main_ctx, ctx[32];
    std::vector<char> stacks;
    uint32_t slot[2][32];
    uint32_t parity[32];
    uint32_t arrived = 0, gen = 0;
    bool done[32];
    int cur = 0;
};

Emu *g_emu = nullptr;

void emu_barrier(Emu *e, uint32_t lane) {
    const uint32_t my = e->gen;
    if (++e->arrived == 32) {
        e->arrived = 0;
        e->gen++;
        return;
    }
    while (e->gen == my) swapcontext(&e->ctx[lane], &e->main_ctx);
}

struct WarpJob {
    bool wide;
    Geo g;
    const uint32_t *signs;
    const uint8_t *ntri, *nth8;
    ListBufs L;
    CountOut out;
    uint32_t gshift, row0, row1, n_warps;
    uint32_t *ticket;
    SegQueue Q;
};
WarpJob *g_job = nullptr;

void lane_main(int lane) {
    WarpJob &J = *g_job;
    const Warp w{(uint32_t)lane, (void *)g_emu};
    if (J.wide) count_list_warp<true>(w, J.g, J.signs, J.ntri, J.nth8, J.L, J.out, J.gshift, J.row0, J.row1, J.ticket, J.Q, J.n_warps);
    else count_list_warp<false>(w, J.g, J.signs, J.ntri, J.nth8, J.L, J.out, J.gshift, J.row0, J.row1, J.ticket, J.Q, J.n_warps);
    g_emu->done[lane] = true;
}

void run_warp(Emu &E, WarpJob &J) {
    g_emu = &E;
    g_job = &J;
    const size_t STK = 256 * 1024;
    E.stacks.resize(32 * STK);
    E.arrived = 0;
    for (int l = 0; l < 32; ++l) {
        E.done[l] = false;
        E.parity[l] = 0;
        getcontext(&E.ctx[l]);
        E.ctx[l].uc_stack.ss_sp = E.stacks.data() + (size_t)l * STK;
        E.ctx[l].uc_stack.ss_size = STK;
        E.ctx[l].uc_link = &E.main_ctx;
        makecontext(&E.ctx[l], (void (*)())lane_main, 1, l);
    }
    for (;;) {
        bool any = false;
        for (int l = 0; l < 32; ++l)
            if (!E.done[l]) {
                any = true;
                swapcontext(&E.main_ctx, &E.ctx[l]);
            }
        if (!any) break;
    }
}

struct HostGridSrc {
    const float *p;
    __host__ __device__ float at(const Geo &g, uint32_t x, uint32_t y, uint32_t lz) const { return p[((uint64_t)lz * g.N + y) * g.N + x]; }
    __host__ __device__ void pair(const Geo &g, uint32_t ux, uint32_t uy, uint32_t uz, uint32_t vx, uint32_t vy, uint32_t vz, uint32_t,
                                  float &a, float &b) const {
        a = at(g, ux, uy, uz);
        b = at(g, vx, vy, vz);
    }
    __host__ __device__ void corner6(const Geo &g, uint32_t x, uint32_t y, uint32_t lz, bool n5, bool n6, bool n10, float &a5, float &b5,
                                     float &a6, float &b6, float &a10, float &b10) const {
        const float s6 = at(g, x + 1, y + 1, lz + 1);
        a5 = n5 ? at(g, x + 1, y, lz + 1) : 0.0f; b5 = s6;
        a6 = s6; b6 = n6 ? at(g, x, y + 1, lz + 1) : 0.0f;
        a10 = n10 ? at(g, x + 1, y + 1, lz) : 0.0f; b10 = s6;
    }
};

/* MarchingCubes<Directed> over an implicit tree: the device source SdfDirSrc restated for the host, same sdf_eval_vec */
struct HostDirSrc {
    SdfProgram prog;
    __host__ __device__ Vec3f vec(const Geo &g, uint32_t x, uint32_t y, uint32_t lz) const {
        return sdf_eval_vec(prog, (float)x * g.inv, (float)y * g.inv, (float)(g.gz0 + lz) * g.inv);
    }
    __host__ __device__ float at(const Geo &g, uint32_t x, uint32_t y, uint32_t lz) const {
        const Vec3f v = vec(g, x, y, lz);
        return (v.x > 0.0f || v.y > 0.0f || v.z > 0.0f) ? 1.0f : -1.0f;
    }
    static __host__ __device__ float comp(const Vec3f &v, uint32_t axis) { return axis == 0 ? v.x : axis == 1 ? v.y : v.z; }
    __host__ __device__ void pair(const Geo &g, uint32_t ux, uint32_t uy, uint32_t uz, uint32_t vx, uint32_t vy, uint32_t vz, uint32_t axis,
                                  float &a, float &b) const {
        a = comp(vec(g, ux, uy, uz), axis);
        b = comp(vec(g, vx, vy, vz), axis);
    }
    __host__ __device__ void corner6(const Geo &g, uint32_t x, uint32_t y, uint32_t lz, bool n5, bool n6, bool n10, float &a5, float &b5,
                                     float &a6, float &b6, float &a10, float &b10) const {
        const Vec3f s6 = vec(g, x + 1, y + 1, lz + 1);
        a5 = n5 ? vec(g, x + 1, y, lz + 1).y : 0.0f; b5 = s6.y;
        a6 = s6.x; b6 = n6 ? vec(g, x, y + 1, lz + 1).x : 0.0f;
        a10 = n10 ? vec(g, x + 1, y + 1, lz).z : 0.0f; b10 = s6.z;
    }
};

}  // namespace

uint32_t isomc_emu_shfl(void *emu, uint32_t lane, uint32_t v, uint32_t src) {
    Emu *e = (Emu *)emu;
    const uint32_t p = e->parity[lane];
    e->slot[p][lane] = v;
    emu_barrier(e, lane);
    const uint32_t r = e->slot[p][src & 31u];
    e->parity[lane] = p ^ 1u;
    return r;
}
void isomc_emu_sync(void *emu, uint32_t lane) { emu_barrier((Emu *)emu, lane); }
uint32_t isomc_emu_atomic_add_u32(uint32_t *p, uint32_t v) { const uint32_t o = *p; *p = o + v; return o; }
/* tasks of the emulated warp that is running (a shuffled share of all tasks); 0xFFFFFFFF = none left */
static std::vector<uint32_t> g_tasks;
uint32_t isomc_emu_next_task(uint32_t *) {
    if (g_tasks.empty()) return 0xFFFFFFFFu;
    const uint32_t t = g_tasks.back();
    g_tasks.pop_back();
    return t;
}
void isomc_emu_atomic_add_u64(unsigned long long *p, unsigned long long v) { *p += v; }

extern "C" {

/*
 * slab: sample layers [z_begin - ghost, z_end] of the size x size x (size+1) lattice, x fastest.
 * n_warps: virtual warps of the count pass (their execution order is shuffled with `seed`).
 * out_totals: [0] vertices owned, [1] of those created before the last cell layer, [2] triangles owned,
 *             [3] active cells owned, [4] list blocks handed out, [5] list entries.
 * Returns 0, or 1 if the list overflowed cap_blocks (outputs invalid then, totals still right).
 */
static int model_extract(uint32_t size, uint32_t z_begin, uint32_t z_end, const float *slab, const SdfProgram *directed,
                         uint32_t n_warps, uint32_t seed, uint32_t vofs, uint32_t cap_blocks, float *xyz, uint64_t cap_v, uint32_t *idx,
                         uint64_t cap_t, uint64_t *out_totals, uint32_t batch = 0, uint64_t *chunk_v = nullptr, uint64_t *chunk_t = nullptr) {
    Geo g;
    g.N = size; g.ncx = size - 1;
    g.nsegx = (g.ncx + 31) / 32; g.nws = (g.nsegx + 2) & ~1u;
    g.ghost = z_begin > 0 ? 1u : 0u;
    g.gz0 = z_begin - g.ghost;
    g.ncl = (g.ncx == 0) ? 0 : (z_end - z_begin + g.ghost);
    g.nsl = g.ncl + 1;
    g.inv = 1.0f / (float)(size - 1);
    g.row_magic = g.ncx ? ((1ull << 40) + g.ncx - 1) / g.ncx : 0;
    g.zper = 0; g.zmagic = 0;
    if (batch) { /* `batch` whole lattices stacked in z (isomc_batch_create): the cell layer between two of them is dead */
        g.zper = size + 1;
        g.zmagic = geo_zmagic(g.zper);
        g.ncl = batch * (size + 1) - 1;
        g.nsl = g.ncl + 1;
    }
    memset(out_totals, 0, 6 * sizeof(uint64_t));
    if (g.ncl == 0) return 0;

    /* K1 restated: inside bit = !(v > 0) */
    const uint64_t nrows_s = (uint64_t)g.nsl * g.N, nrows_c = (uint64_t)g.ncl * g.ncx;
    std::vector<uint32_t> signs(nrows_s * g.nws + 4, 0u);
    HostDirSrc dsrc;
    if (directed) dsrc.prog = *directed;
    for (uint64_t r = 0; r < nrows_s; ++r)
        for (uint32_t x = 0; x < g.N; ++x) {
            const float v = directed ? dsrc.at(g, x, (uint32_t)(r % g.N), (uint32_t)(r / g.N)) : slab[r * g.N + x];
            if (!(v > 0.0f)) signs[r * g.nws + (x >> 5)] |= 1u << (x & 31);
        }

    McTables mt;
    if (isomc_build_tables(&mt)) return -1;
    EmitTab et;
    isomc_build_emit_tab(mt, &et);

    std::vector<uint2> ent((size_t)cap_blocks * LIST_BLOCK);
    std::vector<uint2> segrec(nrows_c * g.nsegx);
    std::vector<uint32_t> segtpre(nrows_c * g.nsegx, 0xEEEEEEEEu);
    std::vector<uint32_t> ent_yz((size_t)cap_blocks * LIST_BLOCK), blkfill(cap_blocks, 0xDEADBEEFu);
    /* poison what the kernels may only read after writing */
    memset(ent.data(), 0xEE, ent.size() * sizeof(uint2));
    memset(segrec.data(), 0xEE, segrec.size() * sizeof(uint2));
    uint32_t ctr = 0;
    ListBufs L{ent.data(), ent_yz.data(), segrec.data(), segtpre.data(), blkfill.data(), &ctr, cap_blocks, nullptr, nullptr};
    /* zeroed before the count, as the library's memset does: the kernel only writes the rows that have active cells */
    std::vector<uint32_t> rowV(nrows_c + 4, 0u), rowT(nrows_c + 4, 0u), rowA(nrows_c + 4, 0u);
    std::vector<unsigned long long> layerTot((size_t)g.ncl * 3 + 4, 0ull);
    CountOut out{rowV.data(), rowT.data(), rowA.data(), layerTot.data()};

    const uint32_t npair = (g.nsegx + 1) / 2; /* lanes per row: every lane scans two segments */
    uint32_t gshift = 0;
    while ((1u << gshift) < npair && gshift < 5) ++gshift;
    std::vector<uint32_t> order(n_warps);
    for (uint32_t i = 0; i < n_warps; ++i) order[i] = i;
    uint64_t st = seed * 0x9E3779B97F4A7C15ull + 1;
    for (uint32_t i = n_warps; i > 1; --i) {
        st = st * 6364136223846793005ull + 1442695040888963407ull;
        std::swap(order[i - 1], order[(st >> 33) % i]);
    }
    std::vector<uint8_t> nth8(256 * 8);
    for (uint32_t m = 0; m < 256; ++m) nth8_fill(nth8.data(), m);
    /* on the GPU a ticket counter hands tasks of count_task_passes() passes to whichever warp asks next; here every task goes
     * to a random emulated warp, and the warps run one after the other in shuffled order */
    uint32_t ticket = 0;
    const uint32_t npass = npair > 32 ? (uint32_t)nrows_c : (uint32_t)((nrows_c + (32u >> gshift) - 1) / (32u >> gshift));
    const uint32_t tp = count_task_passes(npass, n_warps, npair > 32 ? 1u : 32u >> gshift);
    const uint32_t ntask = (npass + tp - 1) / tp;
    std::vector<std::vector<uint32_t>> share(n_warps);
    for (uint32_t t = 0; t < ntask; ++t) {
        st = st * 6364136223846793005ull + 1442695040888963407ull;
        share[(st >> 33) % n_warps].push_back(t);
    }
    static Emu E;
    for (uint32_t wi = 0; wi < n_warps; ++wi) {
        g_tasks = share[order[wi]];
        std::reverse(g_tasks.begin(), g_tasks.end()); /* increasing task order within a warp, as a ticket counter gives */
        WarpJob J;
        J.wide = npair > 32;
        J.g = g; J.signs = signs.data(); J.ntri = mt.ntri; J.nth8 = nth8.data(); J.L = L; J.out = out;
        J.gshift = gshift; J.row0 = 0; J.row1 = (uint32_t)nrows_c; J.ticket = &ticket; J.n_warps = n_warps;
        run_warp(E, J);
    }

    /* k_scan_rows restated */
    std::vector<uint32_t> rowPV(nrows_c + 1), rowPT(nrows_c + 1), chunkV(batch + 1, 0u), chunkT(batch + 1, 0u);
    uint64_t V = 0, T = 0, Act = 0, Vb = 0, Tb = 0;
    for (uint64_t r = 0; r < nrows_c; ++r) {
        if (batch && r % ((uint64_t)g.zper * g.ncx) == 0) { /* a new lattice: ids are chunk-local (k_scan_rows), bases as k_chunk_bases */
            Vb += V; Tb += T;
            chunkV[r / ((uint64_t)g.zper * g.ncx)] = (uint32_t)Vb; chunkT[r / ((uint64_t)g.zper * g.ncx)] = (uint32_t)Tb;
            V = 0; T = 0;
        }
        rowPV[r] = (uint32_t)V; rowPT[r] = (uint32_t)T;
        V += rowV[r]; T += rowT[r]; Act += rowA[r];
    }
    rowPV[nrows_c] = (uint32_t)V; rowPT[nrows_c] = (uint32_t)T;
    if (batch) {
        chunkV[batch] = (uint32_t)(Vb + V); chunkT[batch] = (uint32_t)(Tb + T);
        for (uint32_t b = 0; b <= batch; ++b) { chunk_v[b] = chunkV[b]; chunk_t[b] = chunkT[b]; }
        L.chunkV = chunkV.data(); L.chunkT = chunkT.data();
        V += Vb; T += Tb;
    }
    /* layer totals must agree with the row totals */
    for (uint32_t l = 0; l < g.ncl; ++l) {
        uint64_t sv = 0, stt = 0, sa = 0;
        for (uint32_t y = 0; y < g.ncx; ++y) { sv += rowV[l * g.ncx + y]; stt += rowT[l * g.ncx + y]; sa += rowA[l * g.ncx + y]; }
        if (sv != layerTot[3 * l] || stt != layerTot[3 * l + 1] || sa != layerTot[3 * l + 2]) {
            fprintf(stderr, "list_model: layer %u totals disagree with its rows\n", l);
            return -3;
        }
    }
    const uint64_t gV = g.ghost ? layerTot[0] : 0, gT = g.ghost ? layerTot[1] : 0, gA = g.ghost ? layerTot[2] : 0;
    out_totals[0] = V - gV;
    out_totals[1] = batch ? 0 : rowPV[(uint64_t)(g.ncl - 1) * g.ncx] - gV;
    out_totals[2] = T - gT;
    out_totals[3] = Act - gA;
    out_totals[4] = ctr;
    out_totals[5] = Act;
    if (ctr > cap_blocks) return 1;

    /* every handed-out block must have its fill recorded, and the fills must add up to the active cells */
    uint64_t filled = 0;
    for (uint32_t b = 0; b < ctr; ++b) {
        if (blkfill[b] > LIST_BLOCK) { fprintf(stderr, "list_model: block %u fill %u\n", b, blkfill[b]); return -4; }
        filled += blkfill[b];
    }
    if (filled != Act) { fprintf(stderr, "list_model: fills %llu != active cells %llu\n", (unsigned long long)filled, (unsigned long long)Act); return -5; }

    /* k_emit_list restated: one "lane" per entry */
    EmitArgs A;
    A.rowPV = rowPV.data(); A.rowPT = rowPT.data();
    A.vofs = vofs; A.ghostV = (uint32_t)gV; A.ghostT = (uint32_t)gT;
    A.first_own_layer = g.ghost;
    A.cap_v = cap_v; A.cap_t = cap_t; A.xyz = xyz; A.idx = idx;
    HostGridSrc src{slab};
    uint32_t eid[12 * LIST_BLOCK];
    for (uint32_t b = 0; b < ctr; ++b)
        for (uint32_t j = 0; j < blkfill[b]; ++j) {
            const uint64_t k = (uint64_t)b * LIST_BLOCK + j;
            if (directed) emit_cell(g, dsrc, et, L, A, k, ent[k], ent_yz[k], eid + j, LIST_BLOCK);
            else emit_cell(g, src, et, L, A, k, ent[k], ent_yz[k], eid + j, LIST_BLOCK);
        }
    return 0;
}

int list_model_extract(uint32_t size, uint32_t z_begin, uint32_t z_end, const float *slab, uint32_t n_warps, uint32_t seed,
                       uint32_t vofs, uint32_t cap_blocks, float *xyz, uint64_t cap_v, uint32_t *idx, uint64_t cap_t,
                       uint64_t *out_totals) {
    return model_extract(size, z_begin, z_end, slab, nullptr, n_warps, seed, vofs, cap_blocks, xyz, cap_v, idx, cap_t, out_totals);
}

/* isomc_extract_sdf_batch: `batch` lattices (size^2 x (size+1) samples each, back to back) through one count / scan / emit;
 * chunk b's vertices are [chunk_v[b], chunk_v[b+1]), its triangles [chunk_t[b], chunk_t[b+1]), its indices chunk-local */
int list_model_extract_batch(uint32_t size, uint32_t batch, const float *lattices, uint32_t n_warps, uint32_t seed, uint32_t cap_blocks,
                             float *xyz, uint64_t cap_v, uint32_t *idx, uint64_t cap_t, uint64_t *out_totals, uint64_t *chunk_v,
                             uint64_t *chunk_t) {
    return model_extract(size, 0, size, lattices, nullptr, n_warps, seed, 0, cap_blocks, xyz, cap_v, idx, cap_t, out_totals, batch, chunk_v,
                         chunk_t);
}

/* MarchingCubes<Directed> over an implicit tree (whole lattice): the list kernels' source with the Directed host source */
int list_model_extract_directed(uint32_t size, const isomc_sdf_node *prog, uint32_t n_nodes, uint32_t n_warps, uint32_t seed,
                                uint32_t cap_blocks, float *xyz, uint64_t cap_v, uint32_t *idx, uint64_t cap_t, uint64_t *out_totals) {
    SdfProgram P;
    memset(&P, 0, sizeof P);
    if (n_nodes > ISOMC_SDF_MAX_NODES) return -1;
    memcpy(P.nodes, prog, n_nodes * sizeof(isomc_sdf_node));
    P.n = n_nodes;
    return model_extract(size, 0, size, nullptr, &P, n_warps, seed, 0, cap_blocks, xyz, cap_v, idx, cap_t, out_totals);
}

/* VectorSource::sample_vector through the device evaluator's source, on the host */
int list_model_sample_vector(const isomc_sdf_node *prog, uint32_t n_nodes, const float *xyz, uint64_t n, float *out) {
    SdfProgram P;
    memset(&P, 0, sizeof P);
    if (n_nodes > ISOMC_SDF_MAX_NODES) return -1;
    memcpy(P.nodes, prog, n_nodes * sizeof(isomc_sdf_node));
    P.n = n_nodes;
    for (uint64_t i = 0; i < n; ++i) {
        const Vec3f v = sdf_eval_vec(P, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z;
    }
    return 0;
}

} /* extern "C" */
