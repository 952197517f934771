/*
 * tile_model.cu -- HOST model of the tile path.  TEST INFRASTRUCTURE ONLY (never linked into libisomc_b200.so;
 * built by tests/test_tile_model.py with `nvcc -DISOMC_HOST_MODEL`).
 *
 * It runs the very source the kernels run -- tile_count_item() and tile_emit_item() of
 * isosurface_b200/csrc/isomc_tile.cuh -- on the CPU: an emulated CTA is TILE_NT coroutines; a warp shuffle is
 * "post my value, wait for the other 31 lanes of my warp, read the source lane's"; __syncthreads is a barrier over
 * all coroutines; the TMA ring is filled synchronously (same slot arithmetic).  Work items are dealt to emulated
 * CTAs in shuffled order and the threads of a CTA are resumed in a shuffled order, so nothing can depend on the
 * order in which items or warps happen to run.  The row scan (k_scan_rows) is restated as plain prefix sums.
 * The result is compared with the oracle by the Python test; nothing here is a product path.
 */
#ifndef ISOMC_HOST_MODEL
#define ISOMC_HOST_MODEL
#endif
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <functional>
#include <vector>

#include "../isosurface_b200/csrc/isomc_tile.cuh"

namespace {

struct BlockEmu;
struct WarpEmu {
    uint32_t slot[2][32];
    uint32_t parity[32];
    uint32_t arrived = 0, gen = 0;
    BlockEmu *blk = nullptr;
    uint32_t index = 0;
};
struct BlockEmu {
    ucontext_t main_ctx, ctx[TILE_NT];
    uint32_t nthreads = TILE_NT;
    std::vector<char> stacks;
    WarpEmu warp[TILE_Y];
    uint32_t arrived = 0, gen = 0;
    bool done[TILE_NT];
    std::function<void(uint32_t)> body;
};
BlockEmu *g_blk = nullptr;

void yield_thread(BlockEmu *b, uint32_t tid) { swapcontext(&b->ctx[tid], &b->main_ctx); }

void thread_main(int tid) {
    g_blk->body((uint32_t)tid);
    g_blk->done[tid] = true;
}

/* runs body(tid) for all threads of one emulated CTA; threads are resumed in an order shuffled with `seed` */
void run_cta(BlockEmu &E, uint64_t seed, uint32_t nthreads, std::function<void(uint32_t)> body) {
    g_blk = &E;
    E.nthreads = nthreads;
    const uint32_t NT = nthreads;
    E.body = std::move(body);
    const size_t STK = 192 * 1024;
    E.stacks.resize((size_t)NT * STK);
    E.arrived = 0;
    for (uint32_t wi = 0; wi < TILE_Y; ++wi) {
        E.warp[wi].arrived = 0;
        E.warp[wi].blk = &E;
        E.warp[wi].index = wi;
        memset(E.warp[wi].parity, 0, sizeof E.warp[wi].parity);
    }
    std::vector<uint32_t> order(NT);
    for (uint32_t t = 0; t < NT; ++t) {
        order[t] = t;
        E.done[t] = false;
        getcontext(&E.ctx[t]);
        E.ctx[t].uc_stack.ss_sp = E.stacks.data() + (size_t)t * STK;
        E.ctx[t].uc_stack.ss_size = STK;
        E.ctx[t].uc_link = &E.main_ctx;
        makecontext(&E.ctx[t], (void (*)())thread_main, 1, (int)t);
    }
    uint64_t st = seed * 0x9E3779B97F4A7C15ull + 12345;
    for (;;) {
        for (uint32_t i = NT; i > 1; --i) {
            st = st * 6364136223846793005ull + 1442695040888963407ull;
            std::swap(order[i - 1], order[(st >> 33) % i]);
        }
        bool any = false;
        for (uint32_t i = 0; i < NT; ++i) {
            const uint32_t t = order[i];
            if (!E.done[t]) {
                any = true;
                swapcontext(&E.main_ctx, &E.ctx[t]);
            }
        }
        if (!any) break;
    }
}

Cta make_cta(BlockEmu &E, uint32_t tid) {
    Cta c;
    c.tid = tid;
    c.w.lane = tid & 31u;
    c.w.emu = &E.warp[tid >> 5];
    c.bemu = &E;
    return c;
}

/* sources as the model sees them: same slot ring as the device's TMA source, filled synchronously */
struct HostGridSrc3 {
    static constexpr int NS = 3, NC = 1;
    static constexpr bool ASYNC = true;
    const float *p;
    const float *base() const { return p; }
    void sample(const Geo &g, uint32_t x, uint32_t y, uint32_t lz, float *out) const { out[0] = p[((uint64_t)lz * g.N + y) * g.N + x]; }
};
struct HostGridSrc2 {
    static constexpr int NS = 2, NC = 1;
    static constexpr bool ASYNC = false;
    const float *p;
    const float *base() const { return p; }
    void sample(const Geo &g, uint32_t x, uint32_t y, uint32_t lz, float *out) const { out[0] = p[((uint64_t)lz * g.N + y) * g.N + x]; }
};
struct HostDirSrc {
    static constexpr int NS = 2, NC = 3;
    static constexpr bool ASYNC = false;
    SdfProgram prog;
    const float *base() const { return nullptr; }
    void sample(const Geo &g, uint32_t x, uint32_t y, uint32_t lz, float *out) const {
        const Vec3f v = sdf_eval_vec(prog, (float)x * g.inv, (float)y * g.inv, (float)(g.gz0 + lz) * g.inv);
        out[0] = v.x; out[1] = v.y; out[2] = v.z;
    }
};

}  // namespace

uint32_t isomc_emu_shfl(void *emu, uint32_t lane, uint32_t v, uint32_t src) {
    WarpEmu *e = (WarpEmu *)emu;
    const uint32_t p = e->parity[lane];
    e->slot[p][lane] = v;
    const uint32_t my = e->gen;
    if (++e->arrived == 32) { e->arrived = 0; e->gen++; }
    else while (e->gen == my) yield_thread(e->blk, e->index * 32 + lane);
    const uint32_t r = e->slot[p][src & 31u];
    e->parity[lane] = p ^ 1u;
    return r;
}
void isomc_emu_sync(void *emu, uint32_t lane) {
    WarpEmu *e = (WarpEmu *)emu;
    const uint32_t my = e->gen;
    if (++e->arrived == 32) { e->arrived = 0; e->gen++; return; }
    while (e->gen == my) yield_thread(e->blk, e->index * 32 + lane);
}
void isomc_emu_block_sync(void *bemu, uint32_t tid) {
    BlockEmu *b = (BlockEmu *)bemu;
    const uint32_t my = b->gen;
    if (++b->arrived == b->nthreads) { b->arrived = 0; b->gen++; return; }
    while (b->gen == my) yield_thread(b, tid);
}
uint32_t isomc_emu_atomic_add_u32(uint32_t *p, uint32_t v) { const uint32_t o = *p; *p = o + v; return o; }
uint32_t isomc_emu_next_task(uint32_t *) { return 0xFFFFFFFFu; }
void isomc_emu_atomic_add_u64(unsigned long long *p, unsigned long long v) { *p += v; }

namespace {

template <class Src>
int model_extract(uint32_t size, uint32_t z_begin, uint32_t z_end, const Src &src, uint32_t n_ctas, uint32_t seed, uint32_t zc, uint32_t vofs,
                  uint32_t cap_eb, uint32_t cap_tb, float *xyz, uint64_t cap_v, uint32_t *idx, uint64_t cap_t, uint64_t *out_totals) {
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
    memset(out_totals, 0, 8 * sizeof(uint64_t));
    if (g.ncl == 0) return 0;
    const TileGeo tg = tile_geo(g);

    McTables mt;
    if (isomc_build_tables(&mt)) return -1;
    static EmitTab et;
    isomc_build_emit_tab(mt, &et);

    const uint64_t np = (uint64_t)g.ncl * tg.ppl;
    std::vector<uint32_t> pV(np + 4, 0xDEADBEEFu), pT(np + 4, 0xDEADBEEFu), pE(np + 4, 0xDEADBEEFu), pTp(np + 4, 0xDEADBEEFu);
    std::vector<uint16_t> pA(np + 4, 0xDEADu);
    std::vector<uint2> ent((size_t)cap_eb * ENT_BLOCK);
    std::vector<float> tbuf((size_t)cap_tb * ENT_BLOCK), tq((size_t)cap_eb * ENT_BLOCK * 3);
    memset(ent.data(), 0xEE, ent.size() * sizeof(uint2));
    memset(tq.data(), 0xEE, tq.size() * sizeof(float));
    memset(tbuf.data(), 0xEE, tbuf.size() * sizeof(float));
    uint32_t ctr[2] = {0, 0};
    std::vector<unsigned long long> layerTot((size_t)g.ncl * 3 + 4, 0ull);
    TileBufs B{pV.data(), pT.data(), pE.data(), pTp.data(), pA.data(), ent.data(), tq.data(), tbuf.data(), ctr, cap_eb, cap_tb, layerTot.data()};

    /* work items (chunk-major, then column), dealt to n_ctas emulated CTAs at random; CTAs run in shuffled order */
    uint64_t st = seed * 0x9E3779B97F4A7C15ull + 1;
    auto rnd = [&](uint32_t n) { st = st * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)((st >> 33) % n); };
    const uint32_t nchunks = (g.ncl + zc - 1) / zc, nitems = nchunks * tg.ncols;
    std::vector<std::vector<uint32_t>> share(n_ctas);
    for (uint32_t it = 0; it < nitems; ++it) share[rnd(n_ctas)].push_back(it);
    std::vector<uint32_t> order(n_ctas);
    for (uint32_t i = 0; i < n_ctas; ++i) order[i] = i;
    for (uint32_t i = n_ctas; i > 1; --i) std::swap(order[i - 1], order[rnd(i)]);

    static BlockEmu E;
    {
        auto *S = new CountSmem<Src::NS, Src::NC>();
        memset((void *)S, 0xEE, sizeof *S);
        for (uint32_t ci = 0; ci < n_ctas; ++ci) {
            const std::vector<uint32_t> &items = share[order[ci]];
            if (items.empty()) continue;
            run_cta(E, seed + ci, TILE_NT, [&](uint32_t tid) {
                const Cta c = make_cta(E, tid);
                CountCtx X;
                memset(&X, 0, sizeof X);
                for (uint32_t it : items) {
                    const uint32_t chunk = it / tg.ncols, col = it % tg.ncols;
                    const uint32_t l0 = chunk * zc, l1 = std::min(g.ncl, l0 + zc);
                    tile_count_item(c, g, tg, src, *S, B, &et, col, l0, l1, X);
                }
            });
        }
        delete S;
    }

    /* k_scan_rows restated over row pieces */
    uint64_t V = 0, T = 0, Act = 0;
    for (uint64_t p = 0; p < np; ++p) {
        if (pV[p] == 0xDEADBEEFu || pT[p] == 0xDEADBEEFu || pA[p] == 0xDEADu) {
            fprintf(stderr, "tile_model: piece %llu totals not written\n", (unsigned long long)p);
            return -2;
        }
        const uint32_t v = pV[p], t = pT[p];
        pV[p] = (uint32_t)V; pT[p] = (uint32_t)T;
        V += v; T += t; Act += pA[p];
    }
    pV[np] = (uint32_t)V; pT[np] = (uint32_t)T;
    for (uint32_t l = 0; l < g.ncl; ++l) {
        const uint64_t a = (uint64_t)l * tg.ppl, b = a + tg.ppl;
        uint64_t sa = 0;
        for (uint64_t p = a; p < b; ++p) sa += pA[p];
        const uint64_t sv = (b < np ? pV[b] : V) - pV[a], stt = (b < np ? pT[b] : T) - pT[a];
        if (sv != layerTot[3 * l] || stt != layerTot[3 * l + 1] || sa != layerTot[3 * l + 2]) {
            fprintf(stderr, "tile_model: layer %u totals disagree with its pieces\n", l);
            return -3;
        }
    }
    const uint64_t gV = g.ghost ? layerTot[0] : 0, gT = g.ghost ? layerTot[1] : 0, gA = g.ghost ? layerTot[2] : 0;
    out_totals[0] = V - gV;
    out_totals[1] = pV[(uint64_t)(g.ncl - 1) * tg.ppl] - gV;
    out_totals[2] = T - gT;
    out_totals[3] = Act - gA;
    out_totals[4] = ctr[0];
    out_totals[5] = ctr[1];
    if (ctr[0] > cap_eb || ctr[1] > cap_tb) return 1;

    EmitParams P;
    P.pV = pV.data(); P.pT = pT.data(); P.pE = pE.data(); P.pTp = pTp.data(); P.pA = pA.data();
    P.ent = ent.data(); P.tq = tq.data(); P.tbuf = tbuf.data();
    P.vofs = vofs; P.ghostV = (uint32_t)gV; P.ghostT = (uint32_t)gT; P.first_own_layer = g.ghost;
    P.cap_v = cap_v; P.cap_t = cap_t; P.xyz = xyz; P.idx = idx;
    {
        auto *S = new EmitSmem();
        memset((void *)S, 0xEE, sizeof *S);
        for (int i = 0; i < 256; ++i) S->tri[i] = (mt.tri[i] & 0x0FFFFFFFFFFFFFFFull) | (unsigned long long)mt.ntri[i] << 60;
        memcpy(S->emask, mt.emask, sizeof S->emask);
        memcpy(S->rank3, mt.rank3, sizeof S->rank3);
        for (uint32_t par = 0; par < 2; ++par)
            for (uint32_t e = 0; e < 12; ++e) {
                S->etab[par][e] = tile_edge_loc(par, e);
                S->eofs[par][e] = make_uint2((S->etab[par][e] & 255u) * 4u, (S->etab[par][e] >> 12) * 2u);
            }
        const uint32_t nch2 = (g.ncl + EMIT_ZC - 1) / EMIT_ZC, nit2 = nch2 * tg.ncols_emit;
        std::vector<std::vector<uint32_t>> share2(n_ctas);
        for (uint32_t it = 0; it < nit2; ++it) share2[rnd(n_ctas)].push_back(it);
        for (uint32_t ci = 0; ci < n_ctas; ++ci) {
            const std::vector<uint32_t> &items = share2[order[ci]];
            if (items.empty()) continue;
            run_cta(E, seed + 77 + ci, EMIT_NT, [&](uint32_t tid) {
                const Cta c = make_cta(E, tid);
                for (uint32_t it : items) {
                    const uint32_t chunk = it / tg.ncols_emit, col = it % tg.ncols_emit;
                    const uint32_t l0 = chunk * EMIT_ZC, l1 = std::min(g.ncl, l0 + EMIT_ZC);
                    tile_emit_item(c, g, tg, *S, P, &et, col, l0, l1);
                }
            });
        }
        delete S;
    }
    return 0;
}

}  // namespace

extern "C" {

/*
 * slab: sample layers [z_begin - ghost, z_end] of the size x size x (size+1) lattice, x fastest.
 * n_ctas: emulated CTAs (items are dealt to them at random, execution order shuffled with `seed`); zc: cell layers per
 * counting item; ring: 3 = the TMA source's three-slot ring, 2 = the synchronous sources' two slots.
 * out_totals: [0] vertices owned, [1] of those created before the last cell layer, [2] triangles owned,
 *             [3] active cells owned, [4] entry blocks handed out, [5] t blocks handed out.
 * Returns 0, or 1 if the entry list / t buffer overflowed (outputs invalid then, totals still right).
 */
int tile_model_extract(uint32_t size, uint32_t z_begin, uint32_t z_end, const float *slab, uint32_t n_ctas, uint32_t seed, uint32_t zc,
                       uint32_t ring, uint32_t vofs, uint32_t cap_eb, uint32_t cap_tb, float *xyz, uint64_t cap_v, uint32_t *idx,
                       uint64_t cap_t, uint64_t *out_totals) {
    if (ring == 3)
        return model_extract(size, z_begin, z_end, HostGridSrc3{slab}, n_ctas, seed, zc, vofs, cap_eb, cap_tb, xyz, cap_v, idx, cap_t, out_totals);
    return model_extract(size, z_begin, z_end, HostGridSrc2{slab}, n_ctas, seed, zc, vofs, cap_eb, cap_tb, xyz, cap_v, idx, cap_t, out_totals);
}

/* MarchingCubes<Directed> over an implicit tree (whole lattice) */
int tile_model_extract_directed(uint32_t size, const isomc_sdf_node *prog, uint32_t n_nodes, uint32_t n_ctas, uint32_t seed, uint32_t zc,
                                uint32_t cap_eb, uint32_t cap_tb, float *xyz, uint64_t cap_v, uint32_t *idx, uint64_t cap_t,
                                uint64_t *out_totals) {
    HostDirSrc src;
    memset(&src.prog, 0, sizeof src.prog);
    if (n_nodes > ISOMC_SDF_MAX_NODES) return -1;
    memcpy(src.prog.nodes, prog, n_nodes * sizeof(isomc_sdf_node));
    src.prog.n = n_nodes;
    return model_extract(size, 0, size, src, n_ctas, seed, zc, 0, cap_eb, cap_tb, xyz, cap_v, idx, cap_t, out_totals);
}

/* VectorSource::sample_vector through the device evaluator's source, on the host */
int tile_model_sample_vector(const isomc_sdf_node *prog, uint32_t n_nodes, const float *xyz, uint64_t n, float *out) {
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
