/*
 * isomc_tile.cuh -- the tile path of the MarchingCubes extract (sm_100a), written once as `__host__ __device__`
 * code: the kernels (isomc_tile_kernels.cu) wrap it, and tests/tile_model.cu runs the very same functions on the
 * CPU with emulated CTAs (256 coroutines, emulated warp shuffles and block barriers) against the CPU restatement
 * of the reference -- the image this is developed in has no GPU.
 *
 * Two passes over a lattice cut into TILES of TILE_Y cell rows x TILE_X cells, marched in z:
 *
 *   pass 1  tile_count_item()   the f32 samples are read ONCE: a CTA stages (TILE_Y+1) sample rows of one sample
 *           layer per step in shared memory (cp.async.bulk + mbarrier = TMA bulk copies for aligned device grids, a
 *           ring of three layers; implicit sources evaluate the SDF into the same slots) and, while two layers are
 *           resident, produces everything that needs samples or sign bits:
 *             - the inside bits `!(v > 0)`                  (marching_cubes_impl.rs:32, distance.rs:52-54)
 *             - per 32-cell segment, bit-parallel: active cells, "vertices created" bit planes   (ownership:
 *               SURVEY.md 3.1-9 -- replaces GridKey / IndexCache / add_vertex, index_cache.rs:49-60, mesh.rs:240-251)
 *             - per ACTIVE cell (lane per cell): cube index, triangle count, in-row prefixes -> one 8-byte entry
 *             - per created vertex: the crossing parameter t = -a / (b - a)   (distance.rs:64-69), 4 bytes
 *           and per ROW PIECE (one cell row of one tile) the counts the row scan turns into global bases.
 *   pass 2  tile_emit_item()    reads no samples.  A CTA marches the same tiles with 16-bit EDGE-ID PLANES in shared
 *           memory (one per edge axis and sample layer): phase A, lane per entry, writes the ids of the edges the
 *           cell creates and turns t back into positions p_a*(1-t) + p_b*t (bit-exact: same operands, same order);
 *           phase B, same lanes, reads the ids of the triangles' edges from the planes and stores the indices
 *           (march_cube + extract_indices, marching_cubes_impl.rs:102-117, mesh.rs:91-100).  The cells that created
 *           the edges on the tile's low faces (one halo row, one halo column, the previous layer) only run phase A.
 *
 * Numbering = the reference's: id(cell, e) = base(row piece) + in-piece prefix + rank of e among the edges the cell
 * creates, in first-appearance order of its triangle list; triangle slot likewise.
 */
#ifndef ISOMC_TILE_CUH
#define ISOMC_TILE_CUH

#include <stdint.h>

#include "isomc_cell.cuh"

constexpr uint32_t TILE_X = 512;             /* cells per row piece */
constexpr uint32_t TILE_Y = 8;               /* cell rows per tile = warps per CTA */
constexpr uint32_t TILE_NT = 32 * TILE_Y;    /* threads per CTA */
constexpr uint32_t TILE_PITCH = TILE_X + 4;  /* staged floats per sample row (16-byte multiple) */
constexpr uint32_t TILE_NW = TILE_X / 32 + 2;/* sign words per staged sample row (TILE_X/32 + 1 used) */
constexpr uint32_t ENT_BLOCK = 256;          /* entries (and t values) per allocation block */
constexpr uint32_t EMIT_ZC = 16;             /* cell layers per work item of pass 2 */
constexpr uint32_t EMIT_Y = 4;               /* cell rows per tile of pass 2 (its tiles are its own: row pieces are what the passes share) */
constexpr uint32_t EMIT_NT = 160;            /* threads per CTA of pass 2: ~26 entries per row of a dense field + the halo row */

/* entry of an active cell: x = vrel (13) | tpre (12) << 13 ; y = x in tile (9) | cube index (8) << 9 */
ISOMC_HD uint2 tile_entry_pack(uint32_t vrel, uint32_t tpre, uint32_t tx, uint32_t ci) {
    return make_uint2(vrel | tpre << 13, tx | ci << 9);
}

struct TileGeo {
    uint32_t nxt, nyt;   /* tiles per row / per column of rows */
    uint32_t ncols;      /* nxt * nyt */
    uint32_t ppl;        /* row pieces per cell layer = ncx * nxt */
    uint32_t ncols_emit; /* tile columns of pass 2: nxt * ceil(ncx / EMIT_Y) */
};

static inline TileGeo tile_geo(const Geo &g) {
    TileGeo t;
    t.nxt = g.ncx ? (g.ncx + TILE_X - 1) / TILE_X : 0;
    t.nyt = g.ncx ? (g.ncx + TILE_Y - 1) / TILE_Y : 0;
    t.ncols = t.nxt * t.nyt;
    t.ppl = g.ncx * t.nxt;
    t.ncols_emit = t.nxt * (g.ncx ? (g.ncx + EMIT_Y - 1) / EMIT_Y : 0);
    return t;
}

struct TileBufs {
    uint32_t *pV, *pT;      /* per row piece: vertices created / triangles; after the row scan: exclusive prefixes */
    uint32_t *pE, *pTp;     /* per row piece: position of its first entry / of its first t value */
    uint16_t *pA;           /* per row piece: active cells */
    uint2 *ent;
    float *tq;              /* 3 floats per entry: crossing parameters of the cell's e5, e6, e10 (cells off the low faces) */
    float *tbuf;            /* crossing parameters of the cells ON the low faces (up to 12 edges each), at piece position + rank */
    uint32_t *ctr;          /* [0] entry blocks handed out, [1] t blocks handed out (may exceed the capacity: the host grows and re-runs) */
    uint32_t cap_eb, cap_tb;
    unsigned long long *layerTot; /* per cell layer: vertices, triangles, active cells */
};

/* ---- CTA abstraction (device: the real thing; host model: coroutines) ---------------------------------- */
struct Cta {
    uint32_t tid;
    Warp w;       /* lane + emulated warp */
    void *bemu;   /* emulated block (host model) */
};
#if !defined(__CUDA_ARCH__) && defined(ISOMC_HOST_MODEL)
void isomc_emu_block_sync(void *bemu, uint32_t tid);
#endif
ISOMC_HD void cta_sync(const Cta &c) {
#if defined(__CUDA_ARCH__)
    __syncthreads();
#elif defined(ISOMC_HOST_MODEL)
    isomc_emu_block_sync(c.bemu, c.tid);
#endif
}

/* ---- TMA bulk copy + mbarrier (device only) -------------------------------------------------------------- */
#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t tile_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tile_mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tile_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tile_mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tile_mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tile_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tile_bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tile_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(tile_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tile_mbar_wait(unsigned long long *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(tile_smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}
#endif

ISOMC_HD uint32_t hd_ldg8(const uint8_t *p) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ldg(p);
#else
    return *p;
#endif
}
ISOMC_HD uint32_t hd_ldg16(const uint16_t *p) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ldg(p);
#else
    return *p;
#endif
}

/* ---- pass 1 ------------------------------------------------------------------------------------------------ */

template <int NS, int NC>
struct alignas(16) CountSmem {
    float slot[NS][NC][TILE_Y + 1][TILE_PITCH];
    uint32_t sgn[2][TILE_Y + 1][TILE_NW];
    unsigned long long mbar[NS];
};

struct Cursor { /* warp-private window into the entry list / the t buffer */
    uint32_t pos, end;
};

/* n (> 0, warp-uniform) consecutive slots; ok = false once the space is exhausted (nothing may be written then) */
ISOMC_HD uint32_t cursor_alloc(const Warp &w, Cursor &c, uint32_t *ctr, uint32_t n, uint32_t cap_blocks, bool &ok) {
    if (c.end - c.pos < n) {
        const uint32_t m = (n + ENT_BLOCK - 1) / ENT_BLOCK;
        uint32_t b0 = 0;
        if (w.lane == 0) b0 = hd_atomic_add(ctr, m);
        b0 = w_shfl(w, b0, 0);
        c.pos = b0 * ENT_BLOCK;
        c.end = (b0 + m) * ENT_BLOCK;
    }
    ok = (uint64_t)c.end <= (uint64_t)cap_blocks * ENT_BLOCK;
    const uint32_t r = c.pos;
    c.pos += n;
    return r;
}

ISOMC_HD float crossing_t(float a, float b) { /* Signed::find_crossing_point, distance.rs:64-69 */
    const float delta = hd_sub(b, a);
    return (delta == 0.0f) ? 0.5f : hd_div(-a, delta);
}

/* position of the n-th (0-based) set bit of m; n < popc(m) */
ISOMC_HD uint32_t nth_bit(uint32_t m, uint32_t n) {
    uint32_t pos = 0;
#pragma unroll
    for (uint32_t wd = 16; wd; wd >>= 1) {
        const uint32_t c = hd_popc((m >> pos) & ((1u << wd) - 1u));
        if (n >= c) { n -= c; pos += wd; }
    }
    return pos;
}

struct CountCtx {          /* warp-uniform state of a counting CTA that lives across items */
    Cursor curE, curT;
    uint32_t phase;        /* bit s: parity of the next wait on mbarrier s */
};

/* sample layer L of the item's column -> slot s (synchronous sources: evaluated / loaded by all threads) */
template <class Src, int NS, int NC>
ISOMC_HD void tile_fill_sync(const Cta &c, const Geo &g, const Src &src, CountSmem<NS, NC> &S, uint32_t s, uint32_t x0, uint32_t y0,
                             uint32_t nsr, uint32_t nsx, uint32_t L) {
    const uint32_t total = nsr * nsx;
    for (uint32_t i = c.tid; i < total; i += TILE_NT) {
        const uint32_t r = i / nsx, xx = i - r * nsx;
        float v[NC];
        src.sample(g, x0 + xx, y0 + r, L, v);
#pragma unroll
        for (int k = 0; k < NC; ++k) S.slot[s][k][r][xx] = v[k];
    }
}

/* inside bit of the staged sample (row r, column xx): `!(v > 0)`; Directed: outside iff any component is positive */
template <int NS, int NC>
ISOMC_HD bool tile_inside(const CountSmem<NS, NC> &S, uint32_t s, uint32_t r, uint32_t xx) {
    if (NC == 1) return !(S.slot[s][0][r][xx] > 0.0f);
    return !(S.slot[s][0][r][xx] > 0.0f || S.slot[s][NC > 1 ? 1 : 0][r][xx] > 0.0f || S.slot[s][NC > 2 ? 2 : 0][r][xx] > 0.0f);
}

/*
 * One work item of pass 1: cell layers [l0, l1) of tile column `col`.  All TILE_NT threads of the CTA call it.
 */
template <class Src>
ISOMC_HD void tile_count_item(const Cta &c, const Geo &g, const TileGeo &tg, const Src &src, CountSmem<Src::NS, Src::NC> &S,
                              const TileBufs &B, const EmitTab *tabg, uint32_t col, uint32_t l0, uint32_t l1, CountCtx &X) {
    constexpr int NS = Src::NS, NC = Src::NC;
    const Warp &w = c.w;
    const uint32_t lane = w.lane, warp = c.tid >> 5;
    const uint32_t xt = col % tg.nxt, yt = col / tg.nxt;
    const uint32_t x0 = xt * TILE_X, y0 = yt * TILE_Y;
    const uint32_t nrows = g.ncx - y0 < TILE_Y ? g.ncx - y0 : TILE_Y, nsr = nrows + 1;
    const uint32_t ncellx = g.ncx - x0 < TILE_X ? g.ncx - x0 : TILE_X;
    const uint32_t nseg = (ncellx + 31) / 32, nsx = ncellx + 1, nwt = nseg + 1;

    /* sample layer L (local) lives in slot (L - l0) % NS, its sign words in sgn[(L - l0) & 1] */
    auto issue = [&](uint32_t L) {
        const uint32_t s = (L - l0) % NS;
        (void)s;
#if defined(__CUDA_ARCH__)
        if constexpr (Src::ASYNC) {
            if (warp == 0) { /* one bulk copy per sample row, issued by as many lanes */
                const uint32_t nfl = g.N - x0 < TILE_PITCH ? g.N - x0 : TILE_PITCH; /* floats per row (N % 4 == 0) */
                if (lane == 0) tile_mbar_expect_tx(&S.mbar[s], nsr * nfl * 4u);
                __syncwarp();
                if (lane < nsr)
                    tile_bulk_g2s(&S.slot[s][0][lane][0], src.base() + ((uint64_t)L * g.N + y0 + lane) * g.N + x0, nfl * 4u, &S.mbar[s]);
            }
            return;
        }
#endif
        tile_fill_sync(c, g, src, S, s, x0, y0, nsr, nsx, L);
    };
    auto wait = [&](uint32_t L) {
        const uint32_t s = (L - l0) % NS;
        (void)s;
#if defined(__CUDA_ARCH__)
        if constexpr (Src::ASYNC) {
            tile_mbar_wait(&S.mbar[s], X.phase >> s & 1u);
            X.phase ^= 1u << s;
            return;
        }
#endif
        cta_sync(c);
    };
    /* inside bits of sample layer L from its slot: warp r takes sample row r (a ballot per 32 samples, lane k keeps word k);
     * the tile's last sample row is shared out word by word */
    auto signs = [&](uint32_t L) {
        const uint32_t s = (L - l0) % NS, par = (L - l0) & 1u;
        if (NC == 1) {
            /* a lane takes 4 neighbouring samples (one 16-byte shared load), the nibbles of 8 neighbouring lanes make a word.
             * Samples past the lattice give arbitrary bits: every use masks the cells that do not exist.  Warp r takes sample
             * row r, the tile's last sample row is shared out chunk by chunk. */
            const uint32_t sh = (lane & 7u) * 4u;
            auto chunk = [&](uint32_t r, uint32_t ch) {
                const uint32_t xx = ch * 128 + lane * 4;
                float v0 = 1.0f, v1 = 1.0f, v2 = 1.0f, v3 = 1.0f;
                if (xx < TILE_PITCH) {
                    const float *q = &S.slot[s][0][r][xx];
#if defined(__CUDA_ARCH__)
                    const float4 v = *reinterpret_cast<const float4 *>(q);
                    v0 = v.x; v1 = v.y; v2 = v.z; v3 = v.w;
#else
                    v0 = q[0]; v1 = q[1]; v2 = q[2]; v3 = q[3];
#endif
                }
                uint32_t word = ((!(v0 > 0.0f) ? 1u : 0u) | (!(v1 > 0.0f) ? 2u : 0u) | (!(v2 > 0.0f) ? 4u : 0u) | (!(v3 > 0.0f) ? 8u : 0u)) << sh;
                word |= w_shfl(w, word, lane ^ 1u);
                word |= w_shfl(w, word, lane ^ 2u);
                word |= w_shfl(w, word, lane ^ 4u);
                const uint32_t k = ch * 4 + (lane >> 3);
                if ((lane & 7u) == 0 && k < nwt) S.sgn[par][r][k] = word;
            };
            constexpr uint32_t NCH = (TILE_X + 127) / 128 + 1;
            if (warp < nrows) {
#pragma unroll
                for (uint32_t ch = 0; ch < NCH; ++ch)
                    if (ch * 128 < nsx) chunk(warp, ch);
            }
            for (uint32_t ch = warp; ch < NCH; ch += TILE_Y)
                if (ch * 128 < nsx) chunk(nrows, ch);
            return;
        }
        if (warp < nrows) {
            uint32_t mine = 0;
#pragma unroll
            for (uint32_t k = 0; k < TILE_X / 32 + 1; ++k) {
                if (k < nwt) {
                    const uint32_t xx = k * 32 + lane;
                    const uint32_t word = w_ballot(w, xx < nsx && tile_inside(S, s, warp, xx));
                    if (lane == k) mine = word;
                }
            }
            if (lane < nwt) S.sgn[par][warp][lane] = mine;
        }
        for (uint32_t k = warp; k < nwt; k += TILE_Y) {
            const uint32_t xx = k * 32 + lane;
            const uint32_t word = w_ballot(w, xx < nsx && tile_inside(S, s, nrows, xx));
            if (lane == 0) S.sgn[par][nrows][k] = word;
        }
    };

    issue(l0);
    if (Src::ASYNC) issue(l0 + 1);
    wait(l0);
    signs(l0);
    /* (the first layer's sign words become visible with the barrier inside the loop) */

    for (uint32_t lz = l0; lz < l1; ++lz) {
        if (Src::ASYNC) { if (lz + 2 <= l1) issue(lz + 2); }
        else issue(lz + 1);
        wait(lz + 1);
        signs(lz + 1);
        cta_sync(c);

        const uint32_t sb = (lz - l0) % NS, st = (lz + 1 - l0) % NS, pb = (lz - l0) & 1u, pt = pb ^ 1u;
        if (warp < nrows) {
            const uint32_t r = warp, y = y0 + r;
            const uint32_t piece = (lz * g.ncx + y) * tg.nxt + xt;
            /* lane per 32-cell segment */
            SegClass C;
            seg_clear(C);
            uint32_t nv = 0;
            if (lane < nseg) {
                uint32_t wd[8];
                wd[0] = S.sgn[pb][r][lane]; wd[1] = S.sgn[pb][r][lane + 1];
                wd[2] = S.sgn[pb][r + 1][lane]; wd[3] = S.sgn[pb][r + 1][lane + 1];
                wd[4] = S.sgn[pt][r][lane]; wd[5] = S.sgn[pt][r][lane + 1];
                wd[6] = S.sgn[pt][r + 1][lane]; wd[7] = S.sgn[pt][r + 1][lane + 1];
                if (classify_segment(g, wd, xt * (TILE_X / 32) + lane, y, lz, C)) nv = seg_planes_count(C.p0, C.p1, C.p2, C.p3, 0xFFFFFFFFu);
            }
            const uint32_t na = hd_popc(C.act);
            uint32_t tot;
            const uint32_t pre = w_excl_scan(w, nv | na << 16, tot);
            const uint32_t vpre = pre & 0xFFFFu, apre = pre >> 16, rowV = tot & 0xFFFFu, rowA = tot >> 16;
            uint32_t rowT = 0;
            if (rowA) {
                bool okE = true, okT = true;
                const uint32_t epos = cursor_alloc(w, X.curE, B.ctr, rowA, B.cap_eb, okE);
                uint32_t tpos = 0;
                const bool zlow = (g.gz0 + lz) == 0;
                if (rowV && (zlow || y == 0 || x0 == 0)) tpos = cursor_alloc(w, X.curT, B.ctr + 1, rowV, B.cap_tb, okT); /* rows with cells on a low face */
                const bool ok = okE && okT;
                uint32_t carry = 0;
                for (uint32_t j0 = 0; j0 < rowA; j0 += 32) { /* lane per active cell, in x order */
                    const uint32_t j = j0 + lane;
                    const bool live = j < rowA;
                    /* the cell's segment: the last lane whose exclusive prefix of active cells is <= j (lanes past the row hold rowA) */
                    uint32_t sg = 0;
#pragma unroll
                    for (uint32_t step = 16; step; step >>= 1) {
                        const uint32_t t = w_shfl(w, apre, sg + step);
                        if (live && t <= j) sg += step;
                    }
                    const uint32_t am = w_shfl(w, C.act, sg), ap = w_shfl(w, apre, sg);
                    const uint32_t i = live ? nth_bit(am, j - ap) : 0u;
                    const uint32_t a0 = w_shfl(w, C.a0, sg), b0 = w_shfl(w, C.b0, sg), c0 = w_shfl(w, C.c0, sg), d0 = w_shfl(w, C.d0, sg);
                    const uint32_t nb = w_shfl(w, C.nb, sg), vp = w_shfl(w, vpre, sg);
                    const uint32_t p0 = w_shfl(w, C.p0, sg), p1 = w_shfl(w, C.p1, sg), p2 = w_shfl(w, C.p2, sg), p3 = w_shfl(w, C.p3, sg);
                    const uint32_t ci = seg_cube_index(a0, b0, c0, d0, nb, i);
                    const uint32_t nt = live ? (uint32_t)hd_ldg8(&tabg->ntri[ci]) : 0u;
                    uint32_t tt;
                    const uint32_t tpre = w_excl_scan(w, nt, tt) + carry;
                    carry += tt;
                    const uint32_t vrel = vp + seg_planes_count(p0, p1, p2, p3, (1u << i) - 1u);
                    const uint32_t tx = sg * 32 + i;
                    if (live && ok) {
                        B.ent[epos + j] = tile_entry_pack(vrel, tpre, tx, ci);
                        /* crossing parameters of the edges this cell creates, at their rank */
                        const uint32_t em = hd_ldg16(&tabg->emask[ci]);
                        if (!zlow && y != 0 && (x0 + tx) != 0) { /* creates its crossed e5 (y), e6 (x), e10 (z): all end at corner 6 */
                            const int cx = 0, cy = NC > 2 ? 1 : 0, cz = NC > 2 ? 2 : 0;
                            float *tq = B.tq + 3 * (uint64_t)(epos + j); /* beside the entry: pass 2 fetches both without decoding either */
                            /* all three computed straight-line (an uncrossed edge just gives a value nobody stores) */
                            const float t5 = crossing_t(S.slot[st][cy][r][tx + 1], S.slot[st][cy][r + 1][tx + 1]);
                            const float t6 = crossing_t(S.slot[st][cx][r + 1][tx + 1], S.slot[st][cx][r + 1][tx]);
                            const float t10 = crossing_t(S.slot[sb][cz][r + 1][tx + 1], S.slot[st][cz][r + 1][tx + 1]);
                            if (em >> 5 & 1u) tq[0] = t5;
                            if (em >> 6 & 1u) tq[1] = t6;
                            if (em >> 10 & 1u) tq[2] = t10;
                        } else { /* on a low face: also the edges lying in it */
                            const uint32_t owned = em & tabg->ownmask[cell_flags(g, x0 + tx, y, lz)];
                            float *tp = B.tbuf + tpos + vrel;
                            for (uint32_t m = owned; m; m &= m - 1) {
                                const uint32_t e = hd_ffs0(m), en = tabg->ends[e];
                                const uint32_t axis = ((en ^ en >> 4) & 7u) >> 1, pl = NC > 2 ? axis : 0u;
                                const float a = S.slot[(en >> 2 & 1u) ? st : sb][pl][r + (en >> 1 & 1u)][tx + (en & 1u)];
                                const float b = S.slot[(en >> 6 & 1u) ? st : sb][pl][r + (en >> 5 & 1u)][tx + (en >> 4 & 1u)];
                                tp[hd_popc(tabg->before[ci][e] & owned)] = crossing_t(a, b);
                            }
                        }
                    }
                }
                rowT = carry;
                if (lane == 0) {
                    B.pE[piece] = epos;
                    B.pTp[piece] = tpos;
                    if (rowV) hd_atomic_add64(&B.layerTot[3 * lz + 0], (unsigned long long)rowV);
                    if (rowT) hd_atomic_add64(&B.layerTot[3 * lz + 1], (unsigned long long)rowT);
                    hd_atomic_add64(&B.layerTot[3 * lz + 2], (unsigned long long)rowA);
                }
            }
            if (lane == 0) {
                B.pV[piece] = rowV;
                B.pT[piece] = rowT;
                B.pA[piece] = (uint16_t)rowA;
            }
        }
        cta_sync(c); /* everyone is done with the bottom slot and its sign words before they are refilled */
    }
}

/* ---- pass 2 ------------------------------------------------------------------------------------------------ */

constexpr uint32_t PL_PITCH = TILE_X + 1;
constexpr uint32_t PL_XROW = 0;                          /* X-edge planes: [parity][py 0..EMIT_Y] */
constexpr uint32_t PL_YROW = 2 * (EMIT_Y + 1);           /* Y-edge planes: [parity][cy 0..EMIT_Y-1] */
constexpr uint32_t PL_ZROW = PL_YROW + 2 * EMIT_Y;       /* Z-edge plane:  [py 0..EMIT_Y] */
constexpr uint32_t PL_ROWS = PL_ZROW + EMIT_Y + 1;       
constexpr uint32_t INFO_K = 2 * (EMIT_Y + 1);            /* row pieces a tile layer looks at */

struct LayerInfo {          /* k = 0..EMIT_Y: halo column (x tile - 1), rows -1..EMIT_Y-1;  k = EMIT_Y+1: halo row;  then the own rows */
    uint32_t V[INFO_K], T[INFO_K], E[INFO_K], Tp[INFO_K];
    uint16_t A[INFO_K];
    uint16_t cum[32];       /* phase-A sequence: prefix of the lengths (halo column cells: 0/1 each), padded with the total */
};

struct alignas(16) EmitSmem {
    int16_t plane[PL_ROWS * PL_PITCH + 2];
    LayerInfo li[EMIT_ZC + 1]; /* every layer of the item, the warm-up layer first */
    uint32_t flatB[PL_ROWS + 5]; /* id base of every plane row (+ vofs) */
    unsigned long long tri[256]; /* 15 nibbles + triangle count << 60 */
    uint16_t emask[256];
    uint8_t rank3[256];
    uint32_t etab[2][12];        /* [cell-layer parity][edge]: plane row at ty = 0 | dx << 8 | dy << 9 | (row * PL_PITCH + dx) << 12 */
    uint2 eofs[2][16];           /* the same as byte offsets: x into flatB (row at ty = 0), y into plane (cell (0, 0)) */
};

/* plane location of edge e of the cell (tx, ty) on a cell layer of parity par: row index at ty = 0, dx, dy, flat offset */
static inline uint32_t tile_edge_loc(uint32_t par, uint32_t e) {
    static const uint8_t kind[12] = {0, 1, 0, 1, 0, 1, 0, 1, 2, 2, 2, 2};   /* X, Y, Z plane */
    static const uint8_t top[12] = {0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0};
    static const uint8_t dx[12] = {0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 1, 0};
    static const uint8_t dy[12] = {0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 1};
    const uint32_t p = par ^ top[e];
    uint32_t row = kind[e] == 0 ? PL_XROW + p * (EMIT_Y + 1) : kind[e] == 1 ? PL_YROW + p * EMIT_Y : PL_ZROW;
    row += dy[e];
    return row | (uint32_t)dx[e] << 8 | (uint32_t)dy[e] << 9 | (row * PL_PITCH + dx[e]) << 12;
}

struct EmitParams {
    const uint32_t *pV, *pT, *pE, *pTp; /* pV / pT hold exclusive prefixes now */
    const uint16_t *pA;
    const uint2 *ent;
    const float *tq, *tbuf;
    uint32_t vofs;                 /* local id -> global id */
    uint32_t ghostV, ghostT;       /* vertices / triangles of a slab's ghost layer: local id / slot -> output slot */
    uint32_t first_own_layer;
    uint64_t cap_v, cap_t;
    float *xyz;
    uint32_t *idx;
};

/* position of the vertex on edge e of cell (x, y, gz): p_a*(1-t) + p_b*t per component (distance.rs:64-69, vector.rs:56-79),
 * corner coordinates (i as f32) * inv (primal_grid.rs:50,63-67).  f[axis][0/1] = coordinate of the cell's low / high corner. */
struct CellCorners { float x0, x1, y0, y1, z0, z1; };
ISOMC_HD void tile_vertex_store(float *o, uint32_t en, const CellCorners &f, float t) {
    const float omt = hd_sub(1.0f, t);
    o[0] = hd_add(hd_mul((en & 1u) ? f.x1 : f.x0, omt), hd_mul((en >> 4 & 1u) ? f.x1 : f.x0, t));
    o[1] = hd_add(hd_mul((en >> 1 & 1u) ? f.y1 : f.y0, omt), hd_mul((en >> 5 & 1u) ? f.y1 : f.y0, t));
    o[2] = hd_add(hd_mul((en >> 2 & 1u) ? f.z1 : f.z0, omt), hd_mul((en >> 6 & 1u) ? f.z1 : f.z0, t));
}

/* what a thread works on in round 0 of a layer: its place in the phase-A sequence, its entry and (own cells) the crossing
 * parameters stored beside it -- fetched a whole layer ahead, so that no global load latency is left inside a layer */
struct EmitFetch {
    uint32_t k;        /* info slot of the entry's row piece; INFO_K = none */
    uint2 ea;
    float t5, t6, t10;
};

ISOMC_HD void emit_fetch(const LayerInfo &I, const EmitParams &P, uint32_t j, EmitFetch &F) {
    F.k = INFO_K; F.ea = make_uint2(0u, 0u); F.t5 = F.t6 = F.t10 = 0.0f;
    if (j < I.cum[31]) {
        uint32_t k = 0; /* largest k with cum[k] <= j */
#pragma unroll
        for (uint32_t step = 16; step; step >>= 1)
            if (I.cum[k + step] <= j) k += step;
        const bool xh = k < EMIT_Y + 1;
        const uint32_t pos = xh ? I.E[k] + I.A[k] - 1u : I.E[k] + (j - I.cum[k]);
        F.k = k;
        F.ea = P.ent[pos];
        if (k > EMIT_Y + 1) {
            const float *q = P.tq + 3 * (uint64_t)pos;
            F.t5 = q[0]; F.t6 = q[1]; F.t10 = q[2];
        }
    }
}

/*
 * One work item of pass 2: cell layers [l0, l1) (at most EMIT_ZC) of tile column `col`, plus phase A of layer l0 - 1.
 */
ISOMC_HD void tile_emit_item(const Cta &c, const Geo &g, const TileGeo &tg, EmitSmem &S, const EmitParams &P, const EmitTab *tabg,
                             uint32_t col, uint32_t l0, uint32_t l1) {
    const Warp &w = c.w;
    const uint32_t warp = c.tid >> 5;
    const uint32_t xt = col % tg.nxt, yt = col / tg.nxt;
    const uint32_t x0 = xt * TILE_X, y0 = yt * EMIT_Y;
    const uint32_t la = l0 > 0 ? l0 - 1 : 0, nl = l1 - la;

    cta_sync(c); /* the previous item is done with the shared state */
    /* what the 18 row pieces of every layer hold (an empty piece still has its id base) */
    for (uint32_t i = c.tid; i < nl * INFO_K; i += EMIT_NT) {
        const uint32_t l = i / INFO_K, k = i - l * INFO_K;
        const bool xh = k < EMIT_Y + 1;
        const int32_t row = xh ? (int32_t)k - 1 : (int32_t)k - (int32_t)(EMIT_Y + 2);
        const int64_t y = (int64_t)y0 + row;
        uint32_t v = 0, t = 0, e = 0, tp = 0, a = 0;
        if (y >= 0 && y < (int64_t)g.ncx && (!xh || xt > 0)) {
            const uint32_t p = ((la + l) * g.ncx + (uint32_t)y) * tg.nxt + xt - (xh ? 1u : 0u);
            v = P.pV[p];
            a = P.pA[p];
            if (a) { t = P.pT[p]; e = P.pE[p]; tp = P.pTp[p]; }
        }
        LayerInfo &I = S.li[l];
        I.V[k] = v; I.T[k] = t; I.E[k] = e; I.Tp[k] = tp; I.A[k] = (uint16_t)a;
    }
    cta_sync(c);
    for (uint32_t l = warp; l < nl; l += EMIT_NT / 32) { /* the phase-A sequence of every layer */
        LayerInfo &I = S.li[l];
        const uint32_t k = w.lane;
        uint32_t len = 0;
        if (k < INFO_K) len = k < EMIT_Y + 1 ? (I.A[k] ? 1u : 0u) : I.A[k];
        uint32_t tot;
        const uint32_t ex = w_excl_scan(w, len, tot);
        I.cum[k] = (uint16_t)(k < INFO_K ? ex : tot);
    }
    cta_sync(c);

    /* first layer that creates anything; its round-0 work is fetched here, every later layer's one layer ahead */
    uint32_t lz = la;
    while (lz < l1 && S.li[lz - la].cum[31] == 0) ++lz;
    EmitFetch F;
    if (lz < l1) emit_fetch(S.li[lz - la], P, c.tid, F);

    while (lz < l1) {
        const LayerInfo &I = S.li[lz - la];
        const uint32_t nA = I.cum[31];
        cta_sync(c); /* (A) the planes of the layers below may be overwritten */
        uint32_t lzn = lz + 1; /* next layer that creates anything (nothing can refer to a layer that creates nothing) */
        while (lzn < l1 && S.li[lzn - la].cum[31] == 0) ++lzn;
        EmitFetch Fn;
        Fn.k = INFO_K; Fn.ea = make_uint2(0u, 0u); Fn.t5 = Fn.t6 = Fn.t10 = 0.0f;
        if (lzn < l1) emit_fetch(S.li[lzn - la], P, c.tid, Fn);
        const uint32_t par = lz & 1u, gz = g.gz0 + lz;
        const bool emit = lz >= l0 && lz >= P.first_own_layer;
        /* id base of every plane row: the row piece of the cells that create the edges in it */
        if (c.tid < PL_ROWS) {
            const uint32_t i = c.tid;
            uint32_t k;       /* info slot of the creating piece */
            bool bottom;      /* plane of the cell layer's lower sample layer */
            if (i < PL_YROW) {
                const uint32_t p = i / (EMIT_Y + 1), py = i - p * (EMIT_Y + 1);
                bottom = p == par;
                k = (y0 + py == 0) ? EMIT_Y + 2 : EMIT_Y + 1 + py;
            } else if (i < PL_ZROW) {
                const uint32_t p = (i - PL_YROW) / EMIT_Y, cy = (i - PL_YROW) - p * EMIT_Y;
                bottom = p == par;
                k = EMIT_Y + 2 + cy;
            } else {
                const uint32_t py = i - PL_ZROW;
                bottom = false;
                k = (y0 + py == 0) ? EMIT_Y + 2 : EMIT_Y + 1 + py;
            }
            /* edges in the lower sample layer were created one cell layer down -- except on the lattice's z = 0 face */
            const LayerInfo &J = (bottom && gz != 0 && lz > la) ? S.li[lz - la - 1] : I;
            S.flatB[i] = J.V[k] + P.vofs;
        }
        const float fz0 = hd_mul((float)gz, g.inv), fz1 = hd_mul((float)(gz + 1), g.inv);
        for (uint32_t j0 = 0; j0 < nA; j0 += EMIT_NT) {
            if (j0) emit_fetch(I, P, j0 + c.tid, F); /* (rounds beyond the first: more than EMIT_NT entries in a tile layer) */
            bool have = F.k < INFO_K;
            const uint32_t k = have ? F.k : 0u;
            const bool xh = k < EMIT_Y + 1;
            const uint2 ea = F.ea;
            if (xh && (ea.y & 511u) != TILE_X - 1) have = false; /* the piece's last cell is not the tile's neighbour */
            const int32_t tx = xh ? -1 : (int32_t)(ea.y & 511u);
            const int32_t ty = xh ? (int32_t)k - 1 : (int32_t)k - (int32_t)(EMIT_Y + 2);
            const uint32_t ci = ea.y >> 9 & 255u, vrel = ea.x & 8191u, tpre = ea.x >> 13 & 4095u;
            const uint32_t x = x0 + (uint32_t)tx, y = y0 + (uint32_t)ty;
            const bool own = have && !xh && ty >= 0;
            const int32_t cell = ty * (int32_t)PL_PITCH + tx;
            if (have) {
                const uint32_t em = S.emask[ci];
                const uint32_t fl = (x == 0 ? 1u : 0u) | (y == 0 ? 2u : 0u) | (gz == 0 ? 4u : 0u);
                /* ids are stored relative to the base of the plane row = the own-tile piece of the creating row */
                const uint32_t ko = (uint32_t)(ty + (int32_t)(EMIT_Y + 2));
                const int32_t rel0 = (int32_t)(I.V[k] - I.V[ko]) + (int32_t)vrel;
                const uint64_t vslot0 = (uint64_t)I.V[k] + vrel - P.ghostV;
                const bool put = own && emit;
                CellCorners f;
                f.x0 = hd_mul((float)x, g.inv); f.x1 = hd_mul((float)(x + 1), g.inv);
                f.y0 = hd_mul((float)y, g.inv); f.y1 = hd_mul((float)(y + 1), g.inv);
                f.z0 = fz0; f.z1 = fz1;
                if (fl == 0) { /* a cell off the low faces creates its crossed e5, e6, e10 (a halo cell: those that lie in the tile) */
                    const uint32_t r3 = S.rank3[ci];
                    const uint32_t l5 = S.etab[par][5] >> 12, l6 = S.etab[par][6] >> 12, l10 = S.etab[par][10] >> 12;
                    if ((em >> 5 & 1u) && ty >= 0) {
                        const uint32_t rank = r3 & 3u;
                        S.plane[(int32_t)l5 + cell] = (int16_t)(rel0 + (int32_t)rank);
                        if (put && vslot0 + rank < P.cap_v) tile_vertex_store(P.xyz + 3 * (vslot0 + rank), 0x75u, f, F.t5);  /* corners 5 -> 6 */
                    }
                    if ((em >> 6 & 1u) && tx >= 0) {
                        const uint32_t rank = r3 >> 2 & 3u;
                        S.plane[(int32_t)l6 + cell] = (int16_t)(rel0 + (int32_t)rank);
                        if (put && vslot0 + rank < P.cap_v) tile_vertex_store(P.xyz + 3 * (vslot0 + rank), 0x67u, f, F.t6);  /* corners 6 -> 7 */
                    }
                    if (em >> 10 & 1u) {
                        const uint32_t rank = r3 >> 4 & 3u;
                        S.plane[(int32_t)l10 + cell] = (int16_t)(rel0 + (int32_t)rank);
                        if (put && vslot0 + rank < P.cap_v) tile_vertex_store(P.xyz + 3 * (vslot0 + rank), 0x73u, f, F.t10); /* corners 2 -> 6 */
                    }
                } else { /* on a low face of the lattice: the general tables, range-checked for halo cells */
                    const uint32_t owned = em & (uint32_t)tabg->ownmask[fl];
                    const float *tp = P.tbuf + I.Tp[k] + vrel;
                    for (uint32_t m = owned; m; m &= m - 1) {
                        const uint32_t e = hd_ffs0(m);
                        const uint32_t rank = hd_popc(tabg->before[ci][e] & owned);
                        const uint32_t loc = S.etab[par][e];
                        if ((ty >= 0 || (loc >> 9 & 1u)) && (tx >= 0 || (loc >> 8 & 1u)))
                            S.plane[(int32_t)(loc >> 12) + cell] = (int16_t)(rel0 + (int32_t)rank);
                        if (put && vslot0 + rank < P.cap_v) tile_vertex_store(P.xyz + 3 * (vslot0 + rank), tabg->ends[e], f, tp[rank]);
                    }
                }
            }
            cta_sync(c); /* (B) the ids of this round (and of all earlier cells) are in the planes */
            if (own && emit) {
                unsigned long long tri = S.tri[ci];
                uint32_t nt = (uint32_t)(tri >> 60);
                const uint64_t tslot = (uint64_t)I.T[k] + tpre - P.ghostT;
                if (tslot >= P.cap_t) nt = 0;
                else if (tslot + nt > P.cap_t) nt = (uint32_t)(P.cap_t - tslot);
                uint32_t *o = P.idx + 3 * tslot;
                const uint2 *eo = S.eofs[par];
                const char *fb = reinterpret_cast<const char *>(S.flatB) + 4 * ty, *pl = reinterpret_cast<const char *>(S.plane) + 2 * cell;
                for (uint32_t t = 0; t < nt; ++t, tri >>= 12, o += 3) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const uint2 ofs = eo[(uint32_t)(tri >> (4 * q)) & 15u];
                        o[q] = *reinterpret_cast<const uint32_t *>(fb + ofs.x) + (uint32_t)(int32_t)*reinterpret_cast<const int16_t *>(pl + ofs.y);
                    }
                }
            }
        }
        F = Fn;
        lz = lzn;
    }
}

#endif /* ISOMC_TILE_CUH */
