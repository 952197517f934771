/* isomc_kernels.h -- host-callable launchers of the kernels in isomc_kernels.cu */
#ifndef ISOMC_KERNELS_H
#define ISOMC_KERNELS_H

#include <cuda_runtime.h>
#include <stdint.h>

struct Geo;
struct SdfProgram;
struct McTables;
struct ListBufs;
struct EmitTab;
struct TileGeo;
struct TileBufs;

struct SynthParams {
    int32_t kind;
    float amp[20], freq[20], dx[20], dy[20], dz[20], ph[20]; /* fBm: 5 octaves x 4 waves */
    float cx[64], cy[64], cz[64], r[64];                     /* sphere union */
};

/* ranges: sample rows [row0,row1) for the sign kernels, cell layers [lz0,lz1) for the rest */
cudaError_t isomc_launch_sign_grid(const Geo &g, const float *d_grid, uint32_t *signs, uint32_t row0, uint32_t row1, int sms,
                                   int ctas_per_sm, cudaStream_t st);
/* directed: sample the tree as Directed distances (inside iff no component is positive) */
cudaError_t isomc_launch_sign_sdf(const Geo &g, const SdfProgram &prog, bool directed, uint32_t *signs, uint32_t row0, uint32_t row1,
                                  int sms, int ctas_per_sm, cudaStream_t st);
/* ppl = row pieces per cell layer (cell rows, times the x-tiles of a row on the tile path) */
cudaError_t isomc_launch_scan(const Geo &g, uint32_t ppl, uint32_t *rowV, uint32_t *rowT, const unsigned long long *layerTot,
                              unsigned long long *totals, const uint32_t *list_ctr, uint32_t *list_mark, uint32_t *chunk_end,
                              uint32_t lz0, uint32_t lz1, cudaStream_t st);
/* batched chunks (Geo.zper != 0): per-chunk programs in device memory; chunk totals -> output bases */
cudaError_t isomc_launch_sign_sdf_batch(const Geo &g, const SdfProgram *d_progs, bool directed, uint32_t *signs, uint32_t row0,
                                        uint32_t row1, int sms, cudaStream_t st);
cudaError_t isomc_launch_chunk_bases(uint32_t n, unsigned long long *totals, const uint32_t *list_ctr, uint32_t *chunkV, uint32_t *chunkT,
                                     cudaStream_t st);
cudaError_t isomc_launch_slab_bases(const unsigned long long *gathered, uint32_t rank, uint32_t ghost, uint32_t *vofs,
                                    unsigned long long *ofs64, cudaStream_t st);
/* totals exchange over peer memory (k_slab_exchange): d_peers = device array of every rank's mailbox address */
#define ISOMC_MAX_RANKS 64
#define ISOMC_MAILBOX_BYTES (2 * ISOMC_MAX_RANKS * 4 * sizeof(unsigned long long))
cudaError_t isomc_launch_slab_exchange(unsigned long long *const *d_peers, uint32_t rank, uint32_t n_ranks, uint32_t ghost,
                                       unsigned long long *totals /* [15] = step counter */, uint32_t *vofs, long long timeout_cycles,
                                       cudaStream_t st);
cudaError_t isomc_launch_cube_indices(const Geo &g, const uint32_t *signs, const McTables *tabs, uint8_t *out, int sms,
                                      cudaStream_t st);
cudaError_t isomc_launch_sample_sdf(const SdfProgram &prog, const float *xyz, uint64_t n, float *out, int use_chain, cudaStream_t st);
cudaError_t isomc_launch_synth(const SynthParams &sp, uint32_t size, uint32_t z_first, uint32_t n_layers, float *out,
                               int sms, cudaStream_t st);

/* active-cell-list path (isomc_list_kernels.cu): count + list build, emission incl. vertex positions */
uint32_t isomc_count_list_max_warps(int sms);
cudaError_t isomc_launch_count_list(const Geo &g, const uint32_t *signs, const McTables *tabs, const ListBufs &L, uint32_t *rowV,
                                    uint32_t *rowT, uint32_t *rowA, unsigned long long *layerTot, uint32_t *ticket /* zeroed */,
                                    uint32_t lz0, uint32_t lz1, int sms, cudaStream_t st, int grid_bps = 0);
/* list blocks [*blk_first, *blk_end) (device pointers; blk_first == NULL: from block 0).
 * grid_bps: > 0 CTAs per SM to launch when the stage shares the SMs with others, 0 = as many as fit, < 0 = -(CTAs in total) */
cudaError_t isomc_launch_emit_list_grid(const Geo &g, const float *d_grid, const ListBufs &L, const EmitTab *tab,
                                        const uint32_t *rowPV, const uint32_t *rowPT, const unsigned long long *layerTot,
                                        const uint32_t *vofs, float *xyz, uint32_t *idx, uint64_t cap_v, uint64_t cap_t,
                                        const uint32_t *blk_first, const uint32_t *blk_end, int sms, cudaStream_t st, int grid_bps = 0);
cudaError_t isomc_launch_emit_list_sdf_batch(const Geo &g, const SdfProgram *d_progs, bool directed, const ListBufs &L, const EmitTab *tab,
                                             const uint32_t *rowPV, const uint32_t *rowPT, const unsigned long long *layerTot,
                                             const uint32_t *vofs, float *xyz, uint32_t *idx, uint64_t cap_v, uint64_t cap_t,
                                             const uint32_t *blk_first, const uint32_t *blk_end, int sms, cudaStream_t st, int grid_bps = 0);
cudaError_t isomc_launch_emit_list_sdf(const Geo &g, const SdfProgram &prog, bool directed, const ListBufs &L, const EmitTab *tab,
                                       const uint32_t *rowPV, const uint32_t *rowPT, const unsigned long long *layerTot,
                                       const uint32_t *vofs, float *xyz, uint32_t *idx, uint64_t cap_v, uint64_t cap_t,
                                       const uint32_t *blk_first, const uint32_t *blk_end, int sms, cudaStream_t st, int grid_bps = 0);

/* tile path (isomc_tile_kernels.cu): pass 1 = samples -> entries, crossing parameters, per-piece counts over cell layers
 * [lz0, lz1); pass 2 = entries -> mesh.  ticket: a zeroed u32 per launch */
size_t isomc_tile_emit_smem_bytes();
void isomc_tile_fill_eloc(uint32_t eloc[2][12]);
cudaError_t isomc_launch_tile_count_grid(const Geo &g, const TileGeo &tg, const float *d_grid, const TileBufs &B, const EmitTab *tab,
                                         uint32_t lz0, uint32_t lz1, uint32_t *ticket, int sms, cudaStream_t st);
cudaError_t isomc_launch_tile_count_sdf(const Geo &g, const TileGeo &tg, const SdfProgram &prog, bool directed, const TileBufs &B,
                                        const EmitTab *tab, uint32_t lz0, uint32_t lz1, uint32_t *ticket, int sms, cudaStream_t st);
cudaError_t isomc_launch_tile_emit(const Geo &g, const TileGeo &tg, const TileBufs &B, const EmitTab *tab, const uint32_t *vofs,
                                   float *xyz, uint32_t *idx, uint64_t cap_v, uint64_t cap_t, uint32_t lz0, uint32_t lz1, uint32_t *ticket,
                                   int sms, cudaStream_t st);

/* PointCloud (isomc_points.cu): segA = one u32 per 32-cell segment (in-row prefix of the active-cell count) */
cudaError_t isomc_launch_points_count(const Geo &g, const uint32_t *signs, uint32_t *segA, uint32_t *rowV, uint32_t *rowT,
                                      unsigned long long *layerTot, int sms, cudaStream_t st);
/* central-difference normals of an implicit tree at the vertices of the last extract, interleaved xyz | normal */
cudaError_t isomc_launch_normals_cd(const SdfProgram &inner, const float (*offsets)[3], uint32_t n_offsets, float eps,
                                    const float *xyz, uint64_t n_vertices, float *out, int sms, cudaStream_t st);
cudaError_t isomc_launch_points_emit(const Geo &g, const uint32_t *signs, const uint32_t *segA, const uint32_t *rowPV, float *xyz,
                                     uint64_t cap_v, int sms, cudaStream_t st);
#endif
