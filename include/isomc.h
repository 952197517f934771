/*
 * isomc.h -- C ABI of libisomc_b200.so: B200-native (sm_100a) MarchingCubes extraction.
 *
 * Drop-in boundary for the hot path of the Rust crate swiftcoder/isosurface:
 *
 *     MarchingCubes::<Signed>::new(size).extract(&source, &mut extractor)
 *         (reference src/marching_cubes.rs:46-50,59-82)
 *
 * The reference has no FFI of its own; these are the entry points a `build.rs + extern "C"`
 * shim inside the crate binds (INTEGRATION.md shows that shim).  Conventions:
 *   - plain pointers and sizes only; every call returns an isomc_status (0 = OK, < 0 = error);
 *     nothing unwinds across the boundary; isomc_last_error() gives the message;
 *   - a handle mirrors `&mut self` of MarchingCubes: NOT thread-safe, one extract at a time,
 *     distinct handles are independent (one CUDA stream each);
 *   - result buffers are owned by the handle and stay valid until the next extract/destroy on
 *     it (mirrors the borrow of the reference's builder state); copy-out fills caller memory;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     ISOMC_ERR_CUDA.
 *
 * Lattice contract (reference src/traversal/primal_grid.rs:44-45,59,63-67): for `size` = N the
 * field is sampled on N x N x (N+1) lattice points, coordinate(i) = (i as f32) * (1.0f/(N-1)),
 * and (N-1) x (N-1) x N cells are visited in (z, y, x) order.  Dense grids therefore carry
 * N*N*(N+1) f32 values, x fastest, then y, then z.
 *
 * Output contract (reference src/extractor.rs:72-93, src/mesh.rs:91-100,240-251): vertices as
 * packed xyz f32 in the reference's emission order (first reference while walking triangles in
 * cell order), then 3 u32 indices per triangle in the reference's triangle order.
 */
#ifndef ISOMC_H
#define ISOMC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct isomc isomc_t;

typedef enum isomc_status {
    ISOMC_OK = 0,
    ISOMC_ERR_BAD_ARG = -1,
    ISOMC_ERR_CUDA = -2,            /* no device / driver / kernel failure */
    ISOMC_ERR_OOM = -3,
    ISOMC_ERR_INDEX_OVERFLOW = -4,  /* >= 2^32 vertices or triangles: the u32 index mesh cannot hold it
                                       (the reference silently truncates, src/extractor.rs:90-92) */
    ISOMC_ERR_UNSUPPORTED_SOURCE = -5,
    ISOMC_ERR_NO_RESULT = -6,       /* counts/copy-out before any successful extract */
    ISOMC_ERR_NCCL = -7,
    ISOMC_ERR_BUFFER_TOO_SMALL = -8 /* caller-provided mesh buffers are smaller than the result (which stays on the device) */
} isomc_status;

/* ---- implicit sources: a postfix program over the crate's shapes -------------------------
 * Encodes `Sampler::new(&tree)` for trees built from reference src/implicit/ (scalar side):
 *   SPHERE    a = radius                           sphere.rs:35-39
 *   TORUS     a = radius, b = tube_radius          torus.rs:40-46
 *   CYLINDER  a = radius, b = half_length          cylinder.rs:41-48
 *   PRISM     a,b,c = half_extent                  rectangular_prism.rs:36-41
 *   UNION / INTERSECTION / DIFFERENCE              csg.rs:35-39,68-72,96-100
 *       pop two values (first-pushed = field `a`, second = field `b`):
 *       min(a,b) / max(a,b) / max(b,-a)
 *   TRANSLATE_PUSH a,b,c ... TRANSLATE_POP         examples/common/sources.rs:38-43
 *       the nodes in between are sampled at q = p - (a,b,c)
 * Evaluation is IEEE binary32, no FMA contraction, same operation order as the reference.    */
enum {
    ISOMC_SDF_SPHERE = 1,
    ISOMC_SDF_TORUS = 2,
    ISOMC_SDF_CYLINDER = 3,
    ISOMC_SDF_PRISM = 4,
    ISOMC_SDF_UNION = 16,
    ISOMC_SDF_INTERSECTION = 17,
    ISOMC_SDF_DIFFERENCE = 18,
    ISOMC_SDF_TRANSLATE_PUSH = 32,
    ISOMC_SDF_TRANSLATE_POP = 33
};
#define ISOMC_SDF_MAX_NODES 48
#define ISOMC_SDF_MAX_STACK 8      /* values live at once (CSG nesting depth) */
#define ISOMC_SDF_MAX_TRANSLATE 4  /* nested TRANSLATE_PUSH */

typedef struct isomc_sdf_node {
    uint32_t op;
    float a, b, c;
} isomc_sdf_node;

/* per-extract statistics (new; the reference has none) */
typedef struct isomc_stats {
    uint64_t n_vertices, n_triangles, n_active_cells;
    uint64_t n_samples, n_cells;
    uint64_t algorithmic_bytes;   /* 4*S + 12*V + 12*T (grid sources) */
    float ms_sign, ms_count, ms_scan, ms_emit, ms_total; /* CUDA-event times of the last timed extract */
    uint32_t kernel_launches;     /* kernels launched by the last extract */
    uint32_t emit_reruns;         /* 1 if the output buffers had to grow and emission re-ran */
} isomc_stats;

/* ---- lifecycle:  MarchingCubes::new(size) / Drop ---------------------------------------- */
int32_t isomc_create(uint32_t size, int32_t device, isomc_t **out);
int32_t isomc_destroy(isomc_t *h);
const char *isomc_last_error(const isomc_t *h);      /* h may be NULL: last create() error */
const char *isomc_version(void);
int32_t isomc_device_count(void);   /* CUDA devices this process sees (0: none, no CPU fallback) */

/* ---- extract(&source, ...) --------------------------------------------------------------- */
/* source = implicit tree (evaluated on device; no grid is materialised) */
int32_t isomc_extract_sdf(isomc_t *h, const isomc_sdf_node *prog, uint32_t n_nodes);
/* the same tree sampled as Directed distances: `MarchingCubes::<Directed>::new(size).extract(..)` (reference
 * src/distance.rs:43-45,72-104; sample_vector of src/implicit/{sphere,torus,cylinder,rectangular_prism,csg}.rs): a lattice
 * point is outside iff any of its three axis distances is positive, and each crossing is interpolated from the distance
 * along its own edge's axis.  Same output contract as isomc_extract_sdf. */
int32_t isomc_extract_sdf_directed(isomc_t *h, const isomc_sdf_node *prog, uint32_t n_nodes);
/* source = dense lattice already resident on the handle's device: N*N*(N+1) f32 */
int32_t isomc_extract_grid_device(isomc_t *h, const float *d_grid);
/* source = dense lattice in host memory (H2D copy, then as above) */
int32_t isomc_extract_grid_host(isomc_t *h, const float *h_grid);
/* host lattice in, host mesh out in ONE call -- what `extract(&DenseGrid, &mut IndexedVertices)` does for a host-resident
 * grid (reference src/marching_cubes.rs:59-82 + src/extractor.rs:72-93).  The extract is pipelined in z-chunks: the copy-in
 * of the lattice, the kernels and the copy-out of the finished part of the mesh overlap (use pinned memory for all three
 * buffers to get the overlap).  xyz holds 3*cap_vertices floats, idx 3*cap_triangles u32.  If the mesh is larger the call
 * returns ISOMC_ERR_BUFFER_TOO_SMALL with the result left on the device: read isomc_counts(), grow, isomc_copy_out(). */
int32_t isomc_extract_grid_host_to(isomc_t *h, const float *h_grid, float *xyz, uint64_t cap_vertices, uint32_t *idx,
                                   uint64_t cap_triangles);

/* ---- many chunks in one go (new) --------------------------------------------------------------
 * The crate's usage model is many `size`^3 chunks (reference src/marching_cubes.rs:44-45, README.md:19), each one
 * `MarchingCubes::new(size).extract(&Sampler::new(&tree_b), ..)` over its own tree (a chunk is placed by the TRANSLATE nodes of
 * its tree: the reference has no chunk offset either).  A batch handle extracts up to `n_chunks` of them as ONE kernel sequence
 * with one size read-back, which removes the launch / synchronisation latency that dominates a 32^3 extract.
 * `progs` = the chunks' programs back to back, n_nodes[b] nodes each.  Results: isomc_counts() = totals over the batch,
 * isomc_copy_out() = the chunks' meshes back to back; chunk b owns vertices [v_offsets[b], v_offsets[b+1]) and triangles
 * [t_offsets[b], t_offsets[b+1]) (n_chunks + 1 entries each), and its indices are relative to ITS first vertex -- every chunk
 * is byte for byte what a single isomc_extract_sdf of its program returns. */
int32_t isomc_batch_create(uint32_t size, uint32_t n_chunks, int32_t device, isomc_t **out);
int32_t isomc_extract_sdf_batch(isomc_t *h, const isomc_sdf_node *progs, const uint32_t *n_nodes, uint32_t n_chunks);
/* the D = Directed instantiation per chunk (isomc_extract_sdf_directed; reference src/marching_cubes.rs:38-43, src/distance.rs:72-104) */
int32_t isomc_extract_sdf_batch_directed(isomc_t *h, const isomc_sdf_node *progs, const uint32_t *n_nodes, uint32_t n_chunks);
int32_t isomc_batch_offsets(isomc_t *h, uint64_t *v_offsets, uint64_t *t_offsets);
/* the same for DENSE chunks (voxel worlds cut into size^3 chunks): `n_chunks` lattices of size * size * (size + 1) f32 back to back.
 * Device-resident lattices are used in place and must fill the handle (n_chunks == the handle's capacity); host lattices are
 * copied, and a partly filled batch is padded with empty lattices. */
int32_t isomc_extract_grid_batch_device(isomc_t *h, const float *d_lattices, uint32_t n_chunks);
int32_t isomc_extract_grid_batch_host(isomc_t *h, const float *h_lattices, uint32_t n_chunks);

/* ---- PointCloud::<Signed>::new(size).extract(&source, &mut extractor)  (reference src/point_cloud.rs:50-63) ----
 * One point per active cell (cube index neither 0 nor 255): corners[0].lerp(corners[6], 0.5), in (z, y, x) cell order.
 * Results through the same calls as a mesh: isomc_counts() reports the points as vertices (0 triangles),
 * isomc_copy_out(h, xyz, NULL) / isomc_device_buffers() deliver them.  Whole-lattice handles only. */
int32_t isomc_points_sdf(isomc_t *h, const isomc_sdf_node *prog, uint32_t n_nodes);
/* PointCloud::<Directed> (the D = Directed instantiation, src/distance.rs:72-104): the tree is sampled through sample_vector and a
 * corner is outside iff any component is positive; implicit sources only */
int32_t isomc_points_sdf_directed(isomc_t *h, const isomc_sdf_node *prog, uint32_t n_nodes);
int32_t isomc_points_grid_device(isomc_t *h, const float *d_grid);
int32_t isomc_points_grid_host(isomc_t *h, const float *h_grid);

/* ---- results:  the two Vecs behind extractor::IndexedVertices --------------------------- */
int32_t isomc_counts(isomc_t *h, uint64_t *n_vertices, uint64_t *n_triangles, uint64_t *n_active_cells);
/* the same per cell layer of the last extract: counts[3 * l + 0 / 1 / 2] = vertices created / triangles / active cells of layer
 * z_begin + l, for the handle's own layers (size of them; a slab: z_end - z_begin).  What a host needs to cut z-slabs of equal
 * WORK instead of equal thickness for the next extracts of a similar field (isosurface_b200/sharded.py: balanced_slabs). */
int32_t isomc_layer_counts(isomc_t *h, uint64_t *counts);
int32_t isomc_device_buffers(isomc_t *h, const float **d_xyz, const uint32_t **d_idx);
int32_t isomc_copy_out(isomc_t *h, float *xyz /* 3*V */, uint32_t *idx /* 3*T */);
/* extractor::IndexedInterleavedNormals (reference src/extractor.rs:95-127) for a `CentralDifference` source around an
 * implicit tree (src/source.rs:82-94; epsilon 1e-6 by default there): 6 floats per vertex, position then
 * normal = (f(q+dx)-f(q-dx), f(q+dy)-f(q-dy), f(q+dz)-f(q-dz)) / (2*epsilon), evaluated on the device at the vertices of
 * the last extract.  Translations that enclose the whole program are applied to the vertex first (q = v - offset, as
 * DemoSource does around its CentralDifference, examples/common/sources.rs:55-60).  xyzn holds 6*V floats, idx 3*T (may be NULL). */
int32_t isomc_copy_out_interleaved_normals(isomc_t *h, const isomc_sdf_node *prog, uint32_t n_nodes, float epsilon, float *xyzn,
                                           uint32_t *idx);
/* the same with the POSITION of the CentralDifference adaptor made explicit: only the first `n_outer_translations` enclosing
 * TRANSLATE pairs are outside the adaptor (applied to the vertex before differencing: f((v - o) +- eps)); translations inside
 * it are part of the differenced function (f((v +- eps) - o)), which is what `CentralDifference(Translate(..))` means in the
 * reference (src/source.rs:82-94).  The call above treats every enclosing translation as outside (`Translate(CentralDifference(..))`). */
int32_t isomc_copy_out_interleaved_normals_at(isomc_t *h, const isomc_sdf_node *prog, uint32_t n_nodes, float epsilon,
                                              uint32_t n_outer_translations, float *xyzn, uint32_t *idx);
int32_t isomc_stats_get(isomc_t *h, isomc_stats *out);

/* ---- stream control (benchmarks time with CUDA events on the launching stream) ---------- */
int32_t isomc_get_stream(isomc_t *h, void **cuda_stream);
/* adopt a caller-owned cudaStream_t (e.g. the framework's current stream); NULL = back to the handle's own */
int32_t isomc_set_stream(isomc_t *h, void *cuda_stream);
/* enqueue the whole extract without synchronising; finish with isomc_finish() */
int32_t isomc_enqueue_grid_device(isomc_t *h, const float *d_grid);
int32_t isomc_enqueue_sdf(isomc_t *h, const isomc_sdf_node *prog, uint32_t n_nodes);
int32_t isomc_enqueue_sdf_directed(isomc_t *h, const isomc_sdf_node *prog, uint32_t n_nodes);
int32_t isomc_finish(isomc_t *h);
/* pre-size the output buffers so steady-state extracts never re-run emission */
int32_t isomc_reserve(isomc_t *h, uint64_t n_vertices, uint64_t n_triangles);
/* collect per-kernel CUDA-event timings into isomc_stats on subsequent extracts (0/1) */
int32_t isomc_set_profiling(isomc_t *h, int32_t on);

/* ---- z-slab sharding across GPUs (new) -------------------------------------
 * Rank g of G owns cell layers [z_begin, z_end) of the N cell layers.  It is given sample
 * layers [z_begin - (z_begin > 0), z_end] (one halo layer on the high side, and one extra on the
 * low side so that it can re-derive the ids of the boundary vertices the previous rank owns).
 * Protocol:  slab_count -> exchange totals (NCCL all-gather, 3 x u64 per rank) -> slab_emit.
 * Concatenating the ranks' vertex and index arrays in rank order gives exactly the single-GPU
 * (= reference) mesh.                                                                         */
int32_t isomc_slab_create(uint32_t size, uint32_t z_begin, uint32_t z_end, int32_t device, isomc_t **out);
/* phase 1: classify + count.  d_slab = first sample layer of the slab, (layers)*N*N f32 */
int32_t isomc_slab_count_grid_device(isomc_t *h, const float *d_slab);
int32_t isomc_slab_count_sdf(isomc_t *h, const isomc_sdf_node *prog, uint32_t n_nodes);
int32_t isomc_slab_count_sdf_directed(isomc_t *h, const isomc_sdf_node *prog, uint32_t n_nodes);   /* MarchingCubes<Directed> on a slab */
/* totals[0] = vertices owned by this slab, [1] = of those, vertices created before the slab's
 * last cell layer, [2] = triangles owned by this slab.  Synchronises the stream. */
int32_t isomc_slab_totals(isomc_t *h, uint64_t totals[3]);
/* device pointer to the same three u64 (valid after slab_count is enqueued; for on-stream NCCL) */
int32_t isomc_slab_totals_device(isomc_t *h, const uint64_t **d_totals);
/* phase 2: emit with global numbering.  vertex_base = sum of totals[0] of lower ranks;
 * boundary_base = vertex_base[g-1] + totals[1][g-1] (ignored for z_begin == 0). */
int32_t isomc_slab_emit(isomc_t *h, uint64_t vertex_base, uint64_t boundary_base);
/* phase 2, fully on-stream: d_gathered = all-gathered totals, 3 x u64 per rank, rank-major;
 * the bases are derived on the device, no host round trip between count and emit. */
int32_t isomc_slab_emit_gathered(isomc_t *h, const uint64_t *d_gathered, uint32_t rank, uint32_t n_ranks);
/* the same without the final synchronisation: finish with isomc_finish() (which also reports ISOMC_ERR_INDEX_OVERFLOW if the
 * global ids of this slab do not fit u32) */
int32_t isomc_slab_enqueue_emit_gathered(isomc_t *h, const uint64_t *d_gathered, uint32_t rank, uint32_t n_ranks);

/* ---- the exchange without a collective library: slab totals as peer stores over NVLink / NVSwitch ---------------
 * Every slab handle owns a small mailbox in device memory.  Once a rank knows the mailboxes of all ranks (same process: device
 * pointers, peer access is enabled here; one process per GPU: CUDA IPC handles, exchanged once by whatever the host has --
 * MPI, torch.distributed, a file), isomc_slab_emit_exchanged() replaces "all-gather + isomc_slab_emit_gathered": a one-CTA
 * kernel writes this rank's three totals into every rank's mailbox, waits for the others' and derives the id offset, the emission
 * follows on the same stream.  No host synchronisation, no NCCL call, one tiny launch.  A rank that never publishes makes the
 * others time out (ISOMC_EXCHANGE_TIMEOUT_MS, default 10 s) with ISOMC_ERR_NCCL instead of hanging the device.
 * All ranks must run the same sequence of slab_count / slab_emit_exchanged steps. */
#define ISOMC_IPC_HANDLE_BYTES 64
int32_t isomc_slab_mailbox(isomc_t *h, void **d_mailbox);                     /* this rank's mailbox (device pointer) */
int32_t isomc_slab_mailbox_ipc(isomc_t *h, void *handle /* ISOMC_IPC_HANDLE_BYTES */);   /* the same as a CUDA IPC handle */
int32_t isomc_slab_connect(isomc_t *h, uint32_t rank, uint32_t n_ranks, void *const *mailboxes /* [n_ranks], own entry ignored */);
int32_t isomc_slab_connect_ipc(isomc_t *h, uint32_t rank, uint32_t n_ranks, const void *handles /* n_ranks x ISOMC_IPC_HANDLE_BYTES */);
int32_t isomc_slab_emit_exchanged(isomc_t *h);
int32_t isomc_slab_enqueue_emit_exchanged(isomc_t *h);                        /* without the final synchronisation: isomc_finish() */
/* slab_count + slab_emit_exchanged in one call: memset, sign, count, scan, exchange and emission form ONE launch sequence,
 * replayed from a CUDA graph while source pointer and buffers stay the same (one cudaGraphLaunch per step and rank) */
int32_t isomc_slab_extract_grid_exchanged(isomc_t *h, const float *d_slab);
int32_t isomc_slab_enqueue_extract_grid_exchanged(isomc_t *h, const float *d_slab);   /* finish with isomc_finish() */

/* ---- the same sharding driven from ONE process over the GPUs of a box (new) ----------------
 * `MarchingCubes::new(size)` for a host that owns several devices (the Rust shim, include/isosurface.hpp): one slab handle
 * per listed device, cell layers split evenly; each extract = count on every device, ONE exchange of 3 x u64 per rank on the
 * extraction streams -- peer stores into the ranks' mailboxes when all devices reach each other as peers (NVLink / NVSwitch), else
 * one ncclAllGather (libnccl.so.2 is loaded on first use; ISOMC_EXCHANGE=nccl forces it) --, emission with global ids.
 * devices == NULL: devices 0 .. n_gpus-1.  Listing one device several times runs all slabs there (a single-GPU box can
 * exercise the sharded path; the exchange is then a stream-ordered device copy, NCCL refuses duplicate devices).
 * Concatenating the ranks' parts in rank order (isomc_sharded_copy_out does) gives exactly the unsharded mesh. */
typedef struct isomc_sharded isomc_sharded_t;
int32_t isomc_sharded_create(uint32_t size, uint32_t n_gpus, const int32_t *devices, isomc_sharded_t **out);
int32_t isomc_sharded_destroy(isomc_sharded_t *s);
const char *isomc_sharded_last_error(const isomc_sharded_t *s);   /* s may be NULL: last create() error */
int32_t isomc_sharded_uses_nccl(const isomc_sharded_t *s);        /* 1: the exchange is an NCCL all-gather */
int32_t isomc_sharded_uses_peer_memory(const isomc_sharded_t *s); /* 1: the exchange is peer stores into mailboxes (the default when every pair of devices has peer access) */
/* rank's cell layers [z_begin, z_end) and the sample layers it must be given: n_sample_layers starting at first_sample_layer */
int32_t isomc_sharded_slab(const isomc_sharded_t *s, uint32_t rank, uint32_t *z_begin, uint32_t *z_end, uint32_t *first_sample_layer,
                           uint32_t *n_sample_layers);
int32_t isomc_sharded_handle(isomc_sharded_t *s, uint32_t rank, isomc_t **h);   /* the rank's slab handle (device buffers, stats) */
/* d_slabs[r] = rank r's sample layers on ITS device (n_sample_layers * N * N f32, first = first_sample_layer) */
int32_t isomc_sharded_extract_grid(isomc_sharded_t *s, const float *const *d_slabs);
int32_t isomc_sharded_extract_sdf(isomc_sharded_t *s, const isomc_sdf_node *prog, uint32_t n_nodes);
int32_t isomc_sharded_extract_sdf_directed(isomc_sharded_t *s, const isomc_sdf_node *prog, uint32_t n_nodes);   /* <Directed> */
int32_t isomc_sharded_counts(isomc_sharded_t *s, uint64_t *n_vertices, uint64_t *n_triangles, uint64_t *n_active_cells);   /* whole mesh */
int32_t isomc_sharded_rank_counts(isomc_sharded_t *s, uint32_t rank, uint64_t *n_vertices, uint64_t *n_triangles, uint64_t *n_active_cells);
int32_t isomc_sharded_copy_out(isomc_sharded_t *s, float *xyz /* 3*V */, uint32_t *idx /* 3*T */);

/* ---- debug / parity helpers -------------------------------------------------------------- */
/* per-cell cube_index in the reference's corner order (marching_cubes_impl.rs:26-37) for the
 * last extract: (N-1)*(N-1)*N bytes to host memory, x fastest */
int32_t isomc_debug_cube_indices(isomc_t *h, uint8_t *host_out);
/* evaluate an SDF program on the device at n points (xyz packed) -> host values.  n_nodes | 0x80000000 evaluates
 * through the chain fast path the extract kernels use for left-deep trees (ISOMC_ERR_CUDA if the program has none) */
int32_t isomc_debug_sample_sdf(int32_t device, const isomc_sdf_node *prog, uint32_t n_nodes,
                               const float *h_xyz, uint64_t n_points, float *h_out);

/* ---- synthetic fields used by bench.py and the tests ---------------------- */
enum { ISOMC_FIELD_FBM = 1, ISOMC_FIELD_GYROID = 2, ISOMC_FIELD_SPHERE_UNION = 3 };
/* fills sample layers [z_first, z_first + n_layers) of the size*size*(size+1) lattice into d_out.
 * FBM: 5 octaves x 4 random-phase waves, 4*2^o periods per unit length at size 512, scaled by (size-1)/511. */
int32_t isomc_synth_field(int32_t device, int32_t kind, uint32_t size, uint64_t seed,
                          uint32_t z_first, uint32_t n_layers, float *d_out);

#ifdef __cplusplus
}
#endif
#endif /* ISOMC_H */
