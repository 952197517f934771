//! Reference-side shim: the swiftcoder/isosurface API for the MarchingCubes path, executed by
//! libisomc_b200.so on a B200.  UNCOMPILED HERE (no Rust toolchain in the build image): this file is
//! the binding a maintainer of the crate would add; the C ABI it targets is include/isomc.h, and the
//! same call sequence is exercised from C++ (include/isosurface.hpp) and Python (isosurface_b200/).
//!
//! Replaces, in the reference: `MarchingCubes::<Signed>::extract` (src/marching_cubes.rs:59-82).
//! Keeps: `Sampler` (src/sampler.rs:26-41), `Extractor` (src/extractor.rs:17-20), the implicit shapes.
#![allow(non_camel_case_types)]
use std::os::raw::c_char;

#[repr(C)]
pub struct isomc_t { _private: [u8; 0] }

#[repr(C)]
#[derive(Copy, Clone)]
pub struct isomc_sdf_node { pub op: u32, pub a: f32, pub b: f32, pub c: f32 }

extern "C" {
    fn isomc_create(size: u32, device: i32, out: *mut *mut isomc_t) -> i32;
    fn isomc_destroy(h: *mut isomc_t) -> i32;
    fn isomc_last_error(h: *const isomc_t) -> *const c_char;
    fn isomc_extract_sdf(h: *mut isomc_t, prog: *const isomc_sdf_node, n_nodes: u32) -> i32;
    fn isomc_extract_sdf_directed(h: *mut isomc_t, prog: *const isomc_sdf_node, n_nodes: u32) -> i32;
    fn isomc_extract_grid_host(h: *mut isomc_t, grid: *const f32) -> i32;
    fn isomc_extract_grid_device(h: *mut isomc_t, d_grid: *const f32) -> i32;
    fn isomc_extract_grid_host_to(h: *mut isomc_t, grid: *const f32, xyz: *mut f32, cap_vertices: u64, idx: *mut u32,
                                  cap_triangles: u64) -> i32;
    fn isomc_points_sdf(h: *mut isomc_t, prog: *const isomc_sdf_node, n_nodes: u32) -> i32;
    fn isomc_points_sdf_directed(h: *mut isomc_t, prog: *const isomc_sdf_node, n_nodes: u32) -> i32;
    fn isomc_points_grid_host(h: *mut isomc_t, grid: *const f32) -> i32;
    fn isomc_counts(h: *mut isomc_t, nv: *mut u64, nt: *mut u64, na: *mut u64) -> i32;
    fn isomc_copy_out(h: *mut isomc_t, xyz: *mut f32, idx: *mut u32) -> i32;
    fn isomc_copy_out_interleaved_normals(h: *mut isomc_t, prog: *const isomc_sdf_node, n_nodes: u32, epsilon: f32,
                                          xyzn: *mut f32, idx: *mut u32) -> i32;
    // many chunks per call
    fn isomc_batch_create(size: u32, n_chunks: u32, device: i32, out: *mut *mut isomc_t) -> i32;
    fn isomc_extract_sdf_batch(h: *mut isomc_t, progs: *const isomc_sdf_node, n_nodes: *const u32, n_chunks: u32) -> i32;
    fn isomc_extract_sdf_batch_directed(h: *mut isomc_t, progs: *const isomc_sdf_node, n_nodes: *const u32, n_chunks: u32) -> i32;
    fn isomc_batch_offsets(h: *mut isomc_t, v_offsets: *mut u64, t_offsets: *mut u64) -> i32;
    fn isomc_extract_grid_batch_host(h: *mut isomc_t, h_lattices: *const f32, n_chunks: u32) -> i32;
    // z-slabs: one rank per process (the host brings the exchange) ...
    pub fn isomc_slab_create(size: u32, z_begin: u32, z_end: u32, device: i32, out: *mut *mut isomc_t) -> i32;
    pub fn isomc_slab_count_grid_device(h: *mut isomc_t, d_slab: *const f32) -> i32;
    pub fn isomc_slab_count_sdf(h: *mut isomc_t, prog: *const isomc_sdf_node, n_nodes: u32) -> i32;
    pub fn isomc_slab_totals(h: *mut isomc_t, totals: *mut u64) -> i32;
    pub fn isomc_slab_emit(h: *mut isomc_t, vertex_base: u64, boundary_base: u64) -> i32;
    pub fn isomc_slab_emit_gathered(h: *mut isomc_t, d_gathered: *const u64, rank: u32, n_ranks: u32) -> i32;
    pub fn isomc_slab_mailbox_ipc(h: *mut isomc_t, handle64: *mut u8) -> i32;
    pub fn isomc_slab_connect_ipc(h: *mut isomc_t, rank: u32, n_ranks: u32, handles: *const u8) -> i32;
    pub fn isomc_slab_emit_exchanged(h: *mut isomc_t) -> i32;
    // ... or all ranks of a box driven from this process
    fn isomc_sharded_create(size: u32, n_gpus: u32, devices: *const i32, out: *mut *mut isomc_sharded_t) -> i32;
    fn isomc_sharded_destroy(s: *mut isomc_sharded_t) -> i32;
    fn isomc_sharded_last_error(s: *const isomc_sharded_t) -> *const c_char;
    fn isomc_sharded_slab(s: *const isomc_sharded_t, rank: u32, z_begin: *mut u32, z_end: *mut u32, first_sample_layer: *mut u32,
                          n_sample_layers: *mut u32) -> i32;
    fn isomc_sharded_extract_grid(s: *mut isomc_sharded_t, d_slabs: *const *const f32) -> i32;
    fn isomc_sharded_extract_sdf(s: *mut isomc_sharded_t, prog: *const isomc_sdf_node, n_nodes: u32) -> i32;
    fn isomc_sharded_counts(s: *mut isomc_sharded_t, nv: *mut u64, nt: *mut u64, na: *mut u64) -> i32;
    fn isomc_sharded_copy_out(s: *mut isomc_sharded_t, xyz: *mut f32, idx: *mut u32) -> i32;
}
#[repr(C)]
pub struct isomc_sharded_t { _private: [u8; 0] }
const ERR_BUFFER_TOO_SMALL: i32 = -8;

const SPHERE: u32 = 1; const TORUS: u32 = 2; const CYLINDER: u32 = 3; const PRISM: u32 = 4;
const UNION: u32 = 16; const INTERSECTION: u32 = 17; const DIFFERENCE: u32 = 18;
const TRANSLATE_PUSH: u32 = 32; const TRANSLATE_POP: u32 = 33;

/// A source the device can evaluate: it appends itself to a postfix program.
/// Arbitrary `ScalarSource` closures do not implement this -- rejected at compile time, no CPU path.
pub trait DeviceSource { fn encode(&self, prog: &mut Vec<isomc_sdf_node>); }

fn node(op: u32, a: f32, b: f32, c: f32) -> isomc_sdf_node { isomc_sdf_node { op, a, b, c } }

// In the crate these impls sit next to the shapes (src/implicit/*.rs); shown here on mirror types.
pub struct Sphere { pub radius: f32 }
pub struct Torus { pub radius: f32, pub tube_radius: f32 }
pub struct Cylinder { pub radius: f32, pub half_length: f32 }
pub struct RectangularPrism { pub half_extent: [f32; 3] }
pub struct Union<A, B> { pub a: A, pub b: B }
pub struct Intersection<A, B> { pub a: A, pub b: B }
pub struct Difference<A, B> { pub a: A, pub b: B }
pub struct Translate<S> { pub offset: [f32; 3], pub source: S }
/// Dense lattice: N*N*(N+1) f32, x fastest (the reference samples one z layer past `size`).
pub struct DenseGrid<'a> { pub size: usize, pub data: &'a [f32] }
pub struct Sampler<'a, S> { pub source: &'a S }
impl<'a, S> Sampler<'a, S> { pub fn new(source: &'a S) -> Self { Self { source } } }

impl DeviceSource for Sphere { fn encode(&self, p: &mut Vec<isomc_sdf_node>) { p.push(node(SPHERE, self.radius, 0.0, 0.0)) } }
impl DeviceSource for Torus { fn encode(&self, p: &mut Vec<isomc_sdf_node>) { p.push(node(TORUS, self.radius, self.tube_radius, 0.0)) } }
impl DeviceSource for Cylinder { fn encode(&self, p: &mut Vec<isomc_sdf_node>) { p.push(node(CYLINDER, self.radius, self.half_length, 0.0)) } }
impl DeviceSource for RectangularPrism {
    fn encode(&self, p: &mut Vec<isomc_sdf_node>) { p.push(node(PRISM, self.half_extent[0], self.half_extent[1], self.half_extent[2])) }
}
macro_rules! binary { ($t:ident, $op:expr) => {
    impl<A: DeviceSource, B: DeviceSource> DeviceSource for $t<A, B> {
        fn encode(&self, p: &mut Vec<isomc_sdf_node>) { self.a.encode(p); self.b.encode(p); p.push(node($op, 0.0, 0.0, 0.0)) }
    }
} }
binary!(Union, UNION); binary!(Intersection, INTERSECTION); binary!(Difference, DIFFERENCE);
impl<S: DeviceSource> DeviceSource for Translate<S> {
    fn encode(&self, p: &mut Vec<isomc_sdf_node>) {
        p.push(node(TRANSLATE_PUSH, self.offset[0], self.offset[1], self.offset[2]));
        self.source.encode(p);
        p.push(node(TRANSLATE_POP, 0.0, 0.0, 0.0));
    }
}
/// reference src/source.rs:52-94: central-difference normals around a scalar source (transparent as a scalar source)
pub struct CentralDifference<S> { pub source: S, pub epsilon: f32 }
impl<S> CentralDifference<S> { pub fn new(source: S) -> Self { Self { source, epsilon: 0.000001 } } }
impl<S: DeviceSource> DeviceSource for CentralDifference<S> { fn encode(&self, p: &mut Vec<isomc_sdf_node>) { self.source.encode(p) } }
/// Sources whose `sample_normal` the device can evaluate: a CentralDifference, possibly inside Translate / Sampler.
pub trait DeviceNormals: DeviceSource { fn epsilon(&self) -> f32; }
impl<S: DeviceSource> DeviceNormals for CentralDifference<S> { fn epsilon(&self) -> f32 { self.epsilon } }
impl<S: DeviceNormals> DeviceNormals for Translate<S> { fn epsilon(&self) -> f32 { self.source.epsilon() } }
impl<'a, S: DeviceNormals> DeviceNormals for Sampler<'a, S> { fn epsilon(&self) -> f32 { self.source.epsilon() } }

impl<'a, S: DeviceSource> DeviceSource for Sampler<'a, S> { fn encode(&self, p: &mut Vec<isomc_sdf_node>) { self.source.encode(p) } }

/// reference src/extractor.rs:17-20
pub trait Extractor { fn extract_vertex(&mut self, v: [f32; 3]); fn extract_index(&mut self, index: usize); }

/// reference src/extractor.rs:72-93
pub struct IndexedVertices<'a> { vertices: &'a mut Vec<f32>, indices: &'a mut Vec<u32> }
impl<'a> IndexedVertices<'a> { pub fn new(vertices: &'a mut Vec<f32>, indices: &'a mut Vec<u32>) -> Self { Self { vertices, indices } } }
impl<'a> Extractor for IndexedVertices<'a> {
    fn extract_vertex(&mut self, v: [f32; 3]) { self.vertices.extend_from_slice(&v) }
    fn extract_index(&mut self, index: usize) { self.indices.push(index as u32) }
}

/// reference src/extractor.rs:95-127: x y z nx ny nz per vertex.  On this path the normals of `source` are sampled on the
/// device in one call after the extract instead of one `sample_normal` callback per vertex.
pub struct IndexedInterleavedNormals<'a, S: DeviceNormals> { vertices: &'a mut Vec<f32>, indices: &'a mut Vec<u32>, source: &'a S }
impl<'a, S: DeviceNormals> IndexedInterleavedNormals<'a, S> {
    pub fn new(vertices: &'a mut Vec<f32>, indices: &'a mut Vec<u32>, source: &'a S) -> Self { Self { vertices, indices, source } }
}

/// reference src/distance.rs:39-45: the distance kinds `MarchingCubes` is generic over.  `Signed` = one scalar distance per
/// sample; `Directed` = a signed distance along each cardinal axis (src/distance.rs:72-104), implicit sources only.
pub trait Distance {
    /// the C entry point that samples an implicit tree as this kind of distance
    #[doc(hidden)] unsafe fn extract_sdf(h: *mut isomc_t, prog: *const isomc_sdf_node, n_nodes: u32) -> i32;
    /// ... and the one behind `PointCloud<D>`
    #[doc(hidden)] unsafe fn points_sdf(h: *mut isomc_t, prog: *const isomc_sdf_node, n_nodes: u32) -> i32;
    /// ... and the one behind `BatchedMarchingCubes<D>`
    #[doc(hidden)] unsafe fn extract_sdf_batch(h: *mut isomc_t, progs: *const isomc_sdf_node, n_nodes: *const u32, n_chunks: u32) -> i32;
}
pub struct Signed;
pub struct Directed;
impl Distance for Signed {
    unsafe fn extract_sdf(h: *mut isomc_t, prog: *const isomc_sdf_node, n_nodes: u32) -> i32 { isomc_extract_sdf(h, prog, n_nodes) }
    unsafe fn points_sdf(h: *mut isomc_t, prog: *const isomc_sdf_node, n_nodes: u32) -> i32 { isomc_points_sdf(h, prog, n_nodes) }
    unsafe fn extract_sdf_batch(h: *mut isomc_t, progs: *const isomc_sdf_node, n_nodes: *const u32, n_chunks: u32) -> i32 {
        isomc_extract_sdf_batch(h, progs, n_nodes, n_chunks)
    }
}
impl Distance for Directed {
    unsafe fn extract_sdf(h: *mut isomc_t, prog: *const isomc_sdf_node, n_nodes: u32) -> i32 { isomc_extract_sdf_directed(h, prog, n_nodes) }
    unsafe fn points_sdf(h: *mut isomc_t, prog: *const isomc_sdf_node, n_nodes: u32) -> i32 { isomc_points_sdf_directed(h, prog, n_nodes) }
    unsafe fn extract_sdf_batch(h: *mut isomc_t, progs: *const isomc_sdf_node, n_nodes: *const u32, n_chunks: u32) -> i32 {
        isomc_extract_sdf_batch_directed(h, progs, n_nodes, n_chunks)
    }
}

/// `MarchingCubes<D: Distance>` as in the reference (src/marching_cubes.rs:38-43); `MarchingCubes::<Signed>::new(size)` and
/// `MarchingCubes::<Directed>::new(size)` are the two instantiations, and `extract` is ONE generic method.
pub struct MarchingCubes<D: Distance = Signed> { h: *mut isomc_t, size: usize, _d: std::marker::PhantomData<D> }

impl<D: Distance> MarchingCubes<D> {
    /// `MarchingCubes::new(size)`, reference src/marching_cubes.rs:46-50.  Panics without a CUDA device
    /// (the reference's signature has no error channel).
    pub fn new(size: usize) -> Self {
        let mut h = std::ptr::null_mut();
        let rc = unsafe { isomc_create(size as u32, 0, &mut h) };
        assert!(rc == 0, "isomc_create failed ({}): {:?}", rc, unsafe { std::ffi::CStr::from_ptr(isomc_last_error(std::ptr::null())) });
        Self { h, size, _d: std::marker::PhantomData }
    }

    /// `extract(&source, &mut extractor)`, reference src/marching_cubes.rs:59-82: the tree is sampled as `D` distances on the
    /// device (`Directed`: through the `VectorSource` side of the shapes, src/distance.rs:72-104).
    pub fn extract<S: DeviceSource, E: Extractor>(&mut self, source: &S, extractor: &mut E) {
        let mut prog = Vec::new();
        source.encode(&mut prog);
        self.check(unsafe { D::extract_sdf(self.h, prog.as_ptr(), prog.len() as u32) });
        self.deliver(extractor);
    }

    /// (a dense scalar lattice has no Directed distances: the library rejects it on a `MarchingCubes<Directed>` handle)
    pub fn extract_grid<E: Extractor>(&mut self, grid: &DenseGrid, extractor: &mut E) {
        assert_eq!(grid.size, self.size);
        assert_eq!(grid.data.len(), self.size * self.size * (self.size + 1));
        self.check(unsafe { isomc_extract_grid_host(self.h, grid.data.as_ptr()) });
        self.deliver(extractor);
    }

    /// Host lattice in, `IndexedVertices` out in one pipelined call: copy-in, kernels and copy-out overlap in z-chunks.
    /// The Vecs are grown to the capacity the previous extract needed; if this mesh is larger the result is fetched
    /// with `isomc_copy_out` after growing (the extraction itself is not repeated).
    pub fn extract_grid_into(&mut self, grid: &DenseGrid, sink: &mut IndexedVertices) {
        let (v0, i0) = (sink.vertices.len(), sink.indices.len());
        let (cv, ct) = ((sink.vertices.capacity() - v0) / 3, (sink.indices.capacity() - i0) / 3);
        sink.vertices.resize(v0 + 3 * cv, 0.0);
        sink.indices.resize(i0 + 3 * ct, 0);
        let rc = unsafe { isomc_extract_grid_host_to(self.h, grid.data.as_ptr(), sink.vertices.as_mut_ptr().add(v0), cv as u64,
                                                      sink.indices.as_mut_ptr().add(i0), ct as u64) };
        if rc != ERR_BUFFER_TOO_SMALL { self.check(rc); }
        let (mut nv, mut nt) = (0u64, 0u64);
        self.check(unsafe { isomc_counts(self.h, &mut nv, &mut nt, std::ptr::null_mut()) });
        sink.vertices.resize(v0 + 3 * nv as usize, 0.0);
        sink.indices.resize(i0 + 3 * nt as usize, 0);
        if rc == ERR_BUFFER_TOO_SMALL {
            self.check(unsafe { isomc_copy_out(self.h, sink.vertices.as_mut_ptr().add(v0), sink.indices.as_mut_ptr().add(i0)) });
        }
    }

    /// `extract(&sampler, &mut IndexedInterleavedNormals::new(&mut v, &mut i, &sampler))` (examples/sampler.rs:99-108)
    pub fn extract_with_normals<S: DeviceNormals>(&mut self, source: &S, sink: &mut IndexedInterleavedNormals<S>) {
        let mut prog = Vec::new();
        source.encode(&mut prog);
        self.check(unsafe { isomc_extract_sdf(self.h, prog.as_ptr(), prog.len() as u32) });
        let (mut nv, mut nt) = (0u64, 0u64);
        self.check(unsafe { isomc_counts(self.h, &mut nv, &mut nt, std::ptr::null_mut()) });
        let (v0, i0) = (sink.vertices.len(), sink.indices.len());
        sink.vertices.resize(v0 + 6 * nv as usize, 0.0);
        sink.indices.resize(i0 + 3 * nt as usize, 0);
        let mut nprog = Vec::new();
        sink.source.encode(&mut nprog);
        self.check(unsafe { isomc_copy_out_interleaved_normals(self.h, nprog.as_ptr(), nprog.len() as u32, sink.source.epsilon(),
                                                               sink.vertices.as_mut_ptr().add(v0), sink.indices.as_mut_ptr().add(i0)) });
    }

    /// # Safety: `d_grid` must be a device pointer to N*N*(N+1) f32 on the handle's device.
    pub unsafe fn extract_grid_device<E: Extractor>(&mut self, d_grid: *const f32, extractor: &mut E) {
        self.check(isomc_extract_grid_device(self.h, d_grid));
        self.deliver(extractor);
    }

    fn deliver<E: Extractor>(&mut self, extractor: &mut E) {
        let (mut nv, mut nt) = (0u64, 0u64);
        self.check(unsafe { isomc_counts(self.h, &mut nv, &mut nt, std::ptr::null_mut()) });
        let mut xyz = vec![0f32; 3 * nv as usize];
        let mut idx = vec![0u32; 3 * nt as usize];
        self.check(unsafe { isomc_copy_out(self.h, xyz.as_mut_ptr(), idx.as_mut_ptr()) });
        // protocol of marching_cubes.rs:81 / mesh.rs:91-100,240-251: every vertex first, then every index
        for v in xyz.chunks_exact(3) { extractor.extract_vertex([v[0], v[1], v[2]]) }
        for i in idx { extractor.extract_index(i as usize) }
    }

    fn check(&self, rc: i32) {
        assert!(rc == 0, "isomc error {}: {:?}", rc, unsafe { std::ffi::CStr::from_ptr(isomc_last_error(self.h)) });
    }
}

impl<D: Distance> Drop for MarchingCubes<D> { fn drop(&mut self) { unsafe { isomc_destroy(self.h); } } }

/// reference src/point_cloud.rs:28-63: `PointCloud<D: Distance>`, one vertex per active cell (midpoint of corners 0 and 6), no face data
pub struct PointCloud<D: Distance = Signed> { mc: MarchingCubes<D> }
impl<D: Distance> PointCloud<D> {
    pub fn new(size: usize) -> Self { Self { mc: MarchingCubes::<D>::new(size) } }
    pub fn extract<S: DeviceSource, E: Extractor>(&mut self, source: &S, extractor: &mut E) {
        let mut prog = Vec::new();
        source.encode(&mut prog);
        self.mc.check(unsafe { D::points_sdf(self.mc.h, prog.as_ptr(), prog.len() as u32) });
        self.mc.deliver(extractor);
    }
    /// (a dense scalar lattice has no Directed distances: `PointCloud::<Signed>` only in practice)
    pub fn extract_grid<E: Extractor>(&mut self, grid: &DenseGrid, extractor: &mut E) {
        assert_eq!(grid.size, self.mc.size);
        self.mc.check(unsafe { isomc_points_grid_host(self.mc.h, grid.data.as_ptr()) });
        self.mc.deliver(extractor);
    }
}

/// Many chunks per call (the crate's usage model is one `MarchingCubes::new(size).extract(..)` per chunk,
/// reference src/marching_cubes.rs:44-45, README.md:19): up to `n_chunks` trees through ONE kernel sequence and one size
/// read-back.  Chunk b is delivered to `extractors[b]` exactly as a single `extract` of `sources[b]` would.
pub struct BatchedMarchingCubes<D: Distance = Signed> { h: *mut isomc_t, n_chunks: usize, _d: std::marker::PhantomData<D> }
impl<D: Distance> BatchedMarchingCubes<D> {
    pub fn new(size: usize, n_chunks: usize) -> Self {
        let mut h = std::ptr::null_mut();
        let rc = unsafe { isomc_batch_create(size as u32, n_chunks as u32, 0, &mut h) };
        assert!(rc == 0, "isomc_batch_create failed ({}): {:?}", rc, unsafe { std::ffi::CStr::from_ptr(isomc_last_error(std::ptr::null())) });
        Self { h, n_chunks, _d: std::marker::PhantomData }
    }
    pub fn extract<S: DeviceSource, E: Extractor>(&mut self, sources: &[S], extractors: &mut [E]) {
        assert!(!sources.is_empty() && sources.len() <= self.n_chunks && sources.len() == extractors.len());
        let (mut flat, mut n_nodes) = (Vec::new(), Vec::new());
        for s in sources {
            let before = flat.len();
            s.encode(&mut flat);
            n_nodes.push((flat.len() - before) as u32);
        }
        self.check(unsafe { D::extract_sdf_batch(self.h, flat.as_ptr(), n_nodes.as_ptr(), n_nodes.len() as u32) });
        self.deliver(extractors);
    }
    fn check(&self, rc: i32) {
        assert!(rc == 0, "isomc error {}: {:?}", rc, unsafe { std::ffi::CStr::from_ptr(isomc_last_error(self.h)) });
    }
    fn deliver<E: Extractor>(&mut self, extractors: &mut [E]) {
        let check = |rc: i32| self.check(rc);
        let (mut nv, mut nt) = (0u64, 0u64);
        check(unsafe { isomc_counts(self.h, &mut nv, &mut nt, std::ptr::null_mut()) });
        let (mut xyz, mut idx) = (vec![0f32; 3 * nv as usize], vec![0u32; 3 * nt as usize]);
        check(unsafe { isomc_copy_out(self.h, xyz.as_mut_ptr(), idx.as_mut_ptr()) });
        let (mut vo, mut to) = (vec![0u64; self.n_chunks + 1], vec![0u64; self.n_chunks + 1]);
        check(unsafe { isomc_batch_offsets(self.h, vo.as_mut_ptr(), to.as_mut_ptr()) });
        for (b, ex) in extractors.iter_mut().enumerate() {
            for v in xyz[3 * vo[b] as usize..3 * vo[b + 1] as usize].chunks_exact(3) { ex.extract_vertex([v[0], v[1], v[2]]) }
            for &i in &idx[3 * to[b] as usize..3 * to[b + 1] as usize] { ex.extract_index(i as usize) }
        }
    }
}
/// a lattice holds scalars: dense chunks exist for `Signed` only
impl BatchedMarchingCubes<Signed> {
    /// Dense chunks (a voxel world cut into `size`^3 chunks): `lattices` holds `extractors.len()` lattices of
    /// `size * size * (size + 1)` samples back to back; chunk b goes to `extractors[b]`.
    pub fn extract_grids<E: Extractor>(&mut self, lattices: &[f32], extractors: &mut [E]) {
        let n = extractors.len();
        assert!(n >= 1 && n <= self.n_chunks && lattices.len() % n == 0);
        self.check(unsafe { isomc_extract_grid_batch_host(self.h, lattices.as_ptr(), n as u32) });
        self.deliver(extractors);
    }
}
impl<D: Distance> Drop for BatchedMarchingCubes<D> { fn drop(&mut self) { unsafe { isomc_destroy(self.h); } } }

/// One extract over several GPUs of the box: z-slabs, one exchange of 3 x u64 per rank on the extraction
/// streams (peer stores over NVLink when the devices are peers, else an NCCL all-gather inside the library), global ids
/// written directly.  The delivered mesh is the single-GPU (= reference) mesh.  A one-process-per-GPU host uses the
/// `isomc_slab_*` functions above instead and brings its own exchange (or `isomc_slab_connect_ipc`).
pub struct ShardedMarchingCubes { s: *mut isomc_sharded_t, n: usize }
impl ShardedMarchingCubes {
    pub fn new(size: usize, devices: &[i32]) -> Self {
        let mut s = std::ptr::null_mut();
        let rc = unsafe { isomc_sharded_create(size as u32, devices.len() as u32, devices.as_ptr(), &mut s) };
        assert!(rc == 0, "isomc_sharded_create failed ({}): {:?}", rc, unsafe { std::ffi::CStr::from_ptr(isomc_sharded_last_error(std::ptr::null())) });
        Self { s, n: devices.len() }
    }
    /// (first sample layer, number of sample layers) rank `r` must be given, on its device
    pub fn slab(&self, r: usize) -> (usize, usize) {
        let (mut first, mut count) = (0u32, 0u32);
        let rc = unsafe { isomc_sharded_slab(self.s, r as u32, std::ptr::null_mut(), std::ptr::null_mut(), &mut first, &mut count) };
        assert!(rc == 0);
        (first as usize, count as usize)
    }
    pub fn extract<S: DeviceSource, E: Extractor>(&mut self, source: &S, extractor: &mut E) {
        let mut prog = Vec::new();
        source.encode(&mut prog);
        self.check(unsafe { isomc_sharded_extract_sdf(self.s, prog.as_ptr(), prog.len() as u32) });
        self.deliver(extractor);
    }
    /// # Safety: `d_slabs[r]` must be a device pointer on rank r's device to the sample layers `slab(r)` names.
    pub unsafe fn extract_grid_device<E: Extractor>(&mut self, d_slabs: &[*const f32], extractor: &mut E) {
        assert_eq!(d_slabs.len(), self.n);
        self.check(isomc_sharded_extract_grid(self.s, d_slabs.as_ptr()));
        self.deliver(extractor);
    }
    fn deliver<E: Extractor>(&mut self, extractor: &mut E) {
        let (mut nv, mut nt) = (0u64, 0u64);
        self.check(unsafe { isomc_sharded_counts(self.s, &mut nv, &mut nt, std::ptr::null_mut()) });
        let (mut xyz, mut idx) = (vec![0f32; 3 * nv as usize], vec![0u32; 3 * nt as usize]);
        self.check(unsafe { isomc_sharded_copy_out(self.s, xyz.as_mut_ptr(), idx.as_mut_ptr()) });
        for v in xyz.chunks_exact(3) { extractor.extract_vertex([v[0], v[1], v[2]]) }
        for i in idx { extractor.extract_index(i as usize) }
    }
    fn check(&self, rc: i32) {
        assert!(rc == 0, "isomc error {}: {:?}", rc, unsafe { std::ffi::CStr::from_ptr(isomc_sharded_last_error(self.s)) });
    }
}
impl Drop for ShardedMarchingCubes { fn drop(&mut self) { unsafe { isomc_sharded_destroy(self.s); } } }
