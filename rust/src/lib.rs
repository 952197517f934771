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
    fn isomc_points_grid_host(h: *mut isomc_t, grid: *const f32) -> i32;
    fn isomc_counts(h: *mut isomc_t, nv: *mut u64, nt: *mut u64, na: *mut u64) -> i32;
    fn isomc_copy_out(h: *mut isomc_t, xyz: *mut f32, idx: *mut u32) -> i32;
    fn isomc_copy_out_interleaved_normals(h: *mut isomc_t, prog: *const isomc_sdf_node, n_nodes: u32, epsilon: f32,
                                          xyzn: *mut f32, idx: *mut u32) -> i32;
}
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

pub struct MarchingCubes { h: *mut isomc_t, size: usize }

impl MarchingCubes {
    /// `MarchingCubes::new(size)`, reference src/marching_cubes.rs:46-50.  Panics without a CUDA device
    /// (the reference's signature has no error channel).
    pub fn new(size: usize) -> Self {
        let mut h = std::ptr::null_mut();
        let rc = unsafe { isomc_create(size as u32, 0, &mut h) };
        assert!(rc == 0, "isomc_create failed ({}): {:?}", rc, unsafe { std::ffi::CStr::from_ptr(isomc_last_error(std::ptr::null())) });
        Self { h, size }
    }

    /// `extract(&source, &mut extractor)`, reference src/marching_cubes.rs:59-82.
    pub fn extract<S: DeviceSource, E: Extractor>(&mut self, source: &S, extractor: &mut E) {
        let mut prog = Vec::new();
        source.encode(&mut prog);
        self.check(unsafe { isomc_extract_sdf(self.h, prog.as_ptr(), prog.len() as u32) });
        self.deliver(extractor);
    }

    /// `MarchingCubes::<Directed>::extract` (reference src/distance.rs:72-104): in the crate this is the `D = Directed`
    /// instantiation of the same generic method; the tree is sampled through its `VectorSource` side on the device.
    pub fn extract_directed<S: DeviceSource, E: Extractor>(&mut self, source: &S, extractor: &mut E) {
        let mut prog = Vec::new();
        source.encode(&mut prog);
        self.check(unsafe { isomc_extract_sdf_directed(self.h, prog.as_ptr(), prog.len() as u32) });
        self.deliver(extractor);
    }

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

impl Drop for MarchingCubes { fn drop(&mut self) { unsafe { isomc_destroy(self.h); } } }

/// reference src/point_cloud.rs:33-63: one vertex per active cell (midpoint of corners 0 and 6), no face data
pub struct PointCloud { mc: MarchingCubes }
impl PointCloud {
    pub fn new(size: usize) -> Self { Self { mc: MarchingCubes::new(size) } }
    pub fn extract<S: DeviceSource, E: Extractor>(&mut self, source: &S, extractor: &mut E) {
        let mut prog = Vec::new();
        source.encode(&mut prog);
        self.mc.check(unsafe { isomc_points_sdf(self.mc.h, prog.as_ptr(), prog.len() as u32) });
        self.mc.deliver(extractor);
    }
    pub fn extract_grid<E: Extractor>(&mut self, grid: &DenseGrid, extractor: &mut E) {
        assert_eq!(grid.size, self.mc.size);
        self.mc.check(unsafe { isomc_points_grid_host(self.mc.h, grid.data.as_ptr()) });
        self.mc.deliver(extractor);
    }
}
