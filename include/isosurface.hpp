// isosurface.hpp -- C++ host-side mirror of the swiftcoder/isosurface crate API for the one path this
// repo accelerates, layered over the C ABI in isomc.h (header-only; link libisomc_b200.so).
//
//   Rust (reference)                                          C++ (this header)
//   ---------------------------------------------------------------------------------------------
//   MarchingCubes::<Signed>::new(size)                        isosurface::MarchingCubes mc(size);
//   MarchingCubes::<Directed>::new(size)                      isosurface::MarchingCubes mc(size, 0, isosurface::Distance::Directed);
//   mc.extract(&Sampler::new(&source), &mut extractor)        mc.extract(isosurface::Sampler(source), extractor);
//   implicit::{Sphere,Torus,Cylinder,RectangularPrism}        isosurface::Sphere{r}, Torus{R,r}, Cylinder{r,h}, RectangularPrism{hx,hy,hz}
//   implicit::{Union,Intersection,Difference}                 isosurface::Union(a,b), Intersection(a,b), Difference(a,b)
//   examples/common/sources.rs DemoSource (p - 0.5)           isosurface::Translate(dx,dy,dz, child)
//   extractor::IndexedVertices::new(&mut v, &mut i)           isosurface::IndexedVertices sink(v, i);
//   trait Extractor { extract_vertex; extract_index }         struct Extractor (virtual), replayed in protocol order
//   PointCloud::<Signed>::new(size).extract(..)               isosurface::PointCloud pc(size); pc.extract(source, extractor);
//   source::CentralDifference::new(source)                    isosurface::CentralDifference(source [, epsilon])
//   extractor::IndexedInterleavedNormals::new(&mut v,&mut i,&s)   isosurface::IndexedInterleavedNormals sink(v, i, source);
//
// Reference files: src/marching_cubes.rs:38-82, src/sampler.rs:26-41, src/source.rs:21-28,
// src/extractor.rs:17-127, src/point_cloud.rs:33-63, src/source.rs:52-94, src/implicit/*.rs.  Sources are *encoded* for the device (DeviceSource);
// an arbitrary callable is not a device path and does not compile against this API (no CPU fallback).
#ifndef ISOSURFACE_HPP
#define ISOSURFACE_HPP

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "isomc.h"

namespace isosurface {

struct Error : std::runtime_error {
    int32_t code;
    Error(int32_t c, const std::string &m) : std::runtime_error("isomc error " + std::to_string(c) + ": " + m), code(c) {}
};

using SdfProgram = std::vector<isomc_sdf_node>;

// ---- implicit sources (reference src/implicit/) ---------------------------------------------
struct Sphere { float radius; void encode(SdfProgram &p) const { p.push_back({ISOMC_SDF_SPHERE, radius, 0, 0}); } };
struct Torus { float radius, tube_radius; void encode(SdfProgram &p) const { p.push_back({ISOMC_SDF_TORUS, radius, tube_radius, 0}); } };
struct Cylinder { float radius, half_length; void encode(SdfProgram &p) const { p.push_back({ISOMC_SDF_CYLINDER, radius, half_length, 0}); } };
struct RectangularPrism { float hx, hy, hz; void encode(SdfProgram &p) const { p.push_back({ISOMC_SDF_PRISM, hx, hy, hz}); } };

template <class A, class B, uint32_t OP>
struct Binary {
    A a; B b;
    Binary(A a_, B b_) : a(a_), b(b_) {}
    void encode(SdfProgram &p) const { a.encode(p); b.encode(p); p.push_back({OP, 0, 0, 0}); }
};
template <class A, class B> using UnionT = Binary<A, B, ISOMC_SDF_UNION>;                 // csg.rs:20-39   min(a, b)
template <class A, class B> using IntersectionT = Binary<A, B, ISOMC_SDF_INTERSECTION>;   // csg.rs:54-72   max(a, b)
template <class A, class B> using DifferenceT = Binary<A, B, ISOMC_SDF_DIFFERENCE>;       // csg.rs:82-100  max(b, -a)
template <class A, class B> UnionT<A, B> Union(A a, B b) { return {a, b}; }
template <class A, class B> IntersectionT<A, B> Intersection(A a, B b) { return {a, b}; }
template <class A, class B> DifferenceT<A, B> Difference(A a, B b) { return {a, b}; }

template <class S>
struct TranslateT {  // q = p - (dx,dy,dz), examples/common/sources.rs:38-43
    float dx, dy, dz; S child;
    void encode(SdfProgram &p) const {
        p.push_back({ISOMC_SDF_TRANSLATE_PUSH, dx, dy, dz});
        child.encode(p);
        p.push_back({ISOMC_SDF_TRANSLATE_POP, 0, 0, 0});
    }
};
template <class S> TranslateT<S> Translate(float dx, float dy, float dz, S child) { return {dx, dy, dz, child}; }

// source.rs:52-94: a ScalarSource with central-difference normals; transparent as a scalar source
template <class S>
struct CentralDifferenceT {
    S source; float epsilon;
    void encode(SdfProgram &p) const { source.encode(p); }
};
template <class S> CentralDifferenceT<S> CentralDifference(S source, float epsilon = 0.000001f) { return {source, epsilon}; }
// outer = Translate wrappers OUTSIDE the adaptor: Translate(o, CentralDifference(T)) differentiates f at v - o, while
// CentralDifference(Translate(o, T)) differentiates the translated function; both encode to the same program
template <class S> struct has_central_difference { static constexpr bool value = false; static constexpr uint32_t outer = 0; static float epsilon(const S &) { return 0; } };
template <class S> struct has_central_difference<CentralDifferenceT<S>> {
    static constexpr bool value = true;
    static constexpr uint32_t outer = 0;
    static float epsilon(const CentralDifferenceT<S> &s) { return s.epsilon; }
};
template <class S> struct has_central_difference<TranslateT<S>> {
    static constexpr bool value = has_central_difference<S>::value;
    static constexpr uint32_t outer = has_central_difference<S>::outer + 1;
    static float epsilon(const TranslateT<S> &s) { return has_central_difference<S>::epsilon(s.child); }
};

// Dense lattice source (new): N*N*(N+1) f32, x fastest; host or device memory.
struct DenseGrid { const float *data; uint32_t size; bool on_device; };

// sampler.rs:26-41
template <class S> struct SamplerT { const S &source; };
template <class S> SamplerT<S> Sampler(const S &s) { return {s}; }

// ---- extractors (reference src/extractor.rs) -------------------------------------------------
struct Extractor {
    virtual ~Extractor() = default;
    virtual void extract_vertex(float x, float y, float z) = 0;
    virtual void extract_index(size_t index) = 0;
};
struct IndexedVertices : Extractor {  // extractor.rs:72-93
    std::vector<float> &vertices; std::vector<uint32_t> &indices;
    IndexedVertices(std::vector<float> &v, std::vector<uint32_t> &i) : vertices(v), indices(i) {}
    void extract_vertex(float x, float y, float z) override { vertices.push_back(x); vertices.push_back(y); vertices.push_back(z); }
    void extract_index(size_t index) override { indices.push_back((uint32_t)index); }
};
struct OnlyVertices : Extractor {  // extractor.rs:24-43
    std::vector<float> &vertices;
    explicit OnlyVertices(std::vector<float> &v) : vertices(v) {}
    void extract_vertex(float x, float y, float z) override { vertices.push_back(x); vertices.push_back(y); vertices.push_back(z); }
    void extract_index(size_t) override {}
};

// extractor.rs:95-127: x y z nx ny nz per vertex; the normals of a CentralDifference source are sampled on the device
struct InterleavedNormalsSink {
    std::vector<float> &vertices; std::vector<uint32_t> &indices; SdfProgram program; float epsilon; uint32_t outer_translations;
};
template <class S>
InterleavedNormalsSink IndexedInterleavedNormals(std::vector<float> &v, std::vector<uint32_t> &i, const S &source) {
    static_assert(has_central_difference<S>::value, "IndexedInterleavedNormals needs a CentralDifference source on the device path");
    InterleavedNormalsSink sink{v, i, {}, has_central_difference<S>::epsilon(source), has_central_difference<S>::outer};
    source.encode(sink.program);
    return sink;
}

// ---- MarchingCubes (reference src/marching_cubes.rs:38-82) ------------------------------------
// distance.rs:39-45: Signed = one scalar distance; Directed = a signed distance along each cardinal axis (implicit sources only)
enum class Distance { Signed, Directed };

class MarchingCubes {
  public:
    explicit MarchingCubes(uint32_t size, int32_t device = 0, Distance distance = Distance::Signed) : size_(size), distance_(distance) {
        int32_t rc = isomc_create(size, device, &h_);
        if (rc) throw Error(rc, isomc_last_error(nullptr));
    }
    ~MarchingCubes() { isomc_destroy(h_); }
    MarchingCubes(const MarchingCubes &) = delete;
    MarchingCubes &operator=(const MarchingCubes &) = delete;

    template <class S> void extract(const SamplerT<S> &sampler, Extractor &extractor) { extract(sampler.source, extractor); }
    template <class S> void extract(const S &source, Extractor &extractor) {
        SdfProgram prog;
        source.encode(prog);
        check(distance_ == Distance::Directed ? isomc_extract_sdf_directed(h_, prog.data(), (uint32_t)prog.size())
                                              : isomc_extract_sdf(h_, prog.data(), (uint32_t)prog.size()));
        deliver(extractor);
    }
    void extract(const DenseGrid &grid, Extractor &extractor) {
        if (grid.size != size_) throw Error(ISOMC_ERR_BAD_ARG, "grid size does not match");
        if (distance_ == Distance::Directed) throw Error(ISOMC_ERR_UNSUPPORTED_SOURCE, "a dense scalar lattice has no Directed distances");
        if (!grid.on_device) {
            // host lattice -> host mesh in one pipelined call (copy-in, kernels and copy-out overlap), straight into the
            // Vecs of an IndexedVertices whose capacity is kept from the previous extract
            if (auto *iv = dynamic_cast<IndexedVertices *>(&extractor)) {
                const size_t v0 = iv->vertices.size(), i0 = iv->indices.size();
                iv->vertices.resize(v0 + 3 * (last_v_ + last_v_ / 8 + 1024));
                iv->indices.resize(i0 + 3 * (last_t_ + last_t_ / 8 + 1024));
                int32_t rc = isomc_extract_grid_host_to(h_, grid.data, iv->vertices.data() + v0, (iv->vertices.size() - v0) / 3,
                                                        iv->indices.data() + i0, (iv->indices.size() - i0) / 3);
                if (rc && rc != ISOMC_ERR_BUFFER_TOO_SMALL) check(rc);
                check(isomc_counts(h_, &last_v_, &last_t_, nullptr));
                iv->vertices.resize(v0 + 3 * last_v_);
                iv->indices.resize(i0 + 3 * last_t_);
                if (rc == ISOMC_ERR_BUFFER_TOO_SMALL) check(isomc_copy_out(h_, iv->vertices.data() + v0, iv->indices.data() + i0));
                return;
            }
            check(isomc_extract_grid_host(h_, grid.data));
        } else {
            check(isomc_extract_grid_device(h_, grid.data));
        }
        deliver(extractor);
    }
    // IndexedInterleavedNormals: extract, then positions + central-difference normals from the device
    template <class S> void extract(const S &source, InterleavedNormalsSink &sink) {
        SdfProgram prog;
        source.encode(prog);
        check(isomc_extract_sdf(h_, prog.data(), (uint32_t)prog.size()));
        uint64_t nv = 0, nt = 0;
        check(isomc_counts(h_, &nv, &nt, nullptr));
        const size_t v0 = sink.vertices.size(), i0 = sink.indices.size();
        sink.vertices.resize(v0 + 6 * nv);
        sink.indices.resize(i0 + 3 * nt);
        check(isomc_copy_out_interleaved_normals_at(h_, sink.program.data(), (uint32_t)sink.program.size(), sink.epsilon,
                                                    sink.outer_translations, sink.vertices.data() + v0, sink.indices.data() + i0));
    }
    template <class S> void extract(const SamplerT<S> &sampler, InterleavedNormalsSink &sink) { extract(sampler.source, sink); }
    isomc_t *handle() { return h_; }

  protected:
    void check(int32_t rc) { if (rc) throw Error(rc, isomc_last_error(h_)); }
    void deliver(Extractor &ex) {
        uint64_t nv = 0, nt = 0;
        check(isomc_counts(h_, &nv, &nt, nullptr));
        if (auto *iv = dynamic_cast<IndexedVertices *>(&ex)) {  // bulk fast path: two memcpys
            size_t v0 = iv->vertices.size(), i0 = iv->indices.size();
            iv->vertices.resize(v0 + 3 * nv);
            iv->indices.resize(i0 + 3 * nt);
            check(isomc_copy_out(h_, iv->vertices.data() + v0, iv->indices.data() + i0));
            return;
        }
        std::vector<float> xyz(3 * nv);
        std::vector<uint32_t> idx(3 * nt);
        check(isomc_copy_out(h_, xyz.data(), idx.data()));
        for (uint64_t v = 0; v < nv; ++v) ex.extract_vertex(xyz[3 * v], xyz[3 * v + 1], xyz[3 * v + 2]);  // all vertices first ...
        for (uint64_t i = 0; i < 3 * nt; ++i) ex.extract_index(idx[i]);                                 // ... then all indices
    }
    isomc_t *h_ = nullptr;
    uint32_t size_;
    Distance distance_ = Distance::Signed;
    uint64_t last_v_ = 0, last_t_ = 0;
};

// ---- PointCloud (reference src/point_cloud.rs:33-63): one vertex per active cell, no face data ---------------
class PointCloud : public MarchingCubes {
  public:
    explicit PointCloud(uint32_t size, int32_t device = 0, Distance distance = Distance::Signed) : MarchingCubes(size, device, distance) {}
    template <class S> void extract(const SamplerT<S> &sampler, Extractor &extractor) { extract(sampler.source, extractor); }
    template <class S> void extract(const S &source, Extractor &extractor) {
        SdfProgram prog;
        source.encode(prog);
        check(distance_ == Distance::Directed ? isomc_points_sdf_directed(h_, prog.data(), (uint32_t)prog.size())
                                              : isomc_points_sdf(h_, prog.data(), (uint32_t)prog.size()));
        deliver(extractor);
    }
    void extract(const DenseGrid &grid, Extractor &extractor) {
        if (grid.size != size_) throw Error(ISOMC_ERR_BAD_ARG, "grid size does not match");
        if (distance_ == Distance::Directed) throw Error(ISOMC_ERR_UNSUPPORTED_SOURCE, "a dense scalar lattice has no Directed distances");
        check(grid.on_device ? isomc_points_grid_device(h_, grid.data) : isomc_points_grid_host(h_, grid.data));
        deliver(extractor);
    }
};

// ---- many chunks per call (the crate's usage model: one `MarchingCubes::new(size).extract(..)` per chunk,
// reference src/marching_cubes.rs:44-45, README.md:19).  Up to `n_chunks` implicit trees go through ONE kernel sequence and one
// size read-back; chunk b of a batch is delivered to extractors[b] exactly as MarchingCubes(size).extract(sources[b], ..) would.
class BatchedMarchingCubes {
  public:
    BatchedMarchingCubes(uint32_t size, uint32_t n_chunks, int32_t device = 0, Distance distance = Distance::Signed)
        : n_chunks_(n_chunks), distance_(distance) {
        int32_t rc = isomc_batch_create(size, n_chunks, device, &h_);
        if (rc) throw Error(rc, isomc_last_error(nullptr));
    }
    ~BatchedMarchingCubes() { isomc_destroy(h_); }
    BatchedMarchingCubes(const BatchedMarchingCubes &) = delete;
    BatchedMarchingCubes &operator=(const BatchedMarchingCubes &) = delete;
    uint32_t capacity() const { return n_chunks_; }

    // sources: any container of device sources of ONE type (heterogeneous batches: encode() the programs yourself and call extract_programs)
    template <class Sources> void extract(const Sources &sources, const std::vector<Extractor *> &extractors) {
        std::vector<isomc_sdf_node> flat;
        std::vector<uint32_t> n_nodes;
        for (const auto &src : sources) {
            SdfProgram prog;
            src.encode(prog);
            flat.insert(flat.end(), prog.begin(), prog.end());
            n_nodes.push_back((uint32_t)prog.size());
        }
        extract_programs(flat, n_nodes, extractors);
    }
    void extract_programs(const std::vector<isomc_sdf_node> &flat, const std::vector<uint32_t> &n_nodes, const std::vector<Extractor *> &extractors) {
        const uint32_t b = (uint32_t)n_nodes.size();
        if (b < 1 || b > n_chunks_ || extractors.size() != b) throw Error(ISOMC_ERR_BAD_ARG, "batch size / extractor count mismatch");
        check(distance_ == Distance::Directed ? isomc_extract_sdf_batch_directed(h_, flat.data(), n_nodes.data(), b)
                                              : isomc_extract_sdf_batch(h_, flat.data(), n_nodes.data(), b));
        deliver(b, extractors);
    }
    // dense chunks: `n` host lattices of size * size * (size + 1) floats back to back (a voxel world cut into chunks)
    void extract_grids(const float *lattices, uint32_t n, const std::vector<Extractor *> &extractors) {
        if (n < 1 || n > n_chunks_ || extractors.size() != n) throw Error(ISOMC_ERR_BAD_ARG, "batch size / extractor count mismatch");
        if (distance_ == Distance::Directed) throw Error(ISOMC_ERR_BAD_ARG, "a lattice of scalars has no Directed distances");
        check(isomc_extract_grid_batch_host(h_, lattices, n));
        deliver(n, extractors);
    }
    // the same with the handle's whole capacity of lattices already in device memory (used in place)
    void extract_device_grids(const float *d_lattices, const std::vector<Extractor *> &extractors) {
        if (extractors.size() != n_chunks_) throw Error(ISOMC_ERR_BAD_ARG, "a device batch fills the handle: one extractor per chunk");
        if (distance_ == Distance::Directed) throw Error(ISOMC_ERR_BAD_ARG, "a lattice of scalars has no Directed distances");
        check(isomc_extract_grid_batch_device(h_, d_lattices, n_chunks_));
        deliver(n_chunks_, extractors);
    }

  private:
    void deliver(uint32_t b, const std::vector<Extractor *> &extractors) {
        uint64_t nv = 0, nt = 0;
        check(isomc_counts(h_, &nv, &nt, nullptr));
        std::vector<float> xyz(3 * nv);
        std::vector<uint32_t> idx(3 * nt);
        check(isomc_copy_out(h_, xyz.data(), idx.data()));
        std::vector<uint64_t> vo(n_chunks_ + 1), to(n_chunks_ + 1);
        check(isomc_batch_offsets(h_, vo.data(), to.data()));
        for (uint32_t c = 0; c < b; ++c) {
            Extractor &ex = *extractors[c];
            if (auto *iv = dynamic_cast<IndexedVertices *>(&ex)) {
                iv->vertices.insert(iv->vertices.end(), xyz.begin() + 3 * vo[c], xyz.begin() + 3 * vo[c + 1]);
                iv->indices.insert(iv->indices.end(), idx.begin() + 3 * to[c], idx.begin() + 3 * to[c + 1]);
                continue;
            }
            for (uint64_t v = vo[c]; v < vo[c + 1]; ++v) ex.extract_vertex(xyz[3 * v], xyz[3 * v + 1], xyz[3 * v + 2]);
            for (uint64_t i = 3 * to[c]; i < 3 * to[c + 1]; ++i) ex.extract_index(idx[i]);
        }
    }
    void check(int32_t rc) { if (rc) throw Error(rc, isomc_last_error(h_)); }
    isomc_t *h_ = nullptr;
    uint32_t n_chunks_;
    Distance distance_;
};

// ---- one extract over several GPUs of the box: z-slabs, one exchange of 3 x u64 per rank (peer stores over
// NVLink, or an NCCL all-gather), global ids written directly; the delivered mesh is the single-GPU (= reference) mesh.
class ShardedMarchingCubes {
  public:
    ShardedMarchingCubes(uint32_t size, const std::vector<int32_t> &devices) : size_(size), n_((uint32_t)devices.size()) {
        int32_t rc = isomc_sharded_create(size, n_, devices.data(), &s_);
        if (rc) throw Error(rc, isomc_sharded_last_error(nullptr));
    }
    ~ShardedMarchingCubes() { isomc_sharded_destroy(s_); }
    ShardedMarchingCubes(const ShardedMarchingCubes &) = delete;
    ShardedMarchingCubes &operator=(const ShardedMarchingCubes &) = delete;
    bool uses_peer_memory() const { return isomc_sharded_uses_peer_memory(s_) != 0; }
    bool uses_nccl() const { return isomc_sharded_uses_nccl(s_) != 0; }
    uint32_t ranks() const { return n_; }
    // rank r must be given sample layers [first, first + count) of the size x size x (size+1) lattice, on devices[r]
    void slab(uint32_t rank, uint32_t &first_sample_layer, uint32_t &n_sample_layers) const {
        if (isomc_sharded_slab(s_, rank, nullptr, nullptr, &first_sample_layer, &n_sample_layers)) throw Error(ISOMC_ERR_BAD_ARG, "bad rank");
    }
    template <class S> void extract(const SamplerT<S> &sampler, Extractor &extractor) { extract(sampler.source, extractor); }
    template <class S> void extract(const S &source, Extractor &extractor) {
        SdfProgram prog;
        source.encode(prog);
        check(isomc_sharded_extract_sdf(s_, prog.data(), (uint32_t)prog.size()));
        deliver(extractor);
    }
    // d_slabs[r]: device pointer on devices[r] to rank r's sample layers
    void extract(const std::vector<const float *> &d_slabs, Extractor &extractor) {
        if (d_slabs.size() != n_) throw Error(ISOMC_ERR_BAD_ARG, "one slab pointer per rank");
        check(isomc_sharded_extract_grid(s_, d_slabs.data()));
        deliver(extractor);
    }

  private:
    void check(int32_t rc) { if (rc) throw Error(rc, isomc_sharded_last_error(s_)); }
    void deliver(Extractor &ex) {
        uint64_t nv = 0, nt = 0;
        check(isomc_sharded_counts(s_, &nv, &nt, nullptr));
        std::vector<float> xyz(3 * nv);
        std::vector<uint32_t> idx(3 * nt);
        check(isomc_sharded_copy_out(s_, xyz.data(), idx.data()));
        for (uint64_t v = 0; v < nv; ++v) ex.extract_vertex(xyz[3 * v], xyz[3 * v + 1], xyz[3 * v + 2]);
        for (uint64_t i = 0; i < 3 * nt; ++i) ex.extract_index(idx[i]);
    }
    isomc_sharded_t *s_ = nullptr;
    uint32_t size_, n_;
};

}  // namespace isosurface
#endif
