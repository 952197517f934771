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
    explicit PointCloud(uint32_t size, int32_t device = 0) : MarchingCubes(size, device) {}
    template <class S> void extract(const SamplerT<S> &sampler, Extractor &extractor) { extract(sampler.source, extractor); }
    template <class S> void extract(const S &source, Extractor &extractor) {
        SdfProgram prog;
        source.encode(prog);
        check(isomc_points_sdf(h_, prog.data(), (uint32_t)prog.size()));
        deliver(extractor);
    }
    void extract(const DenseGrid &grid, Extractor &extractor) {
        if (grid.size != size_) throw Error(ISOMC_ERR_BAD_ARG, "grid size does not match");
        check(grid.on_device ? isomc_points_grid_device(h_, grid.data) : isomc_points_grid_host(h_, grid.data));
        deliver(extractor);
    }
};

}  // namespace isosurface
#endif
