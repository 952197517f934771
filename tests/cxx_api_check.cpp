// compile-only check of include/isosurface.hpp (and a run when a GPU is present): mirrors
// benches/isosurface.rs:21-31 -- Torus(0.25, 0.1), MarchingCubes(128), IndexedVertices.
#include <cstdio>
#include "../include/isosurface.hpp"
int main() {
    using namespace isosurface;
    std::vector<float> vertices;
    std::vector<uint32_t> indices;
    try {
        Torus torus{0.25f, 0.1f};
        auto sampler = Sampler(torus);
        IndexedVertices extractor(vertices, indices);
        MarchingCubes mc(128);
        mc.extract(sampler, extractor);
        auto csg = Translate(0.5f, 0.5f, 0.5f, Union(Difference(Sphere{0.25f}, RectangularPrism{0.2f, 0.2f, 0.2f}), Cylinder{0.02f, 0.25f}));
        mc.extract(csg, extractor);
    } catch (const Error &e) {
        std::printf("%s\n", e.what());
        return e.code == ISOMC_ERR_CUDA ? 0 : 1;  // no device: loud failure is the expected behaviour
    }
    // the other extractors of the crate behind the same handle type: PointCloud, IndexedInterleavedNormals, host DenseGrid
    try {
        std::vector<float> pts, vn, hv;
        std::vector<uint32_t> ni, hi;
        auto src = Translate(0.5f, 0.5f, 0.5f, CentralDifference(Intersection(Sphere{0.3f}, RectangularPrism{0.2f, 0.2f, 0.2f})));
        PointCloud pc(64);
        OnlyVertices only(pts);
        pc.extract(Sampler(src), only);
        MarchingCubes mc(64);
        auto sink = IndexedInterleavedNormals(vn, ni, src);
        mc.extract(Sampler(src), sink);
        if (pts.size() / 3 != 4058 || vn.size() / 6 != 4056 || ni.size() / 3 != 8108) { // csgB @0.5, N=64
            std::printf("unexpected counts: points %zu, vertices %zu, triangles %zu\n", pts.size() / 3, vn.size() / 6, ni.size() / 3);
            return 1;
        }
        std::vector<float> dv; std::vector<uint32_t> di;
        MarchingCubes dmc(64, 0, Distance::Directed);   // MarchingCubes::<Directed>
        IndexedVertices dsink(dv, di);
        dmc.extract(Sampler(src), dsink);
        if (di.size() / 3 != ni.size() / 3) { std::printf("Directed: %zu triangles, Signed %zu\n", di.size() / 3, ni.size() / 3); return 1; }
        std::vector<float> grid((size_t)64 * 64 * 65);
        for (size_t i = 0; i < grid.size(); ++i) grid[i] = (float)((i * 2654435761u >> 7) % 1000) - 500.0f;
        IndexedVertices hsink(hv, hi);
        for (int rep = 0; rep < 2; ++rep) { // second call takes the streamed host-to-host path
            hv.clear(); hi.clear();
            mc.extract(DenseGrid{grid.data(), 64, false}, hsink);
        }
        if (hv.empty() || hi.empty()) { std::printf("host grid extract produced nothing\n"); return 1; }
        // many chunks through one kernel sequence: every chunk = the single-chunk mesh
        {
            std::vector<TranslateT<Sphere>> chunks;
            for (int b = 0; b < 5; ++b) chunks.push_back(Translate(0.3f + 0.1f * b, 0.5f, 0.5f, Sphere{0.2f}));
            std::vector<std::vector<float>> bv(5);
            std::vector<std::vector<uint32_t>> bi(5);
            std::vector<IndexedVertices> sinks;
            for (int b = 0; b < 5; ++b) sinks.emplace_back(bv[b], bi[b]);
            std::vector<Extractor *> ptrs;
            for (auto &sk : sinks) ptrs.push_back(&sk);
            BatchedMarchingCubes batch(32, 8);
            batch.extract(chunks, ptrs);
            MarchingCubes one(32);
            for (int b = 0; b < 5; ++b) {
                std::vector<float> ov; std::vector<uint32_t> oi;
                IndexedVertices os(ov, oi);
                one.extract(chunks[b], os);
                if (ov != bv[b] || oi != bi[b]) { std::printf("batched chunk %d differs from the single extract\n", b); return 1; }
            }
        }
        // MarchingCubes<Directed> per chunk: chunk b = the single Directed extract of tree b
        {
            std::vector<TranslateT<Sphere>> chunks;
            for (int b = 0; b < 3; ++b) chunks.push_back(Translate(0.4f + 0.1f * b, 0.5f, 0.5f, Sphere{0.25f}));
            std::vector<std::vector<float>> bv(3);
            std::vector<std::vector<uint32_t>> bi(3);
            std::vector<IndexedVertices> sinks;
            for (int b = 0; b < 3; ++b) sinks.emplace_back(bv[b], bi[b]);
            std::vector<Extractor *> ptrs;
            for (auto &sk : sinks) ptrs.push_back(&sk);
            BatchedMarchingCubes batch(32, 4, 0, Distance::Directed);
            batch.extract(chunks, ptrs);
            MarchingCubes one(32, 0, Distance::Directed);
            for (int b = 0; b < 3; ++b) {
                std::vector<float> ov; std::vector<uint32_t> oi;
                IndexedVertices os(ov, oi);
                one.extract(chunks[b], os);
                if (ov != bv[b] || oi != bi[b]) { std::printf("batched Directed chunk %d differs from the single extract\n", b); return 1; }
            }
        }
        // dense chunks through the same batch handle: chunk b = the host-grid extract of lattice b
        {
            const uint32_t n = 24, B = 3;
            const size_t per = (size_t)n * n * (n + 1);
            std::vector<float> lat(per * B);
            for (size_t i = 0; i < lat.size(); ++i) lat[i] = (float)((i * 2246822519u >> 9) % 1000) - 480.0f;
            std::vector<std::vector<float>> bv(B);
            std::vector<std::vector<uint32_t>> bi(B);
            std::vector<IndexedVertices> sinks;
            for (uint32_t b = 0; b < B; ++b) sinks.emplace_back(bv[b], bi[b]);
            std::vector<Extractor *> ptrs;
            for (auto &sk : sinks) ptrs.push_back(&sk);
            BatchedMarchingCubes batch(n, B);
            batch.extract_grids(lat.data(), B, ptrs);
            MarchingCubes one(n);
            for (uint32_t b = 0; b < B; ++b) {
                std::vector<float> ov; std::vector<uint32_t> oi;
                IndexedVertices os(ov, oi);
                one.extract(DenseGrid{lat.data() + per * b, n, false}, os);
                if (ov != bv[b] || oi != bi[b]) { std::printf("batched dense chunk %u differs from the single extract\n", b); return 1; }
            }
        }
        // the same extract over z-slabs (all on device 0 here; distinct devices exchange over NVLink): identical mesh
        {
            std::vector<float> sv, ov; std::vector<uint32_t> si, oi;
            IndexedVertices ss(sv, si), os(ov, oi);
            const int nd = isomc_device_count();
            std::vector<int32_t> devs = {0, nd > 1 ? 1 : 0, nd > 2 ? 2 : 0};   // distinct devices when the box has them
            ShardedMarchingCubes sharded(64, devs);
            if (nd > 2 && !sharded.uses_peer_memory() && !sharded.uses_nccl()) { std::printf("no inter-device exchange chosen\n"); return 1; }
            sharded.extract(Sampler(src), ss);
            MarchingCubes one(64);
            one.extract(Sampler(src), os);
            if (sv != ov || si != oi) { std::printf("sharded mesh differs from the single-GPU mesh\n"); return 1; }
        }
    } catch (const Error &e) {
        std::printf("%s\n", e.what());
        return 1;
    }
    std::printf("V=%zu T=%zu\n", vertices.size() / 3, indices.size() / 3);
    return 0;
}
