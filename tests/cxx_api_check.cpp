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
    std::printf("V=%zu T=%zu\n", vertices.size() / 3, indices.size() / 3);
    return 0;
}
