"""isosurface_b200 -- B200-native (sm_100a) MarchingCubes extraction behind the API of the Rust
crate swiftcoder/isosurface.  The product is `libisomc_b200.so` (C ABI: include/isomc.h); this
package is the Python host-side mirror of the crate's interface for that one path:

    MarchingCubes(size).extract(Sampler(source), IndexedVertices(vertices, indices))
    PointCloud(size).extract(Sampler(source), OnlyVertices(vertices))

See DESIGN.md for the path, INTEGRATION.md for the Rust-side binding.
"""
from .extractor import ArrayMesh, Extractor, IndexedInterleavedNormals, IndexedVertices, OnlyVertices
from .chunks import BatchedMarchingCubes, ChunkedMarchingCubes
from .marching_cubes import MarchingCubes, PointCloud
from .source import (CentralDifference, Cylinder, DenseGrid, DeviceSource, Difference, Intersection, RectangularPrism, Sampler,
                     Sphere, Torus, Translate, Union)

__all__ = ["MarchingCubes", "PointCloud", "ChunkedMarchingCubes", "BatchedMarchingCubes", "Sampler", "DenseGrid", "DeviceSource", "Sphere", "Torus", "Cylinder",
           "RectangularPrism", "Union", "Intersection", "Difference", "Translate", "Extractor",
           "IndexedVertices", "OnlyVertices", "ArrayMesh", "IndexedInterleavedNormals", "CentralDifference"]
