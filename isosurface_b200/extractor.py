"""Extractors: the output side of the drop-in boundary (reference src/extractor.rs:17-127).

Protocol (reference src/marching_cubes.rs:81, src/mesh.rs:91-100,240-251): every `extract_vertex`
call, in vertex order, then every `extract_index` call, three per triangle, in triangle order.
`IndexedVertices`/`OnlyVertices` take a bulk fast path (one memcpy of each device buffer); any other
`Extractor` is replayed element by element with the same protocol.
"""
import numpy as np


class Extractor:
    def extract_vertex(self, vertex):  # vertex: (x, y, z)
        raise NotImplementedError

    def extract_index(self, index):
        raise NotImplementedError


class IndexedVertices(Extractor):
    """reference src/extractor.rs:72-93: xyz floats appended to `vertices`, u32 indices to `indices`"""

    def __init__(self, vertices, indices):
        self.vertices, self.indices = vertices, indices

    def extract_vertex(self, v):
        self.vertices.extend((v[0], v[1], v[2]))

    def extract_index(self, index):
        self.indices.append(index & 0xFFFFFFFF)  # `index as u32`

    def _bulk(self, xyz, idx):
        self.vertices.extend(xyz.tolist() if isinstance(self.vertices, list) else xyz)
        self.indices.extend(idx.tolist() if isinstance(self.indices, list) else idx)


class OnlyVertices(Extractor):
    """reference src/extractor.rs:24-43: vertices only, face data discarded"""

    def __init__(self, vertices):
        self.vertices = vertices

    def extract_vertex(self, v):
        self.vertices.extend((v[0], v[1], v[2]))

    def extract_index(self, index):
        pass

    def _bulk(self, xyz, idx):
        self.vertices.extend(xyz.tolist() if isinstance(self.vertices, list) else xyz)


class IndexedInterleavedNormals(Extractor):
    """reference src/extractor.rs:95-127: x y z nx ny nz per vertex, normals sampled from `source` (a HermiteSource).
    On the B200 path `source` must be (a Sampler / Translate around) a CentralDifference of an implicit tree; its
    normals are evaluated on the device at the extracted vertices (isomc_copy_out_interleaved_normals)."""

    def __init__(self, vertices, indices, source):
        from .source import locate_central_difference
        self.vertices, self.indices, self.source = vertices, indices, source
        self.central_difference, self.outer_translations = locate_central_difference(source)
        if self.central_difference is None:
            raise TypeError("IndexedInterleavedNormals needs a CentralDifference source on the device path "
                            "(analytic sample_normal implementations are host code in the reference and not a device path)")

    def extract_vertex(self, v):
        raise TypeError("IndexedInterleavedNormals is filled in bulk by MarchingCubes.extract")

    def extract_index(self, index):
        self.indices.append(index & 0xFFFFFFFF)

    def _bulk_normals(self, xyzn, idx):
        self.vertices.extend(xyzn.tolist() if isinstance(self.vertices, list) else xyzn)
        self.indices.extend(idx.tolist() if isinstance(self.indices, list) else idx)


class ArrayMesh(Extractor):
    """Convenience sink holding the mesh as NumPy arrays (`vertices` (V,3) f32, `indices` (T,3) u32)."""

    def __init__(self):
        self.vertices = np.zeros((0, 3), np.float32)
        self.indices = np.zeros((0, 3), np.uint32)

    def _bulk(self, xyz, idx):
        self.vertices = xyz.reshape(-1, 3)
        self.indices = idx.reshape(-1, 3)


def replay(extractor, xyz, idx):
    """Deliver a mesh through the Extractor protocol."""
    if hasattr(extractor, "_bulk"):
        extractor._bulk(xyz, idx)
        return
    for v in xyz.reshape(-1, 3):
        extractor.extract_vertex((float(v[0]), float(v[1]), float(v[2])))
    for i in idx:
        extractor.extract_index(int(i))
