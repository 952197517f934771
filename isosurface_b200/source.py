"""Sources: the input side of the drop-in boundary.

Mirrors the reference's sampling contract (src/source.rs:21-28 `ScalarSource`, src/sampler.rs:26-41
`Sampler`) and its implicit shapes (src/implicit/{sphere,torus,cylinder,rectangular_prism,csg}.rs).
On the B200 path a source is not called back per sample; it *encodes* itself as a device program
(`DeviceSource.encode`) or hands over a dense lattice.  An arbitrary Python callable is not a device
path and is rejected -- there is no CPU fallback.
"""
import numpy as np

from . import _lib


class DeviceSource:
    """A ScalarSource that can be evaluated on the device."""

    def encode(self, prog):  # pragma: no cover - interface
        raise NotImplementedError


class Sphere(DeviceSource):
    """reference src/implicit/sphere.rs:24-39"""

    def __init__(self, radius):
        self.radius = float(radius)

    def encode(self, prog):
        prog.append((_lib.SDF_SPHERE, self.radius, 0.0, 0.0))


class Torus(DeviceSource):
    """reference src/implicit/torus.rs:24-46"""

    def __init__(self, radius, tube_radius):
        self.radius, self.tube_radius = float(radius), float(tube_radius)

    def encode(self, prog):
        prog.append((_lib.SDF_TORUS, self.radius, self.tube_radius, 0.0))


class Cylinder(DeviceSource):
    """reference src/implicit/cylinder.rs:23-48"""

    def __init__(self, radius, half_length):
        self.radius, self.half_length = float(radius), float(half_length)

    def encode(self, prog):
        prog.append((_lib.SDF_CYLINDER, self.radius, self.half_length, 0.0))


class RectangularPrism(DeviceSource):
    """reference src/implicit/rectangular_prism.rs:24-41; half_extent is an (x, y, z) triple"""

    def __init__(self, half_extent):
        self.half_extent = tuple(float(v) for v in half_extent)
        if len(self.half_extent) != 3:
            raise ValueError("half_extent must have 3 components")

    def encode(self, prog):
        prog.append((_lib.SDF_PRISM,) + self.half_extent)


class _Binary(DeviceSource):
    OP = None

    def __init__(self, a, b):
        _require_device_source(a)
        _require_device_source(b)
        self.a, self.b = a, b

    def encode(self, prog):
        self.a.encode(prog)
        self.b.encode(prog)
        prog.append((self.OP, 0.0, 0.0, 0.0))


class Union(_Binary):
    """reference src/implicit/csg.rs:20-39: min(a, b)"""
    OP = _lib.SDF_UNION


class Intersection(_Binary):
    """reference src/implicit/csg.rs:54-72: max(a, b)"""
    OP = _lib.SDF_INTERSECTION


class Difference(_Binary):
    """reference src/implicit/csg.rs:82-100: max(b, -a) -- solid where b is and a is not"""
    OP = _lib.SDF_DIFFERENCE


class Translate(DeviceSource):
    """q = p - offset, then the child (reference examples/common/sources.rs:38-43 uses offset 0.5)"""

    def __init__(self, offset, source):
        _require_device_source(source)
        self.offset = tuple(float(v) for v in (offset if np.ndim(offset) else (offset,) * 3))
        self.source = source

    def encode(self, prog):
        prog.append((_lib.SDF_TRANSLATE_PUSH,) + self.offset)
        self.source.encode(prog)
        prog.append((_lib.SDF_TRANSLATE_POP, 0.0, 0.0, 0.0))


class DenseGrid:
    """A dense f32 lattice source: N x N x (N+1) samples, x fastest (shape (N+1, N, N) as an array).

    New with the B200 path (the reference has no grid source).  The value at lattice point (x,y,z) is
    what `source.sample(Vec3(x,y,z) * 1/(N-1))` would have returned (primal_grid.rs:44-53,61-70).
    `data` may be a NumPy array (host; copied per extract), an object with `data_ptr()`/`is_cuda`
    (a CUDA torch tensor; used in place), or an int device pointer with `on_device=True`.
    """

    def __init__(self, data, size=None, on_device=None):
        self._keep = data
        if hasattr(data, "data_ptr"):
            self.on_device = bool(getattr(data, "is_cuda", False)) if on_device is None else on_device
            if not self.on_device:
                raise TypeError("pass host data as a NumPy array")
            n = _infer_size(tuple(data.shape), data.numel()) if size is None else int(size)
            if data.numel() != n * n * (n + 1) or str(data.dtype) != "torch.float32" or not data.is_contiguous():
                raise ValueError("dense grid must be contiguous float32 with N*N*(N+1) elements")
            self.size, self.ptr = n, int(data.data_ptr())
        elif isinstance(data, int):
            if size is None or not on_device:
                raise ValueError("raw pointers need size= and on_device=True")
            self.size, self.ptr, self.on_device = int(size), data, True
        else:
            arr = np.ascontiguousarray(data, dtype=np.float32)
            n = _infer_size(arr.shape, arr.size) if size is None else int(size)
            if arr.size != n * n * (n + 1):
                raise ValueError("dense grid needs N*N*(N+1) samples, got %d for N=%d" % (arr.size, n))
            self._keep = arr
            self.size, self.ptr, self.on_device = n, arr.ctypes.data, False


def _infer_size(shape, numel):
    """N from a (N+1, N, N) shape, or from N*N*(N+1) elements of a flat buffer"""
    if len(shape) == 3:
        return int(shape[-1])
    n = int(round(numel ** (1.0 / 3.0)))
    for c in (n - 1, n, n + 1):
        if c >= 1 and c * c * (c + 1) == numel:
            return c
    raise ValueError("cannot infer the lattice size from %d samples; pass size=" % numel)


class CentralDifference(DeviceSource):
    """reference src/source.rs:52-94: adds `sample_normal` by central differences (epsilon 1e-6 by default) to a
    ScalarSource.  As a scalar source it is transparent; as the source of an IndexedInterleavedNormals extractor its
    normals are evaluated on the device."""

    def __init__(self, source, epsilon=0.000001):
        _require_device_source(source)
        self.source, self.epsilon = source, float(epsilon)

    def encode(self, prog):
        self.source.encode(prog)


def find_central_difference(source):
    """the CentralDifference adaptor of a source, looking through Sampler and enclosing Translate wrappers"""
    return locate_central_difference(source)[0]


def locate_central_difference(source):
    """(adaptor, number of Translate wrappers OUTSIDE it): `Translate(o, CentralDifference(T))` differentiates f at v - o,
    `CentralDifference(Translate(o, T))` differentiates the translated function -- the encodings are the same program, so the
    position of the adaptor travels beside it (isomc_copy_out_interleaved_normals_at)"""
    s, outer = source, 0
    while True:
        if isinstance(s, CentralDifference):
            return s, outer
        if isinstance(s, Sampler):
            s = s.source
            continue
        if isinstance(s, Translate):
            s, outer = s.source, outer + 1
            continue
        return None, 0


class Sampler:
    """reference src/sampler.rs:26-41: wraps a source for `MarchingCubes.extract`"""

    def __init__(self, source):
        self.source = source


def _require_device_source(s):
    if not isinstance(s, DeviceSource):
        raise TypeError(
            "%r is not a device source: only the crate's implicit shapes, their CSG combinations, Translate "
            "and DenseGrid run on the B200 path (arbitrary callables would need a CPU path, which does not exist)"
            % (s,))


def encode_program(source):
    """source -> packed isomc_sdf_node array"""
    if isinstance(source, Sampler):
        source = source.source
    _require_device_source(source)
    nodes = []
    source.encode(nodes)
    arr = np.zeros(len(nodes), dtype=_lib.NODE_DTYPE)
    for i, nd in enumerate(nodes):
        arr[i] = nd
    return arr
