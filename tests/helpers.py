"""Shared helpers for the parity tests (test infrastructure)."""
import hashlib

import numpy as np

from oracle import oracle as O


def sha(arr, dtype):
    return hashlib.sha256(np.ascontiguousarray(arr).astype(dtype).tobytes()).hexdigest()


# The shapes of SURVEY.md 8(c) / BASELINE.md 3 as oracle programs and as isosurface_b200 sources.
def oracle_prog(name):
    T = lambda *n: [(O.TRANSLATE_PUSH, .5, .5, .5)] + list(n) + [(O.TRANSLATE_POP,)]
    torus = [(O.TORUS, .25, .1)]
    csgA = [(O.SPHERE, .25), (O.PRISM, .2, .2, .2), (O.DIFFERENCE,), (O.CYLINDER, .02, .25), (O.UNION,)]
    csgB = [(O.SPHERE, .3), (O.PRISM, .2, .2, .2), (O.INTERSECTION,)]
    progs = {
        "torus_origin": torus,                       # benches/isosurface.rs:22
        "sphere03": T((O.SPHERE, .3)),
        "sphere05_origin": [(O.SPHERE, .5)],
        "torus": T(*torus),
        "csgA": T(*csgA),                            # examples/sampler.rs:79-85
        "csgB": T(*csgB),                            # examples/sampler.rs:89-92
        "prism": T((O.PRISM, .2, .2, .2)),
        "cylinder": T((O.CYLINDER, .25, .2)),
        "nested": [(O.TRANSLATE_PUSH, .25, .25, .25), (O.SPHERE, .2), (O.TRANSLATE_PUSH, .5, .5, .5), (O.SPHERE, .2),
                   (O.TRANSLATE_POP,), (O.UNION,), (O.TRANSLATE_POP,)],
    }
    return O.program(progs[name])


def iso_source(name):
    import isosurface_b200 as iso
    torus = iso.Torus(.25, .1)
    csgA = iso.Union(iso.Difference(iso.Sphere(.25), iso.RectangularPrism((.2, .2, .2))), iso.Cylinder(.02, .25))
    csgB = iso.Intersection(iso.Sphere(.3), iso.RectangularPrism((.2, .2, .2)))
    srcs = {
        "torus_origin": torus,
        "sphere03": iso.Translate(.5, iso.Sphere(.3)),
        "sphere05_origin": iso.Sphere(.5),
        "torus": iso.Translate(.5, torus),
        "csgA": iso.Translate(.5, csgA),
        "csgB": iso.Translate(.5, csgB),
        "prism": iso.Translate(.5, iso.RectangularPrism((.2, .2, .2))),
        "cylinder": iso.Translate(.5, iso.Cylinder(.25, .2)),
        "nested": iso.Translate(.25, iso.Union(iso.Sphere(.2), iso.Translate(.5, iso.Sphere(.2)))),
    }
    return srcs[name]


def mesh_diff(xyz, idx, oxyz, oidx, tol=0.0):
    """'' if the meshes are identical (indices bit-exact, positions within tol), else a description."""
    msgs = []
    if len(idx) != len(oidx):
        msgs.append("triangle count %d != oracle %d" % (len(idx) // 3, len(oidx) // 3))
    if len(xyz) != len(oxyz):
        msgs.append("vertex count %d != oracle %d" % (len(xyz) // 3, len(oxyz) // 3))
    n = min(len(idx), len(oidx))
    bad = np.nonzero(idx[:n] != oidx[:n])[0]
    if len(bad):
        k = int(bad[0])
        msgs.append("%d index mismatches, first at idx[%d] (tri %d): got %d want %d; next tris got %s want %s"
                    % (len(bad), k, k // 3, idx[k], oidx[k], idx[k - k % 3:k - k % 3 + 6].tolist(),
                       oidx[k - k % 3:k - k % 3 + 6].tolist()))
    n = min(len(xyz), len(oxyz))
    d = np.abs(xyz[:n].astype(np.float64) - oxyz[:n].astype(np.float64))
    d = np.where(np.isnan(d), np.inf, d)
    both_nan = np.isnan(xyz[:n]) & np.isnan(oxyz[:n])
    d = np.where(both_nan, 0.0, d)
    if n and d.max() > tol:
        k = int(np.argmax(d > tol))
        msgs.append("%d position mismatches > %g (max %.3g), first at vertex %d: got %s want %s"
                    % (int((d > tol).sum()), tol, d.max(), k // 3, xyz[k - k % 3:k - k % 3 + 3].tolist(),
                       oxyz[k - k % 3:k - k % 3 + 3].tolist()))
    elif n and tol > 0:
        nb = int((xyz[:n].view(np.uint32) != oxyz[:n].view(np.uint32)).sum())
        if nb:
            msgs.append("NOTE: within tol but %d floats differ in bits" % nb) if False else None
    return "; ".join(m for m in msgs if m)


def mesh_invariants(xyz, idx, closed=True):
    """Oracle-independent checks.  Returns dict of facts, asserts the universal ones."""
    V, T = len(xyz) // 3, len(idx) // 3
    assert idx.size == 0 or int(idx.max()) < V, "index out of range"
    if V:
        ref = np.zeros(V, bool)
        ref[idx] = True
        assert ref.all(), "unreferenced vertex"
    tri = idx.reshape(-1, 3).astype(np.int64)
    # directed edges: each interior edge appears once in each direction on a closed orientable surface
    a = np.concatenate([tri[:, 0], tri[:, 1], tri[:, 2]])
    b = np.concatenate([tri[:, 1], tri[:, 2], tri[:, 0]])
    fwd = a * (V + 1) + b
    rev = b * (V + 1) + a
    uf, cf = np.unique(fwd, return_counts=True)
    facts = {"V": V, "T": T, "directed_edge_dups": int((cf > 1).sum())}
    if closed:
        facts["unpaired"] = int(np.setdiff1d(uf, np.unique(rev)).size)
        E = len(np.unique(np.minimum(a, b) * (V + 1) + np.maximum(a, b)))
        facts["euler"] = V - E + T
    return facts
