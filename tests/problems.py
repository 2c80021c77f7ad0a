"""Shared problem builders for the parity tests (oracle side only builds inputs)."""
import numpy as np

import ikarus_oracle as o
from golden_data import GOLDEN


def cantilever(dim, mat_kind, eas_m, cells=None, E=100.0, nu=0.3, L=10.0, h=2.0):
    """tests/src/testcantileverbeam.hh:83-198 of the reference."""
    if cells is None:
        cells = (10, 1, 1) if dim == 3 else (10, 1)
    bbox = (L, h, h) if dim == 3 else (L, h)
    mesh = o.structured_mesh(cells, bbox)
    lam, mu = o.lame_from_E_nu(E, nu)
    mat = o.Material(mat_kind, lam, mu, plane_strain=(dim == 2))
    kind = o.ElementKind(dim, 1, "gl", eas_m)
    flags = o.fix_nodes(mesh, o.boundary_nodes(mesh, 0, 0.0))
    fext = np.zeros(mesh.n_nodes * dim)
    X = mesh.node_coords
    pts = [(L, h, h), (L, h, 0.0)] if dim == 3 else [(L, h)]
    for p in pts:
        n = np.nonzero(np.all(np.abs(X - np.array(p)) < 1e-9, axis=1))[0][0]
        fext[n * dim + 1] = -1.0  # vec[idx] -= -lambda  (testcantileverbeam.hh:56-80)
    return mesh, kind, mat, flags, fext


def distorted(mesh, amp, seed):
    """Randomly perturb interior-and-boundary vertices (same idea as
    tests/src/testcommon.hh:179-248 createUGGridFromCorners with random distortion)."""
    rng = np.random.default_rng(seed)
    hmin = min(b / c for b, c in zip(mesh.node_coords.max(0), mesh.cells))
    coords = mesh.node_coords + amp * hmin * rng.uniform(-1, 1, mesh.node_coords.shape)
    n1 = mesh.order + 1
    # corner nodes of each element in local lexicographic order
    dim = mesh.dim
    cidx = [sum(((a >> k) & 1) * mesh.order * n1**k for k in range(dim)) for a in range(2**dim)]
    corner_nodes = mesh.elem_nodes[:, cidx]
    return o.Mesh(mesh.dim, mesh.order, coords, mesh.elem_nodes, coords[corner_nodes], mesh.cells)


def unstructured(mesh, seed, drop=None):
    """Makes a structured mesh look like an unstructured (UG/ALUGrid-style) one: cells for which `drop(centre)` is
    true are removed, the remaining elements are shuffled, unused nodes are discarded and the node numbers are
    permuted at random.  Connectivity, valence (1..2^d elements per node) and dof numbering are then arbitrary,
    which is what the pattern builder / gather have to cope with (tests/src/testcommon.hh:179-248 uses UGGrid)."""
    rng = np.random.default_rng(seed)
    keep = np.arange(mesh.n_elem)
    if drop is not None:
        centres = mesh.corner_coords.mean(axis=1)
        keep = np.array([e for e in keep if not drop(centres[e])])
    keep = rng.permutation(keep)
    en = mesh.elem_nodes[keep]
    used = np.unique(en)
    new_id = np.full(mesh.n_nodes, -1, dtype=np.int64)
    new_id[used] = rng.permutation(used.shape[0])
    coords = np.empty((used.shape[0], mesh.dim))
    coords[new_id[used]] = mesh.node_coords[used]
    return o.Mesh(mesh.dim, mesh.order, coords, new_id[en], mesh.corner_coords[keep], mesh.cells)


# tests/src/testinhomogeneousdbc.cpp:22-33 (load control): (total iterations, u_y at (L/2, 0)); lambda ends at 1
ELASTIC_STRIP_EXPECTED = {(c["material"], c["order"]): (c["iterations"], c["u_y"])
                          for c in GOLDEN["elastic_strip"]["cases"]}


def elastic_strip(mat_kind, order):
    """tests/src/testelasticstrip.hh:46-118: YaspGrid 10x10 over 10x10, plane strain, u = 0 at x = 0, u_y = 0 and the
    inhomogeneous value u_x = 10 lambda at x = L.  Materials of testinhomogeneousdbc.cpp:185-190."""
    L = 10.0
    mesh = o.structured_mesh((10, 10), (L, L), order=order)
    lam, mu = o.lame_from_E_nu(100.0, 0.3) if mat_kind == "svk" else (24.0, 6.0)
    mat = o.Material(mat_kind, lam, mu, plane_strain=True)
    kind = o.ElementKind(2, order, "gl")
    flags = o.fix_nodes(mesh, o.boundary_nodes(mesh, 0, 0.0))
    flags |= o.fix_nodes(mesh, o.boundary_nodes(mesh, 0, L), comps=[1])
    value = lambda x, lam_: (10.0 * lam_ if abs(x[0] - L) < 1e-12 else 0.0 * lam_, 0.0 * lam_)
    deriv = lambda x, lam_: (10.0 if abs(x[0] - L) < 1e-12 else 0.0, 0.0)
    probe = int(np.nonzero(np.all(np.abs(mesh.node_coords - np.array([L / 2, 0.0])) < 1e-9, axis=1))[0][0])
    return mesh, kind, mat, flags, value, deriv, 2 * probe + 1


def patch_test_mesh():
    """The five-quad patch of tests/src/testinhomogeneousdbc.cpp:58-73 (Macneal & Harder): vertices and elements in
    insertion order, DUNE local vertex order (x fastest)."""
    X = np.array([[0.0, 0.0], [0.24, 0.0], [0.0, 0.12], [0.24, 0.12], [0.04, 0.02], [0.18, 0.03], [0.08, 0.08],
                  [0.16, 0.08]])
    en = np.array([[0, 1, 4, 5], [5, 1, 7, 3], [6, 7, 2, 3], [0, 4, 2, 6], [4, 5, 6, 7]])
    return o.Mesh(2, 1, X, en, X[en], ())


PATCH_EXPECTED_D = np.array(GOLDEN["plane_stress_patch_test"]["displacements"], float)


def fixed_distorted_quad():
    """createUGGridFromCorners<2>(CornerDistortionFlag::fixedDistorted), tests/src/testcommon.hh:179-215."""
    X = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]]) + np.array([[-0.2, -0.05], [-0.15, 0.05], [0.15, 0.15],
                                                                              [-0.05, -0.1]])
    en = np.array([[0, 1, 2, 3]])
    return o.Mesh(2, 1, X, en, X[en], ())
