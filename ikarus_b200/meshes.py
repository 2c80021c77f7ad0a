"""Synthetic structured meshes with YaspGrid-compatible numbering (host-side input generation only).

YaspGrid numbers vertices and elements lexicographically with x fastest; `lagrange<1>` global index =
vertex index and `power<d>(..., FlatInterleaved)` gives dof = d*node + comp (pinned by
ikarus/python/test/linearelastictest.py:213-231 of the reference).  A z-slab of such a mesh owns a
contiguous block of rows, which is what the element-partitioned multi-GPU path relies on.
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class HexSlab:
    dim: int
    n_nodes: int  # global
    n_dof: int  # global
    corner_coords: np.ndarray  # [nElemLocal, 2^d, d]
    elem_dofs: np.ndarray  # [nElemLocal, nodes*d] global dof ids
    node_begin: int  # owned node range (global ids)
    node_end: int
    n_owned_elems: int  # elements whose lowest z-layer is owned (each element counted on exactly one rank)
    cells: tuple
    h: float
    node_coords_fn: object = None


def structured_q1(cells, bbox, layer_begin=None, layer_end=None):
    """Q1 grid of `cells` over `bbox`.  With layer_begin/layer_end (node layers of the slowest axis) only the
    elements touching an owned node layer are returned (owner-computes with one ghost element layer)."""
    cells = tuple(int(c) for c in cells)
    d = len(cells)
    npts = [c + 1 for c in cells]
    stride = np.cumprod([1] + npts[:-1]).astype(np.int64)
    n_nodes = int(np.prod(npts))
    last = d - 1
    lb = 0 if layer_begin is None else int(layer_begin)
    le = npts[last] if layer_end is None else int(layer_end)
    # element layers touching node layers [lb, le): e_k in [lb-1, le-1] clipped
    ek0, ek1 = max(lb - 1, 0), min(le, cells[last])
    rng = [np.arange(c, dtype=np.int64) for c in cells]
    rng[last] = np.arange(ek0, ek1, dtype=np.int64)
    grids = np.meshgrid(*rng, indexing="ij")
    eidx = np.stack([g.reshape(-1, order="F") for g in grids], axis=-1)
    loc = np.array([[(a >> k) & 1 for k in range(d)] for a in range(2**d)], dtype=np.int64)
    nodes_ijk = eidx[:, None, :] + loc[None, :, :]
    elem_nodes = (nodes_ijk * stride[None, None, :]).sum(-1)
    h = np.array([bbox[k] / cells[k] for k in range(d)])
    corner = nodes_ijk * h[None, None, :]
    elem_dofs = (elem_nodes[:, :, None] * d + np.arange(d)[None, None, :]).reshape(elem_nodes.shape[0], -1)
    owned = int(np.count_nonzero((eidx[:, last] >= lb) & (eidx[:, last] < le)))
    return HexSlab(d, n_nodes, n_nodes * d, np.ascontiguousarray(corner, float), np.ascontiguousarray(elem_dofs),
                   int(lb * stride[last]), int(le * stride[last]), owned, cells, float(h.min()))


def node_coords(cells, bbox):
    cells = tuple(int(c) for c in cells)
    d = len(cells)
    g = [np.linspace(0.0, bbox[k], cells[k] + 1) for k in range(d)]
    mg = np.meshgrid(*g, indexing="ij")
    return np.stack([m.reshape(-1, order="F") for m in mg], axis=-1)


def clamp_face_flags(cells, axis=0, value_index=0):
    """Dirichlet flags fixing every dof of the nodes on the face `axis = value_index` (interleaved dofs)."""
    cells = tuple(int(c) for c in cells)
    d = len(cells)
    npts = [c + 1 for c in cells]
    idx = np.indices(npts).reshape(d, -1, order="F")
    on = idx[axis] == value_index
    flags = np.zeros(int(np.prod(npts)) * d, dtype=bool)
    nodes = np.nonzero(on)[0]
    for c in range(d):
        flags[nodes * d + c] = True
    return flags
