"""Neighbour lists on the GPU (K1) and the CSR structure the edge kernels consume.

Drop-in counterparts of the reference's neighbour-list entry points:
  radius_graph(...)       <- torch_cluster.radius_graph as called at data/transform.py:58-64
  radius_graph_pbc(...)   <- xequinet/data/radius_graph.py:35-192
  NeighborTransform       <- xequinet/data/transform.py:21-69
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, Optional, Tuple

import torch

from . import _lib, keys


class NeighborGraph:
    """Device-resident CSR (by center) + transposed CSR (by neighbor) of one batch."""

    MOLECULE_TILE_MAX_NODES = 64  # graphs up to this size become one work tile each (staged rows)

    def __init__(self, n_nodes: int, n_graphs: int, rowptr, col, offsets=None, cell=None, node_graph=None,
                 capacity: Optional[int] = None, mol_ptr: Optional[torch.Tensor] = None,
                 n_centers: Optional[int] = None, max_tile_nodes: int = 0):
        self.n_nodes = int(n_nodes)
        # nodes that can be centers (rows of the CSR that are walked); the rest (ghost atoms of a spatially
        # sharded run, xequinet_b200/domain.py) only ever appear as neighbours
        self.n_centers = int(n_centers) if n_centers is not None else int(n_nodes)
        self.n_graphs = int(n_graphs)
        self.rowptr = rowptr
        self.col = col
        # capacity mode: n_edges is the allocated capacity, the live count stays in rowptr[N] on the device
        self.capacity = capacity
        self.n_edges = int(capacity) if capacity else int(col.numel())
        self.offsets = offsets  # int8 [E,4] or None
        self.cell = cell  # float32 [G,3,3] or None
        self.node_graph = node_graph  # int32 [N] or None
        self._struct = None
        # provenance: the (edge_index, cell_offsets, cell) tensors this structure was derived from, or None when
        # it was built from positions by K1 in capacity mode; see nn/basic.py::compute_edge_data
        self._src = None
        lib = _lib.get()
        dev = self.rowptr.device
        N, E = self.n_nodes, self.n_edges
        self.t_rowptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
        self.t_row = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
        self.t_eid = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
        self._t_ws = torch.empty(lib.xeq_csr_transpose_workspace_bytes(N, E), dtype=torch.uint8, device=dev)
        tc, tn = lib.xeq_center_tile_edges(), lib.xeq_neighbor_tile_edges()
        # work tiles: one per molecule when the caller vouches that every graph is small (mol_ptr = the
        # batch `ptr` array, int32), else node-aligned blocks of tc / tn edges
        self.mol_ptr = mol_ptr
        self.max_tile_nodes = int(max_tile_nodes) if mol_ptr is not None else 0  # host-known bound of the molecule sizes
        if mol_ptr is not None:
            self.tile_ptr = self.t_tile_ptr = mol_ptr
            self.n_tiles = self.t_n_tiles = int(mol_ptr.numel()) - 1
            self.tile_mode = 1
        else:
            self.n_tiles = lib.xeq_csr_tile_count(self.n_centers, E, tc)
            self.t_n_tiles = lib.xeq_csr_tile_count(N, E, tn)
            self.tile_mode = 0
            self.tile_ptr = torch.empty(lib.xeq_csr_tile_count(N, E, tc) + 1, dtype=torch.int32, device=dev)
            self.t_tile_ptr = torch.empty(self.t_n_tiles + 1, dtype=torch.int32, device=dev)
        self.transpose()

    def transpose(self):
        """(Re)derive the transposed CSR and the work tiles from rowptr/col; launches only."""
        lib = _lib.get()
        N, E = self.n_nodes, self.n_edges
        ws, nbytes = self._t_ws, self._t_ws.numel()
        _lib.check(lib.xeq_csr_transpose(_lib.ptr(self.rowptr), _lib.ptr(self.col), N, E, _lib.ptr(self.t_rowptr),
                                         _lib.ptr(self.t_row), _lib.ptr(self.t_eid), _lib.ptr(ws), nbytes,
                                         _lib.stream()), "xeq_csr_transpose")
        if self.tile_mode == 1:
            return
        # node-aligned work tiles of both structures
        tc, tn = lib.xeq_center_tile_edges(), lib.xeq_neighbor_tile_edges()
        _lib.check(lib.xeq_csr_tile_bounds(_lib.ptr(self.rowptr), self.n_centers, E, tc, _lib.ptr(self.tile_ptr), _lib.stream()),
                   "xeq_csr_tile_bounds")
        _lib.check(lib.xeq_csr_tile_bounds(_lib.ptr(self.t_rowptr), N, E, tn, _lib.ptr(self.t_tile_ptr), _lib.stream()),
                   "xeq_csr_tile_bounds")

    @property
    def struct(self) -> _lib.XeqGraph:
        if self._struct is None:
            g = _lib.XeqGraph()
            g.n_nodes, g.n_edges, g.n_graphs = self.n_nodes, self.n_edges, self.n_graphs
            g.rowptr, g.col = self.rowptr.data_ptr(), self.col.data_ptr()
            g.t_rowptr, g.t_row, g.t_eid = self.t_rowptr.data_ptr(), self.t_row.data_ptr(), self.t_eid.data_ptr()
            g.offsets = self.offsets.data_ptr() if self.offsets is not None else None
            g.cell = self.cell.data_ptr() if self.cell is not None else None
            g.node_graph = self.node_graph.data_ptr() if self.node_graph is not None else None
            g.tile_ptr, g.t_tile_ptr = self.tile_ptr.data_ptr(), self.t_tile_ptr.data_ptr()
            g.n_tiles, g.t_n_tiles, g.tile_mode = self.n_tiles, self.t_n_tiles, self.tile_mode
            g.max_tile_nodes = self.max_tile_nodes
            self._struct = g
        return self._struct

    def edge_index(self) -> torch.Tensor:
        """COO [2,E] int64 in canonical order (row 0 = center, row 1 = neighbor; keys.py:16-17)."""
        counts = (self.rowptr[1:] - self.rowptr[:-1]).long()
        center = torch.repeat_interleave(torch.arange(self.n_nodes, device=self.rowptr.device), counts)
        # capacity mode: col is allocated for the capacity, the live count is rowptr[N]
        return torch.stack([center, self.col[: center.numel()].long()])

    def derived_from(self, edge_index, cell_offsets, cell) -> bool:
        """True when this structure was derived from exactly these tensors (same objects, not modified in place
        since), i.e. when it may stand in for them."""
        if self._src is None:
            return False
        for ref, cur in zip(self._src, (edge_index, cell_offsets, cell)):
            if (ref is None) != (cur is None):
                return False
            if ref is not None and (ref[0] is not cur or ref[1] != cur._version):
                return False
        return True

    def set_source(self, edge_index, cell_offsets, cell) -> None:
        self._src = tuple(None if t is None else (t, t._version) for t in (edge_index, cell_offsets, cell))


def _image_repeats(cell: torch.Tensor, pbc, cutoff: float):
    """Images per axis = ceil(rc * |a_j x a_k| / V), max over graphs (data/radius_graph.py:61-89).
    Tiny host-side computation on the [G,3,3] lattice (the reference also syncs here, :89)."""
    c = cell.detach().to("cpu", torch.float32)
    cross = [torch.cross(c[:, 1], c[:, 2], dim=-1), torch.cross(c[:, 2], c[:, 0], dim=-1),
             torch.cross(c[:, 0], c[:, 1], dim=-1)]
    vol = torch.sum(c[:, 0] * cross[0], dim=-1, keepdim=True)
    reps = []
    for ax in range(3):
        if pbc[ax]:
            inv_min = torch.norm(cross[ax] / vol, p=2, dim=-1)
            reps.append(int(torch.ceil(cutoff * inv_min).max().item()))
        else:
            reps.append(0)
    return reps


def build_graph(pos: torch.Tensor, cutoff: float, ptr: Optional[torch.Tensor] = None,
                batch: Optional[torch.Tensor] = None, cell: Optional[torch.Tensor] = None, pbc=None,
                want_coo: bool = False):
    """K1: radius graph of a (batched, optionally periodic) structure.  Returns
    (NeighborGraph, edge_index or None, cell_offsets or None)."""
    lib = _lib.get()
    if not pos.is_cuda:
        raise RuntimeError("build_graph needs CUDA tensors: xequinet_b200 has no CPU fallback")
    dev = pos.device
    pos32 = pos.detach().to(torch.float32).contiguous()
    N = pos32.shape[0]
    if ptr is None:
        if batch is None:
            ptr = torch.tensor([0, N], dtype=torch.int32, device=dev)
        else:
            G = int(batch.max().item()) + 1 if N else 1
            counts = torch.bincount(batch, minlength=G)
            ptr = torch.zeros(G + 1, dtype=torch.int32, device=dev)
            ptr[1:] = torch.cumsum(counts, 0)
    ptr32 = ptr.to(device=dev, dtype=torch.int32).contiguous()
    G = ptr32.numel() - 1
    if batch is not None:
        node_graph = batch.to(device=dev, dtype=torch.int32).contiguous()
    elif G > 1:
        node_graph = torch.repeat_interleave(torch.arange(G, device=dev, dtype=torch.int32),
                                             (ptr32[1:] - ptr32[:-1]).long())
    else:
        node_graph = None
    periodic = cell is not None
    pbc_arr = (ctypes.c_int32 * 3)(0, 0, 0)
    rep_arr = (ctypes.c_int32 * 3)(0, 0, 0)
    cell32 = None
    if periodic:
        cell32 = cell.detach().to(device=dev, dtype=torch.float32).reshape(-1, 3, 3).contiguous()
        if cell32.shape[0] != G:
            raise ValueError("cell must be [n_graphs, 3, 3]")
        if pbc is None:
            pbc_l = [True, True, True]
        else:
            p = torch.as_tensor(pbc).reshape(-1, 3).cpu()
            # PBC must be the same for all graphs (data/radius_graph.py:50-51)
            assert bool(torch.all(p[0] == p)), "PBC must be the same for all graphs"
            pbc_l = [bool(v) for v in p[0].tolist()]
        reps = _image_repeats(cell32, pbc_l, cutoff)
        for k in range(3):
            pbc_arr[k], rep_arr[k] = int(pbc_l[k]), reps[k]
    nbytes = lib.xeq_radius_graph_workspace_bytes(N, G, int(periodic))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    rowptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
    st = _lib.stream()
    _lib.check(lib.xeq_radius_graph_count(_lib.ptr(pos32), N, _lib.ptr(ptr32), _lib.ptr(node_graph), G,
                                          _lib.ptr(cell32), pbc_arr, rep_arr, float(cutoff), _lib.ptr(rowptr),
                                          _lib.ptr(ws), nbytes, st), "xeq_radius_graph_count")
    sizes = (ptr32[1:] - ptr32[:-1]).max() if N else ptr32[0]
    E, max_nodes = torch.stack([rowptr[-1], sizes]).tolist()  # the one host sync: sizes the edge arrays
    mol_ptr = ptr32 if (0 < max_nodes <= NeighborGraph.MOLECULE_TILE_MAX_NODES) else None
    col = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
    offsets = torch.empty((max(E, 1), 4), dtype=torch.int8, device=dev) if periodic else None
    ei = torch.empty((2, E), dtype=torch.int64, device=dev) if want_coo else None
    co = torch.empty((E, 3), dtype=torch.float32, device=dev) if (want_coo and periodic) else None
    _lib.check(lib.xeq_radius_graph_fill(_lib.ptr(pos32), N, _lib.ptr(ptr32), _lib.ptr(node_graph), G,
                                         _lib.ptr(cell32), pbc_arr, rep_arr, float(cutoff), _lib.ptr(rowptr),
                                         _lib.ptr(col), _lib.ptr(offsets), _lib.ptr(ei) if E else None,
                                         _lib.ptr(co) if (co is not None and E) else None, 0, None, _lib.ptr(ws), nbytes,
                                         st),
               "xeq_radius_graph_fill")
    g = NeighborGraph(N, G, rowptr, col[:E] if E else col[:0], offsets[:E] if (periodic and E) else (offsets[:0] if periodic else None),
                      cell32, node_graph if (periodic and G > 1) else None, mol_ptr=mol_ptr, max_tile_nodes=int(max_nodes))
    return g, ei, co


class StaticGraphBuilder:
    """Capacity-mode K1 for CUDA-graph replay and MD loops: all arrays are allocated once for
    (n_nodes, n_graphs, edge_capacity); build() only launches kernels (no allocation, no host
    sync) and always returns the same NeighborGraph, whose live edge count stays on the device.
    `overflow` (device int32) is raised when a structure has more than `edge_capacity` edges."""

    def __init__(self, n_nodes: int, ptr: torch.Tensor, cutoff: float, edge_capacity: int, cell=None, pbc=None,
                 n_centers: Optional[int] = None):
        lib = _lib.get()
        dev = ptr.device
        self.cutoff = float(cutoff)
        self.N = int(n_nodes)
        self.ptr32 = ptr.to(torch.int32).contiguous()
        self.G = self.ptr32.numel() - 1
        self.cap = int(edge_capacity)
        self.node_graph = (torch.repeat_interleave(torch.arange(self.G, device=dev, dtype=torch.int32),
                                                   (self.ptr32[1:] - self.ptr32[:-1]).long())
                           if self.G > 1 else None)
        self.periodic = cell is not None
        self.pbc_arr = (ctypes.c_int32 * 3)(0, 0, 0)
        self.rep_arr = (ctypes.c_int32 * 3)(0, 0, 0)
        self.cell32 = None
        if self.periodic:
            self.cell32 = cell.detach().to(device=dev, dtype=torch.float32).reshape(-1, 3, 3).contiguous()
            pbc_l = [True, True, True] if pbc is None else [bool(v) for v in torch.as_tensor(pbc).reshape(-1, 3)[0].tolist()]
            reps = _image_repeats(self.cell32, pbc_l, self.cutoff)  # fixed cell: computed once
            for k in range(3):
                self.pbc_arr[k], self.rep_arr[k] = int(pbc_l[k]), reps[k]
        self.nbytes = lib.xeq_radius_graph_workspace_bytes(self.N, self.G, int(self.periodic))
        self.ws = torch.empty(self.nbytes, dtype=torch.uint8, device=dev)
        self.rowptr = torch.zeros(self.N + 1, dtype=torch.int32, device=dev)
        self.col = torch.zeros(self.cap, dtype=torch.int32, device=dev)
        self.offsets = torch.zeros((self.cap, 4), dtype=torch.int8, device=dev) if self.periodic else None
        self.overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        max_nodes = int((self.ptr32[1:] - self.ptr32[:-1]).max().item()) if self.N else 0
        mol_ptr = self.ptr32 if (0 < max_nodes <= NeighborGraph.MOLECULE_TILE_MAX_NODES) else None
        self.graph = NeighborGraph(self.N, self.G, self.rowptr, self.col, self.offsets, self.cell32,
                                   self.node_graph if (self.periodic and self.G > 1) else None, capacity=self.cap,
                                   mol_ptr=mol_ptr, n_centers=n_centers, max_tile_nodes=max_nodes)

    def build(self, pos: torch.Tensor, check_overflow: bool = True) -> "NeighborGraph":
        """Launches K1 on `pos` (float32 [N,3], CUDA).  Outside CUDA-graph capture the overflow flag is read back
        (one host sync; pass check_overflow=False to skip it) and a RuntimeError is raised when the capacity was
        exceeded; under capture the caller checks `self.overflow` after the replay.  Either way the kernels stay
        inside the allocated arrays (rowptr is clamped to the capacity on the device)."""
        lib = _lib.get()
        pos32 = pos.detach()
        if pos32.dtype != torch.float32 or tuple(pos32.shape) != (self.N, 3) or not pos32.is_contiguous():
            raise ValueError(f"StaticGraphBuilder.build: pos must be a contiguous float32 [{self.N}, 3] tensor, got "
                             f"{pos32.dtype} {tuple(pos32.shape)}")
        st = _lib.stream()
        _lib.check(lib.xeq_radius_graph_count(_lib.ptr(pos32), self.N, _lib.ptr(self.ptr32), _lib.ptr(self.node_graph),
                                              self.G, _lib.ptr(self.cell32), self.pbc_arr, self.rep_arr, self.cutoff,
                                              _lib.ptr(self.rowptr), _lib.ptr(self.ws), self.nbytes, st),
                   "xeq_radius_graph_count")
        _lib.check(lib.xeq_radius_graph_fill(_lib.ptr(pos32), self.N, _lib.ptr(self.ptr32), _lib.ptr(self.node_graph),
                                             self.G, _lib.ptr(self.cell32), self.pbc_arr, self.rep_arr, self.cutoff,
                                             _lib.ptr(self.rowptr), _lib.ptr(self.col), _lib.ptr(self.offsets), None, None,
                                             self.cap, _lib.ptr(self.overflow), _lib.ptr(self.ws), self.nbytes, st),
                   "xeq_radius_graph_fill")
        self.graph.transpose()
        if check_overflow and not torch.cuda.is_current_stream_capturing():
            if int(self.overflow.item()) != 0:
                self.overflow.zero_()
                raise RuntimeError(f"StaticGraphBuilder: more than edge_capacity = {self.cap} edges; the list was truncated")
        return self.graph


def graph_from_edge_index(edge_index: torch.Tensor, n_nodes: int, n_graphs: int = 1,
                          cell_offsets: Optional[torch.Tensor] = None, cell: Optional[torch.Tensor] = None,
                          batch: Optional[torch.Tensor] = None, ptr: Optional[torch.Tensor] = None) -> NeighborGraph:
    """CSR structure for a caller-supplied COO edge list (any order; nn/basic.py:67 takes
    `edge_index` from the data dict).  Unsorted lists are first put in canonical order."""
    lib = _lib.get()
    if not edge_index.is_cuda:
        raise RuntimeError("graph_from_edge_index needs CUDA tensors: xequinet_b200 has no CPU fallback")
    dev = edge_index.device
    ei = edge_index.to(torch.int64)
    E = ei.shape[1]
    co = cell_offsets
    if E > 1 and bool((ei[0, 1:] < ei[0, :-1]).any()):
        order = torch.sort(ei[0], stable=True)[1]
        ei = ei[:, order]
        co = co[order] if co is not None else None
    ei = ei.contiguous()
    periodic = cell is not None
    if periodic and co is None:
        raise ValueError("PBC and cell must be both defined or both undefined.")
    co32 = co.to(torch.float32).contiguous() if co is not None else None
    rowptr = torch.empty(n_nodes + 1, dtype=torch.int32, device=dev)
    col = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
    offsets = torch.empty((max(E, 1), 4), dtype=torch.int8, device=dev) if periodic else None
    _lib.check(lib.xeq_csr_from_sorted_coo(_lib.ptr(ei) if E else None, _lib.ptr(co32) if (periodic and E) else None,
                                           n_nodes, E, _lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(offsets),
                                           _lib.stream()), "xeq_csr_from_sorted_coo")
    cell32 = cell.detach().to(device=dev, dtype=torch.float32).reshape(-1, 3, 3).contiguous() if periodic else None
    node_graph = None
    if periodic and n_graphs > 1:
        if batch is None:
            raise ValueError("batch is required for multi-graph periodic input")
        node_graph = batch.to(device=dev, dtype=torch.int32).contiguous()
    mol_ptr, max_nodes = None, 0
    if ptr is not None and n_nodes > 0:
        ptr32 = ptr.to(device=dev, dtype=torch.int32).contiguous()
        max_nodes = int((ptr32[1:] - ptr32[:-1]).max().item())
        if max_nodes <= NeighborGraph.MOLECULE_TILE_MAX_NODES:
            mol_ptr = ptr32
    g = NeighborGraph(n_nodes, n_graphs, rowptr, col[:E], offsets[:E] if periodic else None, cell32, node_graph,
                      mol_ptr=mol_ptr, max_tile_nodes=max_nodes)
    g.sorted_edge_index = ei
    g.set_source(edge_index, cell_offsets, cell)
    return g


# ------------------------------------------------------------------------------------------
# reference-compatible entry points
# ------------------------------------------------------------------------------------------
def radius_graph(x: torch.Tensor, r: float, batch: Optional[torch.Tensor] = None, loop: bool = False,
                 max_num_neighbors: int = 32, flow: str = "source_to_target", num_workers: int = 1,
                 batch_size: Optional[int] = None) -> torch.Tensor:
    """Signature of torch_cluster.radius_graph.  As at the reference's call site
    (data/transform.py:57-64) the neighbour cap is never binding, so it is not applied here;
    row 0 = center, row 1 = neighbor, canonically sorted."""
    if loop:
        raise NotImplementedError("loop=True is not used by XequiNet (data/transform.py:58-64)")
    _, ei, _ = build_graph(x, r, batch=batch, want_coo=True)
    return ei


def radius_graph_pbc(pos: torch.Tensor, n_nodes_per_graph: torch.Tensor, pbc: torch.Tensor, cell: torch.Tensor,
                     cutoff: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """Signature of xequinet.data.radius_graph.radius_graph_pbc (data/radius_graph.py:36-42).
    Returns (edge_index [2,E] int64, cell_offsets [E,3] float), canonically sorted."""
    n = n_nodes_per_graph.to(pos.device)
    ptr = torch.zeros(n.numel() + 1, dtype=torch.int32, device=pos.device)
    ptr[1:] = torch.cumsum(n, 0)
    _, ei, co = build_graph(pos, cutoff, ptr=ptr, cell=cell, pbc=pbc, want_coo=True)
    return ei, co.to(pos.dtype)


class NeighborTransform:
    """xequinet/data/transform.py:21-69 for dict-shaped data: adds `edge_index` (and
    `cell_offsets` under PBC) plus the prebuilt CSR structure the model reuses."""

    def __init__(self, cutoff: float) -> None:
        self.cutoff = cutoff

    def __call__(self, data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        pos = data[keys.POSITIONS]
        has_pbc = keys.PBC in data and bool(torch.as_tensor(data[keys.PBC]).any())
        has_cell = keys.CELL in data
        if has_pbc != has_cell:
            raise ValueError("PBC and cell must be both defined or both undefined.")
        ptr, batch = data.get(keys.BATCH_PTR), data.get(keys.BATCH)
        if has_pbc:
            g, ei, co = build_graph(pos, self.cutoff, ptr=ptr, batch=batch, cell=data[keys.CELL], pbc=data[keys.PBC],
                                    want_coo=True)
            data[keys.CELL_OFFSETS] = co.to(pos.dtype)
        else:
            g, ei, _ = build_graph(pos, self.cutoff, ptr=ptr, batch=batch, want_coo=True)
        data[keys.EDGE_INDEX] = ei
        g.set_source(ei, data.get(keys.CELL_OFFSETS) if has_pbc else None, data[keys.CELL] if has_pbc else None)
        data[keys.GRAPH] = g
        return data


class SkinNeighborTransform(NeighborTransform):
    """Verlet-skin reuse of the neighbour list for MD loops (the reference rebuilds the list at
    every step, interface/ase_calculator.py:86-88).  The list is built with `cutoff + skin` and kept while no atom has
    moved more than `skin / 2` since the build and the cell and batch structure are unchanged.  Edges longer than the
    model's cutoff contribute exactly zero to the message (chi(d) = 0 for d >= r_c, nn/rbf.py:47-48 -- and so do all
    their derivatives), so energies and forces are those of the exact list; `edge_index` is then a superset of
    the reference's edge set.  Cell offsets refer to the unwrapped positions (data/radius_graph.py:186-190), so
    atoms may leave the cell between rebuilds.  One host synchronisation per call (the displacement test)."""

    def __init__(self, cutoff: float, skin: float = 1.0) -> None:
        super().__init__(cutoff)
        if skin < 0:
            raise ValueError("skin must be >= 0")
        self.skin = float(skin)
        self.n_calls = 0
        self.n_builds = 0
        self._ref = None  # (pos, cell, ptr/batch signature) at the last build
        self._cached: Dict[str, torch.Tensor] = {}

    @staticmethod
    def _signature(data) -> tuple:
        ptr, batch = data.get(keys.BATCH_PTR), data.get(keys.BATCH)
        n = int(data[keys.POSITIONS].shape[0])
        g = int(ptr.numel() - 1) if ptr is not None else (int(batch.max().item()) + 1 if (batch is not None and n) else 1)
        return (n, g, keys.CELL in data)

    def needs_rebuild(self, data: Dict[str, torch.Tensor]) -> bool:
        """True when there is no list yet, the structure changed shape, the cell changed, or some atom has moved
        more than skin / 2 since the list was built."""
        if self._ref is None:
            return True
        pos_ref, cell_ref, sig = self._ref
        if sig != self._signature(data) or pos_ref.device != data[keys.POSITIONS].device:
            return True
        if cell_ref is not None and not torch.equal(cell_ref, data[keys.CELL].detach().reshape(-1, 3, 3).to(cell_ref.dtype)):
            return True
        disp2 = ((data[keys.POSITIONS].detach() - pos_ref) ** 2).sum(-1)
        return bool(disp2.max().item() > (0.5 * self.skin) ** 2) if disp2.numel() else False

    def _build(self, data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        wide = NeighborTransform(self.cutoff + self.skin)
        return wide(dict(data))

    def __call__(self, data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        self.n_calls += 1
        if self.needs_rebuild(data):
            built = self._build(data)
            self._cached = {k: built[k] for k in (keys.EDGE_INDEX, keys.GRAPH, keys.CELL_OFFSETS) if k in built}
            cell = data[keys.CELL].detach().reshape(-1, 3, 3).clone() if keys.CELL in data else None
            self._ref = (data[keys.POSITIONS].detach().clone(), cell, self._signature(data))
            self.n_builds += 1
        data.update(self._cached)
        return data
