"""CPU tests of the device math: the headers the sm_100a kernels are built from
(xequinet_b200/csrc/edge_math.cuh, edge_thread.cuh) are compiled for the host in float64
(tests/host_emul/edge_emul.cpp) and checked against torch autograd of the oracle's edge
message -- value, first derivatives (K2b) and second derivatives (K2bb)."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import xpainn_oracle as orc

HERE = Path(__file__).resolve().parent
SRC = HERE / "host_emul" / "edge_emul.cpp"
LIB = HERE / "host_emul" / "libedge_emul.so"
HDRS = [HERE.parent / "xequinet_b200" / "csrc" / n for n in ("edge_math.cuh", "edge_thread.cuh")]


@pytest.fixture(scope="module")
def emul():
    newest = max(p.stat().st_mtime for p in [SRC] + HDRS)
    if not LIB.exists() or LIB.stat().st_mtime < newest:
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", str(LIB), str(SRC)])
    return ctypes.CDLL(str(LIB))


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _graph(kind):
    if kind == "mol":
        d = orc.make_molecule_batch(3, (5, 9), seed=21, dtype=torch.float64)
        ei, co, cell = d["edge_index"], None, None
    else:
        d = orc.make_small_pbc(5, 4.5, seed=5, dtype=torch.float64, triclinic=True)
        ei, co = orc.radius_graph_pbc(d["pos"], torch.tensor([5]), d["pbc"], d["cell"], 5.0)
        cell = d["cell"]
    N = d["pos"].shape[0]
    E = ei.shape[1]
    rowptr = torch.zeros(N + 1, dtype=torch.int32)
    rowptr[1:] = torch.cumsum(torch.bincount(ei[0], minlength=N), 0)
    col = ei[1].to(torch.int32).contiguous()
    # transposed structure: slots grouped by neighbor, ordered by edge id
    order = torch.sort(ei[1], stable=True)[1]
    t_rowptr = torch.zeros(N + 1, dtype=torch.int32)
    t_rowptr[1:] = torch.cumsum(torch.bincount(ei[1], minlength=N), 0)
    t_row = ei[0][order].to(torch.int32).contiguous()
    t_eid = order.to(torch.int32).contiguous()
    offs = None
    if co is not None:
        offs = torch.zeros(E, 4, dtype=torch.int8)
        offs[:, :3] = co.to(torch.int8)
    return d, ei, co, cell, rowptr, col, t_rowptr, t_row, t_eid, offs


@pytest.mark.parametrize("decomposition", [0, 1], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("kind", ["mol", "pbc"])
def test_emulated_kernels_match_autograd(emul, kind, decomposition):
    """decomposition 0 = the per-thread math of the SIMT kernels; 1 = that of the tcgen05 kernels (filter values
    supplied from outside, radial-only d/dr coefficients, weight gradients as outer products over the edges)."""
    emul.emul_set_decomposition(decomposition)
    cfg = orc.XPaiNNConfig(node_dim=32, muls=(32, 32, 32), num_basis=20)
    d, ei, co, cell, rowptr, col, t_rowptr, t_row, t_eid, offs = _graph(kind)
    N, E = d["pos"].shape[0], ei.shape[1]
    C, M, D, H, B = cfg.node_dim, cfg.M, cfg.D, cfg.H_msg, cfg.num_basis
    g = torch.Generator().manual_seed(0)
    rnd = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    pos = d["pos"].clone().requires_grad_(True)
    s = rnd(N, H).requires_grad_(True)
    v = rnd(N, D).requires_grad_(True)  # e3nn layout
    x, V = rnd(N, C), rnd(N, D)
    W = (0.3 * rnd(H, B)).requires_grad_(True)
    b = (0.3 * rnd(H)).requires_grad_(True)
    freq = (torch.pi * torch.arange(1, B + 1, dtype=torch.float64) / cfg.cutoff + 0.1 * rnd(B)).requires_grad_(True)
    gx = rnd(N, C).requires_grad_(True)
    gV = rnd(N, D).requires_grad_(True)
    a_s, a_v, a_pos = rnd(N, H), rnd(N, D), rnd(N, 3)
    cell2 = None if cell is None else cell[0].contiguous()

    xo, Vo = orc.edge_message(x, V, s, v, pos, W, b, freq, ei, cfg, cell, co, d["batch"])
    Phi = (gx * xo).sum() + (gV * Vo).sum()
    first = torch.autograd.grad(Phi, [s, v, pos, W, b, freq], create_graph=True)
    Psi = (a_s * first[0]).sum() + (a_v * first[1]).sum() + (a_pos * first[2]).sum()
    second = torch.autograd.grad(Psi, [gx, gV, s, v, pos, W, b, freq])

    dims = (ctypes.c_int * 5)(C, *cfg.muls, B)
    cm = lambda t: orc.to_cm(t.detach(), cfg).contiguous()
    z = lambda *sh: torch.zeros(*sh, dtype=torch.float64)
    sD, vD, posD, WD, bD, fD = (t.detach().contiguous() for t in (s, v, pos, W, b, freq))
    v_cm, a_v_cm, gV_cm = cm(v), cm(a_v), cm(gV)

    # forward
    ox, oV = x.clone(), cm(V)
    emul.emul_center_pass(dims, ctypes.c_double(cfg.cutoff), N, _ptr(rowptr), _ptr(col), _ptr(offs), _ptr(cell2),
                          _ptr(posD), _ptr(sD), _ptr(v_cm), _ptr(WD), _ptr(bD), _ptr(fD), None, None, None,
                          _ptr(ox), _ptr(oV))
    torch.testing.assert_close(ox, xo.detach(), rtol=1e-11, atol=1e-11)
    torch.testing.assert_close(orc.from_cm(oV, cfg), Vo.detach(), rtol=1e-11, atol=1e-11)

    # first derivatives
    o_s, o_v, o_p, o_W, o_b, o_f = z(N, H), z(N, D), z(N, 3), z(H, B), z(H), z(B)
    gxD = gx.detach().contiguous()
    emul.emul_neighbor_pass(dims, ctypes.c_double(cfg.cutoff), 0, N, E, _ptr(rowptr), _ptr(t_rowptr), _ptr(t_row),
                            _ptr(t_eid), _ptr(offs), _ptr(cell2), _ptr(posD), _ptr(sD), _ptr(v_cm), _ptr(WD), _ptr(bD),
                            _ptr(fD), _ptr(gxD), _ptr(gV_cm), None, None, None,
                            _ptr(o_s), _ptr(o_v), _ptr(o_p), _ptr(o_W), _ptr(o_b), _ptr(o_f))
    for got, ref, name in zip((o_s, orc.from_cm(o_v, cfg), o_p, o_W, o_b, o_f), first, "s v pos W b f".split()):
        torch.testing.assert_close(got, ref.detach(), rtol=1e-9, atol=1e-9, msg=lambda m, n=name: f"first/{n}: {m}")

    # second derivatives: JVP half (o_gx, o_gV) ...
    jx, jV = z(N, C), z(N, D)
    emul.emul_center_pass(dims, ctypes.c_double(cfg.cutoff), N, _ptr(rowptr), _ptr(col), _ptr(offs), _ptr(cell2),
                          _ptr(posD), _ptr(sD), _ptr(v_cm), _ptr(WD), _ptr(bD), _ptr(fD), _ptr(a_s), _ptr(a_v_cm),
                          _ptr(a_pos), _ptr(jx), _ptr(jV))
    torch.testing.assert_close(jx, second[0], rtol=1e-9, atol=1e-9)
    torch.testing.assert_close(orc.from_cm(jV, cfg), second[1], rtol=1e-9, atol=1e-9)
    # ... and reverse half
    emul.emul_neighbor_pass(dims, ctypes.c_double(cfg.cutoff), 1, N, E, _ptr(rowptr), _ptr(t_rowptr), _ptr(t_row),
                            _ptr(t_eid), _ptr(offs), _ptr(cell2), _ptr(posD), _ptr(sD), _ptr(v_cm), _ptr(WD), _ptr(bD),
                            _ptr(fD), _ptr(gxD), _ptr(gV_cm), _ptr(a_s), _ptr(a_v_cm), _ptr(a_pos),
                            _ptr(o_s), _ptr(o_v), _ptr(o_p), _ptr(o_W), _ptr(o_b), _ptr(o_f))
    for got, ref, name in zip((o_s, orc.from_cm(o_v, cfg), o_p, o_W, o_b, o_f), second[2:], "s v pos W b f".split()):
        torch.testing.assert_close(got, ref, rtol=1e-8, atol=1e-8, msg=lambda m, n=name: f"second/{n}: {m}")
