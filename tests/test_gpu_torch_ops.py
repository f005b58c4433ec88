"""The `torch.library` layer (xequinet_b200/torch_ops.py): torch.ops.xeq.* against the autograd.Function route of the
nn modules -- same kernels, so bit-identical values, first and second derivatives --, traced by
torch.compile(fullgraph=True) through the registered fake implementations and autograd formulas, and called from
TorchScript (run/jit_script.py:73 scripts the reference model; its ops must resolve through the dispatcher)."""
from typing import List, Optional

import pytest
import torch

from oracle import xpainn_oracle as orc

pytestmark = pytest.mark.gpu

import xequinet_b200 as xb  # noqa: E402
from xequinet_b200 import gemm, ops  # noqa: E402
from xequinet_b200 import torch_ops as T  # noqa: E402
from xequinet_b200.graph import build_graph  # noqa: E402

DEV = "cuda"


def _inputs(seed=0):
    cfg = orc.CONFIG_DEFAULT
    d = orc.make_molecule_batch(6, (5, 20), seed=3, with_edges=False)
    g, _, _ = build_graph(d["pos"].to(DEV), cfg.cutoff, ptr=d["ptr"].to(DEV), batch=d["batch"].to(DEV))
    dims = ops.Dims(cfg.node_dim, *cfg.muls, cfg.num_basis, cfg.cutoff)
    gen = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=gen).to(DEV)
    N = g.n_nodes
    t = dict(pos=d["pos"].to(DEV), s=r(N, dims.H), v=r(N, dims.D), x=r(N, dims.node_dim), V=r(N, dims.D), W=0.3 * r(dims.H, 20),
             b=0.3 * r(dims.H), freq=(torch.pi * torch.arange(1, 21) / 5.0).float().to(DEV))
    return cfg, d, g, dims, t


def _second_order(fn, t):
    """value, d/d(pos, s, v, W, b, freq) with create_graph, then the gradient of a scalar of those w.r.t. everything."""
    leaves = {k: v.clone().requires_grad_(True) for k, v in t.items()}
    x_out, V_out = fn(leaves)
    phi = (x_out * x_out.detach().cos()).sum() + (V_out * V_out.detach().sin()).sum()
    names = ("pos", "s", "v", "W", "b", "freq")
    first = torch.autograd.grad(phi, [leaves[k] for k in names], create_graph=True)
    psi = sum((f_ * f_.detach().cos()).sum() for f_ in first[:3])  # through d/dpos, d/ds, d/dv (the XPaiNN path)
    second = torch.autograd.grad(psi, [leaves[k] for k in ("pos", "s", "v", "x", "V", "W", "b", "freq")], allow_unused=True)
    return (x_out, V_out), first, second


def test_edge_message_op_matches_function_route_to_second_order():
    cfg, d, g, dims, t = _inputs()
    graph, meta = T.pack_graph(g)
    via_fn = _second_order(lambda L: ops.edge_message(L["x"], L["V"], L["s"], L["v"], L["pos"], L["W"], L["b"], L["freq"], g, dims), t)
    via_op = _second_order(lambda L: torch.ops.xeq.edge_message(graph, meta, T.pack_dims(dims), dims.cutoff, L["pos"], L["s"], L["v"],
                                                                L["x"], L["V"], L["W"], L["b"], L["freq"]), t)
    for a, b in zip(via_fn, via_op):
        for p, q in zip(a, b):
            if p is None or q is None:
                assert p is None and q is None
            else:
                assert torch.equal(p, q)


def test_radius_graph_op():
    d = orc.make_molecule_batch(5, (3, 12), seed=1)
    ei = torch.ops.xeq.radius_graph(d["pos"].to(DEV), 5.0, d["batch"].to(DEV))
    assert torch.equal(ei.cpu(), orc.canonical_sort(d["edge_index"])[0])


def test_linear_and_irreps_linear_ops_match_function_route():
    gen = torch.Generator().manual_seed(1)
    r = lambda *s: torch.randn(*s, generator=gen).to(DEV)
    x, W, b = r(300, 128), r(576, 128), r(576)
    V, w, bo = r(300, 480), r(128 * 128 + 64 * 64 + 32 * 32), r(128)
    for route in ("fn", "op"):
        xs, Ws, Vs, ws = (a.clone().requires_grad_(True) for a in (x, W, V, w))
        if route == "fn":
            y = gemm.linear(xs, Ws, b)
            z = gemm.irreps_linear(Vs, ws, bo, (128, 64, 32))
        else:
            y = T.linear(xs, Ws, b)
            z = torch.ops.xeq.irreps_linear(Vs, ws, bo, [128, 64, 32], False)
        loss = (y * y).sum() + (z * z).sum()
        grads = torch.autograd.grad(loss, [xs, Ws, Vs, ws])
        if route == "fn":
            ref = (y.detach(), z.detach(), *grads)
        else:
            for p, q in zip(ref, (y.detach(), z.detach(), *grads)):
                assert torch.equal(p, q)


def test_ops_trace_under_torch_compile_fullgraph():
    """One message step (Linear -> edge_message -> per-molecule sums) compiled with fullgraph=True: every xeq op has a
    fake implementation and a registered autograd formula, so Dynamo + AOTAutograd trace forward AND backward without
    a graph break; the compiled function returns the eager values and gradients."""
    cfg, d, g, dims, t = _inputs(seed=2)
    graph, meta = T.pack_graph(g)
    pd = T.pack_dims(dims)
    ptr = d["ptr"].to(DEV).to(torch.int32)
    Ws = (0.05 * torch.randn(dims.H, dims.node_dim, generator=torch.Generator().manual_seed(5))).to(DEV)

    def step(pos, x, V, v, Ws, W, b, freq):
        s = torch.ops.xeq.mm(x, Ws, False, True, 1.0)
        x2, V2 = torch.ops.xeq.edge_message(graph, meta, pd, dims.cutoff, pos, s, v, x, V, W, b, freq)
        e_atom = x2.sum(1) + (V2 * V2).sum(1)
        return torch.ops.xeq.segment_sum(e_atom, ptr)

    args = [t["pos"], t["x"], t["V"], t["v"], Ws, t["W"], t["b"], t["freq"]]
    ref_in = [a.clone().requires_grad_(True) for a in args]
    ref = step(*ref_in)
    ref_g = torch.autograd.grad(ref.sum(), ref_in)
    compiled = torch.compile(step, fullgraph=True, backend="aot_eager")
    got_in = [a.clone().requires_grad_(True) for a in args]
    got = compiled(*got_in)
    got_g = torch.autograd.grad(got.sum(), got_in)
    assert torch.equal(got, ref)
    for p, q in zip(ref_g, got_g):
        assert torch.equal(p, q)


def test_ops_are_callable_from_torchscript():
    @torch.jit.script
    def scripted(x: torch.Tensor, W: torch.Tensor, V: torch.Tensor, w: torch.Tensor, muls: List[int], ptr: torch.Tensor):
        y = torch.ops.xeq.mm(x, W, False, True, 1.0)
        bias: Optional[torch.Tensor] = None
        z = torch.ops.xeq.irreps_linear(V, w, bias, muls, False)
        return torch.ops.xeq.segment_sum(y.sum(1) + z.sum(1), ptr)

    gen = torch.Generator().manual_seed(2)
    r = lambda *s: torch.randn(*s, generator=gen).to(DEV)
    x, W, V, w = r(40, 128), r(64, 128), r(40, 480), r(128 * 128 + 64 * 64 + 32 * 32)
    ptr = torch.tensor([0, 10, 25, 40], dtype=torch.int32, device=DEV)
    got = scripted(x, W, V, w, [128, 64, 32], ptr)
    y = gemm.mm_raw(x, W, False, True)
    z = gemm.irreps_linear_raw(V, w, None, (128, 64, 32), False)
    ref = ops.segment_sum(y.sum(1) + z.sum(1), ptr, torch.repeat_interleave(torch.arange(3, device=DEV), torch.tensor([10, 15, 15], device=DEV)))
    assert torch.equal(got, ref)
