"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Dependency stubs that let the *unmodified* reference files
  /root/reference/xequinet/nn/{model,xpainn,o3layer,basic,output,rbf}.py
  /root/reference/xequinet/keys.py
  /root/reference/xequinet/data/radius_graph.py
be imported in this container, where e3nn / torch_scatter / torch_cluster /
PyG / pyscf / ... are not installed.

Only third-party primitives are restated here; every restatement cites the
reference call site whose behaviour it has to reproduce.  The stubs are used
by ``oracle/make_golden.py`` (to generate tests/golden/*.npz from the real
reference code) and by tests that run in this container.  /root/reference
does not exist on the GPU box, so nothing under ``-m gpu`` touches this file.
"""
from __future__ import annotations

import math
import re
import sys
import types
from pathlib import Path
from typing import List, Tuple

import torch
import torch.nn as nn

REFERENCE_ROOT = Path("/root/reference")


# --------------------------------------------------------------------------
# e3nn.o3 (e3nn==0.5.1, environment.yaml:139)
# --------------------------------------------------------------------------
class Irrep:
    def __init__(self, l, p=None):
        if isinstance(l, Irrep):
            l, p = l.l, l.p
        elif isinstance(l, str):
            m = re.fullmatch(r"\s*(\d+)([eoy])\s*", l)
            assert m, f"bad irrep {l!r}"
            l, p = int(m.group(1)), {"e": 1, "o": -1, "y": None}[m.group(2)]
            if p is None:
                p = (-1) ** l
        elif isinstance(l, tuple):
            l, p = l
        self.l, self.p = int(l), int(p)

    @property
    def dim(self) -> int:
        return 2 * self.l + 1

    def __repr__(self):
        return f"{self.l}{'e' if self.p == 1 else 'o'}"

    def __eq__(self, other):
        other = Irrep(other)
        return (self.l, self.p) == (other.l, other.p)

    def __hash__(self):
        return hash((self.l, self.p))

    def __iter__(self):
        yield self.l
        yield self.p


class _MulIr(tuple):
    def __new__(cls, mul, ir):
        return super().__new__(cls, (int(mul), Irrep(ir)))

    @property
    def mul(self):
        return self[0]

    @property
    def ir(self):
        return self[1]

    @property
    def dim(self):
        return self.mul * self.ir.dim


class Irreps(tuple):
    """Parses "128x0e + 64x1o + 32x2e", lists of (mul, "0e") (o3layer.py:21,87)
    and lists of (mul, (l, p)) (model.py:192)."""

    def __new__(cls, irreps=None):
        if isinstance(irreps, Irreps):
            return super().__new__(cls, irreps)
        out = []
        if irreps is None:
            irreps = []
        if isinstance(irreps, Irrep):
            irreps = [(1, irreps)]
        if isinstance(irreps, str):
            if irreps.strip():
                for tok in irreps.split("+"):
                    tok = tok.strip()
                    if "x" in tok:
                        mul, ir = tok.split("x")
                        out.append(_MulIr(int(mul), Irrep(ir)))
                    else:
                        out.append(_MulIr(1, Irrep(tok)))
        else:
            for item in irreps:
                if isinstance(item, (str, Irrep)):
                    out.append(_MulIr(1, Irrep(item)))
                else:
                    mul, ir = item
                    out.append(_MulIr(mul, Irrep(ir)))
        return super().__new__(cls, out)

    @property
    def dim(self) -> int:
        return sum(mi.dim for mi in self)

    @property
    def num_irreps(self) -> int:
        return sum(mi.mul for mi in self)

    @property
    def lmax(self) -> int:
        return max(mi.ir.l for mi in self)

    @property
    def ls(self) -> List[int]:
        return [mi.ir.l for mi in self for _ in range(mi.mul)]

    def simplify(self) -> "Irreps":
        out: List[Tuple[int, Irrep]] = []
        for mul, ir in self:
            if out and out[-1][1] == ir:
                out[-1] = (out[-1][0] + mul, ir)
            elif mul > 0:
                out.append((mul, ir))
        return Irreps(out)

    def slices(self):
        s, i = [], 0
        for mi in self:
            s.append(slice(i, i + mi.dim))
            i += mi.dim
        return s

    def __repr__(self):
        return "+".join(f"{mi.mul}x{mi.ir}" for mi in self)

    def __add__(self, other):
        return Irreps(list(self) + list(Irreps(other)))


def _sh_polynomials(x: torch.Tensor, y: torch.Tensor, z: torch.Tensor, lmax: int):
    """e3nn 0.5.1 o3/_spherical_harmonics.py: real SH, 'integral'-free base
    polynomials; the caller applies sqrt(2l+1) ('component')."""
    out = [[torch.ones_like(x)]]
    if lmax >= 1:
        out.append([x, y, z])
    if lmax >= 2:
        s3 = math.sqrt(3.0)
        out.append(
            [
                s3 * x * z,
                s3 * x * y,
                y * y - 0.5 * (x * x + z * z),
                s3 * y * z,
                (s3 / 2.0) * (z * z - x * x),
            ]
        )
    assert lmax <= 2, "oracle stub restates e3nn spherical harmonics up to l=2"
    return out


class SphericalHarmonics(nn.Module):
    """o3.SphericalHarmonics(irreps_out, normalize=True, normalization="component")
    as called at nn/xpainn.py:49-51,71-74.  When ``irreps_out`` carries
    multiplicities every Y_l is repeated ``mul`` times (layout [u][m])."""

    def __init__(self, irreps_out, normalize: bool, normalization: str = "integral"):
        super().__init__()
        self.irreps_out = Irreps(irreps_out)
        self.normalize = normalize
        assert normalization == "component"
        self.normalization = normalization

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.normalize:
            x = torch.nn.functional.normalize(x, dim=-1)
        polys = _sh_polynomials(x[..., 0], x[..., 1], x[..., 2], self.irreps_out.lmax)
        pieces = []
        for mul, ir in self.irreps_out:
            yl = torch.stack(polys[ir.l], dim=-1) * math.sqrt(2 * ir.l + 1)
            pieces.append(yl.repeat(*([1] * (yl.dim() - 1)), mul))
        return torch.cat(pieces, dim=-1)


class Linear(nn.Module):
    """o3.Linear(irreps_in, irreps_out, biases=True) (nn/xpainn.py:186-187):
    per matching irrep pair  out[z,w,i] = sum_u W[u,w] in[z,u,i] / sqrt(mul_in)
    flat weight, blocks in irreps order, each row-major [u,w]; bias on 0e outputs."""

    def __init__(self, irreps_in, irreps_out, biases: bool = False, **kwargs):
        super().__init__()
        self.irreps_in = Irreps(irreps_in)
        self.irreps_out = Irreps(irreps_out)
        self.paths = []  # (i_in, i_out)
        nw = 0
        for io, (mo, iro) in enumerate(self.irreps_out):
            for ii, (mi, iri) in enumerate(self.irreps_in):
                if iri == iro:
                    self.paths.append((ii, io, nw, mi, mo))
                    nw += mi * mo
        # e3nn orders instructions by (i_in, i_out); identical here because
        # each irrep type appears once on this path.
        self.paths.sort(key=lambda p: (p[0], p[1]))
        off = 0
        fixed = []
        for ii, io, _, mi, mo in self.paths:
            fixed.append((ii, io, off, mi, mo))
            off += mi * mo
        self.paths = fixed
        self.weight = nn.Parameter(torch.randn(off))
        nb = sum(mo for mo, iro in self.irreps_out if iro.l == 0 and iro.p == 1) if biases else 0
        self.bias = nn.Parameter(torch.zeros(nb))
        self.has_bias = biases
        self.register_buffer("output_mask", torch.ones(self.irreps_out.dim))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        sin, sout = self.irreps_in.slices(), self.irreps_out.slices()
        outs = [None] * len(self.irreps_out)
        fan_in = [0] * len(self.irreps_out)
        for ii, io, off, mi, mo in self.paths:
            fan_in[io] += mi
        for ii, io, off, mi, mo in self.paths:
            d = self.irreps_in[ii].ir.dim
            w = self.weight[off : off + mi * mo].view(mi, mo)
            xi = x[..., sin[ii]].reshape(*x.shape[:-1], mi, d)
            y = torch.einsum("uw,...ui->...wi", w, xi) / math.sqrt(fan_in[io])
            y = y.reshape(*x.shape[:-1], mo * d)
            outs[io] = y if outs[io] is None else outs[io] + y
        boff = 0
        for io, (mo, iro) in enumerate(self.irreps_out):
            if outs[io] is None:
                outs[io] = x.new_zeros(*x.shape[:-1], mo * iro.dim)
            if self.has_bias and iro.l == 0 and iro.p == 1:
                outs[io] = outs[io] + self.bias[boff : boff + mo]
                boff += mo
        return torch.cat(outs, dim=-1)


class ElementwiseTensorProduct(nn.Module):
    """o3.ElementwiseTensorProduct(irreps, "Mx0e") (nn/xpainn.py:119-121,191-193,
    nn/o3layer.py:130-132): all-scalar second operand == broadcast multiply of
    each irrep by its own gate (w3j(l,0,l)*sqrt(2l+1) == 1)."""

    def __init__(self, irreps_in1, irreps_in2, **kwargs):
        super().__init__()
        self.irreps_in1 = Irreps(irreps_in1)
        self.irreps_in2 = Irreps(irreps_in2)
        assert all(ir.l == 0 for _, ir in self.irreps_in2)
        assert self.irreps_in1.num_irreps == self.irreps_in2.num_irreps
        reps = []
        for mul, ir in self.irreps_in1:
            reps.extend([ir.dim] * mul)
        self.register_buffer("_reps", torch.tensor(reps, dtype=torch.long), persistent=False)
        self.register_buffer("weight", torch.Tensor())
        self.register_buffer("output_mask", torch.ones(self.irreps_in1.dim))

    def forward(self, x: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
        gate = torch.repeat_interleave(g, self._reps, dim=-1)
        return x * gate


class TensorProduct(nn.Module):
    """o3.TensorProduct(irreps, irreps, "Mx0e", [(i,i,i,"uuu",False,ir.dim)],
    irrep_normalization="component") (nn/o3layer.py:23-29, 89-95): the w3j(l,l,0)
    path with path_weight 2l+1 has coefficient exactly 1 -> per-irrep dot product."""

    def __init__(self, irreps_in1, irreps_in2, irreps_out, instructions, irrep_normalization=None, **kwargs):
        super().__init__()
        self.irreps_in1 = Irreps(irreps_in1)
        self.irreps_in2 = Irreps(irreps_in2)
        self.irreps_out = Irreps(irreps_out)
        for k, ins in enumerate(instructions):
            i1, i2, io, mode, has_w, pw = ins
            assert i1 == i2 == io == k and mode == "uuu" and not has_w
            assert pw == self.irreps_in1[i1].ir.dim and self.irreps_out[io].ir.l == 0
        assert irrep_normalization == "component"
        self.register_buffer("weight", torch.Tensor())
        self.register_buffer("output_mask", torch.ones(self.irreps_out.dim))

    def forward(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        outs = []
        for sl, (mul, ir) in zip(self.irreps_in1.slices(), self.irreps_in1):
            pa = a[..., sl].reshape(*a.shape[:-1], mul, ir.dim)
            pb = b[..., sl].reshape(*b.shape[:-1], mul, ir.dim)
            outs.append((pa * pb).sum(-1))
        return torch.cat(outs, dim=-1)


class ReducedTensorProducts:  # only named at import time of out-of-scope heads
    def __init__(self, *a, **k):
        raise NotImplementedError("out of scope for the XPaiNN energy/forces path")


def _compile_mode(mode):
    def deco(cls):
        return cls

    return deco


# --------------------------------------------------------------------------
# torch_scatter (torch-scatter==2.1.2, environment.yaml:108); nn/output.py:73,124
# --------------------------------------------------------------------------
def scatter_sum(src: torch.Tensor, index: torch.Tensor, dim: int = 0, out=None, dim_size=None):
    assert dim == 0
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() else 0
    res = src.new_zeros((dim_size,) + tuple(src.shape[1:]))
    return res.index_add(0, index, src)


def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    assert reduce in ("sum", "add")
    return scatter_sum(src, index, dim=dim, dim_size=dim_size)


# --------------------------------------------------------------------------
# xequinet.utils (only what nn/ needs): utils/qc.py:222-237
# --------------------------------------------------------------------------
_ELEMENTS = (
    "H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn "
    "Ga Ge As Se Br Kr Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba "
    "La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn"
).split()


def get_embedding_tensor(embed_basis: str = "gfn2-xtb", aux_basis: str = "aux28") -> torch.Tensor:
    pre = REFERENCE_ROOT / "xequinet" / "utils" / "pre_computed" / f"{embed_basis}_{aux_basis}.pt"
    embed_dict = torch.load(pre)
    ten = torch.stack([embed_dict[a] for a in _ELEMENTS])
    ten = torch.cat([torch.zeros(1, ten.shape[-1], dtype=ten.dtype), ten])
    return ten.to(torch.get_default_dtype())


def install() -> None:
    """Put the stubs into sys.modules and /root/reference on sys.path."""
    if "e3nn" in sys.modules and getattr(sys.modules["e3nn"], "__xeq_stub__", False):
        return
    if not REFERENCE_ROOT.exists():
        raise RuntimeError("/root/reference is not present (GPU box?) -- use tests/golden fixtures")

    e3nn = types.ModuleType("e3nn")
    e3nn.__xeq_stub__ = True
    o3 = types.ModuleType("e3nn.o3")
    for name, obj in dict(
        Irrep=Irrep,
        Irreps=Irreps,
        SphericalHarmonics=SphericalHarmonics,
        Linear=Linear,
        ElementwiseTensorProduct=ElementwiseTensorProduct,
        TensorProduct=TensorProduct,
        ReducedTensorProducts=ReducedTensorProducts,
    ).items():
        setattr(o3, name, obj)
    util = types.ModuleType("e3nn.util")
    jit = types.ModuleType("e3nn.util.jit")
    jit.compile_mode = _compile_mode
    e3nn.o3, e3nn.util, util.jit = o3, util, jit
    sys.modules.update({"e3nn": e3nn, "e3nn.o3": o3, "e3nn.util": util, "e3nn.util.jit": jit})

    ts = types.ModuleType("torch_scatter")
    ts.scatter, ts.scatter_sum = scatter, scatter_sum
    sys.modules["torch_scatter"] = ts

    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, str(REFERENCE_ROOT))
    import xequinet  # dependency-free top level (xequinet/__init__.py:1-12)

    utils = types.ModuleType("xequinet.utils")
    qc = types.ModuleType("xequinet.utils.qc")
    qc.ATOM_MASS = [0.0] * 120
    utils.qc = qc
    utils.get_embedding_tensor = get_embedding_tensor
    utils.set_default_units = lambda *a, **k: None
    sys.modules["xequinet.utils"] = utils
    sys.modules["xequinet.utils.qc"] = qc
    xequinet.utils = utils

    data = types.ModuleType("xequinet.data")
    data.NeighborTransform = object
    sys.modules["xequinet.data"] = data
    xequinet.data = data


def reference_resolve_model():
    """The reference's own factory, nn/model.py:310-318, running unmodified code."""
    install()
    from xequinet.nn.model import resolve_model

    return resolve_model


def reference_radius_graph_pbc():
    """The reference's own PBC neighbour search, data/radius_graph.py:35-192."""
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "_xeq_ref_radius_graph", REFERENCE_ROOT / "xequinet" / "data" / "radius_graph.py"
    )
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.radius_graph_pbc, mod.single_radius_graph
