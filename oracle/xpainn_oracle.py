"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the XPaiNN hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product path (xequinet_b200/) never does.

A plain-torch functional restatement of the reference's algorithm, written from
the math in SURVEY.md Appendix A, in the reference's own (e3nn) feature layout:
[N, D] = [mul0 scalars | mul1 x (m=-1,0,1) | mul2 x (m=-2..2)], mul-major, m fastest.
It consumes a state_dict with the reference's parameter names.

Parity status: PINNED.  The reference has no golden vectors of its own
(SURVEY.md 4), so the pin is the reference's own code run in the build container
through dependency stubs (oracle/ref_stubs.py, oracle/make_golden.py); its outputs
are committed under tests/golden/ and this restatement is checked against them in
tests/test_oracle_golden.py.  The third-party primitives (e3nn 0.5.1, torch-scatter
2.1.2, torch-cluster 1.6.3; environment.yaml:105-139) are restated from their
published definitions at the reference's call sites.

Each function cites the reference file:line it follows (paths relative to
/root/reference/xequinet/).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# configuration (nn/model.py:57-70)
# ----------------------------------------------------------------------------
@dataclass(frozen=True)
class XPaiNNConfig:
    node_dim: int = 128
    muls: Tuple[int, int, int] = (128, 64, 32)  # "128x0e + 64x1o + 32x2e"
    num_basis: int = 20
    cutoff: float = 5.0
    action_blocks: int = 3
    hidden_dim: int = 64  # EnergyOut, nn/output.py:83
    embed_dim: int = 56  # aux56

    @property
    def M(self) -> int:  # num_irreps
        return sum(self.muls)

    @property
    def D(self) -> int:  # irreps.dim
        return self.muls[0] + 3 * self.muls[1] + 5 * self.muls[2]

    @property
    def H_msg(self) -> int:  # nn/xpainn.py:108
        return self.node_dim + 2 * self.M

    @property
    def H_upd(self) -> int:  # nn/xpainn.py:184
        return 2 * self.node_dim + self.M

    @property
    def irreps_str(self) -> str:
        return f"{self.muls[0]}x0e + {self.muls[1]}x1o + {self.muls[2]}x2e"

    def model_kwargs(self) -> dict:
        return dict(
            node_dim=self.node_dim,
            node_irreps=self.irreps_str,
            num_basis=self.num_basis,
            cutoff=self.cutoff,
            action_blocks=self.action_blocks,
        )


CONFIG_DEFAULT = XPaiNNConfig()
CONFIG_C4 = XPaiNNConfig(node_dim=256, muls=(256, 128, 64))


# ----------------------------------------------------------------------------
# synthetic parameters: reproducible on any box from a seed (no reference needed)
# ----------------------------------------------------------------------------
def state_dict_spec(cfg: XPaiNNConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, kind) for every learnable entry + buffers, in the reference's
    state_dict naming (observed by instantiating the reference)."""
    C, M, D, B = cfg.node_dim, cfg.M, cfg.D, cfg.num_basis
    Hm, Hu = cfg.H_msg, cfg.H_upd
    nw = sum(m * m for m in cfg.muls)
    spec: List[Tuple[str, Tuple[int, ...], str]] = [
        ("mods.embedding.embedding.1.weight", (C, cfg.embed_dim), "lin"),
        ("mods.embedding.embedding.1.bias", (C,), "bias"),
        ("mods.embedding.rbf.freq", (1, B), "freq"),
    ]
    for i in range(cfg.action_blocks):
        p = f"mods.message_{i}."
        spec += [
            (p + "scalar_mlp.0.weight", (C, C), "lin"),
            (p + "scalar_mlp.0.bias", (C,), "bias"),
            (p + "scalar_mlp.2.weight", (Hm, C), "lin"),
            (p + "scalar_mlp.2.bias", (Hm,), "bias"),
            (p + "rbf_lin.weight", (Hm, B), "lin"),
            (p + "rbf_lin.bias", (Hm,), "bias"),
            (p + "norm.weight", (C,), "gain"),
            (p + "norm.bias", (C,), "bias"),
            (p + "o3norm.affine_weight", (M,), "gain"),
            (p + "o3norm.affine_bias", (cfg.muls[0],), "bias"),
        ]
        p = f"mods.update_{i}."
        spec += [
            (p + "update_U.weight", (nw,), "o3"),
            (p + "update_U.bias", (cfg.muls[0],), "bias"),
            (p + "update_V.weight", (nw,), "o3"),
            (p + "update_V.bias", (cfg.muls[0],), "bias"),
            (p + "dot_lin.weight", (C, M), "lin"),
            (p + "update_mlp.0.weight", (C, C + M), "lin"),
            (p + "update_mlp.0.bias", (C,), "bias"),
            (p + "update_mlp.2.weight", (Hu, C), "lin"),
            (p + "update_mlp.2.bias", (Hu,), "bias"),
            (p + "norm.weight", (C,), "gain"),
            (p + "norm.bias", (C,), "bias"),
            (p + "o3norm.affine_weight", (M,), "gain"),
            (p + "o3norm.affine_bias", (cfg.muls[0],), "bias"),
        ]
    p = "mods.output_energy."
    spec += [
        (p + "out_mlp.0.weight", (cfg.hidden_dim, C), "lin"),
        (p + "out_mlp.0.bias", (cfg.hidden_dim,), "bias"),
        (p + "out_mlp.2.weight", (1, cfg.hidden_dim), "lin"),
        (p + "out_mlp.2.bias", (1,), "bias"),
    ]
    return spec


def synthetic_state_dict(cfg: XPaiNNConfig, seed: int = 1234, dtype=torch.float32, spec=None) -> Dict[str, torch.Tensor]:
    """Default-init-like values plus an N(0, 0.1^2) perturbation of every parameter,
    so zero biases / unit gains cannot hide bugs.  Generated in fp64
    from a CPU generator, then cast, so every box and dtype sees the same numbers."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, shape, kind in (spec if spec is not None else state_dict_spec(cfg)):
        if kind == "lin":  # nn.Linear default: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
            bound = 1.0 / math.sqrt(shape[1])
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound
        elif kind == "o3":  # e3nn o3.Linear: randn
            t = torch.randn(shape, generator=g, dtype=torch.float64)
        elif kind == "gain":
            t = torch.ones(shape, dtype=torch.float64)
        elif kind == "bias":
            t = torch.zeros(shape, dtype=torch.float64)
        elif kind == "freq":  # nn/rbf.py:143
            t = (math.pi * torch.arange(1, shape[1] + 1, dtype=torch.float64) / cfg.cutoff).view(shape)
        else:
            raise ValueError(kind)
        t = t + 0.1 * torch.randn(shape, generator=g, dtype=torch.float64)
        if name == "mods.output_energy.out_mlp.2.weight":
            # keep |F| = O(1) eV/A so the 1e-4 eV/A absolute force tolerance is meaningful
            t = t * 0.01
        sd[name] = t.to(dtype)
    return sd


# ----------------------------------------------------------------------------
# geometry (nn/basic.py:60-140, nn/xpainn.py:66-75, nn/rbf.py:43-57,134-152)
# ----------------------------------------------------------------------------
def edge_vectors(
    pos: torch.Tensor,
    edge_index: torch.Tensor,
    cell: Optional[torch.Tensor] = None,
    cell_offsets: Optional[torch.Tensor] = None,
    batch: Optional[torch.Tensor] = None,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """nn/basic.py:110-131: vector = pos[center] - pos[neighbor] - offsets @ cell[graph]."""
    center, neighbor = edge_index[0], edge_index[1]
    vec = pos.index_select(0, center) - pos.index_select(0, neighbor)
    if cell is not None:
        if cell.dim() == 2:
            cell = cell.unsqueeze(0)
        if cell.shape[0] == 1:
            shifts = torch.einsum("ni,ij->nj", cell_offsets.to(pos.dtype), cell[0])
        else:
            cb = cell.index_select(0, batch.index_select(0, neighbor))
            shifts = torch.einsum("ni,nij->nj", cell_offsets.to(pos.dtype), cb)
        vec = vec - shifts
    dist = torch.linalg.norm(vec, dim=-1)
    return vec, dist


def bessel_rbf(dist: torch.Tensor, freq: torch.Tensor, cutoff: float) -> torch.Tensor:
    """nn/rbf.py:143-150: sqrt(2/rc) * sin(f d) / (d + 1e-5); dist [E,1], freq [1,B]."""
    return math.sqrt(2.0 / cutoff) * torch.sin(freq * dist) / (dist + 1e-5)


def cosine_cutoff(dist: torch.Tensor, cutoff: float) -> torch.Tensor:
    """nn/rbf.py:43-57."""
    return torch.where(dist < cutoff, 0.5 * (torch.cos(math.pi * dist / cutoff) + 1.0), torch.zeros_like(dist))


def spherical_harmonics_l2(vec: torch.Tensor) -> torch.Tensor:
    """e3nn o3.SphericalHarmonics(normalize=True, normalization='component') up to l=2
    on vec[:, [1,2,0]] (nn/xpainn.py:71-74).  Returns the 9 unique numbers [E,9]."""
    u = F.normalize(vec, dim=-1)
    x, y, z = u[:, 1], u[:, 2], u[:, 0]  # e3nn argument order (x,y,z) = (vec_y, vec_z, vec_x)
    s3, s5, s15 = math.sqrt(3.0), math.sqrt(5.0), math.sqrt(15.0)
    return torch.stack(
        [
            torch.ones_like(x),
            s3 * x,
            s3 * y,
            s3 * z,
            s15 * x * z,
            s15 * x * y,
            s5 * (y * y - 0.5 * (x * x + z * z)),
            s15 * y * z,
            (s15 / 2.0) * (z * z - x * x),
        ],
        dim=-1,
    )


# ----------------------------------------------------------------------------
# per-irrep helpers in the e3nn layout (nn/o3layer.py)
# ----------------------------------------------------------------------------
def _blocks(cfg: XPaiNNConfig):
    m0, m1, m2 = cfg.muls
    return [(0, m0, 1), (m0, m1, 3), (m0 + 3 * m1, m2, 5)]  # (offset, mul, 2l+1)


def expand_gate(g: torch.Tensor, cfg: XPaiNNConfig) -> torch.Tensor:
    """ElementwiseTensorProduct(irreps, 'Mx0e') gate expansion: repeat each per-irrep
    gate over its 2l+1 components (nn/xpainn.py:119-121)."""
    m0, m1, m2 = cfg.muls
    return torch.cat(
        [
            g[:, :m0],
            g[:, m0 : m0 + m1].repeat_interleave(3, dim=1),
            g[:, m0 + m1 :].repeat_interleave(5, dim=1),
        ],
        dim=1,
    )


def irrep_dot(a: torch.Tensor, b: torch.Tensor, cfg: XPaiNNConfig) -> torch.Tensor:
    """EquivariantDot / Invariant(squared=True): per-irrep sum_m a*b (nn/o3layer.py:23-29,104-109)."""
    outs = []
    for off, mul, d in _blocks(cfg):
        outs.append((a[:, off : off + mul * d] * b[:, off : off + mul * d]).view(-1, mul, d).sum(-1))
    return torch.cat(outs, dim=1)


def invariant(a: torch.Tensor, cfg: XPaiNNConfig, eps: float = 1e-5) -> torch.Tensor:
    """Invariant(squared=False): sqrt(q + eps^2) - eps (nn/o3layer.py:40-44)."""
    return torch.sqrt(irrep_dot(a, a, cfg) + eps**2) - eps


def equivariant_layer_norm(V: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, cfg: XPaiNNConfig, eps: float = 1e-5):
    """EquivariantLayerNorm.forward (nn/o3layer.py:145-171)."""
    m0 = cfg.muls[0]
    scal = V[:, :m0]
    z = torch.cat([scal - scal.mean(dim=1, keepdim=True), V[:, m0:]], dim=1)
    q = irrep_dot(z, z, cfg)
    rho = torch.reciprocal(torch.sqrt(q.mean(dim=1, keepdim=True) + eps))
    out = z * rho * expand_gate(gamma.unsqueeze(0), cfg)
    return torch.cat([out[:, :m0] + beta.unsqueeze(0), out[:, m0:]], dim=1)


def o3_linear(Vn: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, cfg: XPaiNNConfig) -> torch.Tensor:
    """e3nn o3.Linear(irreps->irreps, biases=True) (nn/xpainn.py:186-187,211-212):
    out[z,w,m] = sum_u W_l[u,w] in[z,u,m] / sqrt(mul_l) (+ bias on 0e)."""
    outs = []
    woff = 0
    for l, (off, mul, d) in enumerate(_blocks(cfg)):
        W = weight[woff : woff + mul * mul].view(mul, mul)
        woff += mul * mul
        x = Vn[:, off : off + mul * d].view(-1, mul, d)
        y = torch.einsum("uw,zui->zwi", W, x) / math.sqrt(mul)
        if l == 0:
            y = y + bias.view(1, mul, 1)
        outs.append(y.reshape(-1, mul * d))
    return torch.cat(outs, dim=1)


# ----------------------------------------------------------------------------
# the model (nn/model.py:26-46, nn/xpainn.py, nn/output.py:114-128)
# ----------------------------------------------------------------------------
def xpainn_features(
    sd: Dict[str, torch.Tensor],
    embed_table: torch.Tensor,
    data: Dict[str, torch.Tensor],
    cfg: XPaiNNConfig = CONFIG_DEFAULT,
):
    """BaseModel.forward up to (not including) the read-out heads: embedding, optional charge / spin conditioning,
    message / update blocks.  Returns (x [N,C], V [N,D] in the e3nn layout, batch [N], G)."""
    pos = data["pos"]
    Z = data["atomic_numbers"].long()
    ei = data["edge_index"]
    batch = data.get("batch")
    if batch is None:
        batch = torch.zeros(pos.shape[0], dtype=torch.long, device=pos.device)
    G = int(data["ptr"].numel() - 1) if "ptr" in data else int(batch.max().item()) + 1
    center, neighbor = ei[0], ei[1]
    m0, m1, m2 = cfg.muls
    M, C = cfg.M, cfg.node_dim

    vec, dist = edge_vectors(pos, ei, data.get("cell"), data.get("cell_offsets"), batch)
    d1 = dist.unsqueeze(-1)
    # XEmbedding.forward, nn/xpainn.py:55-83
    x = F.linear(embed_table.to(pos.dtype)[Z], sd["mods.embedding.embedding.1.weight"], sd["mods.embedding.embedding.1.bias"])
    rbf = bessel_rbf(d1, sd["mods.embedding.rbf.freq"], cfg.cutoff)
    fcut = cosine_cutoff(d1, cfg.cutoff)
    Y = spherical_harmonics_l2(vec)  # [E,9]
    rsh = torch.cat(
        [Y[:, 0:1].repeat(1, m0), Y[:, 1:4].repeat(1, m1), Y[:, 4:9].repeat(1, m2)], dim=1
    )  # [E,D], layout [u][m]
    V = pos.new_zeros(pos.shape[0], cfg.D)
    # nn/model.py:85-96: conditioning modules sit between the embedding and message_0
    if "mods.charge_embedding.linear_q.weight" in sd and "charge" in data:
        x = charge_embedding(sd, "mods.charge_embedding.", x, data["charge"], batch, G)
    if "mods.spin_embedding.linear_q.weight" in sd and "spin" in data:
        x = spin_embedding(sd, "mods.spin_embedding.", x, data["spin"], batch, G)

    for i in range(cfg.action_blocks):
        # XPainnMessage.forward, nn/xpainn.py:128-161
        p = f"mods.message_{i}."
        xn = F.layer_norm(x, (C,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
        vn = equivariant_layer_norm(V, sd[p + "o3norm.affine_weight"], sd[p + "o3norm.affine_bias"], cfg)
        s = F.linear(F.silu(F.linear(xn, sd[p + "scalar_mlp.0.weight"], sd[p + "scalar_mlp.0.bias"])),
                     sd[p + "scalar_mlp.2.weight"], sd[p + "scalar_mlp.2.bias"])
        w = F.linear(rbf, sd[p + "rbf_lin.weight"], sd[p + "rbf_lin.bias"]) * fcut
        h = s.index_select(0, neighbor) * w
        g_state, g_edge, m_s = h[:, :M], h[:, M : 2 * M], h[:, 2 * M :]
        m_e = vn.index_select(0, neighbor) * expand_gate(g_state, cfg) + rsh * expand_gate(g_edge, cfg)
        x = x.index_add(0, center, m_s)
        V = V.index_add(0, center, m_e)
        # XPainnUpdate.forward, nn/xpainn.py:206-231
        p = f"mods.update_{i}."
        xn = F.layer_norm(x, (C,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
        vn = equivariant_layer_norm(V, sd[p + "o3norm.affine_weight"], sd[p + "o3norm.affine_bias"], cfg)
        U = o3_linear(vn, sd[p + "update_U.weight"], sd[p + "update_U.bias"], cfg)
        W = o3_linear(vn, sd[p + "update_V.weight"], sd[p + "update_V.bias"], cfg)
        n = invariant(W, cfg)
        a = F.linear(F.silu(F.linear(torch.cat([xn, n], dim=1), sd[p + "update_mlp.0.weight"], sd[p + "update_mlp.0.bias"])),
                     sd[p + "update_mlp.2.weight"], sd[p + "update_mlp.2.bias"])
        a_vv, a_sv, a_ss = a[:, :M], a[:, M : M + C], a[:, M + C :]
        dV = U * expand_gate(a_vv, cfg)
        t = F.linear(irrep_dot(U, W, cfg), sd[p + "dot_lin.weight"])
        x = x + a_sv * t + a_ss
        V = V + dV

    return x, V, batch, G


def xpainn_energy(
    sd: Dict[str, torch.Tensor],
    embed_table: torch.Tensor,
    data: Dict[str, torch.Tensor],
    cfg: XPaiNNConfig = CONFIG_DEFAULT,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """BaseModel.forward up to the energy head.  Returns (energy [G], atomic_energies [N])."""
    x, V, batch, G = xpainn_features(sd, embed_table, data, cfg)
    # EnergyOut.forward, nn/output.py:114-128
    p = "mods.output_energy."
    e_atom = F.linear(F.silu(F.linear(x, sd[p + "out_mlp.0.weight"], sd[p + "out_mlp.0.bias"])),
                      sd[p + "out_mlp.2.weight"], sd[p + "out_mlp.2.bias"]).reshape(-1)
    energy = torch.zeros(G, dtype=e_atom.dtype, device=e_atom.device).index_add(0, batch, e_atom)
    return energy, e_atom


# ----------------------------------------------------------------------------
# optional conditioning and read-out heads (nn/electronic.py, nn/output.py) -- SURVEY.md 8f rank 4
# ----------------------------------------------------------------------------
def _mlp(sd, p, x):
    return F.linear(F.silu(F.linear(x, sd[p + "0.weight"], sd[p + "0.bias"])), sd[p + "2.weight"], sd[p + "2.bias"])


def _scatter(src, batch, G):
    return src.new_zeros((G,) + tuple(src.shape[1:])).index_add(0, batch, src)


def _conditioning(sd, p, x, feat, batch, G):
    """Common arithmetic of ChargeEmbedding / SpinEmbedding (nn/electronic.py:36-51, 76-90)."""
    norm = torch.maximum(feat, torch.ones_like(feat))
    query = F.linear(x, sd[p + "linear_q.weight"], sd[p + "linear_q.bias"])
    key = F.linear(feat / norm, sd[p + "linear_k.weight"]).index_select(0, batch)
    value = F.linear(feat, sd[p + "linear_v.weight"]).index_select(0, batch)
    dot = (query * key).sum(-1, keepdim=True)
    attn = F.softplus(dot / math.sqrt(x.shape[1]))
    attn_sum = _scatter(attn, batch, G).index_select(0, batch)
    h = attn * value / attn_sum
    # ResidualLayer (nn/basic.py:11-31): two bias-free Linear + SiLU, output scaled by 1/sqrt(2)
    r = F.silu(F.linear(F.silu(F.linear(h, sd[p + "residual.mlp.0.weight"])), sd[p + "residual.mlp.2.weight"]))
    return x + (h + r) / math.sqrt(2)


def charge_embedding(sd, p, x, charge, batch, G):
    """nn/electronic.py:31-51: positive and negative charge are separate features."""
    c = charge.to(x.dtype).reshape(-1)
    return _conditioning(sd, p, x, F.relu(torch.stack([c, -c], dim=-1)), batch, G)


def spin_embedding(sd, p, x, spin, batch, G):
    """nn/electronic.py:71-90 with `spin` as a [G, 1] column."""
    return _conditioning(sd, p, x, spin.to(x.dtype).reshape(-1, 1), batch, G)


def o3_linear_map(V, weight, bias, muls_in, muls_out):
    """e3nn o3.Linear between different multiplicities (nn/output.py:218-222, 282-286), e3nn layout:
    one path per l present on both sides, out[w,m] = sum_u W_l[u,w] in[u,m] / sqrt(mul_in_l); bias on 0e."""
    outs, woff, ioff = [], 0, 0
    for l in range(3):
        mi, mo, d = muls_in[l], muls_out[l], 2 * l + 1
        if mo:
            if mi:
                W = weight[woff : woff + mi * mo].view(mi, mo)
                woff += mi * mo
                y = torch.einsum("uw,zui->zwi", W, V[:, ioff : ioff + mi * d].view(-1, mi, d)) / math.sqrt(mi)
            else:
                y = V.new_zeros(V.shape[0], mo, d)
            if l == 0 and bias is not None and bias.numel():
                y = y + bias.view(1, mo, 1)
            outs.append(y.reshape(-1, mo * d))
        ioff += mi * d
    return torch.cat(outs, dim=1)


def gate(V, muls, eps: float = 1e-5):
    """Gate(irreps, "silu") (nn/o3layer.py:47-75): x * sigmoid(Invariant(x)) per irrep, e3nn layout."""
    outs, off = [], 0
    for l, mul in enumerate(muls):
        d = 2 * l + 1
        if mul:
            blk = V[:, off : off + mul * d].view(-1, mul, d)
            inv = torch.sqrt((blk * blk).sum(-1) + eps**2) - eps
            outs.append((blk * torch.sigmoid(inv).unsqueeze(-1)).reshape(-1, mul * d))
        off += mul * d
    return torch.cat(outs, dim=1)


def xpainn_heads(sd, embed_table, data, cfg: XPaiNNConfig, modes, atom_mass: Optional[torch.Tensor] = None,
                 hidden_dipole=(0, 32, 0), hidden_polar=(64, 0, 16)) -> Dict[str, torch.Tensor]:
    """Every head of `modes` on the features of xpainn_features (nn/output.py; defaults of each constructor)."""
    x, V, batch, G = xpainn_features(sd, embed_table, data, cfg)
    out: Dict[str, torch.Tensor] = {}
    n_atoms = _scatter(torch.ones_like(x[:, 0]), batch, G)
    for mode in modes:
        p = f"mods.output_{mode}."
        if mode == "energy":  # nn/output.py:114-128
            e_atom = _mlp(sd, p + "out_mlp.", x).reshape(-1)
            out["atomic_energies"], out["energy"] = e_atom, _scatter(e_atom, batch, G)
        elif mode == "scalar":  # nn/output.py:65-76
            out["scalar_output"] = _scatter(_mlp(sd, p + "out_mlp.", x).reshape(-1), batch, G)
        elif mode in ("charges", "atomic_charges"):  # nn/output.py:160-180
            q = _mlp(sd, p + "out_mlp.", x).reshape(-1)
            total = data["charge"].to(q.dtype).reshape(-1) if "charge" in data else torch.zeros(G, dtype=q.dtype)
            out["atomic_charges"] = q + ((total - _scatter(q, batch, G)) / n_atoms).index_select(0, batch)
        elif mode == "dipole":  # nn/output.py:226-243
            h = gate(o3_linear_map(V, sd[p + "equi_out_mlp.0.weight"], None, cfg.muls, hidden_dipole), hidden_dipole)
            e = o3_linear_map(h, sd[p + "equi_out_mlp.2.weight"], None, hidden_dipole, (0, 1, 0))[:, [2, 0, 1]]
            out["dipole"] = _scatter(e * _mlp(sd, p + "scalar_out_mlp.", x), batch, G)
        elif mode == "polar":  # nn/output.py:291-327
            h = gate(o3_linear_map(V, sd[p + "equi_out_mlp.0.weight"], sd[p + "equi_out_mlp.0.bias"], cfg.muls,
                                   hidden_polar), hidden_polar)
            e = o3_linear_map(h, sd[p + "equi_out_mlp.2.weight"], sd[p + "equi_out_mlp.2.bias"], hidden_polar, (1, 0, 1))
            sc = _mlp(sd, p + "scalar_out_mlp.", x)
            pol = _scatter(torch.cat([e[:, :1] * sc[:, :1], e[:, 1:] * sc[:, 1:2]], dim=1), batch, G)
            z, d = pol[:, 0], pol[:, 1:6]
            dn = torch.linalg.norm(d, dim=-1)
            r3 = 1 / math.sqrt(3)
            a = torch.zeros(G, 3, 3, dtype=pol.dtype)
            a[:, 0, 0] = r3 * (dn - d[:, 2]) + d[:, 4] + z
            a[:, 1, 1] = r3 * (dn - d[:, 2]) - d[:, 4] + z
            a[:, 2, 2] = r3 * (dn + 2 * d[:, 2]) + z
            a[:, 0, 1] = a[:, 1, 0] = d[:, 0]
            a[:, 1, 2] = a[:, 2, 1] = d[:, 1]
            a[:, 0, 2] = a[:, 2, 0] = d[:, 3]
            out["polarizability"] = a
        elif mode == "spatial":  # nn/output.py:356-373
            m = atom_mass.to(x.dtype)[data["atomic_numbers"].long()].unsqueeze(-1)
            cen = _scatter(m * data["pos"], batch, G) / _scatter(m, batch, G)
            rel = data["pos"] - cen.index_select(0, batch)
            out["spatial_extent"] = _scatter(_mlp(sd, p + "scalar_out_mlp.", x) * (rel * rel).sum(1, keepdim=True), batch, G)
        else:
            raise NotImplementedError(mode)
    return out


def heads_state_dict_spec(cfg: XPaiNNConfig, charge_embed: bool, spin_embed: bool, modes,
                          hidden_dipole=(0, 32, 0), hidden_polar=(64, 0, 16)):
    """state_dict_spec plus the entries of the conditioning modules and heads (names observed on the reference)."""
    C, hd = cfg.node_dim, cfg.hidden_dim
    spec = [e for e in state_dict_spec(cfg) if not e[0].startswith("mods.output_energy.")]
    for on, name, nf in ((charge_embed, "charge", 2), (spin_embed, "spin", 1)):
        if on:
            p = f"mods.{name}_embedding."
            spec += [(p + "linear_q.weight", (C, C), "lin"), (p + "linear_q.bias", (C,), "bias"),
                     (p + "linear_k.weight", (C, nf), "lin"), (p + "linear_v.weight", (C, nf), "lin"),
                     (p + "residual.mlp.0.weight", (C, C), "lin"), (p + "residual.mlp.2.weight", (C, C), "lin")]

    def mlp(p, n_out):
        return [(p + "0.weight", (hd, C), "lin"), (p + "0.bias", (hd,), "bias"),
                (p + "2.weight", (n_out, hd), "lin"), (p + "2.bias", (n_out,), "bias")]

    def nw(mi, mo):
        return sum(a * b for a, b in zip(mi, mo))

    for mode in modes:
        p = f"mods.output_{mode}."
        if mode in ("energy", "scalar", "charges", "atomic_charges"):
            spec += mlp(p + "out_mlp.", 1)
        elif mode == "dipole":
            spec += mlp(p + "scalar_out_mlp.", 1)
            spec += [(p + "equi_out_mlp.0.weight", (nw(cfg.muls, hidden_dipole),), "o3"),
                     (p + "equi_out_mlp.2.weight", (nw(hidden_dipole, (0, 1, 0)),), "o3")]
        elif mode == "polar":
            spec += mlp(p + "scalar_out_mlp.", 2)
            spec += [(p + "equi_out_mlp.0.weight", (nw(cfg.muls, hidden_polar),), "o3"),
                     (p + "equi_out_mlp.0.bias", (hidden_polar[0],), "bias"),
                     (p + "equi_out_mlp.2.weight", (nw(hidden_polar, (1, 0, 1)),), "o3"),
                     (p + "equi_out_mlp.2.bias", (1,), "bias")]
        elif mode == "spatial":
            spec += mlp(p + "scalar_out_mlp.", 1)
        else:
            raise NotImplementedError(mode)
    return spec



def xpainn_energy_forces(sd, embed_table, data, cfg: XPaiNNConfig = CONFIG_DEFAULT, create_graph: bool = False,
                         compute_virial: bool = False):
    """BaseModel.forward with compute_forces=True (nn/basic.py:143-159, 202-238); with compute_virial the strain
    trick of nn/basic.py:93-107 (positions and cell displaced by a symmetrised per-graph strain) and
    virial = -dE/dstrain (nn/basic.py:162-199)."""
    data = dict(data)
    pos = data["pos"].detach().clone().requires_grad_(True)
    data["pos"] = pos
    strain = None
    if compute_virial:
        batch = data.get("batch")
        if batch is None:
            batch = torch.zeros(pos.shape[0], dtype=torch.long)
        G = int(data["ptr"].numel() - 1) if "ptr" in data else int(batch.max().item()) + 1
        strain = torch.zeros((G, 3, 3), dtype=pos.dtype, requires_grad=True)
        symm = 0.5 * (strain + strain.transpose(1, 2))
        data["pos"] = pos + torch.bmm(pos.unsqueeze(1), symm.index_select(0, batch)).squeeze(1)
        if data.get("cell") is not None:
            cell = data["cell"].reshape(-1, 3, 3)
            data["cell"] = cell + torch.bmm(cell, symm)
    energy, e_atom = xpainn_energy(sd, embed_table, data, cfg)
    inputs = [pos] + ([strain] if compute_virial else [])
    grads = torch.autograd.grad([energy], inputs, grad_outputs=[torch.ones_like(energy)],
                                create_graph=create_graph, retain_graph=create_graph)
    out = {"energy": energy, "atomic_energies": e_atom, "forces": -grads[0]}
    if compute_virial:
        out["virial"] = -grads[1]
    return out


# ----------------------------------------------------------------------------
# neighbour lists
# ----------------------------------------------------------------------------
def canonical_sort(edge_index: torch.Tensor, cell_offsets: Optional[torch.Tensor] = None):
    """Lexicographic order on (center, neighbor, ox, oy, oz)."""
    E = edge_index.shape[1]
    cols = [edge_index[0].long(), edge_index[1].long()]
    if cell_offsets is not None:
        co = cell_offsets.round().long()
        cols += [co[:, 0], co[:, 1], co[:, 2]]
    order = torch.arange(E)
    for c in reversed(cols):  # stable sorts, least-significant key first
        order = order[torch.sort(c[order], stable=True)[1]]
    ei = edge_index[:, order]
    return (ei, None) if cell_offsets is None else (ei, cell_offsets[order])


def radius_graph(pos: torch.Tensor, r: float, batch: Optional[torch.Tensor] = None) -> torch.Tensor:
    """torch_cluster.radius_graph(x, r, batch, loop=False, max_num_neighbors=huge)
    as called at data/transform.py:58-64 (torch-cluster 1.6.3, not vendored): all ordered
    pairs a != b of the same graph with sum_d (x_a - x_b)_d^2 < r^2 evaluated in the
    dtype of ``pos``; returned canonically sorted, row 0 = center, row 1 = neighbor."""
    N = pos.shape[0]
    if batch is None:
        batch = torch.zeros(N, dtype=torch.long)
    r2 = torch.tensor(r, dtype=pos.dtype) ** 2
    rows, cols = [], []
    start = 0
    counts = torch.bincount(batch).tolist() if N else []
    for n in counts:
        p = pos[start : start + n]
        d = p.unsqueeze(1) - p.unsqueeze(0)
        d2 = d[..., 0] * d[..., 0]
        d2 = d2 + d[..., 1] * d[..., 1]
        d2 = d2 + d[..., 2] * d[..., 2]
        mask = (d2 < r2) & ~torch.eye(n, dtype=torch.bool)
        idx = mask.nonzero()
        rows.append(idx[:, 0] + start)
        cols.append(idx[:, 1] + start)
        start += n
    if not rows:
        return torch.zeros(2, 0, dtype=torch.long)
    return torch.stack([torch.cat(rows), torch.cat(cols)])


def wrap_positions(pos, cell, n_nodes_per_graph, pbc: List[bool]):
    """data/radius_graph.py:6-32."""
    if not any(pbc):
        return pos, torch.zeros_like(pos)
    cell_pa = cell.repeat_interleave(n_nodes_per_graph, dim=0)
    cell_inv = torch.linalg.inv(cell_pa)
    frac = torch.bmm(pos.unsqueeze(1), cell_inv).squeeze(1)
    shift = torch.zeros_like(pos)
    for i, periodic in enumerate(pbc):
        if periodic:
            shift[:, i] = torch.floor(frac[:, i])
    frac = frac - shift
    pos_wrap = torch.bmm(frac.unsqueeze(1), cell_pa).squeeze(1)
    return pos_wrap, shift


def pbc_repeats(cell: torch.Tensor, pbc: List[bool], cutoff: float) -> List[int]:
    """data/radius_graph.py:61-89: images per axis = ceil(rc * |a_j x a_k| / V), max over graphs."""
    reps = []
    cross = [
        torch.cross(cell[:, 1], cell[:, 2], dim=-1),
        torch.cross(cell[:, 2], cell[:, 0], dim=-1),
        torch.cross(cell[:, 0], cell[:, 1], dim=-1),
    ]
    vol = torch.sum(cell[:, 0] * cross[0], dim=-1, keepdim=True)
    for ax in range(3):
        if pbc[ax]:
            inv_min = torch.norm(cross[ax] / vol, p=2, dim=-1)
            reps.append(int(torch.ceil(cutoff * inv_min).max().item()))
        else:
            reps.append(0)
    return reps


def radius_graph_pbc(pos, n_nodes_per_graph, pbc, cell, cutoff: float):
    """data/radius_graph.py:35-192 restated without the 65 536-column blocking (which only
    changes output order): wrap, replicate images -rep..rep, keep cutoff > D > 0.01 with
    D = sqrt(sum (a - (b + o@cell))^2) on wrapped coordinates, offsets referred back to the
    unwrapped positions.  Returns canonically sorted (edge_index [2,E], cell_offsets [E,3])."""
    pbc_ = [bool(v) for v in pbc[0].tolist()]
    assert bool(torch.all(pbc[0] == pbc))
    reps = pbc_repeats(cell, pbc_, cutoff)
    axes = [torch.arange(-r, r + 1, dtype=pos.dtype) for r in reps]
    offs = torch.cartesian_prod(*axes)  # [n_cells, 3]
    pos_w, shift = wrap_positions(pos, cell, n_nodes_per_graph, pbc_)
    ei_all, off_all = [], []
    start = 0
    for g, n in enumerate(n_nodes_per_graph.tolist()):
        A = pos_w[start : start + n]
        img = offs @ cell[g]  # [n_cells, 3]
        Bp = A.unsqueeze(1) + img.unsqueeze(0)  # [n, n_cells, 3]
        diff = A.view(n, 1, 1, 3) - Bp.view(1, n, -1, 3)
        D = torch.sqrt((diff * diff).sum(-1))  # [n(center), n(neighbor), n_cells]
        idx = ((D < cutoff) & (D > 0.01)).nonzero()
        c, nb, k = idx[:, 0], idx[:, 1], idx[:, 2]
        ei_all.append(torch.stack([c, nb]) + start)
        off_all.append(offs[k] + shift[start + c] - shift[start + nb])
        start += n
    ei = torch.cat(ei_all, dim=1)
    co = torch.cat(off_all, dim=0)
    return canonical_sort(ei, co)


# ----------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------
_Z_QM9 = ([1, 6, 7, 8], [0.51, 0.35, 0.06, 0.08])
_Z_SPICE = ([1, 6, 7, 8, 9, 16, 17, 15, 35], [0.47, 0.33, 0.07, 0.09, 0.01, 0.01, 0.01, 0.005, 0.005])


def _grow_molecule(n_atoms: int, g: torch.Generator) -> torch.Tensor:
    """Random growth: new atom at U(1.0,1.6) A from a random existing atom, rejected if
    closer than 0.9 A to any atom (float64)."""
    pos = torch.zeros(1, 3, dtype=torch.float64)
    while pos.shape[0] < n_atoms:
        anchor = pos[int(torch.randint(pos.shape[0], (1,), generator=g))]
        direction = torch.randn(3, generator=g, dtype=torch.float64)
        direction = direction / direction.norm()
        rad = 1.0 + 0.6 * float(torch.rand(1, generator=g, dtype=torch.float64))
        cand = anchor + rad * direction
        if float((pos - cand).norm(dim=1).min()) >= 0.9:
            pos = torch.cat([pos, cand.unsqueeze(0)])
    return pos


def _margin_ok(pos: torch.Tensor, cutoff: float, margin: float = 1e-4, dmin: float = 0.7) -> bool:
    d = torch.cdist(pos, pos)
    iu = torch.triu_indices(pos.shape[0], pos.shape[0], 1)
    d = d[iu[0], iu[1]]
    return bool(((d - cutoff).abs() >= margin).all() and (d >= dmin).all())


def make_molecule_batch(
    n_mol: int,
    atoms_per_mol=18,
    seed: int = 0,
    z_table=_Z_QM9,
    cutoff: float = 5.0,
    dtype=torch.float32,
    with_edges: bool = True,
) -> Dict[str, torch.Tensor]:
    """c1/c2/c4-shaped batches.  ``atoms_per_mol`` is an int or an inclusive (lo, hi) range."""
    g = torch.Generator().manual_seed(seed)
    zs, probs = torch.tensor(z_table[0]), torch.tensor(z_table[1], dtype=torch.float64)
    pos_l, z_l, batch_l, ptr = [], [], [], [0]
    for m in range(n_mol):
        if isinstance(atoms_per_mol, int):
            n = atoms_per_mol
        else:
            n = int(torch.randint(atoms_per_mol[0], atoms_per_mol[1] + 1, (1,), generator=g))
        while True:
            p = _grow_molecule(n, g)
            if _margin_ok(p, cutoff):
                break
        pos_l.append(p)
        z_l.append(zs[torch.multinomial(probs, n, replacement=True, generator=g)])
        batch_l.append(torch.full((n,), m, dtype=torch.long))
        ptr.append(ptr[-1] + n)
    data = {
        "pos": torch.cat(pos_l).to(dtype),
        "atomic_numbers": torch.cat(z_l).to(torch.int32),
        "batch": torch.cat(batch_l),
        "ptr": torch.tensor(ptr, dtype=torch.long),
    }
    if with_edges:
        data["edge_index"] = radius_graph(data["pos"], cutoff, data["batch"])
    return data


def make_aspirin_batch(n_mol: int, seed: int = 0, cutoff: float = 5.0, dtype=torch.float32, sigma: float = 0.05,
                       with_edges: bool = True):
    """c3: aspirin C9H8O4 (21 atoms), one seed-17 growth geometry + per-frame jitter."""
    g0 = torch.Generator().manual_seed(17)
    base = _grow_molecule(21, g0)
    z = torch.tensor([6] * 9 + [1] * 8 + [8] * 4, dtype=torch.int32)
    g = torch.Generator().manual_seed(seed)
    pos_l = []
    for _ in range(n_mol):
        while True:
            p = base + sigma * torch.randn(21, 3, generator=g, dtype=torch.float64)
            if _margin_ok(p, cutoff):
                break
        pos_l.append(p)
    data = {
        "pos": torch.cat(pos_l).to(dtype),
        "atomic_numbers": z.repeat(n_mol),
        "batch": torch.arange(n_mol).repeat_interleave(21),
        "ptr": torch.arange(0, 21 * n_mol + 1, 21),
    }
    if with_edges:
        data["edge_index"] = radius_graph(data["pos"], cutoff, data["batch"])
    return data


def make_water_box(n_side: int = 15, seed: int = 0, density: float = 0.1002, dtype=torch.float32,
                   jitter: float = 0.25, unwrap_frac: float = 0.2):
    """c5-shaped periodic water box: n_side^3 molecules, O on a jittered simple-cubic grid,
    H at 0.96 A / 104.5 deg with random orientation; some atoms deliberately left outside
    the cell (positions are *not* pre-wrapped).  n_side=15 -> 3375 molecules, 10 125 atoms
    (SURVEY.md 8d names 3334 molecules; a cubic grid needs a cube number)."""
    g = torch.Generator().manual_seed(seed)
    n_mol = n_side**3
    L = (3 * n_mol / density) ** (1.0 / 3.0)
    a = L / n_side
    grid = torch.stack(torch.meshgrid(*[torch.arange(n_side, dtype=torch.float64)] * 3, indexing="ij"), -1).reshape(-1, 3)
    O = (grid + 0.5) * a + jitter * (torch.rand(n_mol, 3, generator=g, dtype=torch.float64) * 2 - 1)
    # random orthonormal frames
    q = torch.randn(n_mol, 3, 3, generator=g, dtype=torch.float64)
    q, _ = torch.linalg.qr(q)
    half = math.radians(104.5) / 2
    h1 = 0.96 * (math.cos(half) * q[:, :, 0] + math.sin(half) * q[:, :, 1])
    h2 = 0.96 * (math.cos(half) * q[:, :, 0] - math.sin(half) * q[:, :, 1])
    pos = torch.stack([O, O + h1, O + h2], dim=1).reshape(-1, 3)
    # leave a fraction of molecules shifted by whole lattice vectors (unwrapped input)
    sh = torch.randint(-1, 2, (n_mol, 3), generator=g).to(torch.float64)
    sel = (torch.rand(n_mol, generator=g, dtype=torch.float64) < unwrap_frac).to(torch.float64).unsqueeze(-1)
    pos = pos + (sh * sel * L).repeat_interleave(3, dim=0)
    N = pos.shape[0]
    return {
        "pos": pos.to(dtype),
        "atomic_numbers": torch.tensor([8, 1, 1], dtype=torch.int32).repeat(n_mol),
        "batch": torch.zeros(N, dtype=torch.long),
        "ptr": torch.tensor([0, N], dtype=torch.long),
        "cell": (torch.eye(3, dtype=torch.float64) * L).unsqueeze(0).to(dtype),
        "pbc": torch.tensor([[True, True, True]]),
    }


def make_small_pbc(n_atoms: int = 12, box: float = 6.0, seed: int = 0, dtype=torch.float32, triclinic: bool = True,
                   pbc=(True, True, True)):
    """A small (possibly smaller-than-cutoff) periodic cell with unwrapped atoms: exercises
    multiple images and self-image edges."""
    g = torch.Generator().manual_seed(seed)
    cell = torch.eye(3, dtype=torch.float64) * box
    if triclinic:
        cell = cell + 0.15 * box * torch.tensor([[0.0, 0.0, 0.0], [0.6, 0.0, 0.0], [-0.4, 0.5, 0.0]], dtype=torch.float64)
    while True:
        frac = torch.rand(n_atoms, 3, generator=g, dtype=torch.float64) * 1.6 - 0.3
        pos = frac @ cell
        pw = (frac - torch.floor(frac)) @ cell
        d = torch.cdist(pw, pw) + 10 * torch.eye(n_atoms, dtype=torch.float64)
        if float(d.min()) > 0.8:
            break
    return {
        "pos": pos.to(dtype),
        "atomic_numbers": torch.tensor([1, 6, 7, 8], dtype=torch.int32)[torch.randint(0, 4, (n_atoms,), generator=g)],
        "batch": torch.zeros(n_atoms, dtype=torch.long),
        "ptr": torch.tensor([0, n_atoms], dtype=torch.long),
        "cell": cell.unsqueeze(0).to(dtype),
        "pbc": torch.tensor([list(pbc)]),
    }


# ----------------------------------------------------------------------------
# stand-alone edge message (one XPainnMessage aggregation), used to check K2/K2b/K2bb
# ----------------------------------------------------------------------------
def edge_message(x, V, s, vn, pos, W_rbf, b_rbf, freq, edge_index, cfg: XPaiNNConfig,
                 cell=None, cell_offsets=None, batch=None):
    """nn/xpainn.py:140-159 restated as a function of (x, V, s, vn, pos, weights), e3nn layout.
    Returns (x_out, V_out)."""
    M = cfg.M
    m0, m1, m2 = cfg.muls
    center, neighbor = edge_index[0], edge_index[1]
    vec, dist = edge_vectors(pos, edge_index, cell, cell_offsets, batch)
    d1 = dist.unsqueeze(-1)
    rbf = bessel_rbf(d1, freq.view(1, -1), cfg.cutoff)
    fcut = cosine_cutoff(d1, cfg.cutoff)
    Y = spherical_harmonics_l2(vec)
    rsh = torch.cat([Y[:, 0:1].repeat(1, m0), Y[:, 1:4].repeat(1, m1), Y[:, 4:9].repeat(1, m2)], dim=1)
    w = F.linear(rbf, W_rbf, b_rbf) * fcut
    h = s.index_select(0, neighbor) * w
    g_state, g_edge, m_s = h[:, :M], h[:, M : 2 * M], h[:, 2 * M :]
    m_e = vn.index_select(0, neighbor) * expand_gate(g_state, cfg) + rsh * expand_gate(g_edge, cfg)
    return x.index_add(0, center, m_s), V.index_add(0, center, m_e)


def to_cm(V: torch.Tensor, cfg: XPaiNNConfig) -> torch.Tensor:
    """e3nn layout [mul][m] -> component-major [m][mul] per l (include/xeq_b200.h)."""
    outs = []
    for off, mul, d in _blocks(cfg):
        outs.append(V[:, off : off + mul * d].reshape(-1, mul, d).transpose(1, 2).reshape(-1, mul * d))
    return torch.cat(outs, dim=1)


def from_cm(V: torch.Tensor, cfg: XPaiNNConfig) -> torch.Tensor:
    outs = []
    for off, mul, d in _blocks(cfg):
        outs.append(V[:, off : off + mul * d].reshape(-1, d, mul).transpose(1, 2).reshape(-1, mul * d))
    return torch.cat(outs, dim=1)
