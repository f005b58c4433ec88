"""Model assembly behind the reference's factory API (xequinet/nn/model.py:18-122, 310-318)."""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Union

import torch
import torch.nn as nn

from ..graph import NeighborTransform
from .basic import compute_edge_data, compute_properties
from .electronic import ChargeEmbedding, SpinEmbedding
from .output import resolve_output
from .xpainn import XEmbedding, XPainnMessage, XPainnUpdate


class BaseModel(nn.Module):
    cutoff_radius: float

    def __init__(self) -> None:
        super().__init__()
        self.mods = nn.ModuleDict()
        self.extra_properties = []

    def forward(self, data: Dict[str, torch.Tensor], compute_forces: bool = True, compute_virial: bool = False):
        data = compute_edge_data(data=data, compute_forces=compute_forces, compute_virial=compute_virial)
        for mod in self.mods.values():
            data = mod(data)
        return compute_properties(data=data, compute_forces=compute_forces, compute_virial=compute_virial,
                                  training=self.training, extra_properties=self.extra_properties)


class XPaiNN(BaseModel):
    """eXtended PaiNN with the hyper-parameters and defaults of nn/model.py:57-70."""

    def __init__(self, **kwargs) -> None:
        super().__init__()
        node_dim: int = kwargs.get("node_dim", 128)
        node_irreps: str = kwargs.get("node_irreps", "128x0e + 64x1o + 32x2e")
        embed_basis: str = kwargs.get("embed_basis", "gfn2-xtb")
        aux_basis: str = kwargs.get("aux_basis", "aux56")
        num_basis: int = kwargs.get("num_basis", 20)
        rbf_kernel: str = kwargs.get("rbf_kernel", "bessel")
        cutoff: float = kwargs.get("cutoff", 5.0)
        cutoff_fn: str = kwargs.get("cutoff_fn", "cosine")
        action_blocks: int = kwargs.get("action_blocks", 3)
        activation: str = kwargs.get("activation", "silu")
        layer_norm: bool = kwargs.get("layer_norm", True)
        charge_embed: bool = kwargs.get("charge_embed", False)
        spin_embed: bool = kwargs.get("spin_embed", False)
        output_modes: Union[str, List[str]] = kwargs.get("output_modes", ["energy"])
        self.cutoff_radius = cutoff
        self.mods["embedding"] = XEmbedding(node_dim=node_dim, node_irreps=node_irreps, embed_basis=embed_basis,
                                            aux_basis=aux_basis, num_basis=num_basis, rbf_kernel=rbf_kernel,
                                            cutoff=cutoff, cutoff_fn=cutoff_fn)
        if charge_embed:
            self.mods["charge_embedding"] = ChargeEmbedding(node_dim=node_dim, activation=activation)
        if spin_embed:
            self.mods["spin_embedding"] = SpinEmbedding(node_dim=node_dim, activation=activation)
        for i in range(action_blocks):
            self.mods[f"message_{i}"] = XPainnMessage(node_dim=node_dim, node_irreps=node_irreps, num_basis=num_basis,
                                                      activation=activation, layer_norm=layer_norm)
            self.mods[f"update_{i}"] = XPainnUpdate(node_dim=node_dim, node_irreps=node_irreps, activation=activation,
                                                    layer_norm=layer_norm)
        if output_modes is None:
            output_modes = ["energy"]
        elif isinstance(output_modes, str) or not isinstance(output_modes, Iterable):
            output_modes = [output_modes]
        for mode in output_modes:
            output = resolve_output(mode, **kwargs)
            self.mods[f"output_{mode}"] = output
            self.extra_properties.extend(output.extra_properties)


def resolve_model(model_name: str, **kwargs) -> BaseModel:
    """nn/model.py:310-318; only XPaiNN is on the B200 path."""
    factory = {"xpainn": XPaiNN}
    if model_name.lower() not in factory:
        raise NotImplementedError(f"Unsupported model {model_name}")
    return factory[model_name.lower()](**kwargs)


def load_model(ckpt_file: str, device: Optional[torch.device] = None, trust_pickle: bool = False):
    """nn/model.py:321-351 for dict-shaped data.  The checkpoint holds a state_dict and a plain config dict, so it is
    read with `weights_only=True`; `trust_pickle=True` opts into the reference's unrestricted `torch.load`."""

    class ModelWithTransform:
        def __init__(self, model, transform, device):
            self.model, self.transform, self.device = model, transform, device

        def __call__(self, data, **kwargs):
            data = {k: (v.to(self.device) if torch.is_tensor(v) else v) for k, v in dict(data).items()}
            return self.model(self.transform(data), **kwargs)

    if device is None:
        device = torch.device("cuda")
    ckpt = torch.load(ckpt_file, map_location=device, weights_only=not trust_pickle)
    cfg = ckpt["config"]
    # nn/model.py:340 calls set_default_units(config["default_units"]): the unit registry only drives the data
    # pipeline's conversions, the model arithmetic is unit-free (shift / scale are baked into the read-out weights,
    # nn/output.py:104-106).  The units the checkpoint was trained in are kept on the returned object.
    units = dict(cfg.get("default_units") or {})
    model = resolve_model(cfg["model_name"], **cfg["model_kwargs"]).to(device)
    model.load_state_dict(ckpt["model"])
    model.eval()
    wrapped = ModelWithTransform(model, NeighborTransform(model.cutoff_radius), device)
    wrapped.default_units = units
    return wrapped
