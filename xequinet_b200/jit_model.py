"""Deployment wrappers behind the reference's interface (xequinet/interface/jit_model.py:12-237, built by
xequinet/run/jit_script.py:28-86): the model an MD engine drives, in the ENGINE's unit system.

  XPaiNNLMP     LAMMPS `pair_style xequinet`: the engine supplies positions, its own neighbour list (`edge_index`,
                `cell_offsets`) and domain decomposition; energy / forces / virial come back in LAMMPS units.
  XPaiNNGMX     GROMACS NNP interface: forward(positions [nm], atomic_numbers, box, pbc) -> energy [kJ/mol]; the
                neighbour list is built here (K1 instead of data/radius_graph.py:195-222) and the engine
                differentiates the returned energy itself.
  XPaiNNDipole  dipole read-out in LAMMPS units.

Same constructor arguments, attribute names (`pos_unit_factor`, `energy_unit_factor`, `forces_unit_factor`,
`net_charge`, `cutoff_radius` in engine units) and call protocol as the reference.  The wrappers run the same modules
as `XPaiNN` on the C-ABI kernels; they are driven from Python (or through `torch.ops.xeq.*`, the dispatcher-visible
form of every kernel, xequinet_b200/torch_ops.py).  The `.jit` archive of run/jit_script.py:73-86 is NOT produced:
`torch.jit.script` of the module tree needs the ops registered from C++ for a libtorch-only host process, which this
package does not ship; `export_deployment` writes the same metadata next to a plain state_dict instead."""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import keys
from .graph import NeighborTransform
from .nn.basic import compute_edge_data, compute_properties
from .nn.model import BaseModel, XPaiNN
from .units import get_default_units, unit_conversion

ELEMENTS = (
    "d H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr Rb Sr Y Zr Nb "
    "Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir Pt Au Hg "
    "Tl Pb Bi Po At Rn").split()


def _with_charge(data: Dict[str, torch.Tensor], net_charge: Optional[int]) -> Dict[str, torch.Tensor]:
    if net_charge is not None:  # interface/jit_model.py:63-66
        data[keys.TOTAL_CHARGE] = torch.tensor([net_charge], device=data[keys.POSITIONS].device)
    return data


class XPaiNNLMP(XPaiNN):
    """interface/jit_model.py:12-89.  Single structure (no batch)."""

    def __init__(self, unit_style: str = "metal", net_charge: Optional[int] = None, **kwargs) -> None:
        super().__init__(**kwargs)
        lammps_units, default_units = keys.LAMMPS_UNIT_STYLE[unit_style], get_default_units()
        self.pos_unit_factor = unit_conversion(lammps_units[keys.POSITIONS], default_units[keys.POSITIONS])  # LAMMPS -> NNP
        self.energy_unit_factor = unit_conversion(default_units.get(keys.TOTAL_ENERGY), lammps_units[keys.TOTAL_ENERGY])
        self.forces_unit_factor = unit_conversion(default_units.get(keys.FORCES), lammps_units[keys.FORCES])
        self.net_charge = net_charge
        self.cutoff_radius /= self.pos_unit_factor  # what the engine builds its neighbour list with (its units)

    def forward(self, data: Dict[str, torch.Tensor], compute_forces: bool = True, compute_virial: bool = False):
        # the reference scales the caller's tensor in place (:62); a fresh tensor keeps the engine's buffer untouched
        data[keys.POSITIONS] = data[keys.POSITIONS] * self.pos_unit_factor
        if keys.CELL in data and self.pos_unit_factor != 1.0:
            data[keys.CELL] = data[keys.CELL] * self.pos_unit_factor
        data = compute_edge_data(data=_with_charge(data, self.net_charge), compute_forces=compute_forces,
                                 compute_virial=compute_virial)
        for mod in self.mods.values():
            data = mod(data)
        result = compute_properties(data=data, compute_forces=compute_forces, compute_virial=compute_virial,
                                    training=self.training, extra_properties=self.extra_properties)
        result[keys.TOTAL_ENERGY] = result[keys.TOTAL_ENERGY] * self.energy_unit_factor
        if compute_forces:
            result[keys.FORCES] = result[keys.FORCES] * self.forces_unit_factor
        if compute_virial:
            result[keys.VIRIAL] = result[keys.VIRIAL] * self.energy_unit_factor
        return result


class XPaiNNDipole(XPaiNN):
    """interface/jit_model.py:92-145; build with output_modes=["dipole"]."""

    def __init__(self, unit_style: str = "metal", net_charge: Optional[int] = None, **kwargs) -> None:
        super().__init__(**kwargs)
        lammps_units, default_units = keys.LAMMPS_UNIT_STYLE[unit_style], get_default_units()
        self.pos_unit_factor = unit_conversion(lammps_units[keys.POSITIONS], default_units[keys.POSITIONS])
        self.dipole_unit_factor = unit_conversion(
            default_units.get(keys.DIPOLE), f"{lammps_units[keys.TOTAL_CHARGE]}*{lammps_units[keys.POSITIONS]}")
        self.net_charge = net_charge
        self.cutoff_radius /= self.pos_unit_factor

    def forward(self, data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        data[keys.POSITIONS] = data[keys.POSITIONS] * self.pos_unit_factor
        if keys.CELL in data and self.pos_unit_factor != 1.0:
            data[keys.CELL] = data[keys.CELL] * self.pos_unit_factor
        data = compute_edge_data(data=_with_charge(data, self.net_charge), compute_forces=False, compute_virial=False)
        for mod in self.mods.values():
            data = mod(data)
        return {keys.DIPOLE: data[keys.DIPOLE] * self.dipole_unit_factor}


class XPaiNNGMX(XPaiNN):
    """interface/jit_model.py:148-216: GROMACS hands over positions [nm] (+ box, pbc) and differentiates the energy."""

    def __init__(self, net_charge: Optional[int] = None, **kwargs) -> None:
        kwargs.pop("unit_style", None)
        super().__init__(**kwargs)
        default_units = get_default_units()
        self.pos_unit_factor = unit_conversion("nm", default_units[keys.POSITIONS])
        self.energy_unit_factor = unit_conversion(default_units.get(keys.TOTAL_ENERGY), "kJ/mol")
        self.forces_unit_factor = unit_conversion(default_units.get(keys.FORCES), "kJ/(mol*nm)")
        self.net_charge = net_charge
        self._neighbors = NeighborTransform(self.cutoff_radius)

    def forward(self, positions: torch.Tensor, atomic_numbers: torch.Tensor, box: Optional[torch.Tensor] = None,
                pbc: Optional[torch.Tensor] = None) -> torch.Tensor:
        positions = positions * self.pos_unit_factor
        data = {keys.POSITIONS: positions, keys.ATOMIC_NUMBERS: atomic_numbers}
        periodic = pbc is not None and bool(pbc.any())
        if periodic:
            if box is None:
                raise ValueError("PBC and cell must be both defined or both undefined.")
            data[keys.CELL] = (box.detach() * self.pos_unit_factor).reshape(1, 3, 3)
            data[keys.PBC] = pbc.reshape(1, 3)
        with torch.no_grad():  # the list is a constant of the differentiation, as at :186-192
            nb = self._neighbors(dict(data, **{keys.POSITIONS: positions.detach()}))
        for k in (keys.EDGE_INDEX, keys.CELL_OFFSETS, keys.GRAPH):
            if k in nb:
                data[k] = nb[k]
        data = compute_edge_data(data=_with_charge(data, self.net_charge), compute_forces=True, compute_virial=False)
        for mod in self.mods.values():
            data = mod(data)
        return data[keys.TOTAL_ENERGY] * self.energy_unit_factor


def resolve_jit_model(mode: str = "lmp", unit_style: str = "metal", net_charge: Optional[int] = None, **kwargs) -> BaseModel:
    """interface/jit_model.py:219-237."""
    factory = {"lmp": XPaiNNLMP, "dipole": XPaiNNDipole, "gmx": XPaiNNGMX}
    if mode not in factory:
        raise NotImplementedError(f"Unsupported mode {mode}")
    return factory[mode](unit_style=unit_style, net_charge=net_charge, **kwargs)


def deployment_metadata(model: BaseModel, fusion_strategy: str = "DYNAMIC,3") -> Dict[str, str]:
    """The `_extra_files` run/jit_script.py:76-83 stores with the archive: what the engine-side plugin reads before
    the first step (its neighbour-list cutoff, the species table)."""
    n_species = ELEMENTS.index("Rn") + 1
    return {"cutoff_radius": str(model.cutoff_radius), "jit_fusion_strategy": fusion_strategy,
            "n_species": str(n_species), "periodic_table": " ".join(ELEMENTS[:n_species])}


def export_deployment(ckpt_file: str, out_file: str, mode: str = "lmp", unit_style: str = "metal",
                      net_charge: Optional[int] = None, trust_pickle: bool = False) -> Dict[str, str]:
    """run/jit_script.py:28-86 without the TorchScript archive: checks that the checkpoint builds and loads into the
    wrapper (strict) and writes {model state_dict, model_kwargs, mode, unit_style, net_charge, metadata}."""
    from .units import set_default_units

    ckpt = torch.load(ckpt_file, map_location="cpu", weights_only=not trust_pickle)
    cfg = ckpt["config"]
    if cfg.get("default_units"):
        set_default_units(dict(cfg["default_units"]))
    model = resolve_jit_model(mode=mode, unit_style=unit_style, net_charge=net_charge, **cfg["model_kwargs"])
    model.load_state_dict(ckpt["model"])
    meta = deployment_metadata(model)
    torch.save({"model": model.state_dict(), "model_kwargs": dict(cfg["model_kwargs"]), "mode": mode,
                "unit_style": unit_style, "net_charge": net_charge, "default_units": dict(get_default_units()),
                "metadata": meta}, out_file)
    return meta


def load_deployment(file: str, device: str = "cuda") -> BaseModel:
    from .units import set_default_units

    blob = torch.load(file, map_location="cpu", weights_only=True)
    set_default_units({k: v for k, v in blob["default_units"].items()
                       if k in (keys.POSITIONS, keys.TOTAL_ENERGY, keys.TOTAL_CHARGE, keys.DIPOLE)})
    model = resolve_jit_model(mode=blob["mode"], unit_style=blob["unit_style"], net_charge=blob["net_charge"],
                              **blob["model_kwargs"])
    model.load_state_dict(blob["model"])
    return model.eval().to(device)
