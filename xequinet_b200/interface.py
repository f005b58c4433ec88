"""MD / geometry-optimisation front end: the reference's ASE calculator
(xequinet/interface/ase_calculator.py:20-118) on the B200 path.

`XequiCalculator` keeps the reference's protocol -- `calculate(atoms, properties)` fills `results` with `energy`
(float, eV), `energies` [N], `forces` [N, 3] (eV/A) and `stress` (Voigt 6, virial / volume) -- and works with any
object that offers the `ase.Atoms` accessors it uses (`get_positions`, `get_atomic_numbers`, `get_cell`, `get_pbc`,
`get_volume`); when ASE is installed it IS an `ase.calculators.calculator.Calculator`.  What changes is the work per
call (interface/ase_calculator.py:86-96 rebuilds the neighbour list on the host path and synchronises several times):

  * neighbour list on the GPU (K1), reused across steps through a Verlet skin (`SkinNeighborTransform`: exact energies
    and forces, edges beyond the cutoff contribute zero);
  * `graph_replay=True`: for a fixed number of atoms and a fixed cell, the whole step (K1 in capacity mode + model +
    forces) is captured once as a CUDA graph (`replay.CapturedStep`) and replayed with one host -> device copy of the
    positions and one device -> host copy of the results;
  * energies / forces of the default model go through the C inference runtime (`runtime.NativeModel` ->
    `xeq_model_energy_forces`: forward and force pass scheduled inside the library, no autograd graph; bit-identical to
    the module path and ~4x faster per call on small molecules when launched eagerly), stress included (the virial
    of the strain trick from the force pass's own records).  Models with conditioning modules or extra heads use the
    module path.

`atoms.wrap()` of the reference (:86) is not needed: K1 wraps internally and returns offsets that refer to the
unwrapped positions (data/radius_graph.py:186-190), so the caller's Atoms object is left untouched."""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from . import keys
from .graph import SkinNeighborTransform
from .nn import load_model, resolve_model
from .replay import CapturedStep

try:  # pragma: no cover - ASE is optional
    from ase.calculators.calculator import Calculator as _AseCalculator, all_changes as _all_changes
except Exception:  # ASE not installed: a minimal stand-in with the same surface
    _all_changes = ["positions", "numbers", "cell", "pbc"]

    class _AseCalculator:  # type: ignore
        def __init__(self, **kwargs):
            self.results: Dict[str, object] = {}
            self.atoms = None

        def calculate(self, atoms=None, properties=None, system_changes=None):
            if atoms is not None:
                self.atoms = atoms


def full_3x3_to_voigt_6_stress(m: np.ndarray) -> np.ndarray:
    """ase.stress.full_3x3_to_voigt_6_stress: (xx, yy, zz, yz, xz, xy) of the symmetrised tensor."""
    m = np.asarray(m, dtype=np.float64).reshape(3, 3)
    return np.array([m[0, 0], m[1, 1], m[2, 2], 0.5 * (m[1, 2] + m[2, 1]), 0.5 * (m[0, 2] + m[2, 0]), 0.5 * (m[0, 1] + m[1, 0])])


class XequiCalculator(_AseCalculator):
    implemented_properties = ["energy", "energies", "forces", "stress"]

    def __init__(self, model: Optional[torch.nn.Module] = None, ckpt_file: Optional[str] = None, device: str = "cuda",
                 skin: float = 1.0, graph_replay: bool = False, **kwargs) -> None:
        super().__init__(**kwargs)
        if (model is None) == (ckpt_file is None):
            raise ValueError("give exactly one of model / ckpt_file")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("xequinet_b200 models run on CUDA devices only: there is no CPU fallback")
        if ckpt_file is not None:
            model = load_model(ckpt_file, self.device).model
        self.model = model.to(self.device).eval()
        for p in self.model.parameters():
            p.requires_grad_(False)
        self.transform = SkinNeighborTransform(self.model.cutoff_radius, skin=skin)
        self.graph_replay = bool(graph_replay)
        try:  # the default module chain runs on the C inference runtime (same results, no autograd)
            from .runtime import NativeModel
            self.native = NativeModel(self.model)
        except NotImplementedError:
            self.native = None
        self._captured: Optional[CapturedStep] = None
        self._captured_sig = None
        self.results = {}

    # ---- atoms -> data dict (datapoint_from_ase, data/datapoint.py) --------------------------------------
    def _data(self, atoms) -> Dict[str, torch.Tensor]:
        pos = torch.as_tensor(np.asarray(atoms.get_positions()), dtype=torch.float32)
        z = torch.as_tensor(np.asarray(atoms.get_atomic_numbers()), dtype=torch.int32)
        d = {keys.POSITIONS: pos, keys.ATOMIC_NUMBERS: z,
             keys.BATCH: torch.zeros(pos.shape[0], dtype=torch.long), keys.BATCH_PTR: torch.tensor([0, pos.shape[0]], dtype=torch.long)}
        pbc = np.asarray(atoms.get_pbc(), dtype=bool)
        if pbc.any():
            d[keys.PBC] = torch.as_tensor(pbc).reshape(1, 3)
            d[keys.CELL] = torch.as_tensor(np.asarray(atoms.get_cell()), dtype=torch.float32).reshape(1, 3, 3)
        return d

    def calculate(self, atoms=None, properties: Optional[List[str]] = None, system_changes=_all_changes) -> None:
        if properties is None:
            properties = self.implemented_properties
        super().calculate(atoms, properties, system_changes)
        atoms = self.atoms if atoms is None else atoms
        want_f, want_s = "forces" in properties, "stress" in properties
        host = self._data(atoms)
        if self.graph_replay and not want_s:
            out = self._replayed(host)
        else:
            data = self.transform({k: v.to(self.device) for k, v in host.items()})
            data.pop(keys.PBC, None)
            if self.native is not None:
                out = self.native(data, compute_forces=want_f, compute_virial=want_s)
            else:
                out = self.model(data, compute_forces=want_f, compute_virial=want_s)
        # one device -> host transfer of everything that was asked for
        self.results["energy"] = float(out[keys.TOTAL_ENERGY].detach().reshape(-1)[0].item())
        self.results["energies"] = out[keys.ATOMIC_ENERGIES].detach().cpu().numpy()
        if want_f:
            self.results["forces"] = out[keys.FORCES].detach().cpu().numpy()
        if want_s:
            assert keys.CELL in host, "stress needs a periodic cell"
            virial = out[keys.VIRIAL].detach().cpu().numpy().reshape(3, 3)
            self.results["stress"] = full_3x3_to_voigt_6_stress(virial) / float(atoms.get_volume())

    def _replayed(self, host: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        sig = (host[keys.POSITIONS].shape[0], tuple(host[keys.ATOMIC_NUMBERS].tolist()),
               None if keys.CELL not in host else tuple(host[keys.CELL].reshape(-1).tolist()))
        if self._captured is None or sig != self._captured_sig:
            example = {k: v.to(self.device) for k, v in host.items()}
            self._captured = CapturedStep(self.native or self.model, example, compute_forces=True, capacity_margin=1.5)
            self._captured_sig = sig
        out = self._captured({keys.POSITIONS: host[keys.POSITIONS], keys.ATOMIC_NUMBERS: host[keys.ATOMIC_NUMBERS],
                              **({keys.CELL: host[keys.CELL]} if keys.CELL in host else {})})
        self._captured.check()
        return out

    # the ASE convenience accessors, for atoms-like objects without ASE
    def get_potential_energy(self, atoms=None) -> float:
        self.calculate(atoms, ["energy"])
        return self.results["energy"]

    def get_forces(self, atoms=None) -> np.ndarray:
        self.calculate(atoms, ["energy", "forces"])
        return self.results["forces"]

    def get_stress(self, atoms=None) -> np.ndarray:
        self.calculate(atoms, ["energy", "forces", "stress"])
        return self.results["stress"]
