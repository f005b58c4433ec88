"""Python face of the whole-model inference runtime (csrc/model_runtime.cu, `xeq_model_*` in include/xeq_b200.h):
XPaiNN energy + forces as ONE C call, the form an MD engine links against (the reference's
deployment path is a TorchScript archive run through libtorch, run/jit_script.py:28-86).

    native = NativeModel(model)                        # flattens the state_dict into the runtime's weight blob
    out = native(data, compute_forces=True)            # {"energy", "atomic_energies", "forces"}: bit-identical to model(data)
    native.save("model.xeqw")                          # what a C / C++ host reads: header + blob (see `save`)

The C runtime schedules the forward pass and the force pass itself (no autograd graph, no per-op Python), on the same
kernels as the nn modules.  This module only exports weights and passes pointers."""
from __future__ import annotations

import ctypes
import struct
from typing import Dict, List

import torch

from . import _lib, keys
from .nn.basic import compute_edge_data
from .nn.layers import LayerNorm

MAGIC = b"XEQW0001"


def weight_order(n_layers: int) -> List[str]:
    """state_dict names in the order of the runtime's blob (include/xeq_b200.h, xeq_model_create)."""
    names = ["mods.embedding.embedding.0.embed_ten", "mods.embedding.embedding.1.weight", "mods.embedding.embedding.1.bias",
             "mods.embedding.rbf.freq"]
    for i in range(n_layers):
        p = f"mods.message_{i}."
        names += [p + k for k in ("norm.weight", "norm.bias", "o3norm.affine_weight", "o3norm.affine_bias",
                                  "scalar_mlp.0.weight", "scalar_mlp.0.bias", "scalar_mlp.2.weight", "scalar_mlp.2.bias",
                                  "rbf_lin.weight", "rbf_lin.bias")]
        p = f"mods.update_{i}."
        names += [p + k for k in ("norm.weight", "norm.bias", "o3norm.affine_weight", "o3norm.affine_bias",
                                  "update_U.weight", "update_U.bias", "update_V.weight", "update_V.bias", "dot_lin.weight",
                                  "update_mlp.0.weight", "update_mlp.0.bias", "update_mlp.2.weight", "update_mlp.2.bias")]
    p = "mods.output_energy.out_mlp."
    return names + [p + "0.weight", p + "0.bias", p + "2.weight", p + "2.bias"]


def export_weights(state_dict: Dict[str, torch.Tensor], n_layers: int) -> torch.Tensor:
    """Flat fp32 blob (CPU): every tensor in `weight_order`, zero-padded to a multiple of 4 floats."""
    parts = []
    for name in weight_order(n_layers):
        t = state_dict[name].detach().to("cpu", torch.float32).reshape(-1)
        pad = -t.numel() % 4
        parts.append(torch.cat([t, t.new_zeros(pad)]) if pad else t)
    return torch.cat(parts).contiguous()


class NativeModel:
    """The default XPaiNN model (layer norms, SiLU, energy head, no charge / spin conditioning) on the C runtime."""

    def __init__(self, model: torch.nn.Module, branch_stream: bool = True) -> None:
        """branch_stream: run the independent branches of the module graph on a second CUDA stream
        (xeq_model_energy_forces_mt); the results do not depend on it."""
        mods = list(model.mods)
        n_layers = sum(1 for m in mods if m.startswith("message_"))
        expect = ["embedding"] + [f"{k}_{i}" for i in range(n_layers) for k in ("message", "update")] + ["output_energy"]
        if mods != expect:
            raise NotImplementedError(f"the inference runtime runs the default module chain {expect}, got {mods}")
        emb, msg0 = model.mods["embedding"], model.mods["message_0"]
        if not isinstance(emb.embedding, torch.nn.Sequential) or not isinstance(msg0.norm, LayerNorm):
            raise NotImplementedError("the inference runtime needs the int2c1e embedding and layer_norm=True")
        sd = model.state_dict()
        dev = sd["mods.embedding.rbf.freq"].device
        if dev.type != "cuda":
            raise RuntimeError("NativeModel needs the model on a CUDA device: there is no CPU fallback")
        m0, m1, m2 = msg0.muls
        self.dims = _lib.XeqDims(msg0.node_dim, m0, m1, m2, msg0.num_basis, float(emb.rbf.cutoff))
        self.n_layers = n_layers
        self.hidden_dim = model.mods["output_energy"].hidden_dim
        table = sd["mods.embedding.embedding.0.embed_ten"]
        self.n_species, self.embed_dim = int(table.shape[0]), int(table.shape[1])
        self.cutoff_radius = float(model.cutoff_radius)
        self.blob = export_weights(sd, n_layers).to(dev)
        lib = _lib.get()
        want = lib.xeq_model_weight_count(ctypes.byref(self.dims), n_layers, self.hidden_dim, self.embed_dim, self.n_species)
        if want != self.blob.numel():
            raise RuntimeError(f"weight blob has {self.blob.numel()} floats, the runtime's layout needs {want}")
        handle = ctypes.c_void_p()
        _lib.check(lib.xeq_model_create(ctypes.byref(self.dims), n_layers, self.hidden_dim, self.embed_dim, self.n_species,
                                        self.blob.data_ptr(), self.blob.numel(), ctypes.byref(handle)), "xeq_model_create")
        self._handle = handle
        self._aux = torch.cuda.Stream(device=dev) if branch_stream else None

    def refresh(self, model: torch.nn.Module) -> None:
        """Re-export the weights into the existing device blob (the runtime reads a COPY of the parameters taken at
        construction: call this after the model's parameters have changed, e.g. after more training)."""
        self.blob.copy_(export_weights(model.state_dict(), self.n_layers))

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            try:
                _lib.get().xeq_model_destroy(h)
            except Exception:  # interpreter shutdown: the library may already be gone
                pass

    def __call__(self, data: Dict[str, torch.Tensor], compute_forces: bool = True, compute_virial: bool = False) -> Dict[str, torch.Tensor]:
        data = compute_edge_data(data, compute_forces=False)  # resolves the neighbour structure and the batch bookkeeping
        graph, ptr32 = data[keys.GRAPH], data["_xeq_ptr32"]
        pos = data[keys.POSITIONS].detach().contiguous()
        z = data[keys.ATOMIC_NUMBERS].to(torch.int32).contiguous()
        N, G, dev = pos.shape[0], ptr32.numel() - 1, pos.device
        energy = torch.empty(G, dtype=torch.float32, device=dev)
        e_atom = torch.empty(N, dtype=torch.float32, device=dev)
        want_f = compute_forces or compute_virial  # the virial comes out of the force pass
        forces = torch.empty((N, 3), dtype=torch.float32, device=dev) if want_f else None
        virial = torch.empty((G, 3, 3), dtype=torch.float32, device=dev) if compute_virial else None
        lib = _lib.get()
        nbytes = lib.xeq_model_workspace_bytes(self._handle, graph.struct, int(want_f))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        aux = self._aux.cuda_stream if self._aux is not None else None
        _lib.check(lib.xeq_model_energy_forces_virial(self._handle, graph.struct, _lib.ptr(pos), _lib.ptr(z), _lib.ptr(ptr32),
                                                      _lib.ptr(energy), _lib.ptr(e_atom), _lib.ptr(forces), _lib.ptr(virial),
                                                      _lib.ptr(ws), nbytes, _lib.stream(), aux), "xeq_model_energy_forces_virial")
        out = {keys.TOTAL_ENERGY: energy, keys.ATOMIC_ENERGIES: e_atom}
        if compute_forces:
            out[keys.FORCES] = forces
        if compute_virial:
            out[keys.VIRIAL] = virial
        return out

    def save(self, path: str) -> None:
        """`XEQW0001` | int32 x 9: node_dim, mul0, mul1, mul2, num_basis, n_layers, hidden_dim, embed_dim, n_species |
        float32 cutoff | uint64 n_weights | the blob: everything xeq_model_create() takes, for a host without Python."""
        d = self.dims
        with open(path, "wb") as f:
            f.write(MAGIC)
            f.write(struct.pack("<9i", d.node_dim, d.mul0, d.mul1, d.mul2, d.num_basis, self.n_layers, self.hidden_dim,
                                self.embed_dim, self.n_species))
            f.write(struct.pack("<fQ", d.cutoff, self.blob.numel()))
            f.write(self.blob.cpu().numpy().tobytes())


def read_weight_file(path: str):
    """(header dict, CPU blob) of a file written by NativeModel.save."""
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError(f"{path} is not an XEQW0001 weight file")
        vals = struct.unpack("<9i", f.read(36))
        cutoff, n = struct.unpack("<fQ", f.read(12))
        blob = torch.frombuffer(bytearray(f.read(4 * n)), dtype=torch.float32).clone()
    names = ("node_dim", "mul0", "mul1", "mul2", "num_basis", "n_layers", "hidden_dim", "embed_dim", "n_species")
    return dict(zip(names, vals), cutoff=cutoff), blob
