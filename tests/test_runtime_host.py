"""CPU tests of the inference runtime's host side (xequinet_b200/runtime.py, `xeq_model_*`): the weight blob covers
every parameter of the default model exactly once and has the length the C layout expects; argument validation."""
import ctypes

import pytest
import torch

import xequinet_b200 as xb
from oracle import xpainn_oracle as orc
from xequinet_b200 import _lib, runtime


@pytest.mark.parametrize("cfg", [orc.CONFIG_DEFAULT, orc.CONFIG_C4, orc.XPaiNNConfig(action_blocks=2)])
def test_weight_blob_covers_the_state_dict(cfg):
    model = xb.resolve_model("xpainn", **cfg.model_kwargs())
    order = runtime.weight_order(cfg.action_blocks)
    assert len(order) == len(set(order))
    params = {k for k, _ in model.named_parameters()}
    assert params <= set(order), params - set(order)
    assert set(order) - params == {"mods.embedding.embedding.0.embed_ten"}
    blob = runtime.export_weights(model.state_dict(), cfg.action_blocks)
    d = _lib.XeqDims(cfg.node_dim, *cfg.muls, cfg.num_basis, cfg.cutoff)
    want = _lib.get().xeq_model_weight_count(ctypes.byref(d), cfg.action_blocks, cfg.hidden_dim, cfg.embed_dim, 87)
    assert blob.numel() == want and blob.numel() % 4 == 0
    # first tensor = the embedding table, last = the read-out bias (padded to 4)
    assert torch.equal(blob[: 87 * cfg.embed_dim], model.state_dict()["mods.embedding.embedding.0.embed_ten"].reshape(-1))
    assert float(blob[-4]) == float(model.state_dict()["mods.output_energy.out_mlp.2.bias"]) and float(blob[-3:].abs().max()) == 0.0


def test_model_create_validates_its_arguments():
    lib = _lib.get()
    d = _lib.XeqDims(128, 128, 64, 32, 20, 5.0)
    h = ctypes.c_void_p()
    n = lib.xeq_model_weight_count(ctypes.byref(d), 3, 64, 56, 87)
    fake = ctypes.c_void_p(4096)  # never dereferenced by create
    assert lib.xeq_model_create(ctypes.byref(d), 3, 64, 56, 87, fake, n - 4, ctypes.byref(h)) != 0
    assert b"layout needs" in lib.xeq_last_error()
    assert lib.xeq_model_create(ctypes.byref(d), 0, 64, 56, 87, fake, n, ctypes.byref(h)) != 0
    assert lib.xeq_model_create(ctypes.byref(d), 3, 64, 56, 87, ctypes.c_void_p(4100), n, ctypes.byref(h)) != 0
    bad = _lib.XeqDims(128, 96, 64, 32, 20, 5.0)
    assert lib.xeq_model_create(ctypes.byref(bad), 3, 64, 56, 87, fake, n, ctypes.byref(h)) != 0
    assert lib.xeq_model_create(ctypes.byref(d), 3, 64, 56, 87, fake, n, ctypes.byref(h)) == 0 and h.value
    lib.xeq_model_destroy(h)
    assert lib.xeq_model_weight_count(ctypes.byref(d), 99, 64, 56, 87) == 0


def test_native_model_rejects_what_the_runtime_does_not_run():
    with pytest.raises(NotImplementedError):
        runtime.NativeModel(xb.resolve_model("xpainn", charge_embed=True))
    with pytest.raises(NotImplementedError):
        runtime.NativeModel(xb.resolve_model("xpainn", output_modes=["energy", "dipole"]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        runtime.NativeModel(xb.resolve_model("xpainn"))
