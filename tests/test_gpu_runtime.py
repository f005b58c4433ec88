"""GPU tests of the whole-model inference runtime (csrc/model_runtime.cu, xeq_model_energy_forces): energy + forces as
ONE C call without autograd -- the deployment form of the path.  It issues the same kernels
with the same arguments and summation order as the nn modules + torch.autograd.grad, so it must agree with them
BIT FOR BIT; the module path itself is pinned to the reference's golden vectors (tests/test_gpu_parity.py)."""
import numpy as np
import pytest
import torch

from helpers import cast_data, force_gate, load_golden
from oracle import xpainn_oracle as orc

pytestmark = pytest.mark.gpu

import xequinet_b200 as xb  # noqa: E402
from xequinet_b200 import runtime  # noqa: E402

DEV = "cuda"


def _dev(data):
    return {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in data.items() if k != "pbc"}


def _model(cfg, seed):
    model = xb.resolve_model("xpainn", **cfg.model_kwargs())
    model.load_state_dict(orc.synthetic_state_dict(cfg, seed), strict=False)
    return model.to(DEV).eval()


@pytest.mark.parametrize("name", ["mol_small", "mol_c4_small", "pbc_small", "pbc_tiny", "pbc_slab", "pbc_two_graphs"])
def test_runtime_is_bit_identical_to_the_module_path_and_matches_the_golden(name):
    z, cfg, data = load_golden(name)
    model = _model(cfg, int(z["sd_seed"]))
    native = runtime.NativeModel(model)
    ref = model(_dev(cast_data(data, torch.float32)), compute_forces=True)
    out = native(_dev(cast_data(data, torch.float32)), compute_forces=True)
    assert set(out) == {"energy", "atomic_energies", "forces"}
    for k in out:
        assert torch.equal(out[k], ref[k].detach()), (k, float((out[k] - ref[k]).abs().max()))
    np.testing.assert_allclose(out["energy"].cpu().numpy(), z["f64:energy"], rtol=1e-5, atol=1e-6)
    force_gate(out["forces"].cpu().numpy(), z["f64:forces"], z["f32:forces"])
    one = runtime.NativeModel(model, branch_stream=False)(_dev(cast_data(data, torch.float32)), compute_forces=True)
    for k in out:  # the second stream only reorders independent launches
        assert torch.equal(out[k], one[k]), k
    e_only = native(_dev(cast_data(data, torch.float32)), compute_forces=False)
    assert set(e_only) == {"energy", "atomic_energies"} and torch.equal(e_only["energy"], out["energy"])


def test_runtime_c1_shape_with_k1_graph_and_cuda_graph_replay():
    """64 x 18 atoms (config c1), neighbour list from K1; then the whole E+F evaluation captured as ONE CUDA graph
    (the runtime allocates nothing and never synchronises) and replayed on moved atoms."""
    cfg = orc.CONFIG_DEFAULT
    model = _model(cfg, 1234)
    native = runtime.NativeModel(model)
    d0 = orc.make_molecule_batch(64, 18, seed=0, with_edges=False)
    data = xb.NeighborTransform(5.0)(_dev(d0))
    ref = model(dict(data), compute_forces=True)
    out = native(dict(data), compute_forces=True)
    for k in out:
        assert torch.equal(out[k], ref[k].detach()), k
    # capture: static positions buffer, same neighbour structure (small displacements keep every pair inside the list)
    static = dict(data)
    static["pos"] = data["pos"].clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        native(dict(static))  # warm-up outside the capture
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        cap = native(dict(static))
    moved = data["pos"] + 1e-3 * torch.randn_like(data["pos"])
    static["pos"].copy_(moved)
    graph.replay()
    eager = native(dict(data, pos=moved))
    for k in cap:
        assert torch.equal(cap[k], eager[k]), k
    assert float((cap["forces"] - out["forces"]).abs().max()) > 0


def test_runtime_edge_cases_and_weight_file(tmp_path):
    cfg = orc.CONFIG_DEFAULT
    model = _model(cfg, 7)
    native = runtime.NativeModel(model)
    # isolated atoms (rows without edges) next to a dimer and a small molecule, three graphs
    pos = torch.tensor([[0.0, 0, 0], [50.0, 0, 0], [51.1, 0, 0], [100.0, 0, 0], [100.9, 0.3, 0], [100.2, 1.0, 0.4]], device=DEV)
    data = {"pos": pos, "atomic_numbers": torch.tensor([8, 1, 1, 6, 1, 8], device=DEV),
            "batch": torch.tensor([0, 1, 1, 2, 2, 2], device=DEV), "ptr": torch.tensor([0, 1, 3, 6], device=DEV)}
    data = xb.NeighborTransform(5.0)(data)
    ref = model(dict(data), compute_forces=True)
    out = native(dict(data), compute_forces=True)
    for k in out:
        assert torch.equal(out[k], ref[k].detach()), k
    assert float(out["forces"][0].abs().max()) == 0.0
    # the runtime reads a copy of the weights: refresh() follows a changed model
    with torch.no_grad():
        model.mods["output_energy"].out_mlp[2].weight.mul_(2.0)
    assert not torch.equal(native(dict(data))["energy"], model(dict(data), compute_forces=False)["energy"].detach())
    native.refresh(model)
    assert torch.equal(native(dict(data))["energy"], model(dict(data), compute_forces=False)["energy"].detach())
    # the weight file a C / C++ host reads
    path = tmp_path / "model.xeqw"
    native.save(str(path))
    hdr, blob = runtime.read_weight_file(str(path))
    assert hdr["node_dim"] == 128 and hdr["n_layers"] == 3 and hdr["n_species"] == 87 and hdr["cutoff"] == 5.0
    assert torch.equal(blob, native.blob.cpu())
    # too small a workspace is refused, not overrun
    from xequinet_b200 import _lib
    lib = _lib.get()
    g = data["_xeq_graph"]
    e = torch.empty(3, device=DEV); ea = torch.empty(6, device=DEV); ws = torch.empty(1024, dtype=torch.uint8, device=DEV)
    rc = lib.xeq_model_energy_forces(native._handle, g.struct, pos.data_ptr(), data["atomic_numbers"].to(torch.int32).data_ptr(),
                                     data["ptr"].to(torch.int32).data_ptr(), e.data_ptr(), ea.data_ptr(), None, ws.data_ptr(), 1024,
                                     _lib.stream())
    assert rc == -3 and b"workspace too small" in lib.xeq_last_error()


def test_python_free_host_process_matches_the_module_path(tmp_path):
    """examples/md_host.cpp: a C++ process that links libxeq_b200.so only (no Python, no torch), reads the weight file
    written by NativeModel.save(), builds its neighbour list with K1 through the C ABI and calls
    xeq_model_energy_forces -- the engine side of the deployment path.  Bit-identical to model(data) in this process."""
    import struct
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    exe = root / "examples" / "md_host"
    if not exe.exists():
        sys.path.insert(0, str(root))
        import __graft_entry__ as entry
        entry.build_md_host()
    ldd = subprocess.run(["ldd", str(exe)], capture_output=True, text=True).stdout
    assert "libxeq_b200.so" in ldd and "torch" not in ldd and "python" not in ldd
    cfg = orc.CONFIG_DEFAULT
    model = _model(cfg, 1234)
    runtime.NativeModel(model).save(str(tmp_path / "model.xeqw"))
    d = orc.make_molecule_batch(1, 100, seed=4, with_edges=False)  # > 64 atoms: edge-block tiles on both sides
    n = d["pos"].shape[0]
    with open(tmp_path / "structure.bin", "wb") as f:
        f.write(struct.pack("<i", n))
        f.write(d["pos"].float().numpy().tobytes())
        f.write(d["atomic_numbers"].to(torch.int32).numpy().tobytes())
    ref = model(xb.NeighborTransform(5.0)(_dev(d)), compute_forces=True)
    for extra in ([], ["graph"]):  # eager launches; the evaluation captured as a CUDA graph by the C++ host itself
        r = subprocess.run([str(exe), str(tmp_path / "model.xeqw"), str(tmp_path / "structure.bin"), str(tmp_path / "out.bin"), "20"] + extra,
                           capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
        print(r.stdout.strip())
        raw = np.fromfile(tmp_path / "out.bin", dtype=np.float32)
        assert raw.size == 1 + n + 3 * n
        assert raw[0] == float(ref["energy"].detach()[0])
        assert np.array_equal(raw[1 : 1 + n], ref["atomic_energies"].detach().cpu().numpy())
        assert np.array_equal(raw[1 + n :].reshape(n, 3), ref["forces"].cpu().numpy())


@pytest.mark.parametrize("which", ["c4", "box"])
def test_runtime_at_benchmark_shapes(which):
    """c4: 128 molecules of 30..70 atoms at 256 channels (two channel slices, mixed staged / unstaged tiles);
    box: a 1536-atom periodic water box (cell-list K1, edge-block tiles, L2 gathers) -- bit-identical to the module path."""
    if which == "c4":
        cfg = orc.CONFIG_C4
        d = orc.make_molecule_batch(128, (30, 70), seed=1, with_edges=False, z_table=orc._Z_SPICE)
    else:
        cfg = orc.CONFIG_DEFAULT
        d = orc.make_water_box(8, seed=2)
    model = _model(cfg, 1234)
    native = runtime.NativeModel(model)
    data = xb.NeighborTransform(5.0)({k: v.to(DEV) for k, v in d.items() if torch.is_tensor(v)})
    data.pop("pbc", None)
    ref = model(dict(data), compute_forces=True)
    out = native(dict(data), compute_forces=True)
    for k in out:
        assert torch.equal(out[k], ref[k].detach()), k
    assert float(out["forces"].abs().max()) > 1e-3


def test_runtime_ghost_nodes_get_forces_but_contribute_no_energy():
    """An MD engine appends ghost atoms after its local atoms and sums the energy over the local ones (seg_ptr covers
    the local atoms only): forces = -d(sum of LOCAL atomic energies)/d(all positions), ghosts included."""
    cfg = orc.CONFIG_DEFAULT
    model = _model(cfg, 1234)
    native = runtime.NativeModel(model)
    d = orc.make_molecule_batch(1, 40, seed=6, with_edges=False)
    data = xb.NeighborTransform(5.0)(_dev(d))
    n, n_loc = 40, 29
    pos = data["pos"].clone().requires_grad_(True)
    ref = model(dict(data, pos=pos), compute_forces=False)
    e_loc = ref["atomic_energies"][:n_loc].sum()
    f_ref = -torch.autograd.grad(e_loc, pos)[0]
    from xequinet_b200 import _lib
    lib = _lib.get()
    g = data["_xeq_graph"]
    seg = torch.tensor([0, n_loc], dtype=torch.int32, device=DEV)
    e, ea, f = torch.empty(1, device=DEV), torch.empty(n, device=DEV), torch.empty(n, 3, device=DEV)
    nbytes = lib.xeq_model_workspace_bytes(native._handle, g.struct, 1)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    _lib.check(lib.xeq_model_energy_forces(native._handle, g.struct, data["pos"].data_ptr(), data["atomic_numbers"].to(torch.int32).data_ptr(),
                                           seg.data_ptr(), e.data_ptr(), ea.data_ptr(), f.data_ptr(), ws.data_ptr(), nbytes, _lib.stream()),
               "xeq_model_energy_forces")
    assert torch.equal(ea, ref["atomic_energies"].detach())
    torch.testing.assert_close(e[0], e_loc.detach(), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(f, f_ref, rtol=1e-5, atol=2e-6)
    assert float(f[n_loc:].abs().max()) > 1e-3  # ghosts do receive forces


@pytest.mark.parametrize("name", ["mol_small", "pbc_small", "pbc_two_graphs"])
def test_graph_from_coo_feeds_the_runtime(name):
    """xeq_graph_from_coo: an engine's sorted COO edge list (+ cell offsets) -> xeq_graph_t in one call; the runtime on
    that structure (edge-block tiles) agrees bit for bit with the module path (molecule tiles for the small molecules)."""
    import ctypes

    from xequinet_b200 import _lib

    z, cfg, data = load_golden(name)
    model = _model(cfg, int(z["sd_seed"]))
    native = runtime.NativeModel(model)
    d = _dev(cast_data(data, torch.float32))
    ref = model(dict(d), compute_forces=True)
    ei, co = orc.canonical_sort(data["edge_index"], data.get("cell_offsets"))
    ei = ei.to(DEV).contiguous()
    periodic = "cell" in d
    co = co.float().to(DEV).contiguous() if periodic else None
    cell = d["cell"].reshape(-1, 3, 3).contiguous() if periodic else None
    N, E, G = d["pos"].shape[0], ei.shape[1], d["ptr"].numel() - 1
    node_graph = d["batch"].to(torch.int32).contiguous()
    lib = _lib.get()
    nbytes = lib.xeq_graph_from_coo_bytes(N, E, int(periodic))
    storage = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    g = _lib.XeqGraph()
    _lib.check(lib.xeq_graph_from_coo(ei.data_ptr(), _lib.ptr(co), _lib.ptr(cell), node_graph.data_ptr(), N, E, G,
                                      storage.data_ptr(), nbytes, ctypes.byref(g), _lib.stream()), "xeq_graph_from_coo")
    assert g.n_nodes == N and g.n_edges == E and g.tile_mode == 0 and bool(g.offsets) == periodic
    seg = d["ptr"].to(torch.int32).contiguous()
    e, ea, f = torch.empty(G, device=DEV), torch.empty(N, device=DEV), torch.empty(N, 3, device=DEV)
    wbytes = lib.xeq_model_workspace_bytes(native._handle, ctypes.byref(g), 1)
    ws = torch.empty(wbytes, dtype=torch.uint8, device=DEV)
    _lib.check(lib.xeq_model_energy_forces(native._handle, ctypes.byref(g), d["pos"].data_ptr(), d["atomic_numbers"].to(torch.int32).data_ptr(),
                                           seg.data_ptr(), e.data_ptr(), ea.data_ptr(), f.data_ptr(), ws.data_ptr(), wbytes, _lib.stream()),
               "xeq_model_energy_forces")
    assert torch.equal(e, ref["energy"].detach()) and torch.equal(ea, ref["atomic_energies"].detach())
    assert torch.equal(f, ref["forces"])
    # too small a buffer is refused
    assert lib.xeq_graph_from_coo(ei.data_ptr(), _lib.ptr(co), _lib.ptr(cell), node_graph.data_ptr(), N, E, G, storage.data_ptr(), 256,
                                  ctypes.byref(g), _lib.stream()) == -3


@pytest.mark.parametrize("name", ["mol_small", "pbc_small", "pbc_slab", "pbc_two_graphs"])
def test_runtime_virial_matches_module_path_and_reference_golden(name):
    """xeq_model_energy_forces_virial: the strain-trick virial (nn/basic.py:93-107, 162-199) assembled from the force
    pass's own records, against the autograd module path and the reference's fp64 virial (tests/golden/virial.npz)."""
    from helpers import GOLDEN

    z, cfg, data = load_golden(name)
    gold = np.load(GOLDEN / "virial.npz")[f"{name}:virial"]
    model = _model(cfg, int(z["sd_seed"]))
    native = runtime.NativeModel(model)
    ref = model(_dev(cast_data(data, torch.float32)), compute_forces=True, compute_virial=True)
    out = native(_dev(cast_data(data, torch.float32)), compute_forces=True, compute_virial=True)
    assert set(out) == {"energy", "atomic_energies", "forces", "virial"}
    assert torch.equal(out["forces"], ref["forces"].detach()) and torch.equal(out["energy"], ref["energy"].detach())
    vir, vref = out["virial"].cpu().numpy(), ref["virial"].detach().cpu().numpy()
    scale = max(1.0, float(np.abs(gold).max()))
    assert np.abs(vir - vref).max() <= 2e-5 * scale, (np.abs(vir - vref).max(), scale)
    assert np.abs(vir - gold).max() <= 3e-4 * scale, (np.abs(vir - gold).max(), scale)
    np.testing.assert_allclose(vir, np.swapaxes(vir, 1, 2), atol=1e-6 * scale)
    v_only = native(_dev(cast_data(data, torch.float32)), compute_forces=False, compute_virial=True)
    assert "forces" not in v_only and np.array_equal(v_only["virial"].cpu().numpy(), vir)
