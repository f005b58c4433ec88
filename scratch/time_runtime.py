"""E+F inference: nn-module path (autograd) vs the C inference runtime (xeq_model_energy_forces), eager and as a
captured CUDA graph (K1 + model + forces).  usage: python scratch/time_runtime.py [c1|c5|c4]"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import xequinet_b200 as xb  # noqa: E402
from oracle import xpainn_oracle as orc  # noqa: E402
from xequinet_b200 import runtime  # noqa: E402
from xequinet_b200.replay import CapturedStep  # noqa: E402

DEV = "cuda"


def timed(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main(which):
    cfg = orc.CONFIG_C4 if which == "c4" else orc.CONFIG_DEFAULT
    if which == "c1":
        d = orc.make_molecule_batch(64, 18, seed=0, with_edges=False)
    elif which == "c4":
        d = orc.make_molecule_batch(128, (30, 70), seed=0, with_edges=False, z_table=orc._Z_SPICE)
    else:
        d = orc.make_water_box(15, seed=0)
    data = {k: v.to(DEV) for k, v in d.items() if torch.is_tensor(v)}
    model = xb.resolve_model("xpainn", **cfg.model_kwargs())
    model.load_state_dict(orc.synthetic_state_dict(cfg, 1234), strict=False)
    model = model.to(DEV).eval()
    native = runtime.NativeModel(model)                         # independent branches on a second stream
    native1 = runtime.NativeModel(model, branch_stream=False)   # single-stream schedule
    nt = xb.NeighborTransform(5.0)
    full = nt(dict(data))
    res = {"workload": which, "atoms": int(data["pos"].shape[0]), "edges": int(full["edge_index"].shape[1])}
    ref = model(dict(full), compute_forces=True)
    out = native(dict(full), compute_forces=True)
    res["bit_identical"] = bool(torch.equal(ref["forces"], out["forces"]) and torch.equal(ref["energy"].detach(), out["energy"]))
    lib_launches = xb._lib.get().xeq_launch_count
    n0 = lib_launches(); model(dict(full), compute_forces=True); n1 = lib_launches(); native(dict(full)); n2 = lib_launches()
    res["own_kernels_module_path"], res["own_kernels_runtime"] = int(n1 - n0), int(n2 - n1)
    res["ms_module_eager"] = timed(lambda: model(dict(full), compute_forces=True))
    res["ms_runtime_eager"] = timed(lambda: native(dict(full), compute_forces=True))
    res["ms_runtime_eager_1stream"] = timed(lambda: native1(dict(full), compute_forces=True))
    o1 = native1(dict(full), compute_forces=True)
    res["streams_bit_identical"] = bool(torch.equal(o1["forces"], out["forces"]) and torch.equal(o1["energy"], out["energy"]))
    inp = {k: v for k, v in data.items()}
    for name, m in (("module", model), ("runtime", native), ("runtime_1stream", native1)):
        try:
            step = CapturedStep(m, inp, compute_forces=True)
            res[f"ms_{name}_graph_with_k1"] = timed(lambda: step(inp))
        except Exception as e:  # ragged shapes etc.
            res[f"ms_{name}_graph_with_k1"] = f"failed: {type(e).__name__}: {e}"[:200]
    print(json.dumps(res))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "c1")
