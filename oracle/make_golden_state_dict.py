"""TEST INFRASTRUCTURE ONLY.  Writes tests/golden/state_dict_keys.json: the (name, shape, dtype) list of the REAL
reference model's state_dict (/root/reference/xequinet/nn/model.py through oracle/ref_stubs.py) for the default and
the 256-channel configuration, in the reference's own order.  tests/test_gpu_bench_shapes.py / test_host_logic.py
load a state_dict of exactly this shape with strict=True.

Run in the build container only (needs /root/reference):  python oracle/make_golden_state_dict.py"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_stubs  # noqa: E402
from oracle import xpainn_oracle as orc  # noqa: E402


def main():
    out = {}
    resolve_model = ref_stubs.reference_resolve_model()
    for name, cfg in (("default", orc.CONFIG_DEFAULT), ("c4", orc.CONFIG_C4)):
        torch.manual_seed(0)
        model = resolve_model("xpainn", **cfg.model_kwargs())
        out[name] = [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in model.state_dict().items()]
        print(name, len(out[name]), "entries,", sum(p.numel() for p in model.parameters()), "parameters")
    (ROOT / "tests" / "golden" / "state_dict_keys.json").write_text(json.dumps(out, indent=0))


if __name__ == "__main__":
    main()
