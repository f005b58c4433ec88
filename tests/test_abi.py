"""CPU tests of the drop-in boundary: libxeq_b200.so builds, loads and exports exactly the
entry points include/xeq_b200.h declares (no compute without a GPU)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib_path():
    from xequinet_b200 import build

    return build.build()


def _declared():
    text = (ROOT / "include" / "xeq_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xeq_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib_path):
    lib = ctypes.CDLL(str(lib_path))
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/xeq_b200.h but not exported"


def test_ctypes_binding_covers_header(lib_path):
    from xequinet_b200 import _lib

    assert sorted(_lib.exported_symbols()) == _declared()
    lib = _lib.get()
    assert lib.xeq_version() >= 100
    assert lib.xeq_radius_graph_workspace_bytes(1000, 4, 1) > 1000 * 24


def test_struct_layout_matches_header():
    from xequinet_b200 import _lib

    assert ctypes.sizeof(_lib.XeqDims) == 24
    assert ctypes.sizeof(_lib.XeqGraph) == 16 + 10 * 8 + 16
    assert _lib.XeqGraph.rowptr.offset == 16


def test_ops_fail_loudly_without_cuda():
    import torch
    from xequinet_b200 import build_graph

    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        build_graph(torch.zeros(4, 3), 5.0)


def test_md_host_example_links_only_the_c_abi(lib_path):
    """examples/md_host.cpp (the engine-side host of the inference runtime) builds against include/xeq_b200.h and
    links libxeq_b200.so + cudart only: no torch, no Python in the process (it is RUN by tests/test_gpu_runtime.py)."""
    import subprocess
    import sys

    sys.path.insert(0, str(ROOT))
    import __graft_entry__ as entry

    exe = entry.build_md_host()
    assert exe.exists()
    ldd = subprocess.run(["ldd", str(exe)], capture_output=True, text=True).stdout
    assert "libxeq_b200.so" in ldd and "torch" not in ldd and "python" not in ldd.lower()
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 1 and "usage" in r.stderr
