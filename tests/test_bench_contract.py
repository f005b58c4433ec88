"""CPU checks of bench.py's bookkeeping: the algorithmic-byte formulas of DESIGN.md / SURVEY.md 8d, and the JSON
line of the reference arm (the CPU oracle timed on the host cores)."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def test_algorithmic_bytes_match_survey():
    import bench
    from oracle import xpainn_oracle as orc

    cfg = orc.CONFIG_DEFAULT
    N, E = 1152, 18698  # c1 (N = 1152, E ~ 18.7k)
    fwd = bench.algorithmic_bytes("edge_fwd", cfg, N, E, periodic=False)
    assert fwd == 9104 * N + 4 * E + 4  # "9104 N + 4 E + 4 B" at the defaults
    assert abs(fwd / 1e6 - 10.6) < 0.1   # "c1: 10.6 MB"
    c4 = orc.CONFIG_C4
    assert bench.algorithmic_bytes("edge_fwd", c4, N, E, periodic=False) == 18192 * N + 4 * E + 4  # "18 192 N + 4 E"
    # periodic graphs add the int8x4 offsets; the derivative kernels read more rows than the forward
    assert bench.algorithmic_bytes("edge_fwd", cfg, N, E, periodic=True) == fwd + 4 * E
    assert bench.algorithmic_bytes("edge_bwd", cfg, N, E, False) > fwd
    assert bench.algorithmic_bytes("edge_bwdbwd", cfg, N, E, False) > bench.algorithmic_bytes("edge_bwd", cfg, N, E, False)


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1",
                          "--warmup", "3"], capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "molecules/s" and line["higher_is_better"] is True
    assert line["steps"] == 1 and line["warmup"] == 3 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "molecules/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("c1")
