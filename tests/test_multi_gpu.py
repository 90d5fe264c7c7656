"""Domain-decomposed path on real GPUs (needs >= 2 of them on one node; skipped otherwise): tools/dd_check.py under
torchrun -- the merged concentrations of 2 ranks against the oracle (rtol 1e-9 per step) and against the single-GPU
run, mass totals and boundary flux sums as sums of the per-rank partial values, for the default options, the exact
per-colour halo exchange, Jacobi sweeps and fp64 sweeps."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2, 4])
def test_domain_decomposition_matches_oracle_and_single_gpu(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29520 + world), str(ROOT / "tools" / "dd_check.py"), "--side", "150", "--K", "4", "--steps", "5"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=850, cwd=ROOT)
    lines = [json.loads(ln) for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert out.returncode == 0, (out.stdout[-3000:], out.stderr[-3000:])
    assert len(lines) >= 4 and all(ln["ok"] for ln in lines), lines
    for ln in lines:
        assert ln["max_rel_diff_vs_oracle"] < 1e-9 and ln["max_rel_diff_vs_single_gpu"] < 1e-9
        assert ln["mass_end_rel_diff"] < 1e-9 and ln["boundary_flux_sums_rel_diff"] < 1e-9
