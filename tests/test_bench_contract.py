"""bench.py's reference arm (the reference's CPU path on the host cores, no GPU needed)
prints ONE JSON line with the keys the driver reads; run here on a tiny sample."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    # (torchrun exports OMP_NUM_THREADS=1 to its workers: the arm must not inherit that)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-cells", "8"], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["value"] > 0 and d["unit"] == "particle-steps/s" and d["higher_is_better"] is True
    # "_ref+port": push + deposit = the reference's own headers (oracle/_ref), sort + exchange = the port
    assert d["cpu_baseline"]["kind"] in ("port", "_ref+port")
    n_cores = len(os.sched_getaffinity(0))
    assert d["cpu_baseline"]["cores"] == min(n_cores, 2 * n_cores)  # every core the process may use
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0", "--ref-cells", "8"], capture_output=True, text=True,
                         timeout=600, env=env)
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]
