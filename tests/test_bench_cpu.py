"""bench.py's CPU arm (`--impl reference`: the oracle port timed in a bounded child process) prints the contract's JSON
line and terminates on a host without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    env = dict(os.environ, B200ST_CPU_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=400, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    if "unavailable" in line:            # a saturated host: the arm must still answer within its budget
        return
    assert line["metric"] == "audio_frames_per_sec_fwd_bwd" and line["unit"] == "frames/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 4
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
