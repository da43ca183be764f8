"""Host-side pieces of bench.py that need no GPU: the reference arm's line (bench.py --impl reference, the restated C
control flow on the host cores) and the watchdog over the legs that follow the timed region."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tail_guard_prints_the_line_when_a_leg_stalls():
    code = (
        "import sys, time; sys.path.insert(0, %r); import bench\n"
        "out = {'metric': 'm', 'value': 1.0, 'e2e': None}\n"
        "g = bench.TailGuard(0, out, 0.3); g.leg = 'e2e'\n"
        "time.sleep(30)\n"
        "print('not reached')\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["value"] == 1.0 and "e2e" in line["tail_note"]


def test_tail_guard_is_silent_when_the_legs_finish():
    code = (
        "import sys, time; sys.path.insert(0, %r); import bench\n"
        "out = {'metric': 'm', 'value': 2.0}\n"
        "g = bench.TailGuard(0, out, 30.0); out['e2e'] = {'value': 3.0}; g.finish()\n"
        "w = bench.TailGuard(1, None, 30.0); w.finish()\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1 and json.loads(lines[0]) == {"metric": "m", "value": 2.0, "e2e": {"value": 3.0}}


def test_reference_arm_line_on_the_small_workload():
    """--impl reference at the n = 2,000 shape: one JSON line, the keys the driver reads."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
