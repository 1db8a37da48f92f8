"""Development probe: schedule sweep after the counter-stride change (fewer cases)."""
import json, os, subprocess, sys
sys.path.insert(0, ".")
child = os.path.join(os.path.dirname(__file__), "r02_sweep_schedule.py")
grid = []
for rows in (1_250_000, 10_000_000):
    for first, growth, cap in ((None, None, None), ("4096", "32", "32768"), ("8192", "64", "65536"), ("16384", "80", "65536"),
                               ("32768", "40", "65536"), ("16384", "160", "131072")):
        grid.append((rows, 64, 100, first, growth, cap))
for rows, nq, k, first, growth, cap in grid:
    env = dict(os.environ)
    if first:
        env.update(VODB_FIRST_ROWS=first, VODB_GROWTH=growth, VODB_CAP=cap)
    r = subprocess.run([sys.executable, child, "child", str(rows), str(nq), str(k)], env=env, capture_output=True, text=True)
    print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
