"""Development probe: schedule sweep after the thread-maximum selection (a select now costs about the same whatever
the list length up to the 32768-entry cache), 64 queries, top-100."""
import os, subprocess, sys
child = os.path.join(os.path.dirname(__file__), "r02_sweep_schedule.py")
for rows in (1_250_000, 10_000_000):
    for first, growth, cap in ((None, None, None), ("8192", "160", "131072"), ("32768", "81", "65536"), ("32768", "320", "262144")):
        env = dict(os.environ)
        if first:
            env.update(VODB_FIRST_ROWS=first, VODB_GROWTH=growth, VODB_CAP=cap)
        r = subprocess.run([sys.executable, child, "child", str(rows), "64", "100"], env=env, capture_output=True, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
