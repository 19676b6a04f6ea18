"""Run each pytest test id of the given files in its own process (a device-side trap poisons the CUDA
context, so isolation keeps one GPU call informative).  Usage: python tools/run_isolated.py <pytest args>"""
import subprocess
import sys

args = sys.argv[1:]
ids = subprocess.run([sys.executable, "-m", "pytest", "--collect-only", "-q", "-m", "gpu", *args],
                     capture_output=True, text=True).stdout.splitlines()
ids = [l.strip() for l in ids if "::" in l]
summary = []
for tid in ids:
    try:
        r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-s", "--no-header", "-p", "no:cacheprovider", tid],
                           capture_output=True, text=True, timeout=300)
        ok = r.returncode == 0
        tail = "" if ok else (r.stdout[-3000:] + r.stderr[-1500:])
    except subprocess.TimeoutExpired:
        ok, tail = False, "TIMEOUT"
    summary.append((tid, ok))
    print(("PASS " if ok else "FAIL ") + tid, flush=True)
    for line in r.stdout.splitlines() if ok else []:
        if line.startswith("["):
            print("    " + line, flush=True)
    if not ok:
        print(tail, flush=True)
print("\n==== summary: %d/%d passed" % (sum(1 for _, o in summary if o), len(summary)))
for tid, ok in summary:
    if not ok:
        print("  FAILED:", tid)
