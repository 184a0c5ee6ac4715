"""Run every GPU test function in its own process (a CUDA fault poisons a context, so one
process per test function keeps failures independent).  Writes gpurun_out/gpu_tests.log."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
os.makedirs("gpurun_out", exist_ok=True)
sel = sys.argv[1:] or ["tests"]
out = subprocess.run([sys.executable, "-m", "pytest", "-m", "gpu", "--collect-only", "-q",
                      "-p", "no:cacheprovider"] + sel, capture_output=True, text=True).stdout
funcs = []
for line in out.splitlines():
    if "::" in line:
        f = line.split("[")[0]
        if f not in funcs:
            funcs.append(f)
log = open("gpurun_out/gpu_tests.log", "w")
summary = []
for f in funcs:
    t = time.time()
    try:
        r = subprocess.run([sys.executable, "-m", "pytest", f, "-q", "-m", "gpu", "-s", "-x",
                            "-p", "no:cacheprovider", "--timeout", "600"],
                           capture_output=True, text=True, timeout=700)
        rc, txt = r.returncode, r.stdout + r.stderr
    except subprocess.TimeoutExpired as e:
        rc, txt = 124, (e.stdout or b"").decode(errors="replace") if isinstance(e.stdout, bytes) else str(e.stdout)
    dt = time.time() - t
    summary.append("%-90s rc=%d %.1fs" % (f, rc, dt))
    log.write("=" * 20 + " " + f + " rc=%d %.1fs\n" % (rc, dt))
    lines = txt.splitlines()
    log.write("\n".join(lines if rc == 0 and len(lines) < 40 else lines[-120:]) + "\n")
    log.flush()
log.write("\n".join(summary) + "\n")
print("\n".join(summary))
