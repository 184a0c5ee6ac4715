"""How close to its 1e-3 bar does __graft_entry__.smoke() run?  The same cycle for several seeds of the
per-env random-start draws (twice each: the result must not depend on timing) and batch sizes.
usage (GPU box): python tools/smoke_scan.py [seeds]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
for B, T in ((8, 2), (32, 2)):
    for seed in range(n):
        runs = [ge.smoke_errors(seed, B, T) for _ in range(2)]
        worst = max(runs[0], key=lambda k: runs[0][k][0])
        same = all(runs[0][k] == runs[1][k] for k in runs[0])
        print("B=%d T=%d seed %d: worst gradient %s %.2e (param %.1e)  l1_w %.2e  repeat identical: %s"
              % (B, T, seed, worst, runs[0][worst][0], max(v[1] for v in runs[0].values()),
                 runs[0]["l1_w"][0], same), flush=True)
