"""BASELINE.json configs[4]: scaling sweep 1K-64K envs/GPU, 18-action head, t_max 5 and 20, on N GPUs.

    python tools/sweep.py --gpus N [--points 1024x5,4096x20,...] [--out profiles/r02_sweep_nN.json]

Runs bench.py once per point (under torchrun for N > 1) and keeps, per point: frames/s, ms per
cycle, the dominant entry and its roofline fraction, every entry's fraction, the end-to-end arm
when requested, CUDA-graph replays and the replica check of the N > 1 runs."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT = "1024x5,2048x5,4096x5,8192x5,16384x5,32768x5,65536x5,1024x20,4096x20,16384x20,65536x20"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--points", default=DEFAULT)
    ap.add_argument("--actions", type=int, default=18)
    ap.add_argument("--e2e", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("--port", type=int, default=29600)
    a = ap.parse_args()
    out, rows = a.out or os.path.join(ROOT, "profiles", "r02_sweep_n%d.json" % a.gpus), []
    for i, pt in enumerate(a.points.split(",")):
        envs, t = (int(v) for v in pt.split("x"))
        frames = envs * t
        steps = 20 if frames <= 100000 else 8 if frames <= 400000 else 4
        cmd = [sys.executable]
        if a.gpus > 1:
            cmd += ["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(a.gpus),
                    "--master-addr", "127.0.0.1", "--master-port", str(a.port + i)]
        cmd += [os.path.join(ROOT, "bench.py"), "--gpus", str(a.gpus), "--envs", str(envs), "--t-max", str(t),
                "--actions", str(a.actions), "--steps", str(steps), "--warmup", "3", "--no-cpu-baseline"]
        if not a.e2e:
            cmd.append("--no-e2e")
        r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=1500)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            rows.append({"envs_per_gpu": envs, "t_max": t, "n_gpus": a.gpus, "failed": (r.stderr or r.stdout)[-600:]})
            print(pt, "FAILED", rows[-1]["failed"][-200:], flush=True)
            continue
        d = json.loads(lines[-1])
        rf = d["roofline"]
        row = {"envs_per_gpu": envs, "t_max": t, "actions": a.actions, "n_gpus": d["n_gpus"],
               "frames_per_s": d["value"], "ms_per_step": d["ms_per_step"], "dominant": rf["entry"],
               "dominant_frac": rf["frac"],
               "entry_frac": {k: round(v["frac"], 3) for k, v in rf["entries"].items()},
               "entry_us_per_launch": {k: round(v["avg_launch_ms"] * 1e3, 1) for k, v in rf["entries"].items()},
               "cycle_frac": rf["chains"].get("cycle", {}).get("frac"),
               "e2e_frames_per_s": (d.get("e2e") or {}).get("value"),
               "cuda_graph_replays": d.get("cuda_graph_replays"), "parity": d.get("parity"),
               "clocks": d.get("clocks")}
        rows.append(row)
        print("%6d envs x t_max %2d  N=%d  %6.2f M frames/s  %8.3f ms  dominant %s %.2f  cycle %.2f" % (
            envs, t, a.gpus, row["frames_per_s"] / 1e6, row["ms_per_step"], row["dominant"],
            row["dominant_frac"], row["cycle_frac"] or 0), flush=True)
        json.dump(rows, open(out, "w"), indent=1)
    json.dump(rows, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
