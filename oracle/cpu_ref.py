"""CPU reference arm: the reference's ps/worker A3C path restated for py3 + torch-CPU fp32.

TEST / BENCH INFRASTRUCTURE ONLY (bench.py's ``cpu_baseline`` and ``--impl reference`` legs).

The literal ``ps_num=1 worker_num=W ./run.sh`` (run.sh:19-31, README.md:25) cannot execute here
(py2-only agent.py, TensorFlow 0.x and gym absent), so this is a port ("kind": "port"):
  * topology of run.sh: 1 shared parameter block ("ps"; torch shared memory, lock-free like
    apply_gradients' use_locking=False, agent.py:321) + W worker processes, one env each,
    one torch thread each (TF's per-op CPU kernels, main.py:45 forces the CPU/NHWC path);
  * per frame, a worker does what agent.py:55-67 does: preprocess with the reference's own
    expression (environment.py:49-53: float64 luma, truncate, cv2.resize), History.add
    (history.py:13-15, float32 shift), a batch-1 forward (agent.py:149), action sampling;
  * every t_max frames: bootstrap forward, n-step returns, loss (network.py:81-94 repaired),
    backward, per-tensor clip_by_norm(40) (agent.py:318-319), RMSProp apply on the shared
    block (main.py:63-65 semantics), lr anneal (agent.py:393-395).
Frames are synthetic 210x160x3 uint8 (the emulator is out of scope on both arms).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _worker(rank, shared, t_max, action_size, cycles_per_step, steps, barrier, counter, seed):
    import torch
    import torch.nn.functional as F
    torch.set_num_threads(1)
    try:
        import cv2
        cv2.setNumThreads(1)
        resize = lambda y: cv2.resize(y, (84, 84))
    except Exception:                       # same fallback order as environment.py:5-12, inverted
        from oracle.preprocess import cv2_resize_linear_u8
        resize = lambda y: cv2_resize_linear_u8(y, (84, 84))
    from oracle import a3c

    names = a3c.PARAM_NAMES
    shapes = a3c.param_shapes(action_size)
    params, rms = shared                    # flat shared tensors
    views, rviews, o = {}, {}, 0
    for k in names:
        n = int(np.prod(shapes[k]))
        views[k] = params[o:o + n].view(shapes[k])
        rviews[k] = rms[o:o + n].view(shapes[k])
        o += n
    rng = np.random.default_rng(seed + rank)
    pool = rng.integers(0, 256, (8, 210, 160, 3), dtype=np.uint8)
    history = np.zeros((4, 84, 84), np.float32)              # history.py:10-11
    step = 0

    def screen(frame):                                       # environment.py:49-53
        y = 0.2126 * frame[:, :, 0] + 0.7152 * frame[:, :, 1] + 0.0722 * frame[:, :, 2]
        return resize(y.astype(np.uint8))

    def forward(p, s):                                       # agent.py:226-252 + network.py:62,79
        x = s.permute(0, 3, 1, 2) / 255.0
        a1 = F.relu(F.conv2d(x, p["l1_w"].permute(3, 2, 0, 1), p["l1_b"], stride=4))
        a2 = F.relu(F.conv2d(a1, p["l2_w"].permute(3, 2, 0, 1), p["l2_b"], stride=2))
        h = F.relu(a2.permute(0, 2, 3, 1).reshape(s.shape[0], -1) @ p["l4_w"] + p["l4_b"])
        return h @ p["p_w"] + p["p_b"], (h @ p["q_w"] + p["q_b"]).reshape(-1)

    for _ in range(4):
        history[:-1] = history[1:]; history[-1] = screen(pool[0])
    fidx = 0
    for _step in range(steps):
        barrier.wait()
        for _c in range(cycles_per_step):
            local = {k: v.clone() for k, v in views.items()}            # theta' <- theta (pull)
            stacks, acts, rews, terms = [], [], [], []
            with torch.no_grad():
                for t in range(t_max):
                    s = torch.from_numpy(np.transpose(history, (1, 2, 0)).copy())[None]
                    logits, _ = forward(local, s)
                    pi = torch.softmax(logits, 1)[0].numpy()
                    a = int(min(np.searchsorted(np.cumsum(pi), rng.random()), action_size - 1))
                    fidx = (fidx + 1) % 8
                    r = float(rng.choice([-1.0, 0.0, 1.0], p=[0.05, 0.9, 0.05]))
                    term = bool(rng.random() < 0.01)
                    stacks.append(s); acts.append(a)
                    rews.append(max(-1.0, min(1.0, r))); terms.append(term)   # agent.py:154
                    history[:-1] = history[1:]; history[-1] = screen(pool[fidx])   # agent.py:156
                s = torch.from_numpy(np.transpose(history, (1, 2, 0)).copy())[None]
                _, vb = forward(local, s)
            R = a3c.nstep_returns(np.array(rews, np.float32)[:, None],
                                  np.array(terms)[:, None], vb.numpy(), 0.99)[:, 0]
            p = {k: v.clone().requires_grad_(True) for k, v in local.items()}
            logits, value = forward(p, torch.cat(stacks))
            total, _, _ = a3c.loss_per_sample(logits, value, torch.tensor(acts),
                                              torch.from_numpy(R.astype(np.float32)), 0.01)
            total.sum().backward()
            lr = a3c.learning_rate(step)
            with torch.no_grad():
                for k in names:                                         # hogwild apply on the "ps"
                    g = p[k].grad
                    g = g * (40.0 / max(float(g.norm()), 40.0))
                    rviews[k].add_((g * g - rviews[k]) * 0.01)
                    views[k].sub_(lr * g / torch.sqrt(rviews[k] + 0.1))
            step += t_max
        with counter.get_lock():
            counter.value += cycles_per_step * t_max
    barrier.wait()


def run(workers, t_max=5, action_size=6, cycles_per_step=20, steps=3, warmup=1, seed=123):
    """Returns per-step wall times (s) and frames per step.  One process per worker."""
    import torch
    import torch.multiprocessing as mp
    from oracle import a3c
    ctx = mp.get_context("spawn")
    flat = torch.from_numpy(a3c.flatten_params(a3c.init_params(action_size, seed))).float()
    params = flat.clone().share_memory_()
    rms = torch.ones_like(flat).share_memory_()
    total = steps + warmup
    barrier = ctx.Barrier(workers + 1)
    counter = ctx.Value("q", 0)
    procs = [ctx.Process(target=_worker, args=(r, (params, rms), t_max, action_size,
                                               cycles_per_step, total, barrier, counter, seed))
             for r in range(workers)]
    for p in procs:
        p.start()
    times = []
    barrier.wait()                                  # step 0 starts
    t0 = time.perf_counter()
    for s in range(total):
        barrier.wait()                              # step s finished (= step s+1 starts)
        t1 = time.perf_counter()
        times.append(t1 - t0)
        t0 = t1
    for p in procs:
        p.join()
    frames_per_step = workers * cycles_per_step * t_max
    return times[warmup:], frames_per_step


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=len(os.sched_getaffinity(0)))
    ap.add_argument("--t-max", type=int, default=5)
    ap.add_argument("--actions", type=int, default=6)
    ap.add_argument("--cycles-per-step", type=int, default=20)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    a = ap.parse_args()
    times, fps = run(a.workers, a.t_max, a.actions, a.cycles_per_step, a.steps, a.warmup)
    print(json.dumps({"frames_per_step": fps, "step_seconds": times, "workers": a.workers,
                      "frames_per_sec": fps * len(times) / sum(times)}))


if __name__ == "__main__":
    main()
