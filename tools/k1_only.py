"""K1 alone (arl_preprocess_push, 4096 frames, 5 frame pools > L2) for ncu captures and timing.
usage: python tools/k1_only.py [iters]"""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cabi = importlib.import_module("async-rl-tensorflow_b200._cabi")
B, R, iters = 4096, 9, int(sys.argv[1]) if len(sys.argv) > 1 else 20
cabi.init("cuda:0")
pools = [torch.randint(0, 256, (B, 210, 160, 3), dtype=torch.uint8, device="cuda") for _ in range(5)]
ring = torch.zeros(B, R, 7056, dtype=torch.uint8, device="cuda")
s = cabi.stream_ptr()
for i in range(5):
    cabi.call("arl_preprocess_push", cabi.ptr(pools[i % 5]), cabi.ptr(ring), B, R, i % R, 1, s)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
for i, (a, b) in enumerate(ev):
    a.record(); cabi.call("arl_preprocess_push", cabi.ptr(pools[i % 5]), cabi.ptr(ring), B, R, i % R, 1, s); b.record()
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in ev)
print("K1 us per launch: min %.1f median %.1f max %.1f  -> %.0f GB/s (median)" %
      (ms[0] * 1e3, ms[len(ms) // 2] * 1e3, ms[-1] * 1e3, B * 87696 / (ms[len(ms) // 2] * 1e-3) / 1e9))
