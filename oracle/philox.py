"""Philox4x32-10 + inverse-CDF action sampler, CPU restatement.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference's sampler ``batch_sample(self.policy)`` (network.py:72) is imported
from ops.py (network.py:4) but never defined there (ops.py:4-46), so there is no
reference RNG stream to match.  The contract both this file and the CUDA kernel
implement:

  counter = (global_env_id, step_lo, step_hi, 0), key = (seed_lo, seed_hi)
  x       = Philox4x32-10(counter, key)[0]                 (Salmon et al. 2011)
  u       = float32(x >> 8) * 2^-24                        in [0, 1)
  c_j     = c_{j-1} + p_j   in float32, index order        (c_{-1} = 0)
  action  = first j with u < c_j, else A-1

It is bit-exact given identical float32 probabilities.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint32 arrays.  Returns 4 uint32 arrays."""
    c0 = np.asarray(c0, np.uint32).copy()
    c1 = np.asarray(c1, np.uint32).copy()
    c2 = np.asarray(c2, np.uint32).copy()
    c3 = np.asarray(c3, np.uint32).copy()
    k0 = np.asarray(k0, np.uint32).copy()
    k1 = np.asarray(k1, np.uint32).copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & MASK).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return c0, c1, c2, c3


def uniform01(env_ids, step, seed):
    """u in [0,1) float32 for each global env id at env-step ``step``."""
    env_ids = np.asarray(env_ids, np.uint32)
    z = np.zeros_like(env_ids)
    x, _, _, _ = philox4x32_10(env_ids, z + np.uint32(step & 0xFFFFFFFF),
                               z + np.uint32((step >> 32) & 0xFFFFFFFF), z,
                               z + np.uint32(seed & 0xFFFFFFFF),
                               z + np.uint32((seed >> 32) & 0xFFFFFFFF))
    return (x >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def sample_actions(probs, env_ids, step, seed):
    """probs float32 [B, A] -> int32 [B]; inverse CDF with float32 running sum."""
    p = np.asarray(probs, np.float32)
    B, A = p.shape
    u = uniform01(env_ids, step, seed)
    c = np.zeros(B, np.float32)
    act = np.full(B, A - 1, np.int32)
    done = np.zeros(B, bool)
    for j in range(A):
        c = (c + p[:, j]).astype(np.float32)
        hit = (~done) & (u < c)
        act[hit] = j
        done |= hit
    return act


def egreedy_actions(q, env_ids, step, seed, ep):
    """agent.py:141-151 with the draw defined by DESIGN.md: Philox block
    (env, step_lo, step_hi, 1): word 0 -> u < ep, word 1 -> floor(x1 * A / 2^32); else argmax
    (ties -> lowest index, numpy argmax == tf.argmax)."""
    q = np.asarray(q, np.float32)
    B, A = q.shape
    env_ids = np.asarray(env_ids, np.uint32)
    z = np.zeros_like(env_ids)
    x0, x1, _, _ = philox4x32_10(env_ids, z + np.uint32(step & 0xFFFFFFFF),
                                 z + np.uint32((step >> 32) & 0xFFFFFFFF), z + np.uint32(1),
                                 z + np.uint32(seed & 0xFFFFFFFF),
                                 z + np.uint32((seed >> 32) & 0xFFFFFFFF))
    u = (x0 >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    rand = ((x1.astype(np.uint64) * np.uint64(A)) >> np.uint64(32)).astype(np.int32)
    greedy = q.argmax(axis=1).astype(np.int32)
    return np.where(u < np.float32(ep), rand, greedy).astype(np.int32)
