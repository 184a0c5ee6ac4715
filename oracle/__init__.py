"""CPU oracle for the A3C worker hot path of datavizweb/async-rl-tensorflow.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is on the product path:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and there only as the checker or as
the CPU reference being timed.  The product (``async-rl-tensorflow_b200``)
never imports this package and fails loudly if its CUDA library is missing.

Pinning status:
  * preprocess / history  -- PINNED against the reference's own executed code
    (``src/environment.py:49-53`` through a stub ``gym``; ``src/history.py``
    imported as-is) via ``tests/golden/*.npz`` made by
    ``oracle/make_golden.py``.
  * network / returns / loss / clip / RMSProp -- "parity unpinned": the
    arithmetic lives in TensorFlow 0.x (un-vendored, un-pinned, absent here)
    and the reference ships no tests or golden vectors for it.  The
    restatement follows the reference call sites cited in each function.
"""
