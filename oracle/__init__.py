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
  * network forward (conv2d / linear wiring: kernel shapes, stride lists, NHWC flatten,
    [in,out] matrices, bias and activation order) -- pinned against the reference's own
    ``src/ops.py`` ``conv2d`` / ``linear`` EXECUTED over ``oracle/tf_stub.py`` with the literal
    arguments of its call sites (``oracle/make_golden_network.py`` ->
    ``tests/golden/network_golden.npz``).
  * heads / loss formulas (softmax, log OF the softmax, entropy, log pi(a), policy / value /
    total loss: ``src/network.py:60-94``) and the scalar formulas of the as-running learner
    (epsilon ``agent.py:142-144``, reward clip ``:154``, 1-step Q targets ``:188-190``, async-Q
    loss ``:310-314``, learning rate ``:395``) -- pinned the same way: the reference's own source
    lines are cut out by their text and exec'd (one sample at a time for the loss block, which
    otherwise broadcasts [N]-[N,1]: SURVEY D3) with ``config.M1`` imported unmodified.
  * TensorFlow's numerics, n-step returns (absent upstream), clip_by_norm / RMSProp (TF ops) --
    "parity unpinned": that
    arithmetic lives in TensorFlow 0.x (un-vendored, un-pinned, absent here; the stub restates
    tf.nn.conv2d / tf.matmul from their documented semantics) and the reference ships no tests
    or golden vectors for it.  The restatement follows the reference call sites cited in each
    function.
"""
