"""CPU oracle for the PPO hot path — TEST INFRASTRUCTURE ONLY.

Nothing in the product package (`mujoco_reinforcement_learning_b200/`) imports this
directory.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it, and only as the checker or as
the timed CPU baseline — never as the thing shipped.

Contents
--------
* `ppo_oracle.py`   — restatement, in stock torch CPU ops (the reference *is* stock
  torch CPU ops), of `src/entities/algorithms/ppo.py:62-154`, the MLP modules
  (`src/models/network_block_creator.py`, `src/models/linear/actor.py`,
  `src/models/critic.py`) and of torchrl 0.6.0's `generalized_advantage_estimate`.
* `naive.py`        — an independent float64 numpy restatement (explicit loops,
  hand-written back-propagation and Adam) that the torch restatement is checked
  against.
* `shims.py` + `make_golden.py` — import the reference **verbatim** from
  `/root/reference/src` behind `sys.modules` shims for its uninstalled third-party
  dependencies and write its outputs to `tests/golden/*.npz`.  These two only run in
  the build container (the GPU box has no `/root/reference`); the fixtures travel.

* `build_ref.py` + `ref_runner.py` — stage the reference's own sources VERBATIM under `oracle/_ref` (git-ignored; travels
  to the GPU box with the snapshot) and run its `PPO.calculate_advantages` / `PPO.train` behind the shims: the timed CPU
  arm of `bench.py` (`cpu_baseline.kind == "reference"`).

Parity pinning status
---------------------
* `PPO.calculate_advantages` pre/post-processing and `PPO.train` (forward, losses,
  backward, Adam, minibatch order, dropped tail): **pinned** by outputs of the
  reference's own code executed here (`tests/golden/ref_*.npz`).
* The GAE recurrence itself lives in `torchrl==0.6.0`
  (`requirements.txt:7`), which is neither in `/root/reference` nor installed, and
  the reference has no tests or golden vectors for it: **parity unpinned** for that
  one function.  It is restated from torchrl's published algorithm and cross-checked
  against an independent float64 double loop and analytic known-answer cases.
"""
