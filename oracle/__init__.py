"""ORACLE — test infrastructure only.

CPU restatement of the reference's (astooke/accel_rl) rollout-sampler + A2C/PPO path, used as the
checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
Nothing under accel_rl_b200/ may import this package: the product path has no CPU fallback.

Pinning status: the reference ships no tests, golden vectors or fixtures.  The
integer/byte/index parts (frame pipeline, action sampling, buffer layout, GAE, minibatch
indexing) are pinned against the reference's OWN source imported in the build container under
stub modules (oracle/ref_harness.py; fixtures under tests/golden/ with their generator
tests/golden/make_golden.py).  The neural-network arithmetic (Theano/Lasagne/cuDNN, absent) is
restated from the published semantics of those libraries: **parity unpinned** for that part.
"""
