"""The multi-process restatement of ActsrvAltOvrlpSampler (oracle/mp_sampler.py: 2*n_parallel simulator processes in two
alternating groups, semaphores, shared buffers — the sampler leg of bench.py's reference arm) produces exactly the
buffers and trajectory records of the single-process OracleSampler, which is pinned to the REAL reference sampler's
golden buffers (tests/test_oracle_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import mp_sampler, sampler as osampler, synth_ale
from tests.golden.make_golden import RULES, POOL_FRAMES, fake_policy_fn


@pytest.mark.parametrize("mbr,mpl", [(True, 27000), (False, 27000), (True, 9)])
def test_mp_sampler_equals_single_process_oracle(mbr, mpl):
    n_parallel, envs_per, T, itrs = 2, 2, 6, 4
    B = 2 * n_parallel * envs_per
    pool = synth_ale.make_pool(POOL_FRAMES, seed=0)
    A = 4
    one = osampler.OracleSampler(B, T, pool, RULES, A, 0.99, mid_batch_reset=mbr, max_path_length=mpl)
    many = mp_sampler.MpOracleSampler(n_parallel, envs_per, T, pool, RULES, A, 0.99, mid_batch_reset=mbr, max_path_length=mpl)
    try:
        rng = np.random.RandomState(3)
        pf = lambda obs: fake_policy_fn(obs, A)
        key = lambda t: (t["env"], t["Length"], t["Return"], t["RawReturn"], t["NonzeroRewards"], round(t["DiscountedReturn"], 9))
        n_traj = 0
        for _ in range(itrs):
            u = rng.rand(T, B)
            b1, t1 = one.obtain_samples(pf, u)
            b2, t2 = many.obtain_samples(pf, u)
            for k in b1:
                assert np.array_equal(b1[k], b2[k]), k
            assert sorted(map(key, t1)) == sorted(map(key, t2))
            n_traj += len(t1)
        assert n_traj > 0
    finally:
        many.shutdown()
