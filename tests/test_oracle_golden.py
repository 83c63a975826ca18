"""Oracle restatement vs the golden fixtures generated from the reference's own code
(tests/golden/make_golden.py).  CPU only."""
import os
import zlib

import numpy as np
import pytest

from oracle import frame as oframe, gae as ogae, sampler as osampler, synth_ale, net as onet

from tests.golden.make_golden import RULES, POOL_FRAMES, fake_policy_fn


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_frame_pipeline_matches_reference(golden_dir):
    g = _load(golden_dir, "frames.npz")
    got = oframe.update_obs_batch(g["stacks"], g["raw1"], g["raw2"], g["reset"])
    assert got.dtype == np.uint8 and got.shape == g["out"].shape
    assert np.array_equal(got, g["out"])


@pytest.mark.parametrize("A", [4, 6, 18])
def test_weighted_sample_n_matches_reference(golden_dir, A):
    g = _load(golden_dir, "sampling.npz")
    got = osampler.weighted_sample_n(g["p%d" % A], g["u%d" % A], A)
    assert got.dtype == np.uint8
    assert np.array_equal(got, g["a%d" % A])


def test_uniform_stream_is_legacy_mt19937(golden_dir):
    g = _load(golden_dir, "sampling.npz")
    np.random.seed(77 + 4)
    assert np.array_equal(np.random.rand(257), g["u4"])


@pytest.mark.parametrize("name,lam", [("gae", 0.95), ("ret", 1.0)])
def test_gae_matches_reference(golden_dir, name, lam):
    g = _load(golden_dir, "gae.npz")
    B, T = g["r"].shape
    adv, ret, _, _ = ogae.process_samples(g["r"].ravel(), g["v"].ravel(), g["d"].ravel(), g["nr"].ravel(), g["lv"],
                                          0.99, lam, T)
    np.testing.assert_allclose(adv.reshape(B, T), g["adv_" + name], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ret.reshape(B, T), g["ret_" + name], rtol=1e-5, atol=1e-6)
    adv, ret, valids, v2 = ogae.process_samples(g["r"].ravel(), g["v"].ravel(), g["d"].ravel(), g["nr"].ravel(),
                                                g["lv"], 0.99, lam, T, use_valids=True)
    assert np.array_equal(valids.reshape(B, T), g["valids"])
    np.testing.assert_allclose(adv.reshape(B, T), g["adv_%s_valid" % name], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ret.reshape(B, T), g["ret_%s_valid" % name], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(v2.reshape(B, T), g["v_valid"], rtol=0, atol=0)


def test_minibatch_indices_match_reference(golden_dir):
    g = _load(golden_dir, "mb_idxs.npz")
    rng = np.random.RandomState(int(g["seed"]))
    for row in g["idx"]:
        got = np.concatenate(list(onet.iterate_mb_idxs(int(g["batch"]), int(g["length"]), rng)))
        assert np.array_equal(got, row)
    # host-side product helper consumes the GLOBAL stream the same way
    from accel_rl_b200.optimizers.util import iterate_mb_idxs, epoch_index_block
    np.random.seed(int(g["seed"]))
    for row in g["idx"]:
        got = np.concatenate([b[0] for b in iterate_mb_idxs(int(g["batch"]), int(g["length"]), shuffle=True)])
        assert np.array_equal(got, row)
    np.random.seed(int(g["seed"]))
    block, n_mb = epoch_index_block(int(g["batch"]), int(g["length"]), 3, True)
    assert n_mb == 6 and np.array_equal(block.reshape(3, -1), g["idx"])


@pytest.mark.parametrize("tag,mbr", [("reset", True), ("nonreset", False), ("overlength", True)])
def test_sampler_restatement_matches_real_reference_sampler(golden_dir, tag, mbr):
    g = _load(golden_dir, "sampler_%s.npz" % tag)
    B, T, itrs = int(g["n_envs"]), int(g["horizon"]), int(g["itrs"])
    pool = synth_ale.make_pool(POOL_FRAMES, seed=0)
    s = osampler.OracleSampler(B, T, pool, RULES, n_actions=4, discount=0.99, mid_batch_reset=mbr,
                               max_path_length=int(g["max_path_length"]))
    for itr in range(itrs):
        buf, infos = s.obtain_samples(lambda o: fake_policy_fn(o, 4), g["uniforms"][itr])
        if itr == 0:
            assert np.array_equal(buf["observations"], g["obs_0"])
        crc = np.array([zlib.crc32(r.tobytes()) for r in buf["observations"]], dtype=np.uint32)
        assert np.array_equal(crc, g["obscrc_%d" % itr]), "observation rows differ at itr %d" % itr
        assert np.array_equal(buf["extra_observations"], g["extra_%d" % itr])
        assert np.array_equal(buf["rewards"], g["rew_%d" % itr])
        assert np.array_equal(buf["dones"], g["done_%d" % itr])
        assert np.array_equal(buf["raw_reward"], g["raw_%d" % itr])
        assert np.array_equal(buf["need_reset"], g["nr_%d" % itr])
        assert np.array_equal(buf["actions"], g["act_%d" % itr])
        np.testing.assert_array_equal(buf["prob"], g["prob_%d" % itr])
        np.testing.assert_array_equal(buf["value"], g["val_%d" % itr])
        ti = sorted([(i["Length"], float(i["Return"]), float(i["RawReturn"]), int(i["NonzeroRewards"]),
                      float(i["DiscountedReturn"])) for i in infos])
        want = g["traj_%d" % itr]
        assert len(ti) == len(want)
        if len(ti):
            np.testing.assert_allclose(np.array(ti, dtype=np.float64), want, rtol=1e-6)


def test_fixture_covers_edge_cases(golden_dir):
    g = _load(golden_dir, "sampler_reset.npz")
    dones = np.concatenate([g["done_%d" % i] for i in range(int(g["itrs"]))])
    nr = np.concatenate([g["nr_%d" % i] for i in range(int(g["itrs"]))])
    assert dones.any() and nr.any() and (dones & ~nr).any(), "fixture must contain life losses and game overs"
    g2 = _load(golden_dir, "sampler_overlength.npz")
    assert sum(len(g2["traj_%d" % i]) for i in range(int(g2["itrs"]))) > 0


def test_net_param_count_matches_reference_comment():
    # accel_rl/policies/atari_cnn_specs.py:22 "3.6M params", :10 "900k params"
    assert onet.n_params(onet.CNN_SPECS[1], (4, 104, 80), 4) == 3620005
    assert onet.n_params(onet.CNN_SPECS[0], (4, 104, 80), 4) == 898613


def test_rgb_frame_oracle_known_answers(golden_dir):
    """north-star RGB mode (builder-defined): the oracle reproduces its committed known-answer fixture, the area
    weights partition every output cell exactly, and flat frames stay flat"""
    from oracle import frame as oframe
    g = np.load(golden_dir + "/frames_rgb.npz")
    stack = np.zeros_like(g["stacks"][0])
    for s in range(g["raw_a"].shape[0]):
        stack = oframe.rgb_update_obs_batch(stack, g["raw_a"][s], g["raw_b"][s], g["reset"][s])
        assert np.array_equal(stack, g["stacks"][s])
    assert (oframe._WY.sum(1) == 5).all() and (oframe._WX.sum(1) == 40).all()
    assert (oframe._WY.sum(0) == 2).all() and (oframe._WX.sum(0) == 21).all()
    for v in (0, 1, 127, 255):
        flat = np.full((210, 160, 3), v, np.uint8)
        assert (oframe.rgb_downsample(oframe.rgb_to_gray(flat)) == v).all()
    assert oframe.rgb_to_gray(np.array([[255, 0, 0], [0, 255, 0], [0, 0, 255]], np.uint8)).tolist() == [77, 149, 29]   # (w*255 + 128) >> 8


@pytest.mark.parametrize("kind", ["adam", "rmsprop"])
def test_update_rules_match_executed_reference(golden_dir, kind):
    """oracle Adam / RMSProp vs OUTPUTS of the reference's own statement of the rules: update_methods_stats.py:11-32,
    :55-87 executed unmodified under oracle/theano_shim.py (tests/golden/make_golden_updates.py), six steps with a
    changing lr_mult and exact-zero gradients.  Differences are 1-2 ulp of the parameter (float32 vs float64
    evaluation of Adam's bias-correction scalar)."""
    g = _load(golden_dir, "update_rules.npz")
    n = g["p0"].size
    opt = onet.Adam(n, 1e-3, epsilon=1e-5) if kind == "adam" else onet.RMSProp(n, 7e-4)
    p = g["p0"].copy()
    for t in range(len(g["grads"])):
        p = opt.step(p, g["grads"][t], float(g["lr_mults"][t]))
        assert p.dtype == np.float32
        np.testing.assert_allclose(p, g[kind][t], rtol=1e-6, atol=3e-8)
        upd, want = p - g["p0"], g[kind][t] - g["p0"]
        assert np.abs(upd - want).max() <= 2e-5 * np.abs(want).max()


def test_theano_shim_update_semantics():
    """the eager stand-in applies a function's updates simultaneously and keeps its shared state across steps"""
    from collections import OrderedDict
    from oracle import theano_shim as S
    sess = S.Session()

    def fn(g, p):
        acc = sess.shared(np.zeros(2, np.float32))
        new_acc = acc + g
        up = OrderedDict()
        up[acc] = new_acc
        up[p] = p - new_acc          # uses the NEW accumulator expression, evaluated from pre-update values
        return up, []
    p = S.Shared(np.array([1.0, 2.0], np.float32))
    for _ in range(3):
        sess.step(fn, np.array([0.5, 1.0], np.float32), p)
    assert len(sess.vars) == 1 and np.array_equal(sess.vars[0].value, [1.5, 3.0])
    assert np.array_equal(p.value, np.array([1.0 - 0.5 - 1.0 - 1.5, 2.0 - 1.0 - 2.0 - 3.0], np.float32))


def test_loss_terms_vs_the_reference_expressions(golden_dir):
    """tests/golden/loss_terms.npz holds the VALUES of the reference's own loss expressions — `pi_loss` of algos/pg/ppo.py
    and a2c.py, the value / entropy terms and pi_kl / v_kl of aac_base.py:60-70, categorical.py's *_sym functions and
    valids_mean — executed on arrays under the eager theano.tensor stand-in (tests/golden/make_golden_losses.py).  The oracle's
    loss terms (oracle/net.py:losses), which the CUDA loss kernel is tested against, must reproduce them: PPO with the clip
    range scaled by lr_mult, exact ties of the two surrogates, A2C, with and without the validity mask."""
    import torch
    from oracle import net as onet
    from accel_rl_b200.distributions.categorical import Categorical
    g = np.load(os.path.join(golden_dir, "loss_terms.npz"))
    t = lambda k: torch.tensor(g[k])
    for tag in g["cases"]:
        algo, v, lr = str(tag).split("_")
        use_valids, lr_mult = v == "v1", float(lr[2:])
        valids = t("valids") if use_valids else None
        pl, vl, el = onet.losses(t("new_prob"), t("new_value"), t("act"), t("adv"), t("ret"), t("old_prob"), algo,
                                 clip_param=0.2, lr_mult=lr_mult, v_coeff=1.0 if algo == "ppo" else 0.25, ent_coeff=0.01,
                                 valids=valids)
        want = g[str(tag)]
        np.testing.assert_allclose([float(pl), float(vl), float(el)], want[:3], rtol=2e-6, atol=1e-8, err_msg=str(tag))
        # pi_kl / v_kl (aac_base.py:68-70): the host diagnostic's formula (algos/pg/aac_base.py:constraint_values)
        kl = Categorical(6).kl(dict(prob=g["old_prob"]), dict(prob=g["new_prob"]))
        dv = (g["new_value"] - g["old_value"]) ** 2
        if use_valids:
            w = g["valids"].astype(np.float32)
            got = [float((kl * w).sum() / w.sum()), float((dv * w).sum() / w.sum())]
        else:
            got = [float(kl.mean()), float(dv.mean())]
        np.testing.assert_allclose(got, want[3:], rtol=2e-6, err_msg=str(tag))
    assert {str(c).split("_")[0] for c in g["cases"]} == {"ppo", "a2c"}


def test_dense_initialiser_vs_the_reference_normc_init(golden_dir):
    """tests/golden/norm_c_init.npz: outputs of the reference's NormCInit.sample (policies/layers.py:9-19) drawn from the
    global numpy stream in the network builder's order (hidden layer, pi head with std 0.01, v head).  The product's
    initial parameter vector (AtariCnnPolicy._init_param_values) and the oracle's (oracle/net.py:init_params) draw the same
    numbers."""
    from accel_rl_b200.policies.pg.atari_cnn_policy import AtariCnnPolicy
    from oracle import net as onet
    g = np.load(os.path.join(golden_dir, "norm_c_init.npz"))
    want = np.concatenate([g["hidden"].ravel(), np.zeros(32, np.float32), g["pi"].ravel(), np.zeros(6, np.float32),
                           g["v"].ravel(), np.zeros(1, np.float32)])
    pol = object.__new__(AtariCnnPolicy)
    pol._conv_filters, pol._hidden_sizes = [], [32]
    pol._shapes = [(96, 32), (32,), (32, 6), (6,), (32, 1), (1,)]
    np.random.seed(31)
    got = pol._init_param_values()
    assert got.dtype == np.float32 and np.array_equal(got, want)
    # the oracle's initialiser on a geometry with the same dense shapes: one 1x1 conv (6 channels) over a 4x4 image
    spec = dict(conv_filters=[6], conv_filter_sizes=[1], conv_strides=[1], conv_pads=[0], hidden_sizes=[32])
    flat = onet.init_params(spec, (3, 4, 4), 6, np.random.RandomState(0), np.random.RandomState(31))
    n_conv = 6 * 3 + 6
    assert np.array_equal(flat[n_conv:], want)
