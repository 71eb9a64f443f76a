"""The oracle is only trusted once it is pinned: every restatement in oracle/
is checked here against the fixtures the REFERENCE's compiled core.pyx
produced (tests/golden/mas_kats.npz), and -- when oracle/_ref is present --
against the compiled reference itself on fresh random inputs."""
import hashlib

import numpy as np
import pytest
import torch

import cases
import oracle


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def expected(kats, name):
    return {k: kats[f"{name}/{k}"] for k in ("t_x", "t_y", "dur", "frame_token", "path_sha256", "input_sha256")}


@pytest.mark.parametrize("name", list(cases.CASES))
def test_c_oracle_matches_reference_golden(kats, name):
    value, t_x, t_y = cases.CASES[name]()
    exp = expected(kats, name)
    assert sha(np.ascontiguousarray(value, np.float32)) == str(exp["input_sha256"]), "input generator drifted"
    paths = np.zeros(value.shape, np.int32)
    bad = oracle.maximum_path_c(paths, np.ascontiguousarray(value, np.float32).copy(), t_x, t_y)
    assert bad == 0
    assert sha(paths) == str(exp["path_sha256"])
    dur, ft = oracle.durations_and_frame_token(paths)
    np.testing.assert_array_equal(dur, exp["dur"])
    np.testing.assert_array_equal(ft, exp["frame_token"])


SMALL = [n for n in cases.CASES if n not in ("cfg1", "large_accum", "lrs2_shape", "wide_513x1030")]


@pytest.mark.parametrize("name", SMALL)
def test_numpy_bit_formulation_matches_golden(kats, name):
    """rolling column + 1 direction bit per cell + no lower band bound == reference."""
    value, t_x, t_y = cases.CASES[name]()
    paths = oracle.maximum_path_numpy(value, t_x, t_y)
    assert sha(paths) == str(kats[f"{name}/path_sha256"])


def test_all_zero_tie_rule(kats):
    np.testing.assert_array_equal(kats["all_zero/dur"][0], [1, 1, 1, 5])


def test_path_structure_properties(kats):
    """every valid frame has exactly one token, tokens monotone, every token >= 1 frame."""
    for name in cases.CASES:
        t_x, t_y = kats[f"{name}/t_x"], kats[f"{name}/t_y"]
        dur, ft = kats[f"{name}/dur"], kats[f"{name}/frame_token"]
        for b in range(len(t_x)):
            f = ft[b, : t_y[b]]
            assert f[0] == 0 and f[-1] == t_x[b] - 1
            d = np.diff(f)
            assert ((d == 0) | (d == 1)).all()
            assert (ft[b, t_y[b]:] == -1).all()
            assert (dur[b, : t_x[b]] >= 1).all() and dur[b].sum() == t_y[b]
            assert (dur[b, t_x[b]:] == 0).all()


def test_rejects_tx_gt_ty():
    v = np.zeros((1, 5, 3), np.float32)
    p = np.zeros((1, 5, 3), np.int32)
    assert oracle.maximum_path_c(p, v, np.asarray([5], np.int32), np.asarray([3], np.int32)) == 1
    assert p.sum() == 0


def test_wrapper_restatement_matches_golden(kats):
    """oracle.maximum_path == reference monotonic_align/__init__.py:8-23 semantics."""
    from face_gan_tts_b200 import synthetic

    value, t_x, t_y = cases.CASES["rand_small"]()
    B, Tx, Ty = value.shape
    mask = synthetic.prefix_mask(torch.from_numpy(t_x), torch.from_numpy(t_y), Tx, Ty)
    path = oracle.maximum_path(torch.from_numpy(value), mask)
    assert path.dtype == torch.float32 and path.shape == (B, Tx, Ty)
    assert sha(path.numpy().astype(np.int32)) == str(kats["rand_small/path_sha256"])


def test_against_compiled_reference_random():
    core = oracle.reference_core("asis")
    if core is None:
        pytest.skip("oracle/_ref not built (no /root/reference and no prebuilt .so)")
    rng = np.random.default_rng(5)
    for trial in range(40):
        B = int(rng.integers(1, 4))
        Tx = int(rng.integers(1, 40))
        Ty = int(rng.integers(Tx, 90))
        kind = trial % 4
        if kind == 0:
            v = rng.standard_normal((B, Tx, Ty)).astype(np.float32)
        elif kind == 1:
            v = rng.integers(-1, 2, (B, Tx, Ty)).astype(np.float32)
        elif kind == 2:
            v = (rng.standard_normal((B, Tx, Ty)) * 3e8 - 4e8).astype(np.float32)
        else:
            v = np.zeros((B, Tx, Ty), np.float32)
        t_x = rng.integers(1, Tx + 1, B).astype(np.int32)
        t_y = np.asarray([rng.integers(t_x[b], Ty + 1) for b in range(B)], np.int32)
        p_ref = np.zeros(v.shape, np.int32)
        core.maximum_path_c(p_ref, v.copy(), t_x, t_y)
        p_c = np.zeros(v.shape, np.int32)
        oracle.maximum_path_c(p_c, v.copy(), t_x, t_y)
        np.testing.assert_array_equal(p_c, p_ref)
        np.testing.assert_array_equal(oracle.maximum_path_numpy(v, t_x, t_y), p_ref)


def test_log_prior_reference_matches_fixture(logprior_fixture):
    fx = logprior_fixture
    lp = oracle.log_prior_reference(torch.from_numpy(fx["mu_x"]), torch.from_numpy(fx["y"])).numpy()
    # torch CPU SGEMM summation order may differ between builds: tolerance, not bits
    np.testing.assert_allclose(lp, fx["log_prior_ref_fp32"], rtol=2e-6)
    np.testing.assert_allclose(lp, fx["log_prior_direct_fp64"], rtol=1e-5)
