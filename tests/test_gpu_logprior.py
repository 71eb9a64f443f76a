"""GPU parity tests for the Grad-TTS log-prior and the fused log-prior + MAS call.

Bar (BASELINE.json north_star): log-prior within 1e-4 RELATIVE of the torch fp32 expression
(reference model/face_tts.py:165-171); the MAS on top of it is bit-exact given the same fp32
value matrix; end-to-end path agreement with (torch fp32 log-prior -> reference MAS) is reported.
"""
import numpy as np
import pytest
import torch

import oracle
import face_gan_tts_b200 as fgt
from face_gan_tts_b200 import synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
REL_TOL = 1e-4      # north_star: "within 1e-4 relative"


def rel_err(a, b):
    return ((a.double() - b.double()).abs() / b.double().abs().clamp_min(1e-30)).max().item()


IMPLS = ["ffma", "auto"]       # "auto" = tcgen05/TMEM kernel where it covers the shape (F in {64,80,96,128}, Tx <= 256), else FFMA


@pytest.mark.parametrize("impl", IMPLS)
def test_log_prior_matches_fixture(logprior_fixture, impl):
    fx = logprior_fixture
    for mk, yk, rk in (("mu_x", "y", "log_prior_ref_fp32"), ("mu_x_f128", "y_f128", "log_prior_ref_fp32_f128")):
        lp = fgt.log_prior(torch.from_numpy(fx[mk]).to(DEV), torch.from_numpy(fx[yk]).to(DEV), impl=impl)
        ref = torch.from_numpy(fx[rk]).to(DEV)
        assert lp.shape == ref.shape and lp.dtype == torch.float32
        assert rel_err(lp, ref) < REL_TOL


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("F", [80, 128])
def test_log_prior_lrs2_shape_vs_torch_fp32_and_fp64(impl, F):
    """configs[1] shape at B=4 (fp64 direct form needs B*F*Tx*Ty doubles)"""
    mu_x, y, t_x, t_y = synthetic.lrs2_batch(B=4, F=F, Tx=190, Ty=1000, seed=1234)
    mu_d, y_d = mu_x.to(DEV), y.to(DEV)
    lp = fgt.log_prior(mu_d, y_d, impl=impl)
    ref32 = oracle.log_prior_reference(mu_d, y_d)             # the reference expression, on the GPU (cuBLAS fp32)
    ref64 = oracle.log_prior_direct(mu_d, y_d)
    assert rel_err(lp, ref32) < REL_TOL
    assert rel_err(lp, ref64) < REL_TOL
    # and our error vs the fp64 truth is of the same order as torch's own
    assert rel_err(lp, ref64) < 10 * max(rel_err(ref32, ref64), 1e-7)


@pytest.mark.parametrize("impl", IMPLS)
def test_log_prior_odd_shapes(impl):
    g = torch.Generator().manual_seed(4)
    for (B, F, Tx, Ty) in [(1, 80, 1, 1), (2, 80, 7, 13), (3, 16, 65, 130), (1, 128, 129, 257), (2, 48, 33, 64)]:
        mu = torch.randn(B, F, Tx, generator=g).to(DEV)
        y = (torch.randn(B, F, Ty, generator=g) * 2 - 5).to(DEV)
        assert rel_err(fgt.log_prior(mu, y, impl=impl), oracle.log_prior_direct(mu, y)) < REL_TOL


def _reference_mas(value_np, t_x_np, t_y_np):
    """int32 paths of the reference MAS: the reference's own compiled core.pyx when it travelled to this box
    (oracle/_ref), else the C restatement that is pinned to it (tests/test_oracle.py)."""
    paths = np.zeros(value_np.shape, np.int32)
    core = oracle.reference_core("asis")
    if core is not None:
        core.maximum_path_c(paths, np.ascontiguousarray(value_np, dtype=np.float32).copy(), t_x_np.astype(np.int32), t_y_np.astype(np.int32))
    else:
        oracle.maximum_path_c(paths, value_np.copy(), t_x_np, t_y_np)
    return paths


def _fused_with_value_dump(mu_d, y_d, t_x, t_y, **kw):
    """The fused call with the tests-only `fused_dump_ptr` hook: also returns the value tiles the in-kernel search
    consumed (NaN where the kernel wrote nothing: beyond the last 32-frame tile of an utterance)."""
    from face_gan_tts_b200 import _lib

    B, _, Tx = mu_d.shape
    dump = torch.full((B, Tx, y_d.shape[2]), float("nan"), device=DEV)
    _lib.set_pointer_option("fused_dump_ptr", dump)
    try:
        res = fgt.log_prior_maximum_path(mu_d, y_d, t_x, t_y, **kw)
        torch.cuda.synchronize()
    finally:
        _lib.set_pointer_option("fused_dump_ptr", None)
    return res, dump


def _agreement(ft, ref_ft):
    valid = ref_ft >= 0
    return float((ft[valid] == ref_ft[valid]).mean())


# (B, F, Tx, Ty): the LRS2 bench shape (a pair of CTAs per utterance at these batch sizes), the reference default
# n_feats = 128 with short texts (one M-tile) and at the LRS2 shape (pair form only), F = 64 / 96, a second CTA with a
# single text row (Tx = 129), a batch too large for the pair form (one CTA with two M-tiles per utterance)
FUSED_SHAPES = [(8, 80, 190, 1000), (32, 80, 190, 1000), (32, 128, 128, 1000), (5, 64, 256, 512), (6, 96, 100, 600),
                (3, 80, 31, 64), (4, 80, 129, 1400), (32, 128, 190, 1000), (5, 96, 200, 600), (3, 80, 256, 1400),
                (80, 80, 190, 1000)]


@pytest.mark.parametrize("B,F,Tx,Ty", FUSED_SHAPES)
def test_fused_kernel_path_is_bit_exact_mas_of_its_own_values(B, F, Tx, Ty):
    """The fused kernel (one CTA per utterance, value tiles handed over in shared memory): (1) the values its search
    consumed are the log-prior within 1e-4 relative of torch fp32; (2) path / durations / frame_token are the BIT-EXACT
    reference MAS of exactly those values; (3) frame-level agreement with torch-fp32 log-prior -> reference MAS."""
    mu_x, y, t_x, t_y = synthetic.lrs2_batch(B=B, F=F, Tx=Tx, Ty=Ty, seed=1234, tx_lo=max(1, Tx // 3), ty_lo=max(Tx, Ty // 3))
    mu_d, y_d = mu_x.to(DEV), y.to(DEV)
    res, dump = _fused_with_value_dump(mu_d, y_d, t_x, t_y, path_dtype=torch.int32, check=True)
    assert not torch.isnan(dump[0, 0, 0]), "the fused kernel did not run (shape fell back to the serial form)"
    ref_lp = oracle.log_prior_reference(mu_d, y_d)
    txn, tyn = t_x.numpy(), t_y.numpy()
    for b in range(B):
        a, c = dump[b, :txn[b], :tyn[b]], ref_lp[b, :txn[b], :tyn[b]]
        assert rel_err(a, c) < REL_TOL
    own = _reference_mas(torch.nan_to_num(dump).cpu().numpy(), txn, tyn)
    np.testing.assert_array_equal(res.path.cpu().numpy(), own)
    dur, ft = oracle.durations_and_frame_token(own)
    np.testing.assert_array_equal(res.durations.cpu().numpy(), dur)
    np.testing.assert_array_equal(res.frame_token.cpu().numpy(), ft)
    ref = _reference_mas(ref_lp.cpu().numpy(), txn, tyn)
    _, ref_ft = oracle.durations_and_frame_token(ref)
    agree = _agreement(ft, ref_ft)
    print(f"\n[fused B={B} F={F} {Tx}x{Ty}] frame-level path agreement with torch-fp32 log-prior -> reference MAS: {agree * 100:.4f}%")
    assert agree > 0.999


@pytest.mark.parametrize("impl,F", [("ffma", 80), ("auto", 128), ("ffma", 128)])
def test_serial_form_bit_exact_on_its_own_value_and_agrees_with_reference_pipeline(impl, F):
    """Shapes / implementations the fused kernel does not take (n_feats = 128 with two M-tiles, the FFMA log-prior):
    log-prior kernel -> [B,Tx,Ty] -> MAS kernel.  Same three checks, at the bench batch size."""
    mu_x, y, t_x, t_y = synthetic.lrs2_batch(B=32, F=F, Tx=190, Ty=1000, seed=1234)
    mu_d, y_d = mu_x.to(DEV), y.to(DEV)
    res = fgt.log_prior_maximum_path(mu_d, y_d, t_x, t_y, path_dtype=torch.int32, impl=impl, check=True)
    lp = fgt.log_prior(mu_d, y_d, impl=impl)
    own = _reference_mas(lp.cpu().numpy(), t_x.numpy(), t_y.numpy())
    np.testing.assert_array_equal(res.path.cpu().numpy(), own)
    dur, ft = oracle.durations_and_frame_token(own)
    np.testing.assert_array_equal(res.durations.cpu().numpy(), dur)
    np.testing.assert_array_equal(res.frame_token.cpu().numpy(), ft)
    ref = _reference_mas(oracle.log_prior_reference(mu_d, y_d).cpu().numpy(), t_x.numpy(), t_y.numpy())
    _, ref_ft = oracle.durations_and_frame_token(ref)
    agree = _agreement(ft, ref_ft)
    print(f"\n[{impl} F={F}] frame-level path agreement with torch-fp32 log-prior -> reference MAS: {agree * 100:.4f}%")
    assert agree > 0.999


def test_prepared_workspace_flag_is_still_accepted():
    """mas_b200_fused_workspace_prepare / MAS_B200_WS_PREPARED were the protocol of the removed two-kernel pipeline; they
    stay in the ABI as no-ops: garbage in the workspace, prepared or not, dense path or not -- same results."""
    from face_gan_tts_b200 import _lib

    L = _lib.lib()
    B, F, Tx, Ty = 32, 80, 190, 1000
    mu_x, y, t_x, t_y = synthetic.lrs2_batch(B=B, F=F, Tx=Tx, Ty=Ty, seed=300)
    mu, yy, tx, ty = mu_x.to(DEV), y.to(DEV), t_x.to(DEV), t_y.to(DEV)
    want = fgt.log_prior_maximum_path(mu, yy, tx, ty, path_dtype=torch.float32)
    ws_bytes = L.mas_b200_fused_workspace_bytes(B, F, Tx, Ty)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=DEV)
    ws.fill_(0xAB)
    assert L.mas_b200_fused_workspace_prepare(ws.data_ptr(), ws_bytes, B, F, Tx, Ty, None) == 0
    path = torch.empty((B, Tx, Ty), device=DEV)
    dur = torch.empty((B, Tx), dtype=torch.int32, device=DEV)
    ft = torch.empty((B, Ty), dtype=torch.int32, device=DEV)
    st = torch.empty((B,), dtype=torch.int32, device=DEV)
    sp = torch.cuda.current_stream().cuda_stream
    for i in range(8):
        dense = (i % 4) != 3
        impl = _lib.LP_AUTO | (_lib.WS_PREPARED if i % 2 else 0)
        rc = L.mas_b200_log_prior_maximum_path(mu.data_ptr(), yy.data_ptr(), tx.data_ptr(), ty.data_ptr(), B, F, Tx, Ty, -1e9,
                                               path.data_ptr() if dense else None, _lib.PATH_F32 if dense else _lib.PATH_NONE,
                                               dur.data_ptr(), ft.data_ptr(), st.data_ptr(), ws.data_ptr(), ws_bytes, impl, sp)
        assert rc == 0
        assert torch.equal(dur, want.durations) and torch.equal(ft, want.frame_token)
        if dense:
            assert torch.equal(path, want.path)


def test_alignment_plan_equals_functional_api_and_reuses_buffers():
    mu_x, y, t_x, t_y = synthetic.lrs2_batch(B=6, F=80, Tx=190, Ty=1000, seed=17)
    mu2, y2, t_x2, t_y2 = synthetic.lrs2_batch(B=6, F=80, Tx=190, Ty=1000, seed=18)
    plan = fgt.AlignmentPlan(6, 80, 190, 1000, device=DEV)
    for (m, yy, a, b) in ((mu_x, y, t_x, t_y), (mu2, y2, t_x2, t_y2), (mu_x, y, t_x, t_y)):
        md, yd, ad, bd = m.to(DEV), yy.to(DEV), a.to(DEV), b.to(DEV)
        want = fgt.log_prior_maximum_path(md, yd, ad, bd)
        got = plan(md, yd, ad, bd, check=True)
        assert got.path.data_ptr() == plan.path.data_ptr()
        assert torch.equal(got.path, want.path) and torch.equal(got.durations, want.durations)
        assert torch.equal(got.frame_token, want.frame_token)
    with pytest.raises(ValueError):
        plan(mu_x.to(DEV)[:, :, :100], y.to(DEV), t_x.to(DEV), t_y.to(DEV))
    with pytest.raises(ValueError):
        plan(mu_x.to(DEV), y.to(DEV), t_x.long().to(DEV), t_y.to(DEV))


def test_fused_call_without_dense_path():
    mu_x, y, t_x, t_y = synthetic.lrs2_batch(B=4, F=80, Tx=190, Ty=1000, seed=99)
    a = fgt.log_prior_maximum_path(mu_x.to(DEV), y.to(DEV), t_x, t_y, dense_path=False)
    b = fgt.log_prior_maximum_path(mu_x.to(DEV), y.to(DEV), t_x, t_y, dense_path=True)
    assert a.path is None
    assert torch.equal(a.durations, b.durations) and torch.equal(a.frame_token, b.frame_token)
    assert torch.equal(b.path.sum(-1).int(), b.durations)


def test_compute_loss_call_site_quantities():
    """The consumers of the path in reference FaceTTS.compute_loss (face_tts.py:176-234), restated:
    logw_/duration loss from durations, mu_y as a gather by frame_token == attn^T @ mu_x^T, prior loss."""
    import math

    mu_x, y, t_x, t_y = synthetic.lrs2_batch(B=4, F=80, Tx=61, Ty=200, seed=5, tx_lo=21, ty_lo=90)
    mu_d, y_d = mu_x.to(DEV), y.to(DEV)
    B, F, Tx = mu_x.shape
    Ty = y.shape[2]
    x_mask = (torch.arange(Tx)[None, None, :] < t_x[:, None, None]).float().to(DEV)
    y_mask = (torch.arange(Ty)[None, None, :] < t_y[:, None, None]).float().to(DEV)
    attn_mask = x_mask.unsqueeze(-1) * y_mask.unsqueeze(2)
    # reference pipeline with the library's drop-in maximum_path (face_tts.py:165-174)
    log_prior = oracle.log_prior_reference(mu_d, y_d)
    attn = fgt.monotonic_align.maximum_path(log_prior, attn_mask.squeeze(1)).detach()
    # oracle on the very same log_prior
    ref = np.zeros(attn.shape, np.int32)
    oracle.maximum_path_c(ref, log_prior.cpu().numpy().copy(), t_x.numpy(), t_y.numpy())
    assert torch.equal(attn.cpu(), torch.from_numpy(ref).float())
    # :176-179 duration target and loss
    logw = torch.randn(B, 1, Tx, device=DEV) * x_mask
    logw_ = torch.log(1e-8 + torch.sum(attn.unsqueeze(1), -1)) * x_mask
    res = fgt.align(log_prior, t_x, t_y)
    assert torch.equal(fgt.durations_to_logw(res.durations, x_mask), logw_)
    dur_loss = torch.sum((logw - logw_) ** 2) / torch.sum(t_x.to(DEV))
    assert torch.isfinite(dur_loss)
    # :217-218 mu_y by GEMM == gather by frame_token
    mu_y = torch.matmul(attn.transpose(1, 2), mu_d.transpose(1, 2)).transpose(1, 2)
    idx = res.frame_token.clamp_min(0).long()
    gathered = torch.gather(mu_d, 2, idx[:, None, :].expand(B, F, Ty)) * y_mask
    assert torch.equal(mu_y, gathered)
    # :233-234 prior loss
    prior = torch.sum(0.5 * ((y_d - mu_y) ** 2 + math.log(2 * math.pi)) * y_mask) / (torch.sum(y_mask) * F)
    assert torch.isfinite(prior)


def test_tcgen05_kernel_is_the_one_running_and_matches_ffma():
    """impl="tcgen05" must not silently fall back: supported shapes run the tensor-core kernel (3xTF32, fp32-class
    accuracy), unsupported ones raise.  Against the FFMA kernel the two agree far inside the 1e-4 bar."""
    from face_gan_tts_b200 import _lib

    # F = 128 (the reference default n_feats) runs the split-M form: one CTA per 128-row M-tile
    for (B, F, Tx, Ty) in [(3, 80, 190, 1000), (2, 64, 129, 136), (2, 96, 256, 420), (5, 80, 37, 68),
                           (3, 128, 190, 1000), (2, 128, 256, 420), (4, 128, 61, 200), (2, 128, 128, 132), (1, 128, 129, 132),
                           # texts longer than two M-tiles: split-M for every n_feats (cfg4: Tx = 512)
                           (2, 80, 512, 640), (1, 96, 300, 304), (2, 64, 257, 260), (1, 128, 513, 516), (1, 80, 1000, 1000),
                           (2, 80, 1, 4), (1, 64, 5, 8), (3, 96, 128, 192)]:
        mu_x, y, _, _ = synthetic.lrs2_batch(B=B, F=F, Tx=Tx, Ty=Ty, seed=7, tx_lo=max(1, Tx // 3), ty_lo=max(Tx // 3, Ty // 3))
        mu_d, y_d = mu_x.to(DEV), y.to(DEV)
        tc = fgt.log_prior(mu_d, y_d, impl="tcgen05")
        ff = fgt.log_prior(mu_d, y_d, impl="ffma")
        ref64 = oracle.log_prior_direct(mu_d, y_d)
        assert rel_err(tc, ref64) < 2e-6 and rel_err(tc, ff) < 2e-6
    mu_x, y, _, _ = synthetic.lrs2_batch(B=2, F=72, Tx=190, Ty=1000, seed=7)
    with pytest.raises(_lib.MasB200Error):
        fgt.log_prior(mu_x.to(DEV), y.to(DEV), impl="tcgen05")          # n_feats not instantiated: raises, no fallback


@pytest.mark.parametrize("B,F,Tx", [(3, 80, 190), (32, 80, 190), (80, 80, 190), (300, 80, 190), (32, 128, 128), (3, 128, 190), (50, 128, 190), (90, 128, 190)])
def test_fused_kernel_vs_serial_form(B, F, Tx):
    """mas_b200_log_prior_maximum_path: the fused kernel (default where the shape is covered) and the serial form
    (fused_impl=1: log-prior kernel -> HBM -> MAS kernel -> path expansion) are two schedules of the same computation.
    Their log-priors differ in the last bits (the fused kernel folds the y^2 / mu^2 terms into the contraction), so the
    paths may differ at near-ties: status and structure must be identical, frames must agree > 99.9 %.  Shapes the fused
    kernel does not cover (F = 128 with two M-tiles in a batch too large for the pair form) take the serial form in both
    modes: identical outputs."""
    from face_gan_tts_b200 import _lib

    mu_x, y, t_x, t_y = synthetic.lrs2_batch(B=B, F=F, Tx=Tx, Ty=1000, seed=11)
    mu_d, y_d = mu_x.to(DEV), y.to(DEV)
    outs = []
    for mode in (0, 1):
        prev = _lib.set_option("fused_impl", mode)
        try:
            r = None
            for _ in range(4):                      # back to back, no host sync in between
                r = fgt.log_prior_maximum_path(mu_d, y_d, t_x, t_y, path_dtype=torch.float32)
            torch.cuda.synchronize()
            outs.append(r)
        finally:
            _lib.set_option("fused_impl", prev)
    a, b = outs
    assert torch.equal(a.status, b.status) and int(a.status.abs().sum()) == 0
    assert torch.equal(a.path.sum(-1).int(), a.durations) and torch.equal(b.path.sum(-1).int(), b.durations)
    assert torch.equal(a.durations.sum(-1), t_y.to(DEV)) and torch.equal(b.durations.sum(-1), t_y.to(DEV))
    valid = b.frame_token >= 0
    assert torch.equal(valid, a.frame_token >= 0)
    agree = float((a.frame_token[valid] == b.frame_token[valid]).float().mean())
    assert agree > 0.999
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    if F == 128 and Tx > 128 and 2 * B > sms:        # no pair form, no one-CTA form: the serial form in both modes
        assert torch.equal(a.path, b.path) and torch.equal(a.frame_token, b.frame_token)


@pytest.mark.parametrize("B,F,Tx,Ty", [(32, 80, 190, 1000), (7, 64, 256, 800), (100, 80, 190, 600)])
def test_pair_form_equals_one_cta_form(B, F, Tx, Ty):
    """A text of 129..256 tokens as a 2-CTA cluster (one M-tile and one DP warp per CTA, halo row and direction words
    crossing with st.async) and as one CTA with two M-tiles: the same values, the same search -- identical outputs, bit for
    bit.  B = 100 forces the pair form on a batch that does not fit the SMs at once (clusters run in waves)."""
    from face_gan_tts_b200 import _lib

    mu_x, y, t_x, t_y = synthetic.lrs2_batch(B=B, F=F, Tx=Tx, Ty=Ty, seed=5, tx_lo=20, ty_lo=max(Tx, Ty // 4))
    mu_d, y_d = mu_x.to(DEV), y.to(DEV)
    outs = []
    for mode in (0, 2):
        prev = _lib.set_option("fused_pair", mode)
        try:
            r, dump = _fused_with_value_dump(mu_d, y_d, t_x, t_y, path_dtype=torch.float32)
            outs.append((r, dump))
        finally:
            _lib.set_option("fused_pair", prev)
    (a, da), (b, db) = outs
    assert not torch.isnan(da[0, 0, 0]) and not torch.isnan(db[0, 0, 0])
    assert torch.equal(torch.nan_to_num(da), torch.nan_to_num(db)), "the two forms computed different log-prior values"
    assert int(a.status.abs().sum()) == 0 and torch.equal(a.status, b.status)
    assert torch.equal(a.path, b.path) and torch.equal(a.durations, b.durations) and torch.equal(a.frame_token, b.frame_token)


def test_long_text_takes_the_serial_form():
    """Tx = 300 (three M-tiles: split-M log-prior kernel + MAS kernel): fused_impl = 0 and 1 are the same code path."""
    from face_gan_tts_b200 import _lib

    mu_x, y, t_x, t_y = synthetic.lrs2_batch(B=8, F=80, Tx=300, Ty=1000, seed=21, tx_lo=100, ty_lo=400)
    mu_d, y_d = mu_x.to(DEV), y.to(DEV)
    outs = []
    for mode in (0, 1):
        prev = _lib.set_option("fused_impl", mode)
        try:
            for _ in range(3):
                r = fgt.log_prior_maximum_path(mu_d, y_d, t_x, t_y, path_dtype=torch.float32)
            torch.cuda.synchronize()
            outs.append(r)
        finally:
            _lib.set_option("fused_impl", prev)
    a, b = outs
    assert torch.equal(a.path, b.path) and torch.equal(a.durations, b.durations) and torch.equal(a.frame_token, b.frame_token)
    assert int(a.status.abs().sum()) == 0 and torch.equal(a.path.sum(-1).int(), a.durations)


@pytest.mark.parametrize("F,Tx", [(80, 190), (128, 128)])
def test_fused_kernel_soak(F, Tx):
    """300 back-to-back fused calls over rotating inputs, no host sync in between (programmatic dependent launch lets
    call i + 1 be placed while call i is still running): every result must equal the first, checked result of its input."""
    B, nset = 32, 3
    sets, want = [], []
    for k in range(nset):
        mu_x, y, t_x, t_y = synthetic.lrs2_batch(B=B, F=F, Tx=Tx, Ty=1000, seed=100 + k)
        s_ = (mu_x.to(DEV), y.to(DEV), t_x.to(DEV), t_y.to(DEV))
        res, dump = _fused_with_value_dump(*s_, path_dtype=torch.float32)
        own = _reference_mas(torch.nan_to_num(dump).cpu().numpy(), t_x.numpy(), t_y.numpy())
        np.testing.assert_array_equal(res.path.cpu().numpy().astype(np.int32), own)
        sets.append(s_)
        want.append((res.durations.clone(), res.frame_token.clone(), res.path.clone()))
    bad = torch.zeros((), dtype=torch.int64, device=DEV)
    for i in range(300):
        r = fgt.log_prior_maximum_path(*sets[i % nset], path_dtype=torch.float32)
        w = want[i % nset]
        bad += (r.durations != w[0]).sum() + (r.frame_token != w[1]).sum() + (r.path != w[2]).sum()
    torch.cuda.synchronize()
    assert int(bad) == 0


@pytest.mark.parametrize("n_sms", [1, 32, 116, 148])
def test_fused_call_completes_beside_a_foreign_kernel_holding_sms(n_sms):
    """The alignment is ONE kernel whose CTAs wait only for each other's warps (never for another launch), so a foreign
    kernel that keeps SMs busy on another stream (NCCL, the decoder of a real training step) can delay it but not
    starve it: with 1 / 32 / 116 / all 148 SMs held for ~2 ms the call still completes with the right answer."""
    from face_gan_tts_b200 import _lib

    mu_x, y, t_x, t_y = synthetic.lrs2_batch(B=32, F=80, Tx=190, Ty=1000, seed=9)
    args = (mu_x.to(DEV), y.to(DEV), t_x.to(DEV), t_y.to(DEV))
    want = fgt.log_prior_maximum_path(*args, path_dtype=torch.float32)
    torch.cuda.synchronize()
    side = torch.cuda.Stream(DEV)
    with torch.cuda.stream(side):
        _lib.check(_lib.lib().mas_b200_debug_occupy_sms(n_sms, 4_000_000, side.cuda_stream), "mas_b200_debug_occupy_sms")
    for _ in range(3):
        got = fgt.log_prior_maximum_path(*args, path_dtype=torch.float32)
    torch.cuda.synchronize()
    assert torch.equal(got.path, want.path) and torch.equal(got.durations, want.durations)
    assert torch.equal(got.frame_token, want.frame_token) and int(got.status.abs().sum()) == 0


def test_fused_call_rejects_bad_items():
    """t_x > t_y is undefined in the reference (core.pyx:34); here the item is rejected and its outputs are zero."""
    mu_x, y, t_x, t_y = synthetic.lrs2_batch(B=4, F=80, Tx=61, Ty=200, seed=5, tx_lo=21, ty_lo=90)
    t_x = t_x.clone(); t_y = t_y.clone()
    t_x[1] = 61; t_y[1] = 40                       # t_x > t_y
    res = fgt.log_prior_maximum_path(mu_x.to(DEV), y.to(DEV), t_x, t_y, path_dtype=torch.int32)
    torch.cuda.synchronize()
    assert res.status.tolist() == [0, 1, 0, 0]
    assert int(res.path[1].sum()) == 0 and int(res.durations[1].sum()) == 0
    assert torch.equal(res.path[0].sum(-1).int(), res.durations[0])


# ---------------------------------------------------------------------------------------------
# ragged zero-copy upload (mas_b200_upload_batch): the e2e path's host -> device transfer
# ---------------------------------------------------------------------------------------------
@pytest.fixture(params=[1, 2, 3], ids=["sm_zero_copy_pull", "copy_engine_2d", "tma_bulk"])
def upload_impl(request):
    prev = fgt._lib.set_option("upload_impl", request.param)
    yield request.param
    fgt._lib.set_option("upload_impl", prev)


@pytest.mark.parametrize("B,F,Tx,Ty", [(32, 80, 190, 1000), (5, 128, 64, 256), (3, 7, 33, 101), (2, 80, 1, 4)])
def test_upload_batch_equals_a_plain_copy_of_padded_inputs(B, F, Tx, Ty, upload_impl):
    mu_x, y, t_x, t_y = synthetic.lrs2_batch(B, F, Tx, Ty, seed=77, tx_lo=1, ty_lo=max(1, Ty // 3))
    # garbage in the padding of the HOST buffers: the device tensors must still be zero-padded
    g = torch.Generator().manual_seed(1)
    junk_x = torch.randn(B, F, Tx, generator=g) * (torch.arange(Tx)[None, None] >= t_x[:, None, None])
    junk_y = torch.randn(B, F, Ty, generator=g) * (torch.arange(Ty)[None, None] >= t_y[:, None, None])
    host = [(mu_x + junk_x).pin_memory(), (y + junk_y).pin_memory(), t_x.pin_memory(), t_y.pin_memory()]
    out = fgt.upload_batch(*host, mu_on_copy_engine=False)
    torch.cuda.synchronize()
    assert torch.equal(out[0].cpu(), mu_x) and torch.equal(out[1].cpu(), y)
    assert torch.equal(out[2].cpu(), t_x) and torch.equal(out[3].cpu(), t_y)
    # reuse of the output buffers (mu_x through the copy engine: a plain copy, host padding included) + the
    # alignment computed from them equals the one from plain copies
    out2 = fgt.upload_batch(*host, out=out, mu_on_copy_engine=True)
    torch.cuda.synchronize()
    assert torch.equal(out2[1].cpu(), y) and torch.equal(out2[0].cpu(), host[0])
    a = fgt.log_prior_maximum_path(out2[0], out2[1], out2[2], out2[3], dense_path=False)
    b = fgt.log_prior_maximum_path(mu_x.to(DEV), y.to(DEV), t_x, t_y, dense_path=False)
    assert torch.equal(a.durations, b.durations) and torch.equal(a.frame_token, b.frame_token)


@pytest.mark.parametrize("B,F,Tx,Ty", [(32, 80, 190, 1000), (5, 128, 64, 256), (3, 7, 33, 101), (2, 80, 1, 4), (300, 8, 21, 40)])
def test_packed_batch_unpacks_to_the_padded_tensors(B, F, Tx, Ty):
    """pack_batch (host collate without padding) -> ONE H2D copy -> mas_b200_unpack_batch: the device tensors are the
    zero-padded tensors of the reference contract, whatever garbage the padded host tensors held beyond the lengths."""
    mu_x, y, t_x, t_y = synthetic.lrs2_batch(B, F, Tx, Ty, seed=78, tx_lo=1, ty_lo=max(1, Ty // 3))
    g = torch.Generator().manual_seed(1)
    junk_x = torch.randn(B, F, Tx, generator=g) * (torch.arange(Tx)[None, None] >= t_x[:, None, None])
    junk_y = torch.randn(B, F, Ty, generator=g) * (torch.arange(Ty)[None, None] >= t_y[:, None, None])
    packed = fgt.pack_batch(mu_x + junk_x, y + junk_y, t_x, t_y)
    assert packed.is_pinned() and packed.numel() == 16 * ((8 * B + 15) // 16) + 4 * F * int(t_x.sum() + t_y.sum())
    out = fgt.upload_packed_batch(packed, B, F, Tx, Ty)
    torch.cuda.synchronize()
    assert torch.equal(out[0].cpu(), mu_x) and torch.equal(out[1].cpu(), y)
    assert torch.equal(out[2].cpu(), t_x) and torch.equal(out[3].cpu(), t_y)
    # buffers reused, alignment from them == alignment from plain copies
    staging = torch.empty((packed.numel() + 64,), dtype=torch.uint8, device=DEV)
    out2 = fgt.upload_packed_batch(packed, B, F, Tx, Ty, out=out, staging=staging)
    torch.cuda.synchronize()
    assert torch.equal(out2[1].cpu(), y)
    if F in (80, 128):
        a = fgt.log_prior_maximum_path(out2[0], out2[1], out2[2], out2[3], dense_path=False)
        b = fgt.log_prior_maximum_path(mu_x.to(DEV), y.to(DEV), t_x, t_y, dense_path=False)
        assert torch.equal(a.durations, b.durations) and torch.equal(a.frame_token, b.frame_token)


def test_upload_batch_rejects_pageable_memory():
    mu_x, y, t_x, t_y = synthetic.lrs2_batch(2, 80, 21, 64, seed=3, tx_lo=5, ty_lo=30)
    with pytest.raises(ValueError):
        fgt.upload_batch(mu_x, y, t_x, t_y)
    L = fgt._lib.lib()
    d = [t.to(DEV) for t in (mu_x, y, t_x, t_y)]
    rc = L.mas_b200_upload_batch(mu_x.data_ptr(), y.data_ptr(), t_x.data_ptr(), t_y.data_ptr(), 2, 80, 21, 64,
                                 d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), None)
    assert rc == fgt._lib.ERR_ARG


@pytest.mark.parametrize("B,Tx", [(6, 190), (5, 100)])
def test_one_sided_duration_gather_writes_every_peer_buffer(B, Tx):
    """Multi-GPU duration gather without a collective: with peer_world > 0 the fused kernel also stores its durations
    into rows [peer_rank * B, +B) of every buffer in `peer_dur_ptrs` (on a multi-GPU box: every rank's symmetric-memory
    gather buffer, face_gan_tts_b200.sharding.OneSidedDurationGather).  Here: three local buffers standing in for three
    ranks' mappings, this process as rank 1 -- pair form (Tx = 190) and one-CTA form, incl. an item the kernel rejects."""
    from face_gan_tts_b200 import _lib

    world, rank = 3, 1
    mu_x, y, t_x, t_y = synthetic.lrs2_batch(B=B, F=80, Tx=Tx, Ty=600, seed=9, tx_lo=20, ty_lo=Tx)
    t_x[2] = int(t_y[2]) + 1 if int(t_y[2]) < Tx else t_x[2]          # may be a bad item (t_x > t_y) -> zeros
    bufs = [torch.full((world * B, Tx), -7, dtype=torch.int32, device=DEV) for _ in range(world)]
    ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=DEV)
    _lib.set_pointer_option("peer_dur_ptrs", ptrs)
    _lib.set_option("peer_rank", rank)
    _lib.set_option("peer_world", world)
    try:
        r = fgt.log_prior_maximum_path(mu_x.to(DEV), y.to(DEV), t_x, t_y, dense_path=False)
        torch.cuda.synchronize()
    finally:
        _lib.set_option("peer_world", 0)
        _lib.set_pointer_option("peer_dur_ptrs", None)
    for b in bufs:
        assert torch.equal(b[rank * B:(rank + 1) * B], r.durations)
        assert bool((b[:rank * B] == -7).all()) and bool((b[(rank + 1) * B:] == -7).all())
    # the same gather as a separate, stream-ordered put (mas_b200_put_durations)
    bufs2 = [torch.full((world * B, Tx), -7, dtype=torch.int32, device=DEV) for _ in range(world)]
    ptrs2 = torch.tensor([b.data_ptr() for b in bufs2], dtype=torch.int64, device=DEV)
    _lib.check(_lib.lib().mas_b200_put_durations(r.durations.data_ptr(), B, Tx, ptrs2.data_ptr(), world, rank,
                                                 torch.cuda.current_stream().cuda_stream), "mas_b200_put_durations")
    torch.cuda.synchronize()
    for b, b2 in zip(bufs, bufs2):
        assert torch.equal(b, b2)


def test_pair_form_fuzz_equals_one_cta_form():
    """Randomised shapes on the M-tile / tile edges (Tx in {129, 130, 191..193, 255, 256}, Ty up to 1404, B up to 74, items
    with t_x = 128 / 129 / Tx): the 2-CTA-cluster form and the one-CTA form must agree bit for bit."""
    import random
    from face_gan_tts_b200 import _lib

    rng = random.Random(7)
    for it in range(24):
        B = rng.choice([1, 2, 3, 5, 9, 33, 74])
        F = rng.choice([64, 80])
        Tx = rng.choice([129, 130, 160, 191, 192, 193, 224, 255, 256])
        Ty = max(rng.choice([260, 288, 320, 516, 1000, 1028, 1404]), ((Tx + 3) // 4) * 4)
        mu_x, y, t_x, t_y = synthetic.lrs2_batch(B=B, F=F, Tx=Tx, Ty=Ty, seed=100 + it, tx_lo=1, ty_lo=min(Ty, max(4, Tx // 2)))
        t_x[0], t_y[0] = Tx, Ty
        if B > 1:
            t_x[1], t_y[1] = 129, max(int(t_y[1]), 129)
        if B > 2:
            t_x[2], t_y[2] = 128, max(int(t_y[2]), 128)
        outs = []
        for mode in (0, 2):
            prev = _lib.set_option("fused_pair", mode)
            try:
                outs.append(fgt.log_prior_maximum_path(mu_x.to(DEV), y.to(DEV), t_x, t_y, path_dtype=torch.int32))
                torch.cuda.synchronize()
            finally:
                _lib.set_option("fused_pair", prev)
        a, b = outs
        assert int(a.status.abs().sum()) == 0 and torch.equal(a.status, b.status), (B, F, Tx, Ty)
        assert torch.equal(a.durations, b.durations) and torch.equal(a.frame_token, b.frame_token) and torch.equal(a.path, b.path), (B, F, Tx, Ty)
