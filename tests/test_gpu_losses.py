"""GPU parity tests for the consumers of the alignment in FaceTTS.compute_loss (SURVEY section 8 rows a1, a6-a8,
f1, f2, f4): sequence_mask, crop, mu_y gather, prior loss, duration loss -- on the index form of the path --
against the dense torch restatement of reference model/face_tts.py:161-218,233-234 (oracle.compute_loss_block,
pinned to the reference's real compute_loss by tests/golden/compute_loss_block.npz).

Bars: masks / crops / gathered mu_y are copies -> BIT-EXACT.  The two scalar losses and the gradients are fp32
reductions in a different order than torch's -> 2e-6 relative (losses), 1e-5 relative to the largest
gradient magnitude (gradients); written next to each assert.
"""
import random

import pytest
import torch

import oracle
import face_gan_tts_b200 as fgt
from face_gan_tts_b200 import losses, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
LOSS_RTOL = 2e-6
GRAD_RTOL = 1e-5


def _batch(B, F, Tx, Ty, seed, tx_lo, ty_lo):
    mu_x, y, t_x, t_y = synthetic.lrs2_batch(B, F, Tx, Ty, seed=seed, tx_lo=tx_lo, ty_lo=ty_lo)
    g = torch.Generator().manual_seed(seed + 7)
    x_mask = (torch.arange(Tx)[None, :] < t_x[:, None]).float().unsqueeze(1)
    logw = torch.randn(B, 1, Tx, generator=g) * x_mask
    return mu_x, logw, x_mask, y, t_x, t_y


def _oracle_block(mu_x, logw, x_mask, y, t_x, t_y, out_size, out_offset, attn=None):
    """dense reference block on the CPU (autograd through torch).  MAS by the pinned C oracle on the torch
    log-prior, or -- `attn` given -- the dense path the GPU produced, so that the CONSUMERS are compared on the
    same alignment (bit-exactness of the path itself is test_gpu_mas / test_gpu_logprior's job; a 1e-7 log-prior
    difference may legitimately move a near-tie)."""
    mu = mu_x.clone().requires_grad_(True)
    lw = logw.clone().requires_grad_(True)
    fn = None if attn is None else (lambda value, mask: attn)
    r = oracle.compute_loss_block(mu, lw, x_mask, y, t_y.long(), t_x.long(), mu_x.shape[1], out_size=out_size,
                                  out_offset=None if out_offset is None else torch.as_tensor(out_offset).long(),
                                  maximum_path_fn=fn)
    return r, mu, lw


def _gpu_path(mu_x, y, t_x, t_y):
    res = fgt.log_prior_maximum_path(mu_x.to(DEV), y.to(DEV), t_x, t_y, dense_path=True)
    return res, res.path.cpu()


def _rel(a, b):
    return abs(float(a.detach() if hasattr(a, "detach") else a) - float(b)) / max(abs(float(b)), 1e-30)


def test_sequence_mask_bit_exact():
    ln = torch.tensor([0, 1, 5, 255, 256, 257, 1000], dtype=torch.int64)
    for T in (1, 7, 256, 257, 1000):
        got = losses.sequence_mask(ln.to(DEV), T)
        ref = oracle.sequence_mask(ln, T).float()
        assert torch.equal(got.cpu(), ref)
    got = losses.sequence_mask(ln.to(DEV))                    # max_length=None -> lengths.max()
    assert got.shape == (7, 1000) and torch.equal(got.cpu(), oracle.sequence_mask(ln).float())


@pytest.mark.parametrize("B,F,Tx,Ty,out_size", [(5, 80, 61, 200, 128), (4, 128, 33, 160, 128), (3, 13, 17, 700, 64)])
def test_crop_frames_matches_reference_loop(B, F, Tx, Ty, out_size):
    mu_x, logw, x_mask, y, t_x, t_y = _batch(B, F, Tx, Ty, 11, max(3, Tx // 3), max(Tx, Ty // 3))
    random.seed(5)
    off = losses.draw_crop_offsets(t_y.tolist(), out_size)
    res, attn = _gpu_path(mu_x, y, t_x, t_y)
    random.seed(5)
    r, _, _ = _oracle_block(mu_x, logw, x_mask, y, t_x, t_y, out_size, None, attn)     # draws with the same stream
    assert r["out_offset"].tolist() == off
    y_cut, ft_cut, cut_len, mask = losses.crop_frames(y.to(DEV), res.frame_token, t_y, off, out_size)
    assert torch.equal(y_cut.cpu(), r["y"])
    assert cut_len.tolist() == [min(int(t), out_size) for t in t_y]
    ref_mask = torch.zeros(B, 1, out_size)
    ref_mask[:, :, :r["y_mask"].shape[-1]] = r["y_mask"]
    assert torch.equal(mask.cpu(), ref_mask)
    # index form of attn_cut: one-hot of ft_cut
    onehot = torch.zeros(B, Tx, out_size)
    ftc = ft_cut.cpu().long()
    for b in range(B):
        for t in range(out_size):
            if ftc[b, t] >= 0:
                onehot[b, ftc[b, t], t] = 1
    assert torch.equal(onehot, r["attn_cut"])


@pytest.mark.parametrize("out_size", [None, 128])
@pytest.mark.parametrize("B,F,Tx,Ty", [(6, 80, 61, 300), (3, 128, 190, 1000), (2, 7, 9, 40)])
def test_alignment_losses_match_dense_reference_block(B, F, Tx, Ty, out_size):
    mu_x, logw, x_mask, y, t_x, t_y = _batch(B, F, Tx, Ty, 21, max(3, Tx // 3), max(Tx, Ty // 4))
    if out_size is not None and Ty < out_size:
        out_size = 16
    off = None
    if out_size is not None:
        random.seed(9)
        off = losses.draw_crop_offsets(t_y.tolist(), out_size)
    _, attn = _gpu_path(mu_x, y, t_x, t_y)
    r, mu_ref, lw_ref = _oracle_block(mu_x, logw, x_mask, y, t_x, t_y, out_size, off, attn)
    w = torch.Generator().manual_seed(3)
    cot = torch.randn(r["mu_y"].shape, generator=w)            # a stand-in for the decoder's gradient into mu_y
    (r["dur_loss"] + 3.0 * r["prior_loss"] + (r["mu_y"] * cot).sum()).backward()

    mu = mu_x.to(DEV).requires_grad_(True)
    lw = logw.to(DEV).requires_grad_(True)
    out = losses.alignment_losses(mu, lw, t_x, y.to(DEV), t_y, out_size=out_size, out_offset=off, dense_path=True)
    (out.dur_loss + 3.0 * out.prior_loss + (out.mu_y * cot.to(DEV)).sum()).backward()
    torch.cuda.synchronize()

    assert torch.equal(out.alignment.path.cpu(), attn), "the same call must give the same path"
    assert torch.equal(out.y.cpu(), r["y"])
    T = r["y_mask"].shape[-1]
    assert torch.equal(out.y_mask.cpu()[:, :, :T], r["y_mask"]) and float(out.y_mask[:, :, T:].sum()) == 0
    assert torch.equal(out.mu_y.detach().cpu(), r["mu_y"].detach()), "mu_y gather must be an exact copy of the GEMM"
    assert _rel(out.dur_loss, r["dur_loss"]) < LOSS_RTOL
    assert _rel(out.prior_loss, r["prior_loss"]) < LOSS_RTOL
    for got, ref in ((mu.grad, mu_ref.grad), (lw.grad, lw_ref.grad)):
        scale = ref.abs().max().item()
        assert (got.cpu() - ref).abs().max().item() <= GRAD_RTOL * scale
    # logf on the device vs torch's CPU log: same function, last-bit rounding may differ
    torch.testing.assert_close(losses.logw_target(out.alignment.durations, t_x).cpu(), r["logw_"], rtol=1e-6, atol=1e-6)


def test_loss_functions_standalone_and_no_grad_path():
    B, F, Tx, Ty = 4, 80, 45, 260
    mu_x, logw, x_mask, y, t_x, t_y = _batch(B, F, Tx, Ty, 31, 15, 100)
    _, attn = _gpu_path(mu_x, y, t_x, t_y)
    r, _, _ = _oracle_block(mu_x, logw, x_mask, y, t_x, t_y, None, None, attn)
    res = fgt.log_prior_maximum_path(mu_x.to(DEV), y.to(DEV), t_x, t_y, dense_path=False)
    d = losses.duration_loss(logw.to(DEV), res.durations, t_x)
    p = losses.prior_loss(mu_x.to(DEV), y.to(DEV), res.frame_token, t_y, res.durations)
    m = losses.gather_mu_y(mu_x.to(DEV), res.frame_token, res.durations)
    assert _rel(d, r["dur_loss"]) < LOSS_RTOL and _rel(p, r["prior_loss"]) < LOSS_RTOL
    assert torch.equal(m.cpu(), r["mu_y"])
    # durations_to_logw (torch expression on the durations) agrees with the kernel's logw_
    torch.testing.assert_close(fgt.durations_to_logw(res.durations, x_mask.to(DEV)).cpu(), r["logw_"], rtol=1e-6, atol=1e-6)


def test_losses_are_deterministic_run_to_run():
    B, F, Tx, Ty = 8, 80, 190, 1000
    mu_x, logw, x_mask, y, t_x, t_y = _batch(B, F, Tx, Ty, 41, 60, 300)
    vals = []
    for _ in range(3):
        mu = mu_x.to(DEV).requires_grad_(True)
        lw = logw.to(DEV).requires_grad_(True)
        out = losses.alignment_losses(mu, lw, t_x, y.to(DEV), t_y, out_size=128, out_offset=[0] * B)
        (out.dur_loss + out.prior_loss + out.mu_y.sum()).backward()
        vals.append((out.dur_loss.item(), out.prior_loss.item(), mu.grad.clone(), lw.grad.clone()))
    for v in vals[1:]:
        assert v[0] == vals[0][0] and v[1] == vals[0][1]
        assert torch.equal(v[2], vals[0][2]) and torch.equal(v[3], vals[0][3])


def test_argument_errors():
    L = fgt._lib.lib()
    assert L.mas_b200_sequence_mask(None, 1, 1, None, None) == fgt._lib.ERR_ARG
    assert L.mas_b200_duration_loss(None, None, None, 1, 1, None, None, None, None) == fgt._lib.ERR_ARG
    assert L.mas_b200_prior_loss(None, None, None, None, 1, 1, 1, 1, None, None, None, 0, None) == fgt._lib.ERR_ARG
    with pytest.raises(ValueError):
        losses.duration_loss(torch.zeros(2, 1, 5, device=DEV), torch.zeros(2, 6, dtype=torch.int32, device=DEV), [5, 5])
