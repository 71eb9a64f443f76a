#!/usr/bin/env python
"""Generate tests/golden/compute_loss_block.npz from the REFERENCE's real FaceTTS.compute_loss.

Run in the build container only (imports /root/reference; the fixture travels, the reference does not):

    python tests/golden/make_compute_loss_golden.py

The unmodified reference model (model/face_tts.py) is imported under four stubs for packages this image
lacks: `pytorch_lightning` (LightningModule = nn.Module + .device), `utils.scheduler`
(imports a class transformers 5 removed), `text` (cleaners need `unidecode`; only text.symbols is used), and the
compiled Cython core registered under the import name model/monotonic_align/__init__.py:5 uses
(oracle/_ref/asis, built from the reference's own core.pyx by oracle/build_ref.py).  Random-init weights, seed
37 (reference config.py:12), n_feats=128 (the reference default; SyncNet's audio branch needs it), eval mode
(no dropout), CPU.

Hooks record what crosses the boundary of the alignment block (face_tts.py:159-218, 233-234):
    in :  mu_x, logw, x_mask (encoder outputs, :157), y, x_lengths, y_lengths, out_size,
          the crop offsets Python's `random.choice` produced (:188)
    out:  log_prior + attn_mask as handed to maximum_path (:173), attn, the (y, y_mask, mu_y) handed to
          decoder.compute_loss (:222), dur_loss, prior_loss, and the gradients of
          dur_loss + prior_loss w.r.t. mu_x and logw.
Two small cases: cropped (out_size=128, some utterances longer, some shorter than the window) and uncropped; plus one
at the LRS2 training shape (B=16, T_text=190, T_mel=1000, out_size=128) in compute_loss_block_lrs2.npz, stored without
the tensors the tests can regenerate (y from the seed, log_prior / attn_mask from the oracle).
"""
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def install_stubs():
    import torch.nn as nn

    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(nn.Module):
        @property
        def device(self):
            p = next(self.parameters(), None)          # parameter-less modules (diffusion.py:27): CPU run
            return p.device if p is not None else torch.device("cpu")

        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

    pl.LightningModule = LightningModule
    sys.modules["pytorch_lightning"] = pl

    utils_pkg = types.ModuleType("utils")
    utils_pkg.__path__ = [os.path.join(REF, "utils")]
    sys.modules["utils"] = utils_pkg
    sched = types.ModuleType("utils.scheduler")
    sys.modules["utils.scheduler"] = sched
    utils_pkg.scheduler = sched

    text_pkg = types.ModuleType("text")
    text_pkg.__path__ = [os.path.join(REF, "text")]
    sys.modules["text"] = text_pkg

    from oracle import build_ref

    build_ref.build(verbose=False)
    core = build_ref.load("asis")
    assert core is not None, "the reference core.pyx did not build"
    for name in ("model.monotonic_align.model", "model.monotonic_align.model.monotonic_align"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    sys.modules["model.monotonic_align.model.monotonic_align.core"] = core
    sys.path.insert(0, REF)


def config():
    return dict(
        add_blank=1, vid_emb_dim=64, n_enc_channels=32, filter_channels=64, filter_channels_dp=32, n_heads=2,
        n_enc_layers=1, enc_kernel=3, enc_dropout=0.0, window_size=4, n_feats=128, dec_dim=8, beta_min=0.05,
        beta_max=20.0, pe_scale=1000.0, spk_emb="face", gamma=0.02, perceptual_loss=1, syncnet_initw=10.0,
        syncnet_initb=-5.0, syncnet_stride=1, syncnet_ckpt=None,
    )


def run_case(model, face_tts_mod, B, Tx, Ty, out_size, seed, slim=False):
    import time

    import cases

    n_vocab = model.n_vocab
    x_len, y_len, x, y, face = cases.compute_loss_inputs(B, Tx, Ty, out_size, seed, n_vocab)

    rec = {}
    enc_hook = model.encoder.register_forward_hook(lambda m, i, o: rec.update(enc=o))
    real_mp = face_tts_mod.monotonic_align.maximum_path
    real_choice = random.choice
    real_dec = model.decoder.compute_loss
    offsets = []

    t_mp = [0.0]

    def mp(value, mask):
        t0 = time.perf_counter()
        out = real_mp(value, mask)
        t_mp[0] += time.perf_counter() - t0
        rec.update(log_prior=value.detach().clone(), attn_mask=mask.detach().clone(), attn=out.detach().clone())
        return out

    def choice(seq):
        v = real_choice(seq)
        offsets.append(int(v))
        return v

    def dec(y_, y_mask_, mu_y_, spk_):
        rec.update(dec_y=y_.detach().clone(), dec_y_mask=y_mask_.detach().clone(), dec_mu_y=mu_y_.detach().clone())
        return real_dec(y_, y_mask_, mu_y_, spk_)

    face_tts_mod.monotonic_align.maximum_path = mp
    model.decoder.compute_loss = dec
    random.choice = choice
    try:
        random.seed(seed)
        torch.manual_seed(seed)
        model.zero_grad()
        t_step0 = time.perf_counter()
        dur_loss, prior_loss, diff_loss, spk_loss = model.compute_loss(x, x_len, y, y_len, spk=face, out_size=out_size)
        t_step = time.perf_counter() - t_step0
        mu_x, logw, x_mask = rec["enc"]
        g_mu, g_lw = torch.autograd.grad(dur_loss + prior_loss, [mu_x, logw])
    finally:
        face_tts_mod.monotonic_align.maximum_path = real_mp
        model.decoder.compute_loss = real_dec
        random.choice = real_choice
        enc_hook.remove()
    # the reference draws an offset only where y_len > out_size (face_tts.py:188); 0 elsewhere
    full_off = None
    if out_size is not None:
        it = iter(offsets)
        full_off = [next(it) if int(n) - out_size > 0 else 0 for n in y_len]
    assert all(torch.isfinite(v) for v in (dur_loss, prior_loss, diff_loss, spk_loss))
    print(f"  case B={B} Tx={Tx} Ty={Ty}: reference compute_loss forward {t_step * 1e3:.1f} ms on this container's CPU, "
          f"of which maximum_path (wrapper + Cython) {t_mp[0] * 1e3:.1f} ms; losses dur {dur_loss.item():.6f} "
          f"prior {prior_loss.item():.6f} diff {diff_loss.item():.6f} spk {spk_loss.item():.6f}")
    if slim:
        # LRS2-sized case: y / log_prior / attn_mask / dec_y are regenerated by the tests (cases.compute_loss_inputs and
        # the oracle's log-prior), not stored
        return dict(
            mu_x=mu_x.detach().numpy(), logw=logw.detach().numpy(), x_mask=x_mask.detach().numpy(),
            x_lengths=x_len.numpy().astype(np.int32), y_lengths=y_len.numpy().astype(np.int32),
            out_size=np.int32(-1 if out_size is None else out_size), shape=np.asarray([B, Tx, Ty], np.int32),
            n_vocab=np.int32(n_vocab),
            offsets=np.asarray(full_off if full_off is not None else [], np.int32), seed=np.int32(seed),
            attn=np.packbits(rec["attn"].numpy().astype(np.uint8), axis=-1), attn_shape=np.asarray(rec["attn"].shape, np.int32),
            dec_y_mask=rec["dec_y_mask"].numpy(), dec_mu_y=rec["dec_mu_y"].numpy(),
            dur_loss=np.float32(dur_loss.item()), prior_loss=np.float32(prior_loss.item()),
            diff_loss=np.float32(diff_loss.item()), spk_loss=np.float32(spk_loss.item()),
            ref_step_ms_cpu=np.float32(t_step * 1e3), ref_maximum_path_ms_cpu=np.float32(t_mp[0] * 1e3),
            grad_mu_x=g_mu.numpy(), grad_logw=g_lw.numpy(),
        )
    return dict(
        mu_x=mu_x.detach().numpy(), logw=logw.detach().numpy(), x_mask=x_mask.detach().numpy(), y=y.numpy(),
        x_lengths=x_len.numpy().astype(np.int32), y_lengths=y_len.numpy().astype(np.int32),
        out_size=np.int32(-1 if out_size is None else out_size),
        offsets=np.asarray(full_off if full_off is not None else [], np.int32), seed=np.int32(seed),
        log_prior=rec["log_prior"].numpy(), attn_mask=rec["attn_mask"].numpy(),
        attn=np.packbits(rec["attn"].numpy().astype(np.uint8), axis=-1), attn_shape=np.asarray(rec["attn"].shape, np.int32),
        dec_y=rec["dec_y"].numpy(), dec_y_mask=rec["dec_y_mask"].numpy(), dec_mu_y=rec["dec_mu_y"].numpy(),
        dur_loss=np.float32(dur_loss.item()), prior_loss=np.float32(prior_loss.item()),
        grad_mu_x=g_mu.numpy(), grad_logw=g_lw.numpy(),
    )


def main():
    install_stubs()
    import model.face_tts as face_tts_mod

    torch.manual_seed(37)
    model = face_tts_mod.FaceTTS(config()).eval()
    out = {}
    for name, (B, Tx, Ty, out_size, seed) in {"cropped": (5, 41, 320, 128, 101), "full": (3, 33, 96, None, 202)}.items():
        for k, v in run_case(model, face_tts_mod, B, Tx, Ty, out_size, seed).items():
            out[f"{name}/{k}"] = v
    path = os.path.join(HERE, "compute_loss_block.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    # BASELINE configs[2] at the bench shape: B=16 (the GAN micro-batch, config.py:112), T_text=190, T_mel=1000, n_feats=128,
    # out_size=128
    big = {f"lrs2/{k}": v for k, v in run_case(model, face_tts_mod, 16, 190, 1000, 128, 303, slim=True).items()}
    path2 = os.path.join(HERE, "compute_loss_block_lrs2.npz")
    np.savez_compressed(path2, **big)
    print("wrote", path2, os.path.getsize(path2), "bytes")
    for k in ("cropped/dur_loss", "cropped/prior_loss", "full/dur_loss", "full/prior_loss", "cropped/offsets"):
        print(k, out[k])


if __name__ == "__main__":
    main()
