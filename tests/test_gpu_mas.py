"""GPU parity tests for Monotonic Alignment Search (run with -m gpu on a B200).

Bar: BIT-EXACT.  The CUDA path (through the C ABI) is compared with
  * the golden fixtures the reference's compiled core.pyx produced (tests/golden), and
  * the pinned C oracle (oracle/mas_oracle.c) on seeded inputs,
under every kernel configuration (rows per lane, DP warps, direction bits in shared
or global memory, TMA-bulk or fallback loader, in-kernel or separate path write).
Nothing here reads /root/reference.
"""
import hashlib
import itertools

import numpy as np
import pytest
import torch

import cases
import oracle
import face_gan_tts_b200 as fgt
from face_gan_tts_b200 import _lib, monotonic_align, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


class options:
    """temporarily set library tuning knobs"""

    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        self.prev = {k: _lib.set_option(k, v) for k, v in self.kw.items()}

    def __exit__(self, *a):
        for k, v in self.prev.items():
            _lib.set_option(k, v)


_REF_CORE = oracle.reference_core("asis")      # the reference's own core.pyx, compiled by oracle/build_ref.py (travels in oracle/_ref)


def oracle_paths(value_np, t_x, t_y):
    """Paths of the reference MAS: the compiled reference itself when it is on this box, else the C restatement that
    tests/test_oracle.py pins to it."""
    p = np.zeros(value_np.shape, np.int32)
    fn = _REF_CORE.maximum_path_c if _REF_CORE is not None else oracle.maximum_path_c
    fn(p, np.ascontiguousarray(value_np, np.float32).copy(), np.asarray(t_x, np.int32), np.asarray(t_y, np.int32))
    return p


def run_align(value_np, t_x, t_y, path_dtype=torch.int32):
    v = torch.from_numpy(np.ascontiguousarray(value_np, np.float32)).to(DEV)
    res = fgt.align(v, torch.from_numpy(np.asarray(t_x, np.int32)), torch.from_numpy(np.asarray(t_y, np.int32)),
                    path_dtype=path_dtype)
    torch.cuda.synchronize()
    return res


def assert_matches_golden(res, kats, name):
    path = res.path.cpu().numpy().astype(np.int32)
    assert sha(path) == str(kats[f"{name}/path_sha256"]), f"{name}: dense path differs from the reference"
    np.testing.assert_array_equal(res.durations.cpu().numpy(), kats[f"{name}/dur"])
    np.testing.assert_array_equal(res.frame_token.cpu().numpy(), kats[f"{name}/frame_token"])
    assert (res.status.cpu().numpy() == 0).all()


@pytest.mark.parametrize("name", list(cases.CASES))
def test_golden_kats_default_plan(kats, name):
    value, t_x, t_y = cases.CASES[name]()
    assert_matches_golden(run_align(value, t_x, t_y), kats, name)


# every (rows-per-lane, DP-warps) instantiation
RW = [(r, w) for r in (1, 2, 4, 8) for w in (1, 2, 3, 4) if 32 * r * w <= 768]   # two ring stages must fit in 227 KB


@pytest.mark.parametrize("R,W", RW)
def test_every_kernel_instantiation_bit_exact(kats, R, W):
    """cfg-shaped + tie / clamp / NaN / edge cases under a forced (R, W): rows beyond 32*R*W
    exercise the multi-pass (global carry line) path."""
    names = ["rand_small", "small_int_ties", "clamp_random", "nan_cell", "inf_cells", "edge_65x129", "lrs2_shape",
             "padding_garbage", "signed_zeros"]
    with options(mas_rows_per_lane=R, mas_dp_warps=W):
        for name in names:
            value, t_x, t_y = cases.CASES[name]()
            assert_matches_golden(run_align(value, t_x, t_y), kats, name)


@pytest.mark.parametrize("opt", [
    dict(mas_force_global_bits=1),
    dict(mas_force_unaligned=1),
    dict(mas_fused_path_write=1),
    dict(mas_fused_path_write=0),
    dict(mas_ctas_per_sm=3),
    dict(mas_ring_stages=2),
    dict(mas_force_global_bits=1, mas_force_unaligned=1, mas_fused_path_write=1, mas_ring_stages=2),
])
def test_kernel_variants_bit_exact(kats, opt):
    with options(**opt):
        for name in ["cfg1", "lrs2_shape", "large_accum", "wide_513x1030", "edge_31x33", "all_zero", "tx_eq_ty"]:
            value, t_x, t_y = cases.CASES[name]()
            assert_matches_golden(run_align(value, t_x, t_y), kats, name)


def test_dropin_maximum_path_matches_reference_wrapper(kats):
    """monotonic_align.maximum_path(value, mask): same signature/return contract as reference
    model/monotonic_align/__init__.py:8-23 (dtype, device, {0,1}); cfg1 = BASELINE configs[0]."""
    value, t_x, t_y = cases.CASES["cfg1"]()
    B, Tx, Ty = value.shape
    mask = synthetic.prefix_mask(torch.from_numpy(t_x), torch.from_numpy(t_y), Tx, Ty).to(DEV)
    v = torch.from_numpy(value).to(DEV)
    v_before = v.clone()
    path = monotonic_align.maximum_path(v, mask)
    assert path.dtype == v.dtype and path.device == v.device and path.shape == v.shape
    assert torch.equal(v, v_before), "value must not be clobbered"
    p = path.cpu().numpy()
    assert set(np.unique(p)) <= {0.0, 1.0}
    assert sha(p.astype(np.int32)) == str(kats["cfg1/path_sha256"])
    # garbage outside the mask is never read (the reference zeroes it with value*mask)
    v2 = torch.where(mask > 0, v, torch.full_like(v, float("nan")))
    assert torch.equal(monotonic_align.maximum_path(v2, mask), path)
    # other dtypes / CPU tensors keep the reference's contract
    p16 = monotonic_align.maximum_path(v.double(), mask.double())
    assert p16.dtype == torch.float64 and torch.equal(p16.float(), path)
    pc = monotonic_align.maximum_path(v.cpu(), mask.cpu())
    assert pc.device.type == "cpu" and torch.equal(pc, path.cpu())


def test_host_buffer_maximum_path_c(kats):
    """core.maximum_path_c(paths, values, t_xs, t_ys): the Cython entry point's argument meaning."""
    value, t_x, t_y = cases.CASES["lrs2_shape"]()
    paths = np.full(value.shape, 7, np.int32)      # need not be pre-zeroed
    vals = np.ascontiguousarray(value, np.float32)
    keep = vals.copy()
    bad = monotonic_align.core.maximum_path_c(paths, vals, t_x, t_y)
    assert bad == 0 and sha(paths) == str(kats["lrs2_shape/path_sha256"])
    np.testing.assert_array_equal(vals, keep)


def test_strided_value_and_float_path(kats):
    value, t_x, t_y = cases.CASES["edge_65x129"]()
    B, Tx, Ty = value.shape
    big = torch.full((B, Tx + 3, Ty + 5), float("nan"), device=DEV)
    big[:, :Tx, :Ty] = torch.from_numpy(value).to(DEV)
    view = big[:, :Tx, :Ty]                               # stride_x = Ty+5: unaligned rows
    res = fgt.align(view, torch.from_numpy(t_x), torch.from_numpy(t_y), path_dtype=torch.float32)
    assert sha(res.path.cpu().numpy().astype(np.int32)) == str(kats["edge_65x129/path_sha256"])


def test_rejected_items_are_reported_not_undefined():
    v = torch.randn(3, 6, 8, device=DEV)
    t_x = torch.tensor([6, 7, 0], dtype=torch.int32)      # ok, t_x > Tx, t_x < 1
    t_y = torch.tensor([8, 8, 8], dtype=torch.int32)
    res = fgt.align(v, t_x, t_y)
    assert res.status.cpu().tolist() == [0, 1, 1]
    assert res.path[1:].abs().sum().item() == 0 and res.durations[1:].sum().item() == 0
    assert (res.frame_token[1:] == -1).all()
    res = fgt.align(v, torch.tensor([6, 5, 4], dtype=torch.int32), torch.tensor([8, 3, 8], dtype=torch.int32))
    assert res.status.cpu().tolist() == [0, 1, 0]         # t_x > t_y: UB in the reference (core.pyx:34)
    with pytest.raises(ValueError):
        fgt.align(v, torch.tensor([6, 5, 4], dtype=torch.int32), torch.tensor([8, 3, 8], dtype=torch.int32), check=True)


_EDGE_TX = [1, 2, 31, 32, 33, 127, 128, 129, 255, 256, 257]      # lane / warp / M-tile / CTA-pass edges
_EDGE_TY = [1, 2, 31, 32, 33, 63, 64, 65, 255, 256, 257]          # 32-frame tile edges


@pytest.mark.parametrize("seed", range(120))
def test_random_fuzz_vs_oracle(seed):
    rng = np.random.default_rng(100 + seed)
    B = int(rng.integers(1, 9))
    if seed % 2:                                                    # sizes on tile edges
        Tx = _EDGE_TX[(seed // 2) % len(_EDGE_TX)]
        Ty = max(Tx, _EDGE_TY[(seed // 2 + seed // 22) % len(_EDGE_TY)]) + int(rng.integers(0, 2)) * 32 * int(rng.integers(0, 8))
    else:
        Tx = int(rng.integers(1, 300))
        Ty = int(rng.integers(Tx, 700))
    kind = seed % 3
    if kind == 0:
        v = rng.standard_normal((B, Tx, Ty)).astype(np.float32)
    elif kind == 1:
        v = rng.integers(-1, 2, (B, Tx, Ty)).astype(np.float32)          # massive ties
    else:
        v = (rng.standard_normal((B, Tx, Ty)) * 3e8 - 4e8).astype(np.float32)   # -1e9 clamp active
    t_x = rng.integers(1, Tx + 1, B).astype(np.int32)
    t_y = np.asarray([rng.integers(t_x[b], Ty + 1) for b in range(B)], np.int32)
    res = run_align(v, t_x, t_y)
    np.testing.assert_array_equal(res.path.cpu().numpy(), oracle_paths(v, t_x, t_y))


def test_long_utterance_streamed_path_vs_oracle():
    """BASELINE configs[3] shape (Tx=512, Ty=4096: direction bits exceed shared memory), B reduced so the
    CPU oracle finishes in seconds; full-B properties are checked in test_full_size_properties."""
    v, t_x, t_y = synthetic.mas_value(4, 512, 4096, seed=77, tx_lo=256, ty_lo=2048)
    res = fgt.align(v.to(DEV), t_x, t_y, path_dtype=torch.int32)
    ref = oracle_paths(v.numpy(), t_x.numpy(), t_y.numpy())
    np.testing.assert_array_equal(res.path.cpu().numpy(), ref)
    dur, ft = oracle.durations_and_frame_token(ref)
    np.testing.assert_array_equal(res.durations.cpu().numpy(), dur)
    np.testing.assert_array_equal(res.frame_token.cpu().numpy(), ft)


def check_path_properties(res, t_x, t_y):
    """size-independent invariants of a MAS path"""
    path, dur, ft = res.path, res.durations.long(), res.frame_token.long()
    B, Tx, Ty = path.shape
    t_x = t_x.to(path.device).long()
    t_y = t_y.to(path.device).long()
    ar_x = torch.arange(Tx, device=path.device)[None]
    ar_y = torch.arange(Ty, device=path.device)[None]
    assert torch.equal(path.sum(2).long(), dur)                               # durations = row sums
    assert torch.equal(path.sum(1).long(), (ar_y < t_y[:, None]).long())      # one token per valid frame
    assert torch.equal(dur.sum(1), t_y)
    assert ((dur >= 1) == (ar_x < t_x[:, None])).all()                        # surjective, padding empty
    valid = ar_y < t_y[:, None]
    assert (ft[~valid] == -1).all()
    d = ft[:, 1:] - ft[:, :-1]
    ok = (d == 0) | (d == 1)
    assert ok[valid[:, 1:]].all()                                             # monotone, no skips
    assert (ft[:, 0] == 0).all()
    last = ft.gather(1, (t_y - 1)[:, None]).squeeze(1)
    assert torch.equal(last, t_x - 1)
    assert torch.equal(path.argmax(1)[valid].long(), ft[valid])               # frame_token consistent


def test_full_size_properties():
    """Full BASELINE sizes: configs[3] (B=64, 512x4096) and a configs[4]-sized batch (B=1024 LRS2 shape)."""
    v, t_x, t_y = synthetic.mas_value(64, 512, 4096, seed=5, tx_lo=256, ty_lo=2048)
    res = fgt.align(v.to(DEV), t_x, t_y, path_dtype=torch.float32)
    check_path_properties(res, t_x, t_y)
    del res, v
    g = torch.Generator(device="cpu").manual_seed(9)
    v = torch.randn(1024, 190, 1000, generator=g)
    t_x, t_y = synthetic.lengths_mas(1024, 190, 1000, 60, 300, seed=10)
    res = fgt.align(v.to(DEV), t_x, t_y, path_dtype=torch.float32)
    check_path_properties(res, t_x, t_y)
    # spot-check 6 utterances of the big batch against the oracle
    idx = [0, 1, 147, 148, 500, 1023]
    ref = oracle_paths(v[idx].numpy(), t_x[idx].numpy(), t_y[idx].numpy())
    np.testing.assert_array_equal(res.path[idx].cpu().numpy().astype(np.int32), ref)


def test_lengths_from_mask_matches_reference_rule():
    t_x = torch.tensor([5, 1, 9], dtype=torch.int32)
    t_y = torch.tensor([12, 7, 20], dtype=torch.int32)
    mask = synthetic.prefix_mask(t_x, t_y, 9, 20).to(DEV)
    from face_gan_tts_b200.alignment import lengths_from_mask

    a, b = lengths_from_mask(mask)
    ex, ey = oracle.lengths_from_mask(mask.cpu().numpy())
    np.testing.assert_array_equal(a.cpu().numpy(), ex)
    np.testing.assert_array_equal(b.cpu().numpy(), ey)
    a2, b2 = lengths_from_mask(mask.bool())
    assert torch.equal(a2, a) and torch.equal(b2, b)


def test_generate_path_matches_reference_expression():
    """reference model/utils.py:27-40, restated with torch ops"""
    torch.manual_seed(3)
    B, Tx, Ty = 4, 23, 96
    t_x = torch.tensor([23, 11, 1, 17], dtype=torch.int32)
    t_y = torch.tensor([96, 40, 9, 96], dtype=torch.int32)
    mask = synthetic.prefix_mask(t_x, t_y, Tx, Ty).to(DEV)
    dur = torch.randint(0, 7, (B, Tx), device=DEV).float()
    cum = torch.cumsum(dur, 1)
    seq = (torch.arange(Ty, device=DEV)[None, None, :] < cum[:, :, None]).float()
    ref = (seq - torch.nn.functional.pad(seq, (0, 0, 1, 0))[:, :-1]) * mask
    out = fgt.generate_path(dur, mask)
    assert out.dtype == mask.dtype and torch.equal(out, ref)


@pytest.mark.parametrize("length_scale", [1.0, 1.1, 0.77])
def test_inference_expansion_matches_reference_forward(length_scale):
    """reference model/face_tts.py:118-129 (w_ceil * length_scale -> y_lengths -> generate_path -> mu_y), with the
    reference's generate_path / sequence_mask restated by the oracle on the CPU (sequential fp32 cumsum)."""
    g = torch.Generator().manual_seed(11)
    B, F, Tx = 5, 80, 37
    t_x = torch.tensor([37, 12, 1, 30, 25])
    x_mask = (torch.arange(Tx)[None, :] < t_x[:, None]).float().unsqueeze(1)
    logw = torch.randn(B, 1, Tx, generator=g) * 0.8 + 0.7
    mu_x = torch.randn(B, F, Tx, generator=g) * x_mask
    w_ceil = torch.ceil(torch.exp(logw) * x_mask) * length_scale                       # :118-119
    y_lengths = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()                    # :120
    Ty = int(y_lengths.max())
    Ty += (-Ty) % 4                                                                      # fix_len_compatibility :122
    y_mask = oracle.sequence_mask(y_lengths, Ty).unsqueeze(1).to(x_mask.dtype)          # :124
    attn_mask = x_mask.unsqueeze(-1) * y_mask.unsqueeze(2)                               # :125
    ref_attn = oracle.generate_path(w_ceil.squeeze(1), attn_mask.squeeze(1))             # :126
    ref_mu_y = torch.matmul(ref_attn.transpose(1, 2), mu_x.transpose(1, 2)).transpose(1, 2)   # :128-129
    path, ft = fgt.generate_path(w_ceil.squeeze(1).to(DEV), attn_mask.squeeze(1).to(DEV), return_index=True)
    assert torch.equal(path.cpu(), ref_attn)
    mu_y, ft2 = fgt.expand_durations(mu_x.to(DEV), w_ceil.to(DEV), t_x, y_lengths, Ty)
    assert torch.equal(ft, ft2) and torch.equal(mu_y.cpu(), ref_mu_y)
    onehot = torch.zeros(B, Tx, Ty)
    for b in range(B):
        for t in range(Ty):
            if ft[b, t] >= 0:
                onehot[b, ft[b, t], t] = 1
    assert torch.equal(onehot, ref_attn)


def test_durations_to_logw_matches_dense_expression(kats):
    value, t_x, t_y = cases.CASES["lrs2_shape"]()
    res = run_align(value, t_x, t_y, path_dtype=torch.float32)
    Tx = value.shape[1]
    x_mask = (torch.arange(Tx)[None, None, :] < torch.from_numpy(t_x)[:, None, None]).float().to(DEV)
    dense = torch.log(1e-8 + torch.sum(res.path.unsqueeze(1), -1)) * x_mask            # face_tts.py:176
    assert torch.equal(fgt.durations_to_logw(res.durations, x_mask), dense)
