"""CPU-side checks of the C-ABI boundary: the library builds/loads without a GPU and
exports every symbol include/mas_b200.h declares; host-side argument validation and
the no-CPU-fallback contract."""
import os
import re
import ctypes

import numpy as np
import pytest

from face_gan_tts_b200 import _lib

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mas_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mas_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mas_b200.h but not exported"
    # and the Python binding covers exactly the header
    assert sorted(_lib.SIGNATURES) == names


def test_abi_version_and_error_strings():
    L = _lib.lib()
    assert L.mas_b200_abi_version() == 1
    assert L.mas_b200_error_string(0) == b"ok"
    for code in (-1, -2, -3, -4, -5):
        assert L.mas_b200_error_string(code)


def test_workspace_sizes_are_monotone_and_aligned():
    L = _lib.lib()
    a = L.mas_b200_workspace_bytes(32, 190, 1000)
    b = L.mas_b200_workspace_bytes(64, 190, 1000)
    c = L.mas_b200_workspace_bytes(64, 512, 4096)
    assert 0 < a < b < c and a % 256 == 0
    assert L.mas_b200_workspace_bytes(0, 1, 1) == 0
    f = L.mas_b200_fused_workspace_bytes(32, 80, 190, 1000)
    assert f >= a + 32 * 190 * 1000 * 4


def test_options_roundtrip():
    prev = _lib.set_option("mas_rows_per_lane", 4)
    assert _lib.get_option("mas_rows_per_lane") == 4
    _lib.set_option("mas_rows_per_lane", prev)
    with pytest.raises(KeyError):
        _lib.get_option("no_such_option")


def test_host_argument_validation_happens_before_any_cuda_call():
    L = _lib.lib()
    assert L.mas_b200_maximum_path(None, 0, 0, None, None, 1, 1, 1, -1e9, None, 0, None, None, None, None, 0, None) == _lib.ERR_ARG
    assert L.mas_b200_log_prior(None, None, 1, 80, 1, 1, None, 0, None) == _lib.ERR_ARG
    assert L.mas_b200_generate_path(None, None, None, 1, 1, 1, None, 1, None) == _lib.ERR_ARG
    assert L.mas_b200_maximum_path_host(None, None, None, None, 1, 1, 1, -1e9) == _lib.ERR_ARG
    # the consumers of the alignment (loss_ops.cu) and the upload
    assert L.mas_b200_sequence_mask(None, 1, 1, None, None) == _lib.ERR_ARG
    assert L.mas_b200_crop_frames(None, None, None, None, 1, 1, 1, 1, None, None, None, None, None) == _lib.ERR_ARG
    assert L.mas_b200_gather_mu_y(None, None, 1, 1, 1, 1, None, None) == _lib.ERR_ARG
    assert L.mas_b200_gather_mu_y_backward(None, None, None, None, None, 1, 1, 1, 1, None, None) == _lib.ERR_ARG
    assert L.mas_b200_prior_loss(None, None, None, None, 1, 1, 1, 1, None, None, None, 0, None) == _lib.ERR_ARG
    assert L.mas_b200_prior_loss_backward(None, None, None, None, None, None, None, 1, 1, 1, 1, None, None) == _lib.ERR_ARG
    assert L.mas_b200_duration_loss(None, None, None, 1, 1, None, None, None, None) == _lib.ERR_ARG
    assert L.mas_b200_generate_path_f32(None, None, None, 1, 1, 1, None, 1, None, None) == _lib.ERR_ARG
    assert L.mas_b200_upload_batch(None, None, None, None, 1, 1, 1, 1, None, None, None, None, None) == _lib.ERR_ARG
    assert L.mas_b200_prior_loss_workspace_bytes(32, 80, 1000) == 8 * 32 * 10 * 4
    assert L.mas_b200_prior_loss_workspace_bytes(0, 80, 1000) == 0


def test_no_cpu_fallback():
    """Without a CUDA device the compute entry points must FAIL, never fall back."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from face_gan_tts_b200 import monotonic_align

    with pytest.raises(RuntimeError):
        monotonic_align.maximum_path(torch.zeros(1, 2, 3), torch.ones(1, 2, 3))
    with pytest.raises(_lib.MasB200Error):
        monotonic_align.core.maximum_path_c(np.zeros((1, 2, 3), np.int32), np.zeros((1, 2, 3), np.float32),
                                            np.array([2], np.int32), np.array([3], np.int32))
    with pytest.raises(ValueError):
        import face_gan_tts_b200 as f

        f.align(torch.zeros(1, 2, 3), [2], [3])


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: no file of the product package may reference it."""
    pkg = os.path.join(ROOT, "face-gan-tts_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "mas_oracle" not in txt, fn


def test_install_shim_registers_reference_import_names():
    import sys
    import face_gan_tts_b200 as f

    ma = f.install()
    try:
        assert sys.modules["model.monotonic_align"] is ma
        assert sys.modules["model.monotonic_align.model.monotonic_align.core"].maximum_path_c is ma.core.maximum_path_c
    finally:
        f.uninstall()
    assert "model.monotonic_align" not in sys.modules


def test_pack_batch_host_layout_matches_numpy():
    """mas_b200_pack_batch_host (no GPU needed): [t_x][t_y][pad16][mu rows, valid part][y rows, valid part]."""
    import numpy as np
    import torch
    import face_gan_tts_b200 as fgt

    B, F, Tx, Ty = 5, 7, 13, 29
    g = torch.Generator().manual_seed(0)
    mu_x = torch.randn(B, F, Tx, generator=g)
    y = torch.randn(B, F, Ty, generator=g)
    t_x = torch.tensor([13, 1, 7, 4, 12], dtype=torch.int32)
    t_y = torch.tensor([29, 5, 17, 4, 28], dtype=torch.int32)
    packed = fgt.pack_batch(mu_x, y, t_x, t_y, pin=False).numpy()
    hdr = (8 * B + 15) // 16 * 16
    assert packed.size == hdr + 4 * F * int(t_x.sum() + t_y.sum())
    ints = packed[:8 * B].view(np.int32)
    assert ints[:B].tolist() == t_x.tolist() and ints[B:].tolist() == t_y.tolist()
    body = packed[hdr:].view(np.float32)
    want = np.concatenate([mu_x[b, f, :t_x[b]].numpy() for b in range(B) for f in range(F)] +
                          [y[b, f, :t_y[b]].numpy() for b in range(B) for f in range(F)])
    np.testing.assert_array_equal(body, want)


def test_tensor_core_issue_is_warp_uniform_in_the_shipped_sass():
    """The tcgen05.mma issue loops must be straight uniform-datapath code.  When ptxas cannot prove that the issuing warp
    is converged (e.g. after an early exit behind a block barrier) it wraps every elect-predicated UTCHMMA in a
    divergent fallback loop (ELECT ... BRA.U.ANY): ~90 instead of ~20 cycles per instruction, which starves the
    alignment search behind it (measured: 52 -> 85 cycles per frame).  Catch that regression in the binary."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    fn, mma, fallback = None, {}, {}
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
        elif "UTCHMMA" in line:
            mma[fn] = mma.get(fn, 0) + 1
        elif "BRA.U.ANY" in line:
            fallback[fn] = fallback.get(fn, 0) + 1
    kernels = [f for f in mma if "lp_mas_fused_kernel" in f or "log_prior_tc_kernel" in f]
    assert len(kernels) >= 8, kernels
    for f in kernels:
        assert fallback.get(f, 0) <= 1, f"{f}: {fallback[f]} divergent fallback loops around {mma[f]} UTCHMMA"


def test_host_side_entry_points_refuse_cpu_tensors_before_touching_the_library():
    """No CPU path: the device halves of the public API raise on host tensors (they never fall back to computing on
    the CPU), and the host halves raise on device-side misuse they can detect without a GPU."""
    import torch
    import face_gan_tts_b200 as fgt

    with pytest.raises(ValueError):
        fgt.unpack_batch(torch.zeros(64, dtype=torch.uint8), 1, 1, 1, 1)             # staging must be a CUDA buffer
    with pytest.raises(ValueError):
        fgt.align(torch.zeros(1, 2, 3), torch.tensor([2]), torch.tensor([3]))
    with pytest.raises(ValueError):
        fgt.log_prior(torch.zeros(1, 80, 4), torch.zeros(1, 80, 8))
    with pytest.raises(ValueError):
        fgt.log_prior_maximum_path(torch.zeros(1, 80, 4), torch.zeros(1, 80, 8), torch.tensor([4]), torch.tensor([8]))
    with pytest.raises(ValueError):
        fgt.pack_batch(torch.zeros(1, 2, 3), torch.zeros(1, 2, 4), torch.tensor([3]), torch.tensor([4]))   # int64 lengths
    with pytest.raises((RuntimeError, ValueError)):
        fgt.monotonic_align.maximum_path(torch.zeros(1, 2, 3), torch.ones(1, 2, 3))
