"""BASELINE configs[2], as far as it can be shown without a GPU: the reference's UNMODIFIED FaceTTS.compute_loss
(model/face_tts.py:142-241), imported from /root/reference after `face_gan_tts_b200.install()`, calls THIS library's
`maximum_path(log_prior, attn_mask.squeeze(1))` (face_tts.py:12,173) -- and the library refuses CPU tensors loudly instead
of falling back.  Runs only where the reference tree exists (the build container); in a subprocess, because the import
stubs for the packages this image lacks would leak into the other tests."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

SCRIPT = r'''
import os, sys, importlib.util
ROOT = sys.argv[1]
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "face-gan-tts_b200")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
spec = importlib.util.spec_from_file_location("mkgolden", os.path.join(ROOT, "tests", "golden", "make_compute_loss_golden.py"))
mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
mk.install_stubs()                      # pytorch_lightning / utils.scheduler / text stubs + the reference on sys.path
import face_gan_tts_b200 as fgt
ma = fgt.install()                      # the drop-in under the names the reference imports
import model.face_tts as face_tts_mod   # reference code, unmodified
assert face_tts_mod.monotonic_align is ma, "the reference did not pick up the drop-in"
assert face_tts_mod.__file__.startswith("/root/reference/"), face_tts_mod.__file__
import cases
torch.manual_seed(37)
net = face_tts_mod.FaceTTS(mk.config()).eval()
B, Tx, Ty = 3, 33, 96
x_len, y_len, x, y, face = cases.compute_loss_inputs(B, Tx, Ty, None, 202, net.n_vocab)
seen = {}
real = ma.maximum_path
def spy(value, mask):
    seen.update(value=tuple(value.shape), mask=tuple(mask.shape), dtype=str(value.dtype), mask_vals=sorted(set(mask.unique().tolist())))
    return real(value, mask)
ma.maximum_path = spy
try:
    net.compute_loss(x, x_len, y, y_len, spk=face, out_size=None)
except (RuntimeError, ValueError) as ex:
    msg = str(ex)
else:
    raise SystemExit("compute_loss on CPU tensors did not fail: a CPU fallback exists")
finally:
    ma.maximum_path = real
assert "CUDA" in msg and ("no CPU fallback" in msg or "no CPU path" in msg), msg
assert seen["value"] == seen["mask"] and seen["value"][0] == B and seen["dtype"] == "torch.float32", seen
assert seen["value"][1] == int(x_len.max()) and seen["value"][2] == int(y_len.max()), (seen, x_len, y_len)
assert seen["mask_vals"] == [0.0, 1.0], seen
print("OK", seen)
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "model")), reason="the reference tree is only in the build container")
def test_unmodified_reference_compute_loss_calls_the_drop_in_and_there_is_no_cpu_fallback():
    r = subprocess.run([sys.executable, "-c", SCRIPT, ROOT], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    assert r.stdout.strip().splitlines()[-1].startswith("OK")
