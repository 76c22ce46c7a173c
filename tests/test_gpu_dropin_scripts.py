"""-m gpu: the reference's UNMODIFIED scripts run on the B200 modules ("scripts run unchanged").

dropin/run_unchanged.py puts dropin/ in front of the reference checkout on sys.path and executes
train_gta2cityscapes_multi.py (multi-level branch, 2 iterations) and evaluate_cityscapes.py (one snapshot, one frame)
exactly as they are on disk.  The checkout is looked up at /root/reference (build container) or baseline/_ref/reference
(the git-ignored staging directory that travels to the GPU box: `tools/stage_reference.sh`); the tests skip where neither
exists -- the reference's sources are never part of this repository."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from gpu_util import gpu

pytestmark = gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference():
    for cand in (os.environ.get("ASN_REFERENCE", ""), "/root/reference", os.path.join(ROOT, "baseline", "_ref", "reference")):
        if cand and os.path.exists(os.path.join(cand, "train_gta2cityscapes_multi.py")):
            return cand
    pytest.skip("reference checkout not available on this machine (see tools/stage_reference.sh)")


def _run(ref, script, args, cwd, env_extra=None):
    env = dict(os.environ, PYTHONWARNINGS="ignore", **(env_extra or {}))
    cmd = [sys.executable, os.path.join(ROOT, "dropin", "run_unchanged.py"), "--reference", ref, script, "--"] + args
    return subprocess.run(cmd, cwd=cwd, env=env, capture_output=True, text=True, timeout=900)


def _fake_checkpoint(path):
    """a `Scale.`-prefixed DeepLab checkpoint like the one train...:202-215 expects, random weights"""
    from adaptsegnet_b200.model.deeplab_multi import DeeplabMulti
    torch.manual_seed(1338)
    sd = DeeplabMulti(19).state_dict()
    torch.save({"Scale." + k: v for k, v in sd.items()}, path)


@pytest.mark.parametrize("lazy", ["1", "0"])
def test_train_script_multi_level_runs_unchanged(tmp_path, lazy):
    ref = _reference()
    ckpt = tmp_path / "init.pth"
    _fake_checkpoint(str(ckpt))
    args = ["--level", "multi-level", "--gan", "Vanilla", "--warper", "", "--num-steps", "2", "--num-steps-stop", "2",
            "--input-size", "256,128", "--input-size-target", "192,96", "--num-workers", "0", "--restore-from", str(ckpt),
            "--snapshot-dir", str(tmp_path / "snap"), "--data-list", "/nonexistent/gta5.txt", "--data-list-target",
            "/nonexistent/city.txt"]
    r = _run(ref, "train", args, str(tmp_path), {"ASN_LAZY_LOGITS": lazy})
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("iter =")]
    assert len(lines) == 2, r.stdout[-2000:]
    vals = [float(v) for v in re.findall(r"= ([0-9.]+|nan)", lines[-1].split(",", 1)[1])]
    assert len(vals) == 6 and all(np.isfinite(vals)) and vals[0] > 0 and 0.3 < vals[4] < 1.5   # D loss ~ log 2
    # the snapshots the script writes at num_steps_stop load back key for key
    sd = torch.load(str(tmp_path / "snap" / "multi_level" / "GTA5_2.pth"), map_location="cpu")
    from adaptsegnet_b200.model.deeplab_multi import DeeplabMulti
    assert set(sd) == set(DeeplabMulti(19).state_dict())
    assert os.path.exists(str(tmp_path / "snap" / "multi_level" / "GTA5_2_D1.pth"))
    if lazy == "1":
        test_train_script_multi_level_runs_unchanged.lazy_lines = lines
    elif hasattr(test_train_script_multi_level_runs_unchanged, "lazy_lines"):
        # Tier-B handles and materialised tensors print the same losses.  The two runs are two PROCESSES of a script that
        # sets cudnn.benchmark = True (train...:228): cuDNN's autotuner may pick different TF32 algorithms for the trunk in
        # each, and the random-init trunk (segmentation loss ~9, logits of order 10) amplifies that to a few 1e-3 relative --
        # measured 5.5e-3 once in four runs on the pool's boxes, while libasn_b200 itself is bitwise run-to-run
        # deterministic (tools/determinism_check.py).  Hence 2e-2 here; the lazy-vs-materialised equality of the library's
        # own kernels is gated in-process, bit for bit, by tests/test_gpu_lazy.py.
        a = re.findall(r"= ([0-9.]+)", test_train_script_multi_level_runs_unchanged.lazy_lines[0].split(",", 1)[1])
        b = re.findall(r"= ([0-9.]+)", lines[0].split(",", 1)[1])
        assert np.allclose([float(x) for x in a], [float(x) for x in b], rtol=2e-2, atol=2e-3)


def test_evaluate_script_runs_unchanged(tmp_path):
    ref = _reference()
    from adaptsegnet_b200.model.deeplab_multi import DeeplabMulti
    torch.manual_seed(7)
    os.makedirs(tmp_path / "snapshots" / "multi_level")
    torch.save(DeeplabMulti(19).state_dict(), str(tmp_path / "snapshots" / "multi_level" / "GTA5_5000.pth"))
    args = ["--level", "multi-level", "--num-steps-stop", "5000", "--save-pred-every", "5000", "--multi-gpu",
            "--save", str(tmp_path / "result"), "--data-list", "/nonexistent/val.txt"]
    r = _run(ref, "evaluate", args, str(tmp_path), {"ASN_SYNTHETIC_FRAMES": "1"})
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    out_dir = tmp_path / "result" / "multi_level" / "step5000"
    pngs = sorted(p for p in os.listdir(out_dir) if not p.endswith("_color.png"))
    assert pngs, os.listdir(out_dir)
    from PIL import Image
    pred = np.array(Image.open(str(out_dir / pngs[0])))
    assert pred.shape == (1024, 2048) and pred.dtype == np.uint8 and pred.max() < 19
    # the same frame through the fused evaluation path gives the same labels (bit-exact tail; same logits)
    sys.path.insert(0, os.path.join(ROOT, "dropin"))
    from dataset.cityscapes_dataset import cityscapesDataSet
    from adaptsegnet_b200.evaluate import predict_labels
    mean = np.array((104.00698793, 116.66876762, 122.67891434), dtype=np.float32)
    img = torch.from_numpy(cityscapesDataSet("/x", "/nonexistent/val.txt", crop_size=(1024, 512), mean=mean, scale=False,
                                             mirror=False, set="val")[0][0])[None].cuda()
    model = DeeplabMulti(19).cuda().eval()
    model.load_state_dict(torch.load(str(tmp_path / "snapshots" / "multi_level" / "GTA5_5000.pth")))
    fused = predict_labels(model, img)[0].cpu().numpy()
    assert (fused != pred).mean() < 1e-3   # identical kernels up to cuDNN algorithm choice in the two processes
