"""CPU-side checks of the drop-in boundary (no GPU): the reference's UNMODIFIED training script, launched through
dropin/run_unchanged.py, imports, parses its arguments, builds the B200 modules through its own import statements,
loads a `Scale.`-prefixed checkpoint, draws batches from the synthetic dataset stand-ins and reaches libasn_b200 -- where,
on a machine without a GPU, the call is refused loudly (there is no CPU fallback to fall into silently)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_synthetic_datasets_have_the_reference_item_layout():
    sys.path.insert(0, os.path.join(ROOT, "dropin"))
    try:
        from dataset.cityscapes_dataset import cityscapesDataSet
        from dataset.gta5_dataset import GTA5DataSet
    finally:
        sys.path.remove(os.path.join(ROOT, "dropin"))
    mean = np.array((104.00698793, 116.66876762, 122.67891434), dtype=np.float32)
    g = GTA5DataSet("/x", "/nonexistent/list.txt", max_iters=10, crop_size=(64, 32), scale=False, mirror=False, mean=mean)
    img, lab, size, name = g[3]                                  # dataset/gta5_dataset.py:71
    assert img.shape == (3, 32, 64) and img.dtype == np.float32 and lab.shape == (32, 64) and lab.dtype == np.float32
    assert set(np.unique(lab)) <= set(range(19)) | {255.0} and tuple(size) == (32, 64, 3) and isinstance(name, str)
    assert len(g) >= 10 and -123 <= img.min() and img.max() <= 151
    c = cityscapesDataSet("/x", "/nonexistent/list.txt", max_iters=4, crop_size=(64, 32), scale=False, mirror=False,
                          mean=mean, set="train")                # train...:333-337
    item = c[0]
    assert len(item) == 3 and item[0].shape == (3, 32, 64) and item[0].dtype == np.float32   # train...:418
    batch = next(iter(torch.utils.data.DataLoader(g, batch_size=1)))
    assert batch[0].shape == (1, 3, 32, 64) and batch[1].long().dtype == torch.int64


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "train_gta2cityscapes_multi.py")),
                    reason="reference checkout not mounted")
def test_unchanged_train_script_reaches_the_library_and_is_refused_on_cpu(tmp_path):
    from adaptsegnet_b200.model.deeplab_multi import DeeplabMulti
    torch.manual_seed(0)
    torch.save({"Scale." + k: v for k, v in DeeplabMulti(19).state_dict().items()}, str(tmp_path / "init.pth"))
    cmd = [sys.executable, os.path.join(ROOT, "dropin", "run_unchanged.py"), "--reference", REF, "train", "--", "--cpu",
           "--level", "multi-level", "--gan", "Vanilla", "--warper", "", "--num-steps", "1", "--num-steps-stop", "1",
           "--input-size", "128,64", "--input-size-target", "128,64", "--num-workers", "0", "--restore-from",
           str(tmp_path / "init.pth"), "--snapshot-dir", str(tmp_path / "snap"), "--data-list", "/nonexistent/a",
           "--data-list-target", "/nonexistent/b"]
    r = subprocess.run(cmd, cwd=str(tmp_path), capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, PYTHONWARNINGS="ignore"))
    assert r.returncode != 0
    assert "adaptsegnet_b200 has no CPU fallback" in r.stderr, r.stderr[-2000:]
    assert "train_gta2cityscapes_multi.py" in r.stderr and "pred1, pred2 = model(images)" in r.stderr   # train...:597
