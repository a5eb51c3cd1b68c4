"""GPU: the tools/* entry points end to end (reference flags in, reference on-disk formats out) against the CPU oracle."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import decode_oracle, track_oracle
from tests import gpu_util
from tests.golden import recipes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_flownet_main_inference_save_flow(tmp_path):
    """tools/flownet/main.py --inference --save_flow: .flo files in the reference's folder layout hold the oracle's flows."""
    import cv2
    from flowtrack.pytorch_b200.flownet.utils import flow_utils
    fm = _load("tools/flownet/main.py", "ft_tools_flownet_main")
    rng = np.random.default_rng(1)
    frames = tmp_path / "frames"
    frames.mkdir()
    for i in range(3):
        cv2.imwrite(str(frames / ("f%02d.png" % i)), rng.integers(0, 255, (140, 200, 3), dtype=np.uint8))
    orc = recipes.flow_oracle_model("FlowNet2S")
    ckpt = str(tmp_path / "FlowNet2-S_checkpoint.pth.tar")
    torch.save({"arch": "FlowNet2S", "epoch": 7, "state_dict": orc.state_dict(), "best_EPE": 1.0}, ckpt)
    rc = fm.main(["--inference", "--save_flow", "--save_flow_image", "--model", "FlowNet2S", "--resume", ckpt, "--save", str(tmp_path / "work"),
                  "--name", "run", "--inference_dataset_root", str(frames)])
    assert rc == 0
    folder = tmp_path / "work" / "inference" / "run.epoch-6-flow-field"
    assert sorted(os.listdir(folder)) == ["000000.flo", "000000.png", "000001.flo", "000001.png"]
    ds = fm.ImagesFromFolder(fm.build_parser().parse_args(["--inference"]), False, root=str(frames), iext="png")
    for i in range(2):
        with torch.no_grad():
            want = orc(ds[i][0][0][None])[0].permute(1, 2, 0).numpy()
        got = flow_utils.readFlow(str(folder / ("%06d.flo" % i)))
        assert got.shape == (128, 192, 2)
        assert np.abs(got - want).max() <= gpu_util.FLOW_TOL * max(1.0, np.abs(want).max())
        png = cv2.imread(str(folder / ("%06d.png" % i)))[:, :, ::-1]
        from oracle import flowvis_oracle
        d = np.abs(png.astype(np.int32) - flowvis_oracle.flow_to_image(got).astype(np.int32))
        assert d.max() <= 1


@pytest.mark.parametrize("flip", [False, True])
def test_pose_main_validate(tmp_path, flip):
    """tools/pose/main.py run_type=valid over a .npz of crops, with and without the flip test (tools/pose/main.py:289-299)."""
    pm = _load("tools/pose/main.py", "ft_tools_pose_main")
    orc = recipes.pose_oracle_model(50)
    n, hw = 5, (128, 96)
    g = torch.Generator().manual_seed(8)
    x = torch.randn(n, 3, hw[0], hw[1], generator=g)
    center = np.stack([np.linspace(90, 300, n), np.linspace(100, 250, n)], 1)
    scale = np.linspace(120, 260, n)
    data = str(tmp_path / "crops.npz")
    np.savez(data, inputs=x.numpy(), center=center, scale=scale, area=np.full(n, 5000.0), score=np.linspace(1, 0.5, n), image=np.zeros(n, np.int64))
    ckpt = str(tmp_path / "deconv_resnet50_best.pth")
    torch.save({"epoch": 3, "model": "deconv", "state_dict": orc.state_dict(), "best_loss": 0.1}, ckpt)
    preds = pm.main(data=data, backbone="resnet50", resume=ckpt, flip_test=flip, test_batch_size=2, checkpoint_path=str(tmp_path / "ck"))
    with torch.no_grad():
        hm = orc(x).numpy()
        if flip:
            hm = track_oracle.flip_test_merge(hm, orc(torch.flip(x, dims=[3])).numpy(), "coco")
    xy, sc = decode_oracle.final_preds(hm, center, scale, True)
    assert preds.shape == (n, 17, 3)
    assert np.abs(preds[:, :, :2] - xy).max() <= gpu_util.KEYPOINT_TOL_PX
    assert np.abs(preds[:, :, 2:] - sc).max() < gpu_util.HEATMAP_TOL
    assert os.path.exists(os.path.join(str(tmp_path / "ck"), "coco", "exp_01", "preds.npy"))
