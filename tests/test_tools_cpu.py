"""CPU: host logic of the tools/* entry points (argument surface, ImagesFromFolder pairing / centre crop, option parsing).
The forward passes of these tools are covered on the GPU by tests/test_gpu_tools.py."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_flownet_main_flags_and_images_from_folder(tmp_path):
    import cv2
    fm = _load("tools/flownet/main.py", "ft_tools_flownet_main")
    rng = np.random.default_rng(0)
    for i in range(4):
        cv2.imwrite(str(tmp_path / ("f%02d.png" % i)), rng.integers(0, 255, (140, 200, 3), dtype=np.uint8))
    args = fm.build_parser().parse_args(["--inference", "--save_flow", "--model", "FlowNet2S", "--inference_dataset_root", str(tmp_path)])
    assert args.inference and args.save_flow and args.inference_size == [-1, -1] and args.rgb_max == 255.0 and args.inference_batch_size == 1
    ds = fm.ImagesFromFolder(args, False, root=str(tmp_path), iext="png")
    assert len(ds) == 3 and args.inference_size == [128, 192]                  # consecutive pairs, largest multiple of 64
    (images,), (target,) = ds[1]
    assert tuple(images.shape) == (3, 2, 128, 192) and tuple(target.shape) == (3, 2, 128, 192)[:1] + (2, 128, 192)
    a = cv2.imread(str(tmp_path / "f01.png"))[:, :, ::-1]
    assert np.array_equal(images[:, 0].numpy().transpose(1, 2, 0), a[6:134, 4:196].astype(np.float32))      # centre crop, RGB order
    b = cv2.imread(str(tmp_path / "f02.png"))[:, :, ::-1]
    assert np.array_equal(images[:, 1].numpy().transpose(1, 2, 0), b[6:134, 4:196].astype(np.float32))
    with pytest.raises(SystemExit):
        fm.main(["--model", "FlowNet2S"])                                      # training is out of scope: --inference is required


def test_pose_main_options_and_npz_loader(tmp_path):
    pm = _load("tools/pose/main.py", "ft_tools_pose_main")
    kw = pm._cli(["--backbone=resnet101", "--flip_test=True", "--input_res=(384,288)", "--oks_threshold=0.8"])
    assert kw == {"backbone": "resnet101", "flip_test": True, "input_res": (384, 288), "oks_threshold": 0.8}
    opt = pm.DefaultConfig()
    opt.parse(kw)
    assert opt.backbone == "resnet101" and opt.flip_test is True and opt.adjust_coords is True and opt.dataset == "coco"
    path = str(tmp_path / "crops.npz")
    np.savez(path, inputs=np.zeros((5, 3, 64, 48), np.float32), center=np.ones((5, 2)), scale=np.full(5, 64.0), area=np.ones(5), score=np.ones(5))
    batches = list(pm.NpzCrops(path, 2))
    assert [b[0].shape[0] for b in batches] == [2, 2, 1] and set(batches[0][3]) == {"center", "scale", "area", "score"}
    assert pm.num_joints == {"mpii": 16, "aic": 14, "coco": 17}


def test_ncu_capture_is_tied_to_the_kernel_sources(tmp_path, monkeypatch):
    """bench.read_traffic_capture: the committed conv capture is accepted only while the conv kernel's own sources (build.kernel_hash
    over ft_conv.cu / ft_tc.cuh / ft_common.cuh + flags) are the ones it was taken with and the library was built from the tree."""
    import json
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    from flowtrack.pytorch_b200 import build as ft_build
    h = ft_build.kernel_hash(ft_build.CONV_SOURCES)
    assert h == ft_build.kernel_hash(tuple(reversed(ft_build.CONV_SOURCES))) and h != ft_build.kernel_hash(ft_build.CORR_SOURCES)
    with open(os.path.join(root, "profiles", "r02_conv_traffic.json")) as f:
        tr = json.load(f)
    assert tr["kernel_sources"] == list(ft_build.CONV_SOURCES) and tr["dram_bytes_per_launch"] > 0
    monkeypatch.setattr(ft_build, "read_info", lambda: {"source_sha256": ft_build.source_hash()})
    traffic, note = bench.read_traffic_capture(None)
    if tr["kernel_sources_sha256"] == h:          # the committed capture belongs to the committed conv sources
        assert traffic == tr["dram_bytes_per_launch"] and note["kernel_sources_sha256"] == h
    else:
        assert traffic is None and "other conv kernel sources" in note["refused"]
    monkeypatch.setattr(ft_build, "kernel_hash", lambda names: "0" * 64)
    traffic, note = bench.read_traffic_capture(None)
    assert traffic is None and "other conv kernel sources" in note["refused"]
    monkeypatch.setattr(ft_build, "read_info", lambda: {"source_sha256": "stale"})
    traffic, note = bench.read_traffic_capture(None)
    assert traffic is None and "not built from the sources" in note["refused"]
