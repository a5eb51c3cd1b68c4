"""tests/golden/make_golden.py -- generate the golden fixtures from the REAL reference.

Run once in the build container (needs /root/reference; the GPU box has no reference):

    python tests/golden/make_golden.py

It imports the reference's own modules (lib/pose unmodified; lib/flownet through a stub layer for the
long-removed `torch.utils.ffi` and the three `_ext` modules, whose C-ABI callables are backed by the
oracle's CPU restatements because the reference has no CPU implementation of them), loads the
ORACLE's seeded weights into them (strict `load_state_dict`, which also proves the key layouts are
identical), runs them on seeded inputs and stores inputs-by-seed + outputs in tests/golden/*.npz.

tests/test_oracle_golden.py then checks oracle == these fixtures without needing the reference.
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import ops_oracle  # noqa: E402
from tests.golden import recipes  # noqa: E402


def _keys_digest(sd):
    return hashlib.sha256("\n".join("%s %s" % (k, tuple(v.shape)) for k, v in sd.items()).encode()).hexdigest()


def import_reference_pose():
    sys.path.insert(0, os.path.join(REF, "lib", "pose"))
    import models as ref_pose_models  # noqa
    import utils.evaluation as ref_eval  # noqa
    import utils.transforms as ref_tf  # noqa
    sys.path.pop(0)
    return ref_pose_models, ref_eval, ref_tf


def import_reference_flow():
    """fake torch.utils.ffi + pre-registered _ext modules exposing the six legacy names."""
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "utils" or k.startswith("utils.")]:
        del sys.modules[k]  # lib/pose's top-level `models`/`utils` would shadow lib/flownet's
    ffi = types.ModuleType("torch.utils.ffi")
    ffi._wrap_function = lambda fn, _ffi: fn
    sys.modules["torch.utils.ffi"] = ffi

    def corr_fwd(in1, in2, r1, r2, out, pad, k, md, s1, s2, mult):
        res = ops_oracle.correlation_c(in1, in2, pad, k, md, s1, s2)
        out.resize_(res.shape).copy_(res)
        return 1

    def resample_fwd(in1, in2, out, k):
        out.copy_(ops_oracle.resample2d_c(in1, in2))
        return 1

    def cnorm_fwd(in1, out, deg):
        out.copy_(ops_oracle.channelnorm_c(in1))
        return 1

    def reg(name, **fns):
        m = types.ModuleType(name)
        m.lib = types.SimpleNamespace(**fns)
        m.ffi = None
        sys.modules[name] = m

    reg("correlation_package._ext.correlation._correlation", Correlation_forward_cuda=corr_fwd, Correlation_backward_cuda=None)
    reg("networks.resample2d_package._ext.resample2d._resample2d", Resample2d_cuda_forward=resample_fwd, Resample2d_cuda_backward=None)
    reg("networks.channelnorm_package._ext.channelnorm._channelnorm", ChannelNorm_cuda_forward=cnorm_fwd, ChannelNorm_cuda_backward=None)
    sys.path.insert(0, os.path.join(REF, "lib", "flownet", "networks"))
    sys.path.insert(0, os.path.join(REF, "lib", "flownet"))
    import warnings
    warnings.filterwarnings("ignore")
    from model import models as ref_flow_models  # noqa
    return ref_flow_models


def make_pose(ref_models, out):
    for depth, (hw, nb) in recipes.POSE_CASES.items():
        orc = recipes.pose_oracle_model(depth)
        ref = ref_models.deconv("resnet%d" % depth, 17, False)
        ref.load_state_dict(orc.state_dict(), strict=True)
        ref.eval()
        x = recipes.pose_input(depth)
        with torch.no_grad():
            y = ref(x)
        out["pose%d_heatmap" % depth] = y.numpy()
        out["pose%d_meta" % depth] = np.array([depth, hw[0], hw[1], nb, sum(p.numel() for p in ref.parameters())], dtype=np.int64)
        out["pose%d_keys" % depth] = np.array(_keys_digest(ref.state_dict()))
        print("pose", depth, tuple(y.shape), float(y.abs().max()))


def make_flow(ref_flow, out):
    class Args:
        rgb_max = 255.0
        fp16 = False
        grads = {}

    for name in recipes.FLOW_NAMES:
        orc = recipes.flow_oracle_model(name)
        ref = getattr(ref_flow, name)(Args()).eval()
        ref.load_state_dict(orc.state_dict(), strict=True)
        x = recipes.flow_input()
        y = ref(x).detach()   # FlowNet2.forward registers hooks -> needs grad mode
        out["flow_%s_out" % name] = y.numpy()
        out["flow_%s_params" % name] = np.array(sum(p.numel() for p in ref.parameters()), dtype=np.int64)
        out["flow_%s_keys" % name] = np.array(_keys_digest(ref.state_dict()))
        print(name, tuple(y.shape), float(y.abs().max()), int(out["flow_%s_params" % name]))


def make_decode(ref_eval, ref_tf, out):
    hm = recipes.decode_heatmaps()
    center = np.array([[100.0, 120.0], [320.5, 200.25], [50.0, 60.0]])
    scale = np.array([200.0, 310.0, 96.5])
    # torch-0.4 semantics for `indexes.div(w)` on an integer tensor (evaluation.py:15): floor division.
    orig_div = torch.Tensor.div
    torch.Tensor.div = lambda self, other, **kw: (orig_div(self, other, rounding_mode="floor")
                                                  if not self.is_floating_point() else orig_div(self, other, **kw))
    try:
        coords, scores = ref_eval.final_preds(hm, center, scale, adjust_coords=True)
        coords0, _ = ref_eval.final_preds(hm, center, scale, adjust_coords=False)
    finally:
        torch.Tensor.div = orig_div
    out["decode_hm_seed"] = np.array(5)
    out["decode_center"], out["decode_scale"] = center, scale
    out["decode_coords_adjust"], out["decode_coords_plain"], out["decode_scores"] = coords, coords0, scores
    pts = np.array([[[0.0, 0.0], [10.25, 20.75], [47.0, 63.0]]], dtype=np.float32)
    out["tf_pts"] = pts
    out["tf_out"] = ref_tf.transform_preds(pts.copy(), center[:1], scale[:1], (64, 48))
    print("decode ok", coords.shape, scores.shape)


def make_track(ref_tf, ref_eval=None):
    """Person crops through the reference's own transform_image (transforms.py:231-240) with pose_est's centre/scale
    rule (net_utils.py:49-50); stored in a separate small file (reference_track.npz)."""
    img = recipes.track_frame()
    boxes = recipes.TRACK_BOXES
    res = recipes.TRACK_RES
    centers = np.stack((boxes[:, [0, 2]].mean(1), boxes[:, [1, 3]].mean(1)), axis=1)
    scales = np.maximum(boxes[:, 3] - boxes[:, 1], (boxes[:, 2] - boxes[:, 0]) / res[1] * res[0])
    crops = np.stack([ref_tf.transform_image(img, centers[i], scales[i], res).numpy() for i in range(len(boxes))])
    crops_f = np.stack([ref_tf.transform_image(img.astype(np.float32), centers[i], scales[i], res).numpy() for i in range(len(boxes))])
    hm = recipes.flip_heatmaps()
    swapped = {ds: ref_tf.swaplr_image(hm[:, :k].copy(), ds) for ds, k in (("coco", 17), ("mpii", 16), ("aic", 14))}
    extra = {}
    if ref_eval is not None:      # OKS-NMS through the reference's own nms_oks / compute_oks (evaluation.py:60-101)
        preds = recipes.oks_poses()
        for th, kt in ((0.9, 0.0), (0.5, 0.2), (0.3, 0.5)):
            extra["oks_keep_%g_%g" % (th, kt)] = np.array(ref_eval.nms_oks(preds, th, recipes.COCO_DELTA, kt), np.int64)
        j = np.array([p["joints"] for p in preds])
        extra["oks_pairwise0"] = ref_eval.compute_oks(j[1:], j[0], np.full(len(preds) - 1, 5000.0), recipes.COCO_DELTA, ground_truth=False, threshold=0.2)
    # box NMS through the reference's own cpu_nms (nms.c compiled unmodified against oracle/ref_shims, `make -C oracle ref`)
    from oracle import track_oracle
    for n, seed, th in recipes.NMS_CASES:
        k = track_oracle.ref_cpu_nms(recipes.nms_boxes(n, seed), th)
        assert k is not None, "oracle/_ref/libref_nms.so missing: run `make -C oracle ref`"
        extra["nms_keep_%d_%d_%g" % (n, seed, th)] = k
    path = os.path.join(HERE, "reference_track.npz")
    np.savez_compressed(path, crops_u8=crops, crops_f32=crops_f, centers=centers, scales=scales,
                        **{"swap_" + ds: v for ds, v in swapped.items()}, **extra)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB", crops.shape, crops.dtype)


def make_flowvis():
    """Middlebury colour coding through the reference's own flowlib.flow_to_image (lib/flownet/utils/flowlib.py:243-282;
    its matplotlib / pypng imports are display-only and stubbed) -> reference_flowvis.npz."""
    import importlib.util
    import types
    for name in ("png", "matplotlib", "matplotlib.colors", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    spec = importlib.util.spec_from_file_location("ref_flowlib", os.path.join(REF, "lib", "flownet", "utils", "flowlib.py"))
    flowlib = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(flowlib)
    out = {}
    for i, flow in enumerate(recipes.flowvis_flows()):
        out["flowvis_img%d" % i] = flowlib.flow_to_image(flow.copy())
        print("flowvis", i, flow.shape, out["flowvis_img%d" % i].shape, out["flowvis_img%d" % i].dtype)
    out["flowvis_wheel"] = flowlib.make_color_wheel()
    path = os.path.join(HERE, "reference_flowvis.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def main():
    out = {}
    if "--flowvis-only" in sys.argv:
        make_flowvis()
        return
    make_flowvis()
    ref_pose, ref_eval, ref_tf = import_reference_pose()
    if "--track-only" in sys.argv:
        make_track(ref_tf, ref_eval)
        return
    make_track(ref_tf, ref_eval)
    make_pose(ref_pose, out)
    make_decode(ref_eval, ref_tf, out)
    ref_flow = import_reference_flow()
    make_flow(ref_flow, out)
    path = os.path.join(HERE, "reference_outputs.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
