"""Validation entry point of the reference's tools/pose/main.py (`run_type='valid'`, :207-209 -> validate(), :254-372) on the
B200 path.  Training (`train()`, optimisers, tensorboard, checkpoints writers) and the COCO / MPII dataset classes are out
of scope; what is kept:

  * the option names of tools/pose/config.py (dataset, model, backbone, input_res, test_batch_size, checkpoint_path, exp_id,
    resume, adjust_coords, flip_test, oks_threshold, kpt_threshold), passed fire-style as `--name=value`;
  * `validate(valid_loader, valid_data, model, criterion, opt)` with the reference's loop: forward, optional flip test
    (:289-299), final_preds(output, center, scale, opt.adjust_coords) (:316), `all_preds` [N, J, 3], metas, and
    `valid_data.evaluate(all_preds, metas, work_dir)` when the dataset object offers it;
  * the checkpoint format ({'state_dict', 'epoch', 'best_loss', ...}, :176-181).

`valid_loader` is any iterable of (data [B,3,H,W], heatmaps-or-None, target_weight-or-None, meta) with meta['center'] [B,2],
meta['scale'] [B] (the tuples the reference datasets yield).  Without a dataset package the CLI reads a `.npz` of crops:

    python tools/pose/main.py --data=crops.npz --backbone=resnet50 --resume=deconv_resnet50_best.pth --flip_test=True

(arrays `inputs` [N,3,H,W] float32, `center` [N,2], `scale` [N]; optional `area`, `score`, `image`) and writes
<checkpoint_path>/<dataset>/<exp_id>/preds.npy.  The forward, the flip merge and the decode run on the GPU."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from flowtrack.pytorch_b200.pose import models  # noqa: E402
from flowtrack.pytorch_b200.pose.utils.evaluation import final_preds, flip_test_forward, nms_oks  # noqa: E402,F401

num_joints = {"mpii": 16, "aic": 14, "coco": 17}


class DefaultConfig(object):
    """The inference-side options of tools/pose/config.py with the values of its commented COCO block (:8-62)."""
    dataset = "coco"
    data = ""                 # .npz of crops (this package ships no dataset classes)
    use_gpu = True
    input_res = (256, 192)
    model = "deconv"
    backbone = "resnet50"
    with_bg = False
    pretrained = False
    test_batch_size = 32
    checkpoint_path = "./checkpoints"
    exp_id = "exp_01"
    resume = None
    run_type = "valid"
    adjust_coords = True
    flip_test = False
    oks_threshold = 0.9
    kpt_threshold = 0.2

    def parse(self, kwargs):
        for k, v in kwargs.items():
            if not hasattr(self, k):
                import warnings
                warnings.warn("Warning: opt has not attribut %s" % k)
            setattr(self, k, v)


class AverageMeter(object):
    def __init__(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count if self.count != 0 else 0


class NpzCrops(object):
    """Iterable with the tuple layout of the reference's validation datasets, over a .npz of ready-made crops."""

    def __init__(self, path, batch_size):
        z = np.load(path)
        self.inputs, self.center, self.scale = z["inputs"].astype(np.float32), z["center"], z["scale"]
        self.extra = {k: z[k] for k in ("area", "score", "image") if k in z.files}
        self.batch_size = batch_size

    def __len__(self):
        return (len(self.inputs) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        for s in range(0, len(self.inputs), self.batch_size):
            e = min(len(self.inputs), s + self.batch_size)
            meta = {"center": torch.from_numpy(self.center[s:e]), "scale": torch.from_numpy(self.scale[s:e])}
            meta.update({k: torch.from_numpy(np.asarray(v[s:e])) for k, v in self.extra.items()})
            yield torch.from_numpy(self.inputs[s:e]), None, None, meta


def validate(valid_loader, valid_data, model, criterion, opt, writer_dict=None):
    """tools/pose/main.py:254-372.  Returns (loss average or None, results dict, ap) like the reference when `valid_data`
    can evaluate, and always leaves the predictions in `validate.all_preds` ([N, J, 3] = x, y, score in image pixels)."""
    losses = AverageMeter()
    model.eval()
    preds_all, metas = [], {"image": [], "area": [], "score": []}
    annos, ref_scale = [], []
    with torch.no_grad():
        for data, heatmaps, target_weight, meta in valid_loader:
            inputs = data.cuda()
            if opt.flip_test:
                # :289-299 -- mirrored crops in the same batch, left/right joints swapped back, average (ft_flip_w / ft_flip_merge)
                output = flip_test_forward(model, inputs, opt.dataset)
            else:
                output = model(inputs)
            if criterion is not None and heatmaps is not None:
                loss = criterion(output, heatmaps.cuda(non_blocking=True), target_weight.cuda(non_blocking=True) if target_weight is not None else None, None)
                losses.update(float(loss), inputs.size(0))
            c, s = meta["center"].numpy(), meta["scale"].numpy()
            if opt.with_bg:
                output = output[:, :-1].contiguous()
            preds, scores = final_preds(output, c, s, opt.adjust_coords)                    # :316
            preds_all.append(np.concatenate((preds, scores), axis=2).astype(np.float32))
            if opt.dataset == "coco":
                for k in ("area", "score", "image"):
                    if k in meta:
                        metas[k].extend(meta[k].numpy().tolist())
            elif opt.dataset == "mpii" and "joints" in meta:
                annos.append(meta["joints"].numpy())
                ref_scale.append(meta["ref_scale"].numpy())
    all_preds = np.concatenate(preds_all, axis=0) if preds_all else np.zeros((0, num_joints[opt.dataset], 3), np.float32)
    validate.all_preds, validate.metas = all_preds, metas
    if annos:
        metas["joints"], metas["ref_scale"] = np.concatenate(annos, axis=0), np.concatenate(ref_scale, axis=0)
    results, ap = {}, None
    if valid_data is not None and hasattr(valid_data, "evaluate"):
        results, ap = valid_data.evaluate(all_preds, metas, opt.work_dir)
    return (losses.avg if losses.count else None), dict(results), ap


def main(**kwargs):
    opt = DefaultConfig()
    opt.parse(kwargs)
    opt.work_dir = os.path.join(opt.checkpoint_path, opt.dataset, opt.exp_id)
    os.makedirs(opt.work_dir, exist_ok=True)
    num_classes = num_joints[opt.dataset] + (1 if opt.with_bg else 0)
    if opt.model != "deconv":
        raise SystemExit("only model='deconv' (PoseResNet) is on the B200 hot path (got %r)" % opt.model)
    print("==> creating model '{}', backbone = {}".format(opt.model, opt.backbone))
    model = getattr(models, opt.model)(opt.backbone, num_classes=num_classes, pretrained=opt.pretrained)
    opt.model_name = "{}_{}".format(opt.model, opt.backbone)
    if opt.resume:
        model_path = opt.resume if os.path.exists(opt.resume) else os.path.join(opt.work_dir, opt.resume)
        if os.path.exists(model_path):
            checkpoint = torch.load(model_path, map_location="cpu")
            model.load_state_dict(checkpoint["state_dict"])
            print("=> loaded checkpoint '{}' (epoch {})".format(opt.resume, checkpoint.get("epoch")))
        else:
            print("=> no checkpoint found at '{}'".format(opt.resume))
    model = model.cuda()
    print("    Total params: %.4fM" % (sum(p.numel() for p in model.parameters()) / 1000000.0))
    if opt.run_type != "valid":
        raise SystemExit("run_type=%r: only 'valid' is implemented (the training loop is out of scope)" % opt.run_type)
    if not opt.data:
        raise SystemExit("pass --data=<crops.npz> (inputs [N,3,H,W], center [N,2], scale [N]); the COCO/MPII dataset classes are out of scope")
    loader = NpzCrops(opt.data, opt.test_batch_size)
    _, results, ap = validate(loader, None, model, None, opt)
    out = os.path.join(opt.work_dir, "preds.npy")
    np.save(out, validate.all_preds)
    print("wrote %s %s" % (out, validate.all_preds.shape))
    if "area" in loader.extra and "score" in loader.extra:                               # per-image OKS-NMS as coco.py:232-277 applies it
        delta = 2 * np.array([.26, .25, .25, .35, .35, .79, .79, .72, .72, .62, .62, 1.07, 1.07, .87, .87, .89, .89]) / 10.0
        img = loader.extra.get("image", np.zeros(len(validate.all_preds), np.int64))
        kept = 0
        for im in np.unique(img):
            idx = np.nonzero(img == im)[0]
            pred = [{"score": float(loader.extra["score"][i]), "joints": validate.all_preds[i], "area": float(loader.extra["area"][i])} for i in idx]
            kept += len(nms_oks(pred, opt.oks_threshold, delta, opt.kpt_threshold))
        print("OKS-NMS (threshold %.2f): kept %d of %d poses" % (opt.oks_threshold, kept, len(validate.all_preds)))
    return validate.all_preds


def _cli(argv):
    kw = {}
    for a in argv:
        if not a.startswith("--") or "=" not in a:
            raise SystemExit("options are passed as --name=value (tools/pose/config.py names)")
        k, v = a[2:].split("=", 1)
        try:
            import ast
            v = ast.literal_eval(v)
        except Exception:
            pass
        kw[k] = v
    return kw


if __name__ == "__main__":
    main(**_cli(sys.argv[1:]))
