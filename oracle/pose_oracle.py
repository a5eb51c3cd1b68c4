"""oracle/pose_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain-torch fp32 CPU restatement of the reference PoseResNet ("DeconvResnet"):

  trunk        lib/pose/models/resnet.py:16-36   (7x7/2 stem, BN, ReLU, 3x3/2 max-pool, 4 stages)
  Bottleneck   lib/pose/models/blocks.py:83-120  (1x1 -> 3x3(stride) -> 1x1, BN after each, ReLU after
                                                  the first two, projection shortcut when the shape
                                                  changes, add, ReLU)
  head         lib/pose/models/pose_deconv.py:19-30 (3 x [ConvTranspose2d k4 s2 p1 no-bias, BN, ReLU], 1x1 conv)
  forward      lib/pose/models/pose_deconv.py:32-46
  init         lib/pose/models/resnet.py:38-49, pose_deconv.py:48-63
  factory      lib/pose/models/pose_deconv.py:165-179

The state_dict key layout equals the reference's (checked by tests/test_oracle_golden.py against the
fixtures made from the real reference by tests/golden/make_golden.py).  The arithmetic is PyTorch's own
conv/BN/pool kernels, as in the reference (third-party: torch, un-vendored).

`condition_weights` is the synthetic-weight recipe of SURVEY.md section 8(d): the reference's default
init yields ~1e-4 heatmaps (vacuous parity) so tests and the bench use conditioned weights.
"""
import torch
import torch.nn as nn

STAGES = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}


class BottleneckOracle(nn.Module):
    def __init__(self, cin, planes, stride):
        super().__init__()
        cout = planes * 4
        self.conv1 = nn.Conv2d(cin, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, cout, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(cout)
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride, bias=False), nn.BatchNorm2d(cout))
        else:
            self.downsample = nn.Sequential()

    def forward(self, x):
        y = torch.relu(self.bn1(self.conv1(x)))
        y = torch.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        return torch.relu(y + self.downsample(x))


class PoseResNetOracle(nn.Module):
    def __init__(self, depth=50, num_classes=17):
        super().__init__()
        self.depth = depth
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        cin = 64
        for si, (planes, nblk) in enumerate(zip((64, 128, 256, 512), STAGES[depth])):
            blocks = []
            for b in range(nblk):
                blocks.append(BottleneckOracle(cin, planes, (2 if si > 0 else 1) if b == 0 else 1))
                cin = planes * 4
            setattr(self, "layer%d" % (si + 1), nn.Sequential(*blocks))
        head = []
        for i in range(3):
            head += [nn.ConvTranspose2d(2048 if i == 0 else 256, 256, 4, stride=2, padding=1, bias=False),
                     nn.BatchNorm2d(256), nn.ReLU(inplace=True)]
        self.deconv = nn.Sequential(*head)
        self.heatmap = nn.Conv2d(256, num_classes, 1)

    def forward(self, x):
        x = self.maxpool(torch.relu(self.bn1(self.conv1(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.heatmap(self.deconv(x))

    def init_net(self):
        """Reference default init (resnet.py:44-49 then pose_deconv.py:51-63)."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        for m in self.deconv.modules():
            if isinstance(m, nn.ConvTranspose2d):
                nn.init.normal_(m.weight, std=0.001)
        nn.init.normal_(self.heatmap.weight, std=0.001)
        nn.init.constant_(self.heatmap.bias, 0)


def deconv(backbone, num_classes, pretrained=False):
    """Same call as the reference factory `models.deconv('resnet50', 17, False)`."""
    assert backbone.startswith("resnet"), "the hot path covers the Bottleneck ResNet trunks only"
    m = PoseResNetOracle(int(backbone[len("resnet"):]), num_classes)
    m.init_net()
    return m


@torch.no_grad()
def condition_weights(model, seed=0, calib_hw=(256, 192), calib_batch=8):
    """make activations O(1) everywhere so a 1e-3 max-abs heatmap gate means something.
    Works on any module with the reference state_dict layout (oracle or product)."""
    g = torch.Generator().manual_seed(seed)
    for name, m in model.named_modules():
        if isinstance(m, nn.BatchNorm2d):
            lo, hi = (0.1, 0.3) if name.endswith("bn3") else (0.5, 1.5)
            m.weight.copy_(torch.rand(m.weight.shape, generator=g) * (hi - lo) + lo)
            m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
            m.momentum = 1.0
        elif isinstance(m, nn.ConvTranspose2d):
            std = (2.0 / (m.in_channels * 4)) ** 0.5
            m.weight.copy_(torch.randn(m.weight.shape, generator=g) * std)
    model.heatmap.weight.copy_(torch.randn(model.heatmap.weight.shape, generator=g) / 16.0)
    model.heatmap.bias.copy_(torch.randn(model.heatmap.bias.shape, generator=g) * 0.1)
    x = torch.randn(calib_batch, 3, calib_hw[0], calib_hw[1], generator=g)
    model.train()
    model(x)  # one train-mode pass: running stats := batch stats (momentum 1.0)
    model.eval()
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.momentum = 0.1
    return model
