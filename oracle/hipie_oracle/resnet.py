"""Oracle: detectron2 ResNet-50 (FrozenBN, STRIDE_IN_1X1 False), config #1 only (test infrastructure).

Restates /root/reference/detectron2/modeling/backbone/resnet.py: BasicStem :329-366, BottleneckBlock :105-205,
make_stage / build_resnet_backbone :614-694, and layers/batch_norm.py FrozenBatchNorm2d :13-118 (eps 1e-5,
y = x * (w * rsqrt(var + eps)) + (b - mean * w * rsqrt(var + eps))).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class FrozenBatchNorm2d(nn.Module):
    def __init__(self, num_features, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.register_buffer("weight", torch.ones(num_features))
        self.register_buffer("bias", torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features) - eps)

    def forward(self, x):
        scale = self.weight * (self.running_var + self.eps).rsqrt()
        bias = self.bias - self.running_mean * scale
        return x * scale.reshape(1, -1, 1, 1).to(x.dtype) + bias.reshape(1, -1, 1, 1).to(x.dtype)


class ConvNorm(nn.Conv2d):
    """detectron2.layers.Conv2d with a `norm` child (wrappers.py:70-110)."""

    def __init__(self, cin, cout, k, stride=1, padding=0):
        super().__init__(cin, cout, k, stride=stride, padding=padding, bias=False)
        self.norm = FrozenBatchNorm2d(cout)

    def forward(self, x):
        return self.norm(F.conv2d(x, self.weight, None, self.stride, self.padding))


class BottleneckBlock(nn.Module):
    def __init__(self, cin, cout, bottleneck, stride):
        super().__init__()
        self.shortcut = ConvNorm(cin, cout, 1, stride=stride) if cin != cout else None
        self.conv1 = ConvNorm(cin, bottleneck, 1, stride=1)          # stride_in_1x1 = False
        self.conv2 = ConvNorm(bottleneck, bottleneck, 3, stride=stride, padding=1)
        self.conv3 = ConvNorm(bottleneck, cout, 1)

    def forward(self, x):
        out = F.relu_(self.conv1(x))
        out = F.relu_(self.conv2(out))
        out = self.conv3(out)
        sc = self.shortcut(x) if self.shortcut is not None else x
        return F.relu_(out + sc)


class BasicStem(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = ConvNorm(3, 64, 7, stride=2, padding=3)

    def forward(self, x):
        return F.max_pool2d(F.relu_(self.conv1(x)), kernel_size=3, stride=2, padding=1)


class ResNet50(nn.Module):
    def __init__(self):
        super().__init__()
        self.stem = BasicStem()
        cin, bott, cout = 64, 64, 256
        for name, n, first_stride in (("res2", 3, 1), ("res3", 4, 2), ("res4", 6, 2), ("res5", 3, 2)):
            blocks = []
            for i in range(n):
                blocks.append(BottleneckBlock(cin, cout, bott, first_stride if i == 0 else 1))
                cin = cout
            setattr(self, name, nn.Sequential(*blocks))
            bott *= 2
            cout *= 2
        self.size_divisibility = 0
        self.num_channels = [512, 1024, 2048]
        self.strides = [8, 16, 32]

    def forward(self, x):
        x = self.res2(self.stem(x))
        r3 = self.res3(x)
        r4 = self.res4(r3)
        r5 = self.res5(r4)
        return {"res3": r3, "res4": r4, "res5": r5}
