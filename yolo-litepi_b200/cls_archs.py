"""Layer plans of the other classifier architectures the reference can be started with
(``build_classifier``, ``src/vntsr/pipeline/e2e.py:320-347``: ``--clf_arch {resnet18, efficientnet, mobilenetv2,
shufflenetv2}``): torchvision ``resnet18``, ``mobilenet_v2`` and ``efficientnet_b0`` with the classification head
swapped for ``num_classes`` (e2e.py:323-330), lowered from their ``state_dict`` to the same ``lp_op_desc`` list the
detector and ShuffleNetV2 use (BatchNorm folded, split-f16 NHWC activations, channels padded to multiples of 16 so
that the 1x1 / 3x3 convs run on the tcgen05 kernel where their shape allows it).

These run layer by layer (``lp_classify`` without the fused ShuffleNetV2 program): parity first -- logits within 1e-2
and top-1 equal to torchvision on the CPU -- they are the SURVEY 8(f)4 row, not the benchmarked configuration.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import numpy as np

from . import _lib as L
from .plan import Plan, View, _fold_bn

G = 16      # channel padding granularity (tensor-core K / N granularity)


def _np_sd(state_dict: dict) -> Dict[str, np.ndarray]:
    return {k: (v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)) for k, v in state_dict.items()}


class _Builder:
    def __init__(self, sd: dict, in_size: int, mean: float, std: float, bn_eps: float = 1e-5):
        self.sd, self.P, self.S, self.mean, self.std, self.eps = _np_sd(sd), Plan(), in_size, mean, std, bn_eps
        self.img = View(self.P.buf(in_size, in_size, 3, L.FMT_U8), (0,), (3,), False)

    def hw(self, v: View) -> int:
        return self.P.bufs[v.buf]["h"]

    def view(self, hw: int, c: int) -> View:
        return self.P.new_view(hw, hw, [c], gran=G)

    def conv_bn(self, name: str, conv: str, bn: Optional[str], src: View, stride: int, act: int, res: Optional[View] = None,
                flags: int = 0, stem: bool = False) -> View:
        w = self.sd[conv + ".weight"]
        if bn is not None:
            w, b = _fold_bn(w, self.sd, bn, self.eps)
        else:
            b = self.sd[conv + ".bias"].astype(np.float32)
        k = w.shape[2]
        hi = self.S if stem else self.hw(src)
        ho = (hi + 2 * (k // 2) - k) // stride + 1
        dst = self.view(ho, w.shape[0])
        self.P.conv(name, w, b, src, dst, stride, act, res=res, flags=flags,
                    kind=L.OP_STEM_U8 if stem else L.OP_CONV, in_mean=self.mean if stem else 0.0, in_std=self.std if stem else 1.0)
        return dst

    def dw_bn(self, name: str, conv: str, bn: str, src: View, stride: int, act: int) -> View:
        w, b = _fold_bn(self.sd[conv + ".weight"], self.sd, bn, self.eps)            # [C, 1, k, k]
        c, k = w.shape[0], w.shape[2]
        assert c == src.logical
        ho = (self.hw(src) + 2 * (k // 2) - k) // stride + 1
        dst = self.view(ho, c)
        wp = np.zeros((k * k, src.phys), np.float32)
        bp = np.zeros(src.phys, np.float32)
        cm = src.chan_map()
        wp[:, cm] = w.reshape(c, k * k).T
        bp[cm] = b
        self.P.simple(L.OP_DWCONV3, name, src, dst, ksize=k, stride=stride, w=wp, b=bp, act=act)
        self.P.macs[-1] = ho * ho * k * k * c
        return dst

    def squeeze_excite(self, name: str, prefix: str, x: View) -> View:
        """torchvision ops/misc.py SqueezeExcitation: x * sigmoid(fc2(silu(fc1(mean_hw(x)))))"""
        c = x.logical
        pooled = self.view(1, c)
        self.P.simple(L.OP_GLOBAL_MEAN, name + ".avgpool", x, pooled)
        h = self.conv_bn(name + ".fc1", prefix + ".fc1", None, pooled, 1, L.ACT_SILU)
        g = self.conv_bn(name + ".fc2", prefix + ".fc2", None, h, 1, L.ACT_SIGMOID)
        out = self.view(self.hw(x), c)
        self.P.simple(L.OP_SCALE, name + ".scale", x, out, res=g)
        return out

    def finish(self, x: View, fcw: np.ndarray, fcb: np.ndarray) -> Plan:
        self.P.mean_fc(x, fcw, fcb)
        self.P.meta.update(num_classes=int(fcw.shape[0]), in_size=self.S)
        return self.P


def build_resnet18_plan(state_dict: dict, in_size: int = 64, mean: float = 0.18, std: float = 0.34) -> Plan:
    """torchvision resnet.py: conv1 7x7/2 + bn + relu, maxpool 3x3/2, 4 stages of 2 BasicBlocks
    (out = relu(bn2(conv2(relu(bn1(conv1(x))))) + identity), identity = 1x1/2 conv + bn where the shape changes), avgpool, fc."""
    B = _Builder(state_dict, in_size, mean, std)
    sd = B.sd
    x = B.conv_bn("conv1", "conv1", "bn1", B.img, 2, L.ACT_RELU, stem=True)
    y = B.view((B.hw(x) + 2 - 3) // 2 + 1, x.logical)
    B.P.simple(L.OP_MAXPOOL, "maxpool", x, y, ksize=3, stride=2)
    x = y
    for li in range(1, 5):
        bi = 0
        while f"layer{li}.{bi}.conv1.weight" in sd:
            pre = f"layer{li}.{bi}"
            stride = 2 if (li > 1 and bi == 0) else 1
            idt = x
            if pre + ".downsample.0.weight" in sd:
                idt = B.conv_bn(pre + ".downsample", pre + ".downsample.0", pre + ".downsample.1", x, stride, L.ACT_NONE)
            t = B.conv_bn(pre + ".conv1", pre + ".conv1", pre + ".bn1", x, stride, L.ACT_RELU)
            x = B.conv_bn(pre + ".conv2", pre + ".conv2", pre + ".bn2", t, 1, L.ACT_RELU, res=idt, flags=L.OPF_RES_BEFORE_ACT)
            bi += 1
    return B.finish(x, sd["fc.weight"], sd["fc.bias"])


def build_mobilenetv2_plan(state_dict: dict, in_size: int = 64, mean: float = 0.18, std: float = 0.34) -> Plan:
    """torchvision mobilenetv2.py: 3x3/2 stem, 17 InvertedResidual blocks ([1x1 expand + ReLU6], dw 3x3 + ReLU6, 1x1 project,
    + x when stride 1 and the width is unchanged), 1x1 to 1280 + ReLU6, mean, Linear."""
    B = _Builder(state_dict, in_size, mean, std)
    sd = B.sd
    x = B.conv_bn("features.0", "features.0.0", "features.0.1", B.img, 2, L.ACT_RELU6, stem=True)
    i = 1
    while f"features.{i}.conv.0.0.weight" in sd:
        pre = f"features.{i}.conv"
        inp = x
        # layers of the block: Conv2dNormActivation items have sub-index .0/.1; the final projection is a bare Conv2d + BN
        idx, t = 0, x
        while f"{pre}.{idx}.0.weight" in sd:
            w = sd[f"{pre}.{idx}.0.weight"]
            if w.shape[1] == 1 and w.shape[2] == 3:               # depthwise 3x3
                stride = _dw_stride_mbv2(i)
                t = B.dw_bn(f"features.{i}.dw", f"{pre}.{idx}.0", f"{pre}.{idx}.1", t, stride, L.ACT_RELU6)
            else:                                                  # 1x1 expansion
                t = B.conv_bn(f"features.{i}.expand", f"{pre}.{idx}.0", f"{pre}.{idx}.1", t, 1, L.ACT_RELU6)
            idx += 1
        proj_w = sd[f"{pre}.{idx}.weight"]
        use_res = _dw_stride_mbv2(i) == 1 and proj_w.shape[0] == inp.logical
        x = B.conv_bn(f"features.{i}.project", f"{pre}.{idx}", f"{pre}.{idx + 1}", t, 1, L.ACT_NONE, res=inp if use_res else None)
        i += 1
    x = B.conv_bn(f"features.{i}", f"features.{i}.0", f"features.{i}.1", x, 1, L.ACT_RELU6)
    return B.finish(x, sd["classifier.1.weight"], sd["classifier.1.bias"])


_MBV2_STRIDES = None


def _dw_stride_mbv2(block_index: int) -> int:
    """stride of the depthwise conv of features[block_index] (inverted_residual_setting of mobilenetv2.py: t, c, n, s)"""
    global _MBV2_STRIDES
    if _MBV2_STRIDES is None:
        s_list = []
        for t, c, n, s in [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]:
            for k in range(n):
                s_list.append(s if k == 0 else 1)
        _MBV2_STRIDES = s_list
    return _MBV2_STRIDES[block_index - 1]


def build_efficientnet_b0_plan(state_dict: dict, in_size: int = 64, mean: float = 0.18, std: float = 0.34) -> Plan:
    """torchvision efficientnet.py (b0): 3x3/2 stem + SiLU, MBConv blocks ([1x1 expand + SiLU], dw kxk + SiLU, SqueezeExcitation,
    1x1 project, + x when stride 1 and the width is unchanged; stochastic depth is the identity in eval), 1x1 to 1280 + SiLU,
    mean, Linear."""
    B = _Builder(state_dict, in_size, mean, std)
    sd = B.sd
    x = B.conv_bn("features.0", "features.0.0", "features.0.1", B.img, 2, L.ACT_SILU, stem=True)
    # MBConvConfig(expand_ratio, kernel, stride, input_channels, out_channels, num_layers) of efficientnet_b0
    strides = {1: 1, 2: 2, 3: 2, 4: 2, 5: 1, 6: 2, 7: 1}
    si = 1
    while f"features.{si}.0.block.0.0.weight" in sd:
        bi = 0
        while f"features.{si}.{bi}.block.0.0.weight" in sd:
            pre = f"features.{si}.{bi}.block"
            nm = f"features.{si}.{bi}"
            inp, t, idx = x, x, 0
            stride = strides[si] if bi == 0 else 1
            while True:
                if f"{pre}.{idx}.fc1.weight" in sd:                # SqueezeExcitation
                    t = B.squeeze_excite(nm + ".se", f"{pre}.{idx}", t)
                elif f"{pre}.{idx}.0.weight" in sd:
                    w = sd[f"{pre}.{idx}.0.weight"]
                    last = f"{pre}.{idx + 1}.0.weight" not in sd and f"{pre}.{idx + 1}.fc1.weight" not in sd
                    if w.shape[1] == 1 and w.shape[2] > 1:
                        t = B.dw_bn(nm + ".dw", f"{pre}.{idx}.0", f"{pre}.{idx}.1", t, stride, L.ACT_SILU)
                    elif last:                                     # projection: no activation
                        use_res = stride == 1 and w.shape[0] == inp.logical
                        t = B.conv_bn(nm + ".project", f"{pre}.{idx}.0", f"{pre}.{idx}.1", t, 1, L.ACT_NONE, res=inp if use_res else None)
                    else:
                        t = B.conv_bn(nm + ".expand", f"{pre}.{idx}.0", f"{pre}.{idx}.1", t, 1, L.ACT_SILU)
                else:
                    break
                idx += 1
            x = t
            bi += 1
        si += 1
    x = B.conv_bn(f"features.{si}", f"features.{si}.0", f"features.{si}.1", x, 1, L.ACT_SILU)
    return B.finish(x, sd["classifier.1.weight"], sd["classifier.1.bias"])


def torchvision_model(arch: str, num_classes: int):
    """The module ``build_classifier`` constructs (e2e.py:322-335), random-init."""
    import torch.nn as nn
    from torchvision import models
    if arch == "resnet18":
        m = models.resnet18(weights=None)
        m.fc = nn.Linear(m.fc.in_features, num_classes)
    elif arch == "efficientnet":
        m = models.efficientnet_b0(weights=None)
        m.classifier[1] = nn.Linear(m.classifier[1].in_features, num_classes)
    elif arch == "mobilenetv2":
        m = models.mobilenet_v2(weights=None)
        m.classifier[1] = nn.Linear(m.classifier[1].in_features, num_classes)
    elif arch == "shufflenetv2":
        m = models.shufflenet_v2_x1_0(weights=None)
        m.fc = nn.Linear(m.fc.in_features, num_classes)
    else:
        raise ValueError(f"Unknown architecture: {arch}")
    return m


PLAN_BUILDERS: Dict[str, Callable] = {
    "resnet18": build_resnet18_plan,
    "mobilenetv2": build_mobilenetv2_plan,
    "efficientnet": build_efficientnet_b0_plan,
}
