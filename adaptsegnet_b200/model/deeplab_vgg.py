"""DeepLab-VGG16 variant behind the reference's ``model/deeplab_vgg.py`` surface (BASELINE config 4).

The reference file does not construct on Python 3 (`range(23)+range(24,30)`, model/deeplab_vgg.py:34 -- SURVEY.md Q10);
this restates it with the one-line fix.  Trunk = torchvision's VGG-16 feature extractor with pool4/pool5 removed, conv5
dilated by 2, and fc6/fc7 as dilated 3x3 convolutions to 1024 channels (:24-46): unchanged PyTorch modules, timed, not
rewritten.  Hot path = the ASPP classifier on the 1024-channel features, on the same sm_100a head kernels as
DeeplabMulti; the reference's `Classifier_Module.forward` returns from inside its loop (:17-21), so only branches 0 and 1
(dilation 6, 12) are summed although all four own weights -- `active_branches=2` reproduces that (Q9).
State-dict keys (`features.N.*`, `classifier.conv2d_list.i.*`) equal the reference's."""
from __future__ import annotations

import torch.nn as nn

from .deeplab_multi import Classifier_Module as _Head

_RATES = (6, 12, 18, 24)


class Classifier_Module(_Head):
    """model/deeplab_vgg.py:6-21: four branches are built, the first two are summed"""

    def __init__(self, dims_in, dilation_series, padding_series, num_classes):
        super().__init__(dims_in, dilation_series, padding_series, num_classes, active_branches=min(2, len(dilation_series)))


class DeeplabVGG(nn.Module):
    def __init__(self, num_classes, vgg16_caffe_path=None, pretrained=False):
        super().__init__()
        from torchvision import models
        vgg = models.vgg16()
        if pretrained:
            import torch
            vgg.load_state_dict(torch.load(vgg16_caffe_path))
        features = list(vgg.features.children())
        features = [features[i] for i in list(range(23)) + list(range(24, 30))]   # pool4 / pool5 removed (Q10 fix)
        for i in (23, 25, 27):                                                     # conv5_x: dilation 2
            features[i].dilation = (2, 2)
            features[i].padding = (2, 2)
        fc6 = nn.Conv2d(512, 1024, kernel_size=3, padding=4, dilation=4)
        fc7 = nn.Conv2d(1024, 1024, kernel_size=3, padding=4, dilation=4)
        self.features = nn.Sequential(*(features + [fc6, nn.ReLU(inplace=True), fc7, nn.ReLU(inplace=True)]))
        self.classifier = Classifier_Module(1024, list(_RATES), list(_RATES), num_classes)

    def forward(self, x):
        """-> (N, num_classes, H/8, W/8) logits; the caller upsamples (evaluate_cityscapes.py:164-166)"""
        return self.classifier(self.features(x))

    def low_res_logits(self, x):
        """(None, logits): same shape of result as ResNetMulti.low_res_logits, for the single-level trainer"""
        return None, self.forward(x)

    def optim_parameters(self, args):
        return self.parameters()
