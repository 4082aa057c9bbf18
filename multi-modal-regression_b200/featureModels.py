# -*- coding: utf-8 -*-
"""Feature trunks with the reference's constructor strings, NameError behaviour and attribute names
(`features`, `avgpool`, `classifier`, `original_model` — they are state_dict keys of saved
checkpoints; reference featureModels.py:11-67).  Stock torchvision / cuDNN: outside the kernel claim.
Pretrained weights are required, as in the reference (`pretrained=True`); a box without network
access (the benchmark box) opts into random initialisation with BDPOSE_ALLOW_RANDOM_TRUNK=1."""
import os
import warnings

from torch import nn

_RESNET_DEPTH = {'layer2': (6, 28), 'layer3': (7, 14), 'layer4': (8, 7)}   # children kept, pool size


def _torchvision(name):
    import torchvision.models as tv
    ctor = getattr(tv, name)
    try:
        return ctor(weights='DEFAULT')
    except Exception as e:
        if os.environ.get('BDPOSE_ALLOW_RANDOM_TRUNK', '0') != '1':
            raise RuntimeError("featureModels: pretrained %s weights could not be loaded (%s); set "
                               "BDPOSE_ALLOW_RANDOM_TRUNK=1 to train from a randomly initialised "
                               "trunk" % (name, e)) from e
        warnings.warn("featureModels: pretrained %s weights unavailable (%s) - RANDOM trunk "
                      "initialisation (BDPOSE_ALLOW_RANDOM_TRUNK=1)" % (name, e))
        return ctor(weights=None)


class resnet_model(nn.Module):
    def __init__(self, model_type='resnet50', layer_type='layer4'):
        super().__init__()
        if model_type not in ('resnet50', 'resnet101'):
            raise NameError('Unknown model_type passed')
        if layer_type not in _RESNET_DEPTH:
            raise NameError('Uknown layer_type passed')
        keep, pool = _RESNET_DEPTH[layer_type]
        trunk = list(_torchvision(model_type).children())
        self.features = nn.Sequential(*trunk[:keep])
        self.avgpool = nn.AvgPool2d(pool, stride=1)

    def forward(self, x):
        return self.avgpool(self.features(x)).flatten(1)


class vgg_model(nn.Module):
    def __init__(self, model_type='vgg13', layer_type='fc6'):
        super().__init__()
        if model_type not in ('vgg13', 'vgg16'):
            raise NameError('Unknown model_type passed')
        if layer_type not in ('fc6', 'fc7'):
            raise NameError('Uknown layer_type passed')
        self.original_model = _torchvision(model_type + '_bn')
        self.features = self.original_model.features
        fc = list(self.original_model.classifier.children())
        self.classifier = nn.Sequential(*(fc[:2] if layer_type == 'fc6' else fc[:-2]))

    def forward(self, x):
        return self.classifier(self.features(x).flatten(1))
