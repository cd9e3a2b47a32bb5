"""PeerLearningNet with the reference's surface (model/methods/PeerLearningNet.py:8-20): two copies of a base model built
through the registry (``config.base_model.name`` — BCNN in configs/PeerLearning_BCNN_S{1,2}.yaml), the second with a freshly
initialised classifier; ``forward`` returns both logit tensors.  The base model is whatever ``MODEL`` holds under that name,
i.e. the B200-native BCNN / CBCNN / MPN, so this caller of the hot path inherits the kernels unchanged."""
import copy

import torch.nn as nn

from ..registry import MODEL
from ..utils import initialize_weights


@MODEL.register
class PeerLearningNet(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.base_model = MODEL.get(config.base_model.name)(config.base_model)     # PeerLearningNet.py:13
        self.base_model2 = copy.deepcopy(self.base_model)                           # :14
        self.base_model2.classifier.apply(initialize_weights)                       # :15

    def forward(self, x):
        return self.base_model(x), self.base_model2(x)                              # :17-20
