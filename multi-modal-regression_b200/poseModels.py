"""Drop-in for the reference's poseModels module (poseModels.py:1-53): plain FC(+BN+ReLU)
regressors without category logic.  Same layer shape as res_3layer, so model_3layer runs on the
same head kernels (one group)."""
import torch
from torch import nn
import torch.nn.functional as F

from binDeltaModels import _mlp3


class model_3layer(_mlp3):
    """fc3(relu(bn2(fc2(relu(bn1(fc1 x)))))) — poseModels.py:10-25"""

    def __init__(self, N0, N1, N2, N3):
        super().__init__(N0, N1, N2, N3)


class model_2layer(nn.Module):
    """tanh(fc2(relu(bn1(fc1 x)))) — poseModels.py:29-40"""

    def __init__(self, N0, N1, N2):
        super().__init__()
        self.fc1 = nn.Linear(N0, N1, bias=False)
        self.bn1 = nn.BatchNorm1d(N1)
        self.fc2 = nn.Linear(N1, N2)

    def forward(self, x):
        x = F.relu(self.bn1(self.fc1(x)))
        return torch.tanh(self.fc2(x))


class model_1layer(nn.Module):
    """poseModels.py:44-52"""

    def __init__(self, N0, N1):
        super().__init__()
        self.fc = nn.Linear(N0, N1)

    def forward(self, x):
        return self.fc(x)
