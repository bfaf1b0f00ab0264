"""ORACLE (test infrastructure, not product code).

Plain-PyTorch restatement of the reference's Cnn14 ConvBlock stack (mst/panns.py:27-85,
126-209): conv3x3(no bias) -> BatchNorm2d -> ReLU, twice, average pooling; six blocks; mean over
bins, max + mean over frames, linear head.  Same sub-module names as the reference so that state
dicts are interchangeable.  Used as the float32/float64 reference of the tensor-core path."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class OracleConvBlock(nn.Module):
    def __init__(self, in_channels, out_channels, use_batchnorm=True):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, 1, bias=False)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, 1, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(out_channels) if use_batchnorm else nn.Identity()
        self.bn2 = nn.BatchNorm2d(out_channels) if use_batchnorm else nn.Identity()

    def forward(self, x, pool_size):
        x = F.relu(self.bn1(self.conv1(x)))
        x = F.relu(self.bn2(self.conv2(x)))
        return F.avg_pool2d(x, pool_size)


class OracleCnn14(nn.Module):
    POOLS = [(2, 2), (4, 4), (4, 2), (4, 2), (4, 2), (2, 2)]

    def __init__(self, num_classes, n_inputs=1, use_batchnorm=True):
        super().__init__()
        ch = [n_inputs, 64, 128, 256, 512, 1024, 2048]
        for i in range(6):
            setattr(self, f"conv_block{i + 1}", OracleConvBlock(ch[i], ch[i + 1], use_batchnorm))
        self.fc = nn.Linear(2048, num_classes)

    def forward(self, x):
        for i, p in enumerate(self.POOLS):
            x = getattr(self, f"conv_block{i + 1}")(x, p)
        x = torch.mean(x, dim=2)
        x1, _ = torch.max(x, dim=2)
        return self.fc(x1 + torch.mean(x, dim=2))
