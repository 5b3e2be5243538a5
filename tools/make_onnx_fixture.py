#!/usr/bin/env python
"""Generates tests/golden/resnet_torch_export.onnx + resnet_torch_export_state.npz: a small ResNet with the tensor
contract of reference src/infer/trt.cc:144-150,193-227 (input [B,86,9,9] -> policy [B,2187], value, draw), with
batch-norm layers and non-trivial running statistics, exported by torch.onnx.export (legacy TorchScript exporter,
which serialises the ModelProto in C++).  The `onnx` Python package is not installed here; the exporter only
needs it for a final pass that attaches onnxscript functions, which this model has none of, so that pass is
stubbed.  The file pins nshogi-engine_b200/onnx_io.py's reader against a real exporter's bytes.  (In eval mode the
exporter folds conv + batch-norm itself; graphs that still hold BatchNormalization nodes are covered by a graph
tests/test_onnx_io.py assembles.)  The .npz holds the PyTorch state dict, a random input and the model's outputs."""
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


class Block(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv1 = nn.Conv2d(c, c, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(c)
        self.conv2 = nn.Conv2d(c, c, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(c)

    def forward(self, x):
        h = torch.relu(self.bn1(self.conv1(x)))
        return torch.relu(x + self.bn2(self.conv2(h)))


class Stem(nn.Module):
    def __init__(self, cin, c):
        super().__init__()
        self.conv = nn.Conv2d(cin, c, 3, padding=1, bias=False)
        self.bn = nn.BatchNorm2d(c)

    def forward(self, x):
        return torch.relu(self.bn(self.conv(x)))


class Policy(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, 27, 1)

    def forward(self, x):
        return torch.flatten(self.conv(x), 1)


class Value(nn.Module):
    def __init__(self, c, h):
        super().__init__()
        self.conv = nn.Conv2d(c, 1, 1, bias=False)
        self.bn = nn.BatchNorm2d(1)
        self.fc1 = nn.Linear(81, h)
        self.fc2 = nn.Linear(h, 2)

    def forward(self, x):
        v = torch.flatten(torch.relu(self.bn(self.conv(x))), 1)
        o = torch.sigmoid(self.fc2(torch.relu(self.fc1(v))))
        return o[:, 0:1], o[:, 1:2]


class Net(nn.Module):
    def __init__(self, cin=86, c=16, blocks=2, h=8):
        super().__init__()
        self.stem = Stem(cin, c)
        self.blocks = nn.ModuleList([Block(c) for _ in range(blocks)])
        self.policy = Policy(c)
        self.value = Value(c, h)

    def forward(self, x):
        x = self.stem(x)
        for b in self.blocks:
            x = b(x)
        value, draw = self.value(x)
        return self.policy(x), value, draw


def main():
    torch.manual_seed(20240203)
    net = Net()
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.normal_(0, 0.2)
                m.running_mean.normal_(0, 0.3)
                m.running_var.uniform_(0.5, 2.0)
    net.eval()
    from torch.onnx._internal.torchscript_exporter import onnx_proto_utils
    onnx_proto_utils._add_onnxscript_fn = lambda proto, custom_opsets: proto  # needs `onnx`; nothing to attach here
    path = os.path.join(OUT, "resnet_torch_export.onnx")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.onnx.export(net, (torch.zeros(2, 86, 9, 9),), path, dynamo=False, input_names=["input"],
                          output_names=["policy", "value", "draw"], opset_version=17,
                          dynamic_axes={"input": {0: "batch"}, "policy": {0: "batch"}, "value": {0: "batch"}, "draw": {0: "batch"}})
    x = torch.randn(3, 86, 9, 9)
    with torch.no_grad():
        p, v, d = net(x)
    np.savez_compressed(os.path.join(OUT, "resnet_torch_export_state.npz"),
                        x=x.numpy(), policy=p.numpy(), value=v.numpy(), draw=d.numpy(),
                        **{k: t.numpy() for k, t in net.state_dict().items() if "num_batches" not in k})
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    sys.exit(main())
