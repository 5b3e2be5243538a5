"""Weight import (the part that does not need ONNX tooling): a PyTorch ResNet with
batch-norm layers -> BN folded into the canonical blob -> NSBW file -> the executor, checked against
the PyTorch model's own inference-mode forward."""
import json
import os
import subprocess

import numpy as np
import pytest

import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "nshogi-engine_b200", "host")


def make_model(C, blocks, H=256, seed=0):
    import torch
    import torch.nn as nn

    torch.manual_seed(seed)

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1, self.bn1 = nn.Conv2d(C, C, 3, padding=1, bias=False), nn.BatchNorm2d(C)
            self.conv2, self.bn2 = nn.Conv2d(C, C, 3, padding=1, bias=False), nn.BatchNorm2d(C)

        def forward(self, x):
            y = torch.relu(self.bn1(self.conv1(x)))
            return torch.relu(self.bn2(self.conv2(y)) + x)

    class Stem(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv, self.bn = nn.Conv2d(86, C, 3, padding=1, bias=False), nn.BatchNorm2d(C)

    class Policy(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = nn.Conv2d(C, 27, 1)

    class Value(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv, self.bn = nn.Conv2d(C, 1, 1, bias=False), nn.BatchNorm2d(1)
            self.fc1, self.fc2 = nn.Linear(81, H), nn.Linear(H, 2)

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.stem, self.blocks = Stem(), nn.ModuleList([Block() for _ in range(blocks)])
            self.policy, self.value = Policy(), Value()

        def forward(self, x):
            x = torch.relu(self.stem.bn(self.stem.conv(x)))
            for b in self.blocks:
                x = b(x)
            pol = self.policy.conv(x).reshape(-1, 2187)
            v = torch.relu(self.value.bn(self.value.conv(x))).reshape(-1, 81)
            o = torch.sigmoid(self.value.fc2(torch.relu(self.value.fc1(v))))
            return pol, o[:, 0], o[:, 1]

    net = Net()
    with torch.no_grad():   # non-trivial running statistics and affine parameters; damped second convs
        for m in net.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.2)
                m.running_var.uniform_(0.5, 1.5)
                m.weight.uniform_(0.7, 1.3)
                m.bias.normal_(0, 0.1)
        for b in net.blocks:
            b.conv2.weight.mul_(0.25)
    return net.eval()


def test_bn_fold_matches_torch_eval_forward(pkg, nb, orc, synth):
    import torch

    wio = pkg.weights_io
    C, blocks = 128, 2
    net = make_model(C, blocks)
    blob = wio.blob_from_state(net.state_dict(), C, blocks)
    desc = nb.net_desc(C, blocks)
    assert blob.size == nb.random_blob(desc, 1).size
    pos = synth.random_positions(5, seed=3)
    planes = orc.expand(orc.pack(pos), 5)
    with torch.no_grad():
        tp, tw, td = net(torch.from_numpy(planes).reshape(-1, 86, 9, 9))
    fp, fw, fd = helpers.forward_torch(desc, blob, planes)          # folded blob through the canonical forward
    assert np.max(np.abs(fp - tp.numpy())) < 2e-4 and np.max(np.abs(fw - tw.numpy())) < 1e-5
    op, ow, od = orc.forward(desc, blob, planes, emulate_bf16=False)  # and through the oracle
    assert np.max(np.abs(op - tp.numpy())) < 2e-4 and np.max(np.abs(od - td.numpy())) < 1e-5


def test_nsbw_round_trip_and_errors(pkg, tmp_path):
    wio = pkg.weights_io
    net = make_model(128, 1)
    blob = wio.blob_from_state(net.state_dict(), 128, 1)
    path = str(tmp_path / "net.nsbw")
    wio.write_nsbw(path, blob, 128, 1)
    meta, back = wio.read_nsbw(path)
    assert meta == {"channels": 128, "blocks": 1, "value_hidden": 256, "in_channels": 86}
    assert np.array_equal(back, blob)
    state = dict(net.state_dict())
    del state["blocks.0.bn1.running_var"]
    with pytest.raises(KeyError):
        wio.blob_from_state(state, 128, 1)
    with pytest.raises(ValueError):
        wio.blob_from_state(net.state_dict(), 256, 1)
    open(path, "wb").write(b"nope")
    with pytest.raises(ValueError):
        wio.read_nsbw(path)


@pytest.mark.gpu
@pytest.mark.parametrize("C", [128, 256])
def test_imported_net_matches_torch_on_gpu(pkg, nb, orc, synth, C, tmp_path):
    """The imported (BN-folded) net on the executor vs the PyTorch model's fp32 inference forward."""
    import torch

    wio = pkg.weights_io
    blocks = 3
    net = make_model(C, blocks, seed=C)
    blob = wio.blob_from_state(net.state_dict(), C, blocks)
    desc = nb.net_desc(C, blocks)
    n = 16
    pos = synth.random_positions(n, seed=8)
    fb = orc.pack(pos)
    planes = orc.expand(fb, n)
    with torch.no_grad():
        tp, tw, td = (t.numpy() for t in net(torch.from_numpy(planes).reshape(-1, 86, 9, 9)))
    policy = np.zeros((n, nb.POLICY_SIZE), dtype=np.float32)
    win, draw = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32)
    with nb.Context(desc, batch_max=n, blob=blob) as ctx:
        ctx.eval_async(0, fb, n, policy, win, draw)
        ctx.await_(0)
    off, idx = synth.random_legal_moves(n, seed=2, edge_rows=False)
    pg, _ = orc.decode(policy, win, draw, off, idx, nb.DECODE_PROBS)
    pt, _ = orc.decode(tp, tw, td, off, idx, nb.DECODE_PROBS)
    assert np.max(np.abs(pg - pt)) < 2e-2                     # SURVEY.md §8c tolerance (bf16 in, fp32 accumulate)
    assert np.max(np.abs(win - tw)) < 1e-2 and np.max(np.abs(draw - td)) < 1e-2
    # the same file through the C++ loader (infer::B200::load) and the reference-shaped micro-benchmark
    path = str(tmp_path / "net.nsbw")
    wio.write_nsbw(path, blob, C, blocks)
    subprocess.check_call(["make", "-C", HOST, "-s", "all"])
    out = subprocess.run([os.path.join(HOST, "nsb_host_bench"), "--selfcheck", "--repeat", "20", "--channels", str(C),
                          "--blocks", str(blocks), "--batch", "64", "--weights", path], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert json.loads(out.stdout.strip().splitlines()[-1])["ok"]
    # ... and as the reference's own model format: an ONNX file handed to load(), as TensorRT::load takes it
    # (trt.cc:109-232).  The executor is constructed for the default 10 x 128 net; the file decides the shape.
    onnx_path = str(tmp_path / "net.onnx")
    pkg.onnx_io.write_onnx(onnx_path, blob, C, blocks)
    out = subprocess.run([os.path.join(HOST, "nsb_host_bench"), "--selfcheck", "--repeat", "20", "--batch", "64",
                          "--weights", onnx_path], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["ok"] and line["net"] == f"{blocks}x{C}"
