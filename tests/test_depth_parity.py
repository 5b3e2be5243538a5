"""Parity AT DEPTH (-m gpu): the nets BASELINE.json's configs benchmark - 10 x 128 at batch 256 (config 2), 20 x 256
(configs 3 / 4), 40 x 256 (config 5) - through the C ABI against the oracle at both precisions, plus a 10-block
batch-norm net imported from an ONNX file the way infer::B200::load takes it.

Tolerances (the north star's: "policy/value outputs within a stated fp tolerance of the reference TensorRT FP32 path",
src/infer/trt.cc:144-161 = fp32 I/O, TF32 allowed):
    decoded policy probabilities  max |p - p_fp32|      <= 2e-2
    per-row KL(p_fp32 || p)                              <= 1e-3
    win rate, draw rate           max |v - v_fp32|      <= 1e-2
and against the bf16-emulating oracle (same rounding points; only the fp32 accumulation order differs, which can flip a
bf16 rounding once in a while, and every flip is carried through the layers that follow):
    logits                        max |l - l_bf16|      <= TOL_LOGIT_BF16_PER_RMS * max(1, rms of the fp32 logits) * layers / 21
    win rate, draw rate           max |v - v_bf16|      <= TOL_VALUE_BF16_AT_21_LAYERS * layers / 21
The measured values are written to gpurun_out/drift_table.json (DESIGN.md §4 quotes them).  Seeds are chosen so that the
value head is alive (a random-init 1-channel value conv is dead for about every second seed: then win / draw are
constants and the value tolerance would be vacuous); the test asserts it."""
import json
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "nshogi-engine_b200", "host")

TOL_PROB_VS_FP32 = 2e-2
TOL_KL_VS_FP32 = 1e-3
TOL_VALUE_VS_FP32 = 1e-2
TOL_LOGIT_BF16_PER_RMS = 3e-2     # at 21 layers and logits of rms <= 1 (the bound test_full_size_batch_invariance uses)
TOL_VALUE_BF16_AT_21_LAYERS = 3e-3   # win / draw vs the bf16-emulating oracle, scaled by layers / 21 like the logits

# (channels, blocks, weight seed, n, NSB_TRUNK128, slots)
CASES = [
    (128, 10, 1, 256, "classic", 1),   # config 2 on the one-CTA-per-SM kernel
    (128, 10, 1, 256, "duo", 4),       # config 2 on the kernel the bench's pipeline launches
    (128, 10, 1, 5, "duo", 2),         # odd n: a CTA with one idle position
    (256, 20, 1234, 1, None, 1),       # configs 3 / 4: CTA-pair kernel, the second CTA of the pair idle
    (256, 20, 1234, 5, None, 1),
    (256, 20, 1234, 64, None, 3),
    (256, 40, 5, 4, None, 1),          # config 5
    (256, 40, 5, 7, None, 1),
]


def _record(name, m):
    out = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(out):
        return
    path = os.path.join(out, "drift_table.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    table[name] = m
    json.dump(table, open(path, "w"), indent=1, sort_keys=True)


def _check(m, name):
    _record(name, m)
    assert m["prob_vs_fp32"] <= TOL_PROB_VS_FP32, (name, m)
    assert m["kl_vs_fp32"] <= TOL_KL_VS_FP32, (name, m)
    assert m["win_vs_fp32"] <= TOL_VALUE_VS_FP32 and m["draw_vs_fp32"] <= TOL_VALUE_VS_FP32, (name, m)
    tol_logit = TOL_LOGIT_BF16_PER_RMS * max(1.0, m["logit_rms_fp32"]) * m["layers"] / 21.0
    assert m["logit_vs_bf16"] <= tol_logit, (name, tol_logit, m)
    assert m["value_vs_bf16"] <= TOL_VALUE_BF16_AT_21_LAYERS * m["layers"] / 21.0, (name, m)


@pytest.mark.parametrize("channels,blocks,seed,n,kernel,slots", CASES)
def test_depth_parity_vs_oracle(nb, orc, synth, monkeypatch, channels, blocks, seed, n, kernel, slots):
    if kernel:
        monkeypatch.setenv("NSB_TRUNK128", kernel)
    else:
        monkeypatch.delenv("NSB_TRUNK128", raising=False)
    monkeypatch.delenv("NSB_TRUNK256", raising=False)
    desc = nb.net_desc(channels, blocks)
    blob = nb.random_blob(desc, seed)
    pos = synth.random_positions(n, seed=2024)
    fb = orc.pack(pos)
    off, idx = synth.random_legal_moves(n, seed=3, edge_rows=False)
    policy = np.zeros((n, nb.POLICY_SIZE), dtype=np.float32)
    win, draw = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32)
    legal = np.zeros(int(off[-1]), dtype=np.float32)
    w2, d2 = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32)
    with nb.Context(desc, batch_max=n, slots=slots, blob=blob) as ctx:
        name = ctx.trunk_kernel_name()
        ctx.eval_async(0, fb, n, policy, win, draw)                     # the Infer contract: dense logits
        ctx.await_(0)
        ctx.eval_positions_decode_async(slots - 1, pos, n, off, idx, nb.DECODE_PROBS, legal, w2, d2, None)   # the fused path
        ctx.await_(slots - 1)
    if channels == 256:
        assert name.startswith("trunk_pair_kernel")
    else:
        assert name.startswith("trunk_duo_kernel" if kernel == "duo" else "trunk_fused_kernel<128>")
    m = helpers.drift_metrics(nb, orc, desc, blob, fb, n, (policy, win, draw), off, idx)
    if n >= 4:
        assert m["win_spread_fp32"] > 1e-2, ("dead value head: pick another seed", m)
    _check(m, f"{blocks}x{channels}@{n}" + (f":{kernel}" if kernel else ""))
    # the fused path (stage 1 + decode in the same launch) returns the decode of those very logits
    want, _ = orc.decode(policy, win, draw, off, idx, nb.DECODE_PROBS)
    assert np.allclose(legal, want, rtol=1e-6, atol=1e-9) and np.array_equal(w2, win) and np.array_equal(d2, draw)


def _trained_like_net(C, blocks, seed):
    """A ResNet with batch-norm layers whose running statistics and affine parameters are far from the identity
    (tests/test_weights_io.make_model draws them), in inference mode: what a trained model file looks like to a loader."""
    import torch
    from test_weights_io import make_model
    net = make_model(C, blocks, seed=seed)
    with torch.no_grad():   # heads with the output range of a trained net: logits of rms ~2, value spread > 0.1
        net.policy.conv.weight.mul_(5.0)
        net.value.fc1.weight.mul_(2.0)
        net.value.fc2.weight.mul_(12.0)
    return net


def test_ten_block_batchnorm_net_through_onnx_load(pkg, nb, orc, synth, tmp_path):
    """10 x 128 with batch-norm -> ONNX (torch.onnx.export, the reference's model format, src/context.h:93) ->
    onnx_import.h (what infer::B200::load runs; nsb_host_unit --onnx-blob exposes it) -> the executor, against the
    PyTorch model's own fp32 inference forward: BN fold + bf16 weights and activations over 21 layers, north-star
    tolerances.  Then the same file through infer::B200::load itself (nsb_host_bench --weights)."""
    import subprocess
    import warnings

    import torch

    C, blocks, n = 128, 10, 64
    net = _trained_like_net(C, blocks, seed=3)
    onnx_path = str(tmp_path / "net10.onnx")

    class Wrap(torch.nn.Module):       # the reference's tensor contract: policy [B,2187], value [B,1], draw [B,1]
        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def forward(self, x):
            p, w, d = self.inner(x)
            return p, w.reshape(-1, 1), d.reshape(-1, 1)

    from torch.onnx._internal.torchscript_exporter import onnx_proto_utils
    onnx_proto_utils._add_onnxscript_fn = lambda proto, custom_opsets: proto   # needs the absent `onnx` package; nothing to attach
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.onnx.export(Wrap(net).eval(), (torch.zeros(2, 86, 9, 9),), onnx_path, dynamo=False, input_names=["input"],
                          output_names=["policy", "value", "draw"], opset_version=17,
                          dynamic_axes={"input": {0: "batch"}, "policy": {0: "batch"}, "value": {0: "batch"}, "draw": {0: "batch"}})
    subprocess.check_call(["make", "-C", HOST, "-s", "all"])
    raw = str(tmp_path / "net10.blob")
    out = subprocess.run([os.path.join(HOST, "nsb_host_unit"), "--onnx-blob", onnx_path, raw], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.startswith("ok 86 128 10 256"), out.stdout + out.stderr
    blob = np.fromfile(raw, dtype=np.float32, offset=16)
    desc = nb.net_desc(C, blocks)
    pos = synth.random_positions(n, seed=77)
    fb = orc.pack(pos)
    planes = orc.expand(fb, n)
    with torch.no_grad():
        tp, tw, td = (t.numpy() for t in net(torch.from_numpy(planes).reshape(-1, 86, 9, 9)))
    assert np.ptp(tw) > 1e-2, "dead value head: pick another seed"
    policy = np.zeros((n, nb.POLICY_SIZE), dtype=np.float32)
    win, draw = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32)
    with nb.Context(desc, batch_max=n, slots=2, blob=blob) as ctx:
        ctx.eval_async(0, fb, n, policy, win, draw)
        ctx.await_(0)
    off, idx = synth.random_legal_moves(n, seed=2, edge_rows=False)
    pg, _ = orc.decode(policy, win, draw, off, idx, nb.DECODE_PROBS)
    pt, _ = orc.decode(tp, tw, td, off, idx, nb.DECODE_PROBS)
    m = {"layers": 22, "n": n, "prob_vs_fp32": float(np.max(np.abs(pg - pt))), "win_vs_fp32": float(np.max(np.abs(win - tw))),
         "draw_vs_fp32": float(np.max(np.abs(draw - td))), "logit_vs_fp32": float(np.max(np.abs(policy - tp))),
         "logit_rms_fp32": float(np.sqrt(np.mean(tp.astype(np.float64) ** 2))), "win_spread_fp32": float(np.ptp(tw)),
         "reference": "PyTorch fp32 forward of the batch-norm model (weights NOT bf16-exact: includes weight rounding)"}
    _record("10x128-batchnorm-onnx@64", m)
    assert m["prob_vs_fp32"] <= TOL_PROB_VS_FP32 and m["win_vs_fp32"] <= TOL_VALUE_VS_FP32 and m["draw_vs_fp32"] <= TOL_VALUE_VS_FP32, m
    out = subprocess.run([os.path.join(HOST, "nsb_host_bench"), "--selfcheck", "--repeat", "10", "--batch", "64", "--weights", onnx_path],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["ok"] and line["net"] == "10x128"
