"""ONNX import / export: the reference's model files (reference src/infer/trt.cc:109-232: input
`input` [B,86,9,9]; outputs `policy`, `value`, `draw`) <-> the canonical blob, without the `onnx` package.
The reader is pinned against bytes written by torch.onnx.export (tests/golden/resnet_torch_export.onnx,
tools/make_onnx_fixture.py)."""
import os

import numpy as np
import pytest


@pytest.fixture(scope="module")
def oi(pkg):
    return pkg.onnx_io


def test_reads_torch_exported_model(pkg, oi, nb, orc, golden_dir):
    """The exporter's file -> blob == the blob built from the PyTorch state dict (batch-norm folded by the exporter in
    fp32 vs by weights_io in fp64), and the oracle's forward of that blob reproduces the PyTorch model's outputs."""
    meta, blob = oi.read_onnx(os.path.join(golden_dir, "resnet_torch_export.onnx"))
    assert meta == {"channels": 16, "blocks": 2, "value_hidden": 8, "in_channels": 86}
    g = np.load(os.path.join(golden_dir, "resnet_torch_export_state.npz"))
    state = {k: g[k] for k in g.files if k not in ("x", "policy", "value", "draw")}
    want = pkg.weights_io.blob_from_state(state, 16, 2, value_hidden=8)
    assert blob.shape == want.shape
    assert np.allclose(blob, want, rtol=2e-5, atol=2e-6)
    desc = nb.net_desc(16, 2, value_hidden=8)
    policy, win, draw = orc.forward(desc, blob, g["x"].reshape(len(g["x"]), 86, 81), emulate_bf16=False)
    assert np.max(np.abs(policy - g["policy"])) < 2e-4
    assert np.max(np.abs(win - g["value"][:, 0])) < 1e-5 and np.max(np.abs(draw - g["draw"][:, 0])) < 1e-5


def test_write_then_read_is_exact(pkg, oi, nb, tmp_path):
    desc = nb.net_desc(128, 2)
    blob = nb.random_blob(desc, 42)
    path = str(tmp_path / "net.onnx")
    oi.write_onnx(path, blob, 128, 2)
    meta, back = oi.read_onnx(path)
    assert meta == {"channels": 128, "blocks": 2, "value_hidden": 256, "in_channels": 86}
    assert np.array_equal(back.view(np.uint32), blob.view(np.uint32))
    g = oi.parse_model(open(path, "rb").read())
    assert g.inputs == ["input"] and g.outputs == ["policy", "value", "draw"] and g.opset == 17


def _bn_graph(oi, rng, C=8, H=4, fused_fc2=True, extra=b""):
    """A graph with BatchNormalization nodes left in (not folded), MatMul + Add dense layers and the (value, draw)
    pair cut with Split, assembled from the module's own protobuf writers."""
    state, nodes, inits = {}, [], []

    def add_init(name, arr):
        inits.append(oi._ld(5, oi._tensor(name, np.asarray(arr, dtype=np.float32))))

    def conv_bn(prefix, x, cin, cout, k):
        w = rng.normal(0, 0.3, size=(cout, cin, k, k)).astype(np.float32)
        state[f"{prefix}.conv.weight" if prefix in ("stem", "value") else f"{prefix}.weight"] = w
        bn = {"weight": rng.uniform(0.5, 1.5, cout), "bias": rng.normal(0, 0.2, cout),
              "running_mean": rng.normal(0, 0.3, cout), "running_var": rng.uniform(0.5, 2.0, cout)}
        bn = {kk: v.astype(np.float32) for kk, v in bn.items()}
        bnp = {"stem": "stem.bn", "value": "value.bn"}.get(prefix, prefix.replace("conv", "bn"))
        for kk, v in bn.items():
            state[f"{bnp}.{kk}"] = v
        add_init(prefix + ".w", w)
        p = (k - 1) // 2
        nodes.append(oi._node("Conv", [x, prefix + ".w"], [prefix + ".c"],
                              oi._attr_ints("kernel_shape", [k, k]) + oi._attr_ints("pads", [p] * 4)))
        for kk in ("weight", "bias", "running_mean", "running_var"):
            add_init(f"{prefix}.bn.{kk}", bn[kk])
        eps_attr = oi._ld(5, oi._st(1, "epsilon") + oi._varint((2 << 3) | 5) + np.float32(1e-3).tobytes() + oi._vi(20, 1))
        nodes.append(oi._node("BatchNormalization", [prefix + ".c"] + [f"{prefix}.bn.{kk}" for kk in
                              ("weight", "bias", "running_mean", "running_var")], [prefix + ".n"], eps_attr))
        return prefix + ".n"

    x = conv_bn("stem", "input", 86, C, 3)
    nodes.append(oi._node("Relu", [x], ["x0"]))
    h = conv_bn("blocks.0.conv1", "x0", C, C, 3)
    nodes.append(oi._node("Relu", [h], ["h0"]))
    y = conv_bn("blocks.0.conv2", "h0", C, C, 3)
    nodes.append(oi._node("Add", [y, "x0"], ["s0"]))
    nodes.append(oi._node("Relu", ["s0"], ["x1"]))
    pw, pb = rng.normal(0, 0.3, size=(27, C, 1, 1)).astype(np.float32), rng.normal(0, 0.1, 27).astype(np.float32)
    state["policy.conv.weight"], state["policy.conv.bias"] = pw, pb
    add_init("p.w", pw)
    add_init("p.b", pb)
    nodes.append(oi._node("Conv", ["x1", "p.w", "p.b"], ["p.c"], oi._attr_ints("kernel_shape", [1, 1])))
    inits.append(oi._ld(5, oi._tensor("p.shape", np.asarray([-1, 2187], dtype=np.int64))))
    nodes.append(oi._node("Reshape", ["p.c", "p.shape"], ["policy"]))
    v = conv_bn("value", "x1", C, 1, 1)
    nodes.append(oi._node("Relu", [v], ["v.r"]))
    nodes.append(oi._node("Flatten", ["v.r"], ["v.f"], oi._attr_int("axis", 1)))
    w1, b1 = rng.normal(0, 0.2, size=(H, 81)).astype(np.float32), rng.normal(0, 0.1, H).astype(np.float32)
    w2, b2 = rng.normal(0, 0.2, size=(2, H)).astype(np.float32), rng.normal(0, 0.1, 2).astype(np.float32)
    state.update({"value.fc1.weight": w1, "value.fc1.bias": b1, "value.fc2.weight": w2, "value.fc2.bias": b2})
    add_init("fc1.wt", w1.T)            # MatMul takes [in, out]
    add_init("fc1.b", b1)
    nodes.append(oi._node("MatMul", ["v.f", "fc1.wt"], ["v.m"]))
    nodes.append(oi._node("Add", ["fc1.b", "v.m"], ["v.h"]))
    nodes.append(oi._node("Relu", ["v.h"], ["v.hr"]))
    if fused_fc2:
        add_init("fc2.w", w2)
        add_init("fc2.b", b2)
        nodes.append(oi._node("Gemm", ["v.hr", "fc2.w", "fc2.b"], ["v.o"], oi._attr_int("transB", 1)))
        nodes.append(oi._node("Sigmoid", ["v.o"], ["v.s"]))
        nodes.append(oi._node("Split", ["v.s"], ["value", "draw"], oi._attr_int("axis", 1)))
    graph = (b"".join(nodes) + extra + oi._st(2, "g") + b"".join(inits) + oi._ld(11, oi._value_info("input", ["N", 86, 9, 9])) +
             b"".join(oi._ld(12, oi._value_info(o, ["N", d])) for o, d in (("policy", 2187), ("value", 1), ("draw", 1))))
    model = oi._vi(1, 8) + oi._ld(7, graph) + oi._ld(8, oi._st(1, "") + oi._vi(2, 13))
    return model, state


def test_batchnorm_nodes_matmul_and_split(pkg, oi):
    rng = np.random.default_rng(5)
    model, state = _bn_graph(oi, rng)
    meta, blob = oi.blob_from_graph(oi.parse_model(model))
    assert meta == {"channels": 8, "blocks": 1, "value_hidden": 4, "in_channels": 86}
    want = pkg.weights_io.blob_from_state(state, 8, 1, value_hidden=4, eps=float(np.float32(1e-3)))   # the nodes' own epsilon (an fp32 attribute)
    assert np.array_equal(blob.view(np.uint32), want.view(np.uint32))
    other = pkg.weights_io.blob_from_state(state, 8, 1, value_hidden=4, eps=1e-5)
    assert not np.array_equal(blob, other)


def test_unsupported_graphs_fail_loudly(oi):
    rng = np.random.default_rng(6)
    # a node the executor would not run (an extra activation on a side branch)
    extra = oi._node("Relu", ["x1"], ["dangling"])
    model, _ = _bn_graph(oi, rng, extra=extra)
    with pytest.raises(ValueError, match="trunk output must feed"):
        oi.blob_from_graph(oi.parse_model(model))
    # missing outputs
    model, _ = _bn_graph(oi, rng, fused_fc2=False)
    with pytest.raises(ValueError, match="'value'"):
        oi.blob_from_graph(oi.parse_model(model))
    with pytest.raises(ValueError):
        oi.parse_model(b"\x0a\x03abc")
    # a 5x5 stem
    g = oi.parse_model(_bn_graph(oi, rng)[0])
    stem = [n for n in g.nodes if n.op == "Conv" and n.inputs[0] == "input"][0]
    g.initializers[stem.inputs[1]] = np.zeros((8, 86, 5, 5), dtype=np.float32)
    with pytest.raises(ValueError, match="3x3"):
        oi.blob_from_graph(g)


def _cpp_blob(path, tmp_path):
    import subprocess
    unit = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nshogi-engine_b200", "host", "nsb_host_unit")
    out = str(tmp_path / "blob.bin")
    r = subprocess.run([unit, "--onnx-blob", path, out], capture_output=True, text=True, errors="replace", timeout=60)
    if r.returncode != 0:
        return r.returncode, r.stdout, None, None
    raw = open(out, "rb").read()
    head = np.frombuffer(raw[:16], dtype="<i4")
    return 0, r.stdout, {"in_channels": int(head[0]), "channels": int(head[1]), "blocks": int(head[2]),
                         "value_hidden": int(head[3])}, np.frombuffer(raw[16:], dtype="<f4")


def test_cpp_reader_of_infer_b200_load_matches_python(pkg, oi, nb, golden_dir, tmp_path):
    """host/onnx_import.h (what infer::B200::load runs on a .onnx path) == onnx_io.py, bit for bit: on the
    torch-exported file, on a graph with BatchNormalization / MatMul / Split, on a written 128-channel net; and it
    rejects what the Python reader rejects."""
    cases = [os.path.join(golden_dir, "resnet_torch_export.onnx")]
    model, _ = _bn_graph(oi, np.random.default_rng(5))
    p = str(tmp_path / "bn.onnx")
    open(p, "wb").write(model)
    cases.append(p)
    p = str(tmp_path / "w128.onnx")
    oi.write_onnx(p, nb.random_blob(nb.net_desc(128, 1), 9), 128, 1)
    cases.append(p)
    for path in cases:
        meta, blob = oi.read_onnx(path)
        rc, out, cmeta, cblob = _cpp_blob(path, tmp_path)
        assert rc == 0, out
        assert cmeta == meta
        assert np.array_equal(cblob.view(np.uint32), blob.view(np.uint32))
    bad, _ = _bn_graph(oi, np.random.default_rng(6), extra=oi._node("Relu", ["x1"], ["dangling"]))
    p = str(tmp_path / "bad.onnx")
    open(p, "wb").write(bad)
    rc, out, _, _ = _cpp_blob(p, tmp_path)
    assert rc == 3 and "trunk output must feed" in out
    open(p, "wb").write(b"\x0a\x03abc")
    assert _cpp_blob(p, tmp_path)[0] == 3


def test_damaged_files_are_rejected_not_crashing(oi, golden_dir, tmp_path):
    """Byte-level fuzz of the torch-exported file: truncations, flipped bytes, spliced garbage.  Both readers must
    either load the file (a flipped weight byte is still a valid model) or reject it with their error type -
    ValueError / exit code 3 - never crash, hang or raise something else."""
    data = open(os.path.join(golden_dir, "resnet_torch_export.onnx"), "rb").read()
    rng = np.random.default_rng(11)
    cases = [data[:k] for k in (0, 1, 7, 100, 5000, len(data) - 1)]
    for _ in range(40):
        b = bytearray(data)
        # mutate the structural part (the graph's nodes sit at the front of the file) and anywhere
        for _ in range(int(rng.integers(1, 6))):
            k = int(rng.integers(0, 4000 if rng.random() < 0.7 else len(b)))
            b[k] = int(rng.integers(0, 256))
        cases.append(bytes(b))
    for _ in range(10):
        k = int(rng.integers(0, len(data)))
        cases.append(data[:k] + bytes(rng.integers(0, 256, size=64, dtype=np.uint8)) + data[k:])
    loaded = rejected = 0
    for i, c in enumerate(cases):
        try:
            oi.blob_from_bytes(c)
            py_ok = True
        except ValueError:
            py_ok = False
        p = str(tmp_path / "fuzz.onnx")
        open(p, "wb").write(c)
        rc, out, _, _ = _cpp_blob(p, tmp_path)
        assert rc in (0, 3), (i, rc, out)
        assert (rc == 0) == py_ok, (i, rc, py_ok, out)      # the two readers agree on what is loadable
        loaded += py_ok
        rejected += not py_ok
    assert rejected >= 10


def test_tensor_dims_are_validated(oi, tmp_path):
    """A TensorProto whose dims are negative (or huge) must be rejected, not reinterpreted: dims [-1, -1] multiply to 1
    and would otherwise pass the value-count check with a one-element payload."""
    neg = b"\xff" * 9 + b"\x01"                                   # int64 -1 as a varint
    def tensor(dims_bytes):
        return dims_bytes + oi._vi(2, 1) + oi._st(8, "w") + oi._ld(9, np.float32(1.0).tobytes())
    for dims in (b"\x08" + neg + b"\x08" + neg, b"\x08\x00", b"\x08" + oi._varint(1 << 30)):
        with pytest.raises(ValueError):
            oi._parse_tensor(tensor(dims))
        model = oi._ld(7, oi._ld(5, tensor(dims)))                  # ModelProto.graph.initializer
        with pytest.raises(ValueError):
            oi.blob_from_bytes(model)
        p = str(tmp_path / "dims.onnx")
        open(p, "wb").write(model)
        rc, out, _, _ = _cpp_blob(p, tmp_path)
        assert rc == 3 and "out of range" in out, (rc, out)
    name, arr = oi._parse_tensor(tensor(b"\x08\x01"))              # a well-formed [1] tensor still parses
    assert name == "w" and arr.shape == (1,) and arr[0] == 1.0
