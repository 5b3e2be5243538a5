"""ONNX <-> canonical weight blob, without the `onnx` package.

The reference's executor loads an external ONNX file through TensorRT (reference src/infer/trt.cc:109-232); the
file's contract is pinned there: input tensor `input` [B,86,9,9] (:144-150), outputs `policy` (2187 values per
sample, :193-214), `value` and `draw` (:215-227).  This module

  * reads such a file (`read_onnx`): a small protobuf wire-format decoder, then a walk over the graph that
    recognises the ResNet this executor runs - stem conv3x3, residual blocks (conv-[bn]-relu-conv-[bn]-add-relu),
    policy head conv1x1(27), value head conv1x1(1)-[bn]-relu-fc-relu-fc-sigmoid -> value / draw - folds
    batch-norm (weights_io.fold_bn) and returns the canonical fp32 blob of DESIGN.md §5.  Anything else in the
    graph is an error, never silently skipped: a net this executor cannot run must not load.
  * writes one (`write_onnx`) from a blob, so the same weights can be given to the reference's TensorRT executor
    for a side-by-side run.

Only the subset of ONNX these graphs use is understood (Conv, BatchNormalization, Relu, Add, Flatten, Reshape,
Gemm, MatMul, Sigmoid, Split, Slice, Gather, Squeeze, Unsqueeze, Identity, Constant; float / double / float16 /
int32 / int64 tensors).  Field numbers follow onnx.proto (ONNX IR version 3-10); the reader is pinned against a
file produced by torch.onnx.export (tests/golden/resnet_torch_export.onnx, tools/make_golden.py).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import weights_io

# ---------------------------------------------------------------------------------------------------------------
# protobuf wire format
# ---------------------------------------------------------------------------------------------------------------


def _read_varint(buf: bytes, i: int) -> Tuple[int, int]:
    shift = val = 0
    while True:
        b = buf[i]
        i += 1
        val |= (b & 0x7F) << shift
        if b < 0x80:
            return val, i
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _fields(buf: bytes):
    """Yield (field number, wire type, value): int for varint / fixed, bytes for length-delimited."""
    i, n = 0, len(buf)
    while i < n:
        key, i = _read_varint(buf, i)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _read_varint(buf, i)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, i)[0]
            i += 8
        elif wt == 2:
            ln, i = _read_varint(buf, i)
            v = bytes(buf[i:i + ln])
            if len(v) != ln:
                raise ValueError("truncated protobuf message")
            i += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, i)[0]
            i += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fno, wt, v


def _s(v) -> str:
    """Names are matched, never interpreted: bytes that are not UTF-8 survive as surrogates (the C++ reader compares
    raw bytes)."""
    return v.decode("utf-8", "surrogateescape")


def _signed(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _repeated_int64(acc: List[int], wt: int, v):
    if wt == 0:
        acc.append(_signed(v))
    else:  # packed
        i = 0
        while i < len(v):
            x, i = _read_varint(v, i)
            acc.append(_signed(x))


def _varint(n: int) -> bytes:
    n &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _vi(fno: int, n: int) -> bytes:
    return _varint(fno << 3) + _varint(n)


def _ld(fno: int, payload: bytes) -> bytes:
    return _varint((fno << 3) | 2) + _varint(len(payload)) + payload


def _st(fno: int, s: str) -> bytes:
    return _ld(fno, s.encode())


# ---------------------------------------------------------------------------------------------------------------
# ONNX messages (the subset needed)
# ---------------------------------------------------------------------------------------------------------------
_DTYPES = {1: np.dtype("<f4"), 6: np.dtype("<i4"), 7: np.dtype("<i8"), 10: np.dtype("<f2"), 11: np.dtype("<f8")}


@dataclass
class Node:
    op: str
    inputs: List[str]
    outputs: List[str]
    attrs: Dict[str, object] = field(default_factory=dict)
    name: str = ""


@dataclass
class Graph:
    nodes: List[Node]
    initializers: Dict[str, np.ndarray]
    inputs: List[str]     # graph inputs that are not initializers
    outputs: List[str]
    opset: int = 0


def _parse_tensor(buf: bytes) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    dtype, name, raw = 1, "", None
    floats: List[float] = []
    doubles: List[float] = []
    i32: List[int] = []
    i64: List[int] = []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            _repeated_int64(dims, wt, v)
        elif fno == 2:
            dtype = v
        elif fno == 8:
            name = _s(v)
        elif fno == 9:
            raw = v
        elif fno == 4:
            floats.extend(struct.unpack(f"<{len(v) // 4}f", v) if wt == 2 else struct.unpack("<f", struct.pack("<I", v)))
        elif fno == 10:
            doubles.extend(struct.unpack(f"<{len(v) // 8}d", v) if wt == 2 else struct.unpack("<d", struct.pack("<Q", v)))
        elif fno == 5:
            _repeated_int64(i32, wt, v)
        elif fno == 7:
            _repeated_int64(i64, wt, v)
        elif fno == 14 and v == 1:
            raise ValueError(f"tensor {name!r}: external data is not supported (export with weights embedded)")
    if dtype not in _DTYPES:
        raise ValueError(f"tensor {name!r}: unsupported ONNX data type {dtype}")
    dt = _DTYPES[dtype]
    if raw is not None:
        arr = np.frombuffer(raw, dtype=dt).copy()
    elif dtype == 1:
        arr = np.asarray(floats, dtype=dt)
    elif dtype == 11:
        arr = np.asarray(doubles, dtype=dt)
    elif dtype == 7:
        arr = np.asarray(i64, dtype=dt)
    elif dtype == 6:
        arr = np.asarray(i32, dtype=dt)
    else:  # float16 bits travel in int32_data
        arr = np.asarray(i32, dtype="<u2").view("<f2")
    if any(d <= 0 or d > (1 << 24) for d in dims):
        raise ValueError(f"tensor {name!r}: dimension out of range in shape {dims}")
    n = int(np.prod(dims, dtype=object)) if dims else 1
    if arr.size != n:
        raise ValueError(f"tensor {name!r}: {arr.size} values for shape {dims}")
    return name, arr.reshape(dims)


def _parse_attr(buf: bytes) -> Tuple[str, object]:
    name, atype = "", 0
    f = i = s = t = None
    floats: List[float] = []
    ints: List[int] = []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = _s(v)
        elif fno == 20:
            atype = v
        elif fno == 2:
            f = struct.unpack("<f", struct.pack("<I", v))[0]
        elif fno == 3:
            i = _signed(v)
        elif fno == 4:
            s = v
        elif fno == 5:
            t = _parse_tensor(v)[1]
        elif fno == 7:
            floats.extend(struct.unpack(f"<{len(v) // 4}f", v) if wt == 2 else struct.unpack("<f", struct.pack("<I", v)))
        elif fno == 8:
            _repeated_int64(ints, wt, v)
    if atype == 1 or (atype == 0 and f is not None):
        return name, f
    if atype == 2 or (atype == 0 and i is not None):
        return name, i
    if atype == 3 or (atype == 0 and s is not None):
        return name, s.decode(errors="replace")
    if atype == 4 or (atype == 0 and t is not None):
        return name, t
    if atype == 6 or (atype == 0 and floats):
        return name, floats
    if atype == 7 or atype == 0:
        return name, ints
    raise ValueError(f"attribute {name!r}: unsupported attribute type {atype}")


def _parse_node(buf: bytes) -> Node:
    n = Node("", [], [])
    for fno, _, v in _fields(buf):
        if fno == 1:
            n.inputs.append(_s(v))
        elif fno == 2:
            n.outputs.append(_s(v))
        elif fno == 3:
            n.name = _s(v)
        elif fno == 4:
            n.op = _s(v)
        elif fno == 5:
            k, val = _parse_attr(v)
            n.attrs[k] = val
        elif fno == 7 and v not in (b"", b"ai.onnx"):
            raise ValueError(f"node {n.name!r}: operator domain {_s(v)!r} is not supported")
    return n


def _value_info_name(buf: bytes) -> str:
    for fno, _, v in _fields(buf):
        if fno == 1:
            return _s(v)
    return ""


def parse_model(data: bytes) -> Graph:
    graph_buf, opset = None, 0
    for fno, _, v in _fields(data):
        if fno == 7:
            graph_buf = v
        elif fno == 8:
            dom, ver = "", 0
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    dom = _s(v2)
                elif f2 == 2:
                    ver = v2
            if dom in ("", "ai.onnx"):
                opset = ver
    if graph_buf is None:
        raise ValueError("not an ONNX ModelProto (no graph)")
    nodes: List[Node] = []
    inits: Dict[str, np.ndarray] = {}
    inputs: List[str] = []
    outputs: List[str] = []
    for fno, _, v in _fields(graph_buf):
        if fno == 1:
            nodes.append(_parse_node(v))
        elif fno == 5:
            name, arr = _parse_tensor(v)
            inits[name] = arr
        elif fno == 11:
            inputs.append(_value_info_name(v))
        elif fno == 12:
            outputs.append(_value_info_name(v))
    inputs = [x for x in inputs if x not in inits]
    return Graph(nodes, inits, inputs, outputs, opset)


# ---------------------------------------------------------------------------------------------------------------
# graph -> canonical blob
# ---------------------------------------------------------------------------------------------------------------
_PASS_THROUGH = ("Flatten", "Reshape", "Identity", "Squeeze", "Unsqueeze")


class _Walker:
    def __init__(self, g: Graph):
        self.g = g
        self.const: Dict[str, np.ndarray] = dict(g.initializers)
        self.nodes: List[Node] = []
        for n in g.nodes:
            if n.op == "Constant":
                if "value" not in n.attrs:
                    raise ValueError("Constant node without a tensor value")
                self.const[n.outputs[0]] = np.asarray(n.attrs["value"])
            else:
                self.nodes.append(n)
        self.consumers: Dict[str, List[Node]] = {}
        for n in self.nodes:
            for x in n.inputs:
                if x and x not in self.const:
                    self.consumers.setdefault(x, []).append(n)
        self.used: set = set()

    def cons(self, name: str, op: Optional[str] = None) -> List[Node]:
        c = self.consumers.get(name, [])
        return [n for n in c if n.op == op] if op else list(c)

    def sole(self, name: str, op: str, what: str) -> Node:
        c = self.cons(name)
        if len(c) != 1 or c[0].op != op:
            raise ValueError(f"{what}: expected a single {op} after {name!r}, found {[n.op for n in c]}")
        self.used.add(id(c[0]))
        return c[0]

    def weight(self, name: str, what: str) -> np.ndarray:
        if name not in self.const:
            raise ValueError(f"{what}: {name!r} is not a constant tensor")
        return np.asarray(self.const[name], dtype=np.float64)

    def conv(self, node: Node, ksize: int, what: str) -> Tuple[np.ndarray, np.ndarray, str]:
        """Conv [+ BatchNormalization] -> (folded weight, folded bias, output tensor name)."""
        self.used.add(id(node))
        a = node.attrs
        w = self.weight(node.inputs[1], what)
        if w.ndim != 4 or w.shape[2] != ksize or w.shape[3] != ksize:
            raise ValueError(f"{what}: expected a {ksize}x{ksize} convolution, weight shape {w.shape}")
        pad = (ksize - 1) // 2
        pads = list(a.get("pads", [0, 0, 0, 0]))
        auto_pad = a.get("auto_pad", "NOTSET")
        if not (pads == [pad] * 4 or (auto_pad in ("SAME_UPPER", "SAME_LOWER")) or (pad == 0 and auto_pad == "VALID")):
            raise ValueError(f"{what}: needs 'same' padding ({pad}), got pads={pads} auto_pad={auto_pad}")
        if list(a.get("strides", [1, 1])) != [1, 1] or list(a.get("dilations", [1, 1])) != [1, 1] or a.get("group", 1) != 1:
            raise ValueError(f"{what}: stride / dilation / group must be 1")
        b = self.weight(node.inputs[2], what) if len(node.inputs) > 2 and node.inputs[2] else None
        out = node.outputs[0]
        nxt = self.cons(out)
        bn = None
        if len(nxt) == 1 and nxt[0].op == "BatchNormalization":
            n = nxt[0]
            self.used.add(id(n))
            if n.attrs.get("training_mode", 0):
                raise ValueError(f"{what}: batch-norm in training mode; export the model in eval mode")
            bn = {k: self.weight(n.inputs[i], what) for i, k in ((1, "weight"), (2, "bias"), (3, "running_mean"), (4, "running_var"))}
            w, b = weights_io.fold_bn(w, b, bn, eps=float(n.attrs.get("epsilon", 1e-5)))
            out = n.outputs[0]
        else:
            w, b = weights_io.fold_bn(w, b, None)
        return w, b, out

    def relu(self, name: str, what: str) -> str:
        return self.sole(name, "Relu", what).outputs[0]

    def through(self, name: str, what: str) -> str:
        """Follow shape-only nodes (flatten / reshape / squeeze ...) to the tensor they end in."""
        while True:
            c = self.cons(name)
            if len(c) == 1 and c[0].op in _PASS_THROUGH:
                self.used.add(id(c[0]))
                name = c[0].outputs[0]
                continue
            return name

    def dense(self, name: str, what: str) -> Tuple[np.ndarray, np.ndarray, str]:
        """Gemm, or MatMul + Add -> (weight [out, in], bias [out], output tensor name)."""
        c = self.cons(name)
        if len(c) != 1:
            raise ValueError(f"{what}: expected one fully connected layer after {name!r}, found {[n.op for n in c]}")
        n = c[0]
        return self.dense_node(n, what)

    def dense_node(self, n: Node, what: str) -> Tuple[np.ndarray, np.ndarray, str]:
        self.used.add(id(n))
        if n.op == "Gemm":
            a = n.attrs
            if a.get("transA", 0) or float(a.get("alpha", 1.0)) != 1.0 or float(a.get("beta", 1.0)) != 1.0:
                raise ValueError(f"{what}: Gemm with transA / alpha / beta is not supported")
            w = self.weight(n.inputs[1], what)
            if not a.get("transB", 0):
                w = w.T
            b = self.weight(n.inputs[2], what).reshape(-1) if len(n.inputs) > 2 and n.inputs[2] else np.zeros(w.shape[0])
            return w, b, n.outputs[0]
        if n.op == "MatMul":
            w = self.weight(n.inputs[1], what).T
            out = n.outputs[0]
            add = self.cons(out)
            b = np.zeros(w.shape[0])
            if len(add) == 1 and add[0].op == "Add":
                other = [x for x in add[0].inputs if x != out]
                if len(other) == 1 and other[0] in self.const:
                    self.used.add(id(add[0]))
                    b = self.weight(other[0], what).reshape(-1)
                    out = add[0].outputs[0]
            return w, b, out
        raise ValueError(f"{what}: expected Gemm or MatMul, found {n.op}")

    def int_const(self, name: str, what: str) -> int:
        if name not in self.const:
            raise ValueError(f"{what}: index {name!r} is not a constant")
        v = np.asarray(self.const[name]).reshape(-1)
        if v.size != 1:
            raise ValueError(f"{what}: expected a scalar index")
        return int(v[0])


def blob_from_graph(g: Graph) -> Tuple[dict, np.ndarray]:
    """Canonical blob + net description from a parsed ONNX graph of the supported ResNet."""
    wk = _Walker(g)
    if "input" not in g.inputs:
        raise ValueError(f"graph input 'input' not found (reference src/infer/trt.cc:144-150); inputs: {g.inputs}")
    for o in ("policy", "value", "draw"):
        if o not in g.outputs:
            raise ValueError(f"graph output {o!r} not found (reference src/infer/trt.cc:193-227); outputs: {g.outputs}")
    state: Dict[str, np.ndarray] = {}

    stem = wk.cons("input", "Conv")
    if len(stem) != 1 or len(wk.cons("input")) != 1:
        raise ValueError("the stem must be one 3x3 convolution on 'input'")
    w, b, x = wk.conv(stem[0], 3, "stem")
    state["stem.conv.weight"], state["stem.conv.bias"] = w, b
    C, cin = w.shape[0], w.shape[1]
    x = wk.relu(x, "stem")

    blocks = 0
    while wk.cons(x, "Add"):
        what = f"block {blocks}"
        add = wk.cons(x, "Add")
        c1 = [n for n in wk.cons(x, "Conv")]
        if len(add) != 1 or len(c1) != 1 or len(wk.cons(x)) != 2:
            raise ValueError(f"{what}: a residual block input feeds exactly one Conv and one Add")
        w1, b1, h = wk.conv(c1[0], 3, what + " conv1")
        h = wk.relu(h, what + " conv1")
        c2 = wk.sole(h, "Conv", what + " conv2")
        w2, b2, y = wk.conv(c2, 3, what + " conv2")
        if sorted(add[0].inputs) != sorted([x, y]) or len(wk.cons(y)) != 1:
            raise ValueError(f"{what}: the skip connection must add the block input to conv2's output")
        wk.used.add(id(add[0]))
        for k, v in (("conv1.weight", w1), ("conv1.bias", b1), ("conv2.weight", w2), ("conv2.bias", b2)):
            if v.ndim == 4 and v.shape[:2] != (C, C):
                raise ValueError(f"{what}: expected {C}->{C} channels, weight shape {v.shape}")
            state[f"blocks.{blocks}.{k}"] = v
        x = wk.relu(add[0].outputs[0], what)
        blocks += 1
    if blocks == 0:
        raise ValueError("no residual block found after the stem")

    heads = wk.cons(x)
    if len(heads) != 2 or any(n.op != "Conv" for n in heads):
        raise ValueError(f"the trunk output must feed the policy and value 1x1 convolutions, found {[n.op for n in heads]}")
    by_out = {wk.weight(n.inputs[1], "head").shape[0]: n for n in heads}
    if set(by_out) != {27, 1}:
        raise ValueError(f"head convolutions must have 27 (policy) and 1 (value) output channels, found {sorted(by_out)}")

    w, b, p = wk.conv(by_out[27], 1, "policy head")
    state["policy.conv.weight"], state["policy.conv.bias"] = w, b
    if wk.through(p, "policy head") != "policy":
        raise ValueError("policy head: conv1x1(27) must reach the output 'policy' through shape-only nodes "
                         "(plane-major logits, index = plane * 81 + square)")

    w, b, v = wk.conv(by_out[1], 1, "value head")
    state["value.conv.weight"], state["value.conv.bias"] = w, b
    v = wk.through(wk.relu(v, "value head"), "value head")
    w1, b1, v = wk.dense(v, "value head fc1")
    if w1.shape[1] != 81:
        raise ValueError(f"value head fc1: expected 81 inputs, weight shape {w1.shape}")
    state["value.fc1.weight"], state["value.fc1.bias"] = w1, b1
    H = w1.shape[0]
    v = wk.relu(v, "value head fc1")
    rows: Dict[str, Tuple[np.ndarray, float]] = {}
    nxt = wk.cons(v)
    if len(nxt) == 1:
        w2, b2, o = wk.dense_node(nxt[0], "value head fc2")
        if w2.shape != (2, H):
            raise ValueError(f"value head fc2: expected weight [2, {H}], got {w2.shape}")
        o = wk.sole(o, "Sigmoid", "value head").outputs[0]
        picks = wk.cons(o)
        for n in picks:
            wk.used.add(id(n))
            if n.op == "Split" and n.attrs.get("axis", 0) in (1, -1) and len(n.outputs) == 2:
                ends = [(0, n.outputs[0]), (1, n.outputs[1])]
            elif n.op == "Gather" and n.attrs.get("axis", 0) in (1, -1):
                ends = [(wk.int_const(n.inputs[1], "value head gather"), n.outputs[0])]
            elif n.op == "Slice":
                axes = wk.int_const(n.inputs[3], "value head slice") if len(n.inputs) > 3 and n.inputs[3] else 0
                if axes not in (1, -1):
                    raise ValueError("value head: Slice must cut axis 1")
                ends = [(wk.int_const(n.inputs[1], "value head slice"), n.outputs[0])]
            else:
                raise ValueError(f"value head: cannot split the (value, draw) pair with {n.op}")
            for idx, name in ends:
                out = wk.through(name, "value head")
                if out not in ("value", "draw") or idx not in (0, 1):
                    raise ValueError(f"value head: component {idx} ends in {out!r}, expected 'value' or 'draw'")
                rows[out] = (w2[idx], float(b2[idx]))
    elif len(nxt) == 2:
        for n in nxt:
            w2, b2, o = wk.dense_node(n, "value head fc2")
            if w2.shape != (1, H):
                raise ValueError(f"value head fc2: expected weight [1, {H}], got {w2.shape}")
            out = wk.through(wk.sole(o, "Sigmoid", "value head").outputs[0], "value head")
            rows[out] = (w2[0], float(b2[0]))
    if set(rows) != {"value", "draw"}:
        raise ValueError(f"value head: could not resolve the outputs 'value' and 'draw' (found {sorted(rows)})")
    state["value.fc2.weight"] = np.stack([rows["value"][0], rows["draw"][0]])
    state["value.fc2.bias"] = np.asarray([rows["value"][1], rows["draw"][1]])

    unused = [f"{n.op}({n.name or n.outputs[0]})" for n in wk.nodes if id(n) not in wk.used]
    if unused:
        raise ValueError(f"unsupported graph: nodes outside the recognised ResNet: {unused[:8]}")
    blob = weights_io.blob_from_state(state, channels=C, blocks=blocks, value_hidden=H, in_channels=cin)
    return {"channels": C, "blocks": blocks, "value_hidden": H, "in_channels": cin}, blob


def blob_from_bytes(data: bytes) -> Tuple[dict, np.ndarray]:
    """parse_model + blob_from_graph; a damaged file is a ValueError like an unsupported one, never a stray
    IndexError / struct.error from the decoder."""
    try:
        return blob_from_graph(parse_model(data))
    except ValueError:
        raise
    except (IndexError, KeyError, AttributeError, struct.error, UnicodeDecodeError, OverflowError, MemoryError, TypeError) as e:
        raise ValueError(f"malformed ONNX file: {type(e).__name__}: {e}") from e


def read_onnx(path: str) -> Tuple[dict, np.ndarray]:
    with open(path, "rb") as f:
        return blob_from_bytes(f.read())


# ---------------------------------------------------------------------------------------------------------------
# canonical blob -> ONNX
# ---------------------------------------------------------------------------------------------------------------


def split_blob(blob: np.ndarray, channels: int, blocks: int, value_hidden: int = 256, in_channels: int = 86) -> Dict[str, np.ndarray]:
    """The blob's tensors by name (layout: weights_io.blob_from_state / DESIGN.md §5)."""
    C, H = channels, value_hidden
    shapes = [("stem.conv", (C, in_channels, 3, 3))]
    for i in range(blocks):
        shapes += [(f"blocks.{i}.conv1", (C, C, 3, 3)), (f"blocks.{i}.conv2", (C, C, 3, 3))]
    shapes += [("policy.conv", (27, C, 1, 1)), ("value.conv", (1, C, 1, 1)), ("value.fc1", (H, 81)), ("value.fc2", (2, H))]
    out, o = {}, 0
    blob = np.asarray(blob, dtype=np.float32).reshape(-1)
    for name, shp in shapes:
        n = int(np.prod(shp))
        out[name + ".weight"] = blob[o:o + n].reshape(shp)
        o += n
        out[name + ".bias"] = blob[o:o + shp[0]]
        o += shp[0]
    if o != blob.size:
        raise ValueError(f"blob has {blob.size} floats, the net needs {o}")
    return out


def _tensor(name: str, arr: np.ndarray) -> bytes:
    arr = np.ascontiguousarray(arr)
    if arr.dtype == np.int64:
        dt, raw = 7, arr.astype("<i8").tobytes()
    else:
        dt, raw = 1, arr.astype("<f4").tobytes()
    return b"".join(_vi(1, d) for d in arr.shape) + _vi(2, dt) + _st(8, name) + _ld(9, raw)


def _attr_int(name: str, v: int) -> bytes:
    return _ld(5, _st(1, name) + _vi(3, v) + _vi(20, 2))


def _attr_ints(name: str, vs) -> bytes:
    return _ld(5, _st(1, name) + b"".join(_vi(8, v) for v in vs) + _vi(20, 7))


def _node(op: str, inputs, outputs, attrs: bytes = b"") -> bytes:
    return _ld(1, b"".join(_st(1, x) for x in inputs) + b"".join(_st(2, x) for x in outputs) + _st(3, outputs[0]) +
               _st(4, op) + attrs)


def _value_info(name: str, dims) -> bytes:
    shape = b"".join(_ld(1, _st(2, d) if isinstance(d, str) else _vi(1, d)) for d in dims)
    return _st(1, name) + _ld(2, _ld(1, _vi(1, 1) + _ld(2, shape)))


def model_bytes(blob: np.ndarray, channels: int, blocks: int, value_hidden: int = 256, in_channels: int = 86,
                opset: int = 17) -> bytes:
    t = split_blob(blob, channels, blocks, value_hidden, in_channels)
    nodes, inits = [], []

    def conv(name, x, out, k):
        inits.append(_ld(5, _tensor(name + ".weight", t[name + ".weight"])))
        inits.append(_ld(5, _tensor(name + ".bias", t[name + ".bias"])))
        p = (k - 1) // 2
        nodes.append(_node("Conv", [x, name + ".weight", name + ".bias"], [out],
                           _attr_ints("dilations", [1, 1]) + _attr_int("group", 1) + _attr_ints("kernel_shape", [k, k]) +
                           _attr_ints("pads", [p, p, p, p]) + _attr_ints("strides", [1, 1])))

    def relu(x, out):
        nodes.append(_node("Relu", [x], [out]))

    conv("stem.conv", "input", "stem.c", 3)
    relu("stem.c", "x0")
    x = "x0"
    for i in range(blocks):
        conv(f"blocks.{i}.conv1", x, f"b{i}.c1", 3)
        relu(f"b{i}.c1", f"b{i}.h")
        conv(f"blocks.{i}.conv2", f"b{i}.h", f"b{i}.c2", 3)
        nodes.append(_node("Add", [x, f"b{i}.c2"], [f"b{i}.sum"]))
        relu(f"b{i}.sum", f"x{i + 1}")
        x = f"x{i + 1}"
    conv("policy.conv", x, "policy.c", 1)
    nodes.append(_node("Flatten", ["policy.c"], ["policy"], _attr_int("axis", 1)))
    conv("value.conv", x, "value.c", 1)
    relu("value.c", "value.r")
    nodes.append(_node("Flatten", ["value.r"], ["value.f"], _attr_int("axis", 1)))
    inits.append(_ld(5, _tensor("value.fc1.weight", t["value.fc1.weight"])))
    inits.append(_ld(5, _tensor("value.fc1.bias", t["value.fc1.bias"])))
    nodes.append(_node("Gemm", ["value.f", "value.fc1.weight", "value.fc1.bias"], ["value.h"], _attr_int("transB", 1)))
    relu("value.h", "value.hr")
    for row, out in ((0, "value"), (1, "draw")):   # fc2 row 0 = win rate, row 1 = draw rate
        inits.append(_ld(5, _tensor(f"{out}.fc2.weight", t["value.fc2.weight"][row:row + 1])))
        inits.append(_ld(5, _tensor(f"{out}.fc2.bias", t["value.fc2.bias"][row:row + 1])))
        nodes.append(_node("Gemm", ["value.hr", f"{out}.fc2.weight", f"{out}.fc2.bias"], [f"{out}.logit"], _attr_int("transB", 1)))
        nodes.append(_node("Sigmoid", [f"{out}.logit"], [out]))
    graph = (b"".join(nodes) + _st(2, "nsb_resnet") + b"".join(inits) +
             _ld(11, _value_info("input", ["batch", in_channels, 9, 9])) +
             _ld(12, _value_info("policy", ["batch", 27 * 81])) + _ld(12, _value_info("value", ["batch", 1])) +
             _ld(12, _value_info("draw", ["batch", 1])))
    return _vi(1, 8) + _st(2, "nsb") + _st(3, "0.2") + _ld(7, graph) + _ld(8, _st(1, "") + _vi(2, opset))


def write_onnx(path: str, blob: np.ndarray, channels: int, blocks: int, value_hidden: int = 256, in_channels: int = 86):
    """An ONNX file with the tensor contract of reference src/infer/trt.cc:144-150,193-227 holding these weights
    (fp32, batch-norm already folded), loadable by the reference's TensorRT executor."""
    with open(path, "wb") as f:
        f.write(model_bytes(blob, channels, blocks, value_hidden, in_channels))
