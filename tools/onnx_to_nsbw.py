#!/usr/bin/env python
"""ONNX model file of the reference (src/infer/trt.cc:109-232 loads it through TensorRT) -> NSBW weight file for
infer::B200::load (host/infer_b200.h), or back:
    onnx_to_nsbw.py model.onnx model.nsbw
    onnx_to_nsbw.py --to-onnx model.nsbw model.onnx
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402


def main(argv):
    pkg = graft.load_package()
    if len(argv) == 3 and argv[0] == "--to-onnx":
        meta, blob = pkg.weights_io.read_nsbw(argv[1])
        pkg.onnx_io.write_onnx(argv[2], blob, meta["channels"], meta["blocks"], meta["value_hidden"], meta["in_channels"])
        print(f"{argv[2]}: {meta}")
        return 0
    if len(argv) != 2:
        print(__doc__)
        return 2
    meta, blob = pkg.onnx_io.read_onnx(argv[0])
    pkg.weights_io.write_nsbw(argv[1], blob, meta["channels"], meta["blocks"], meta["value_hidden"], meta["in_channels"])
    print(f"{argv[1]}: {meta}, {blob.size} floats")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
