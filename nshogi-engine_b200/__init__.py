"""nshogi-engine_b200 — B200-native leaf-evaluation executor for nyashiki/nshogi-engine.

The product is ``libnsb.so`` (csrc/, C ABI in include/nsb.h) plus the C++ host mirror of the
reference's ``infer::Infer`` / ``evaluate::Evaluator`` (host/).  The Python modules are plumbing
for tests and benchmarks: ``binding`` (ctypes), ``infer`` (Python twin of the Infer interface),
``synth`` (seeded synthetic inputs), ``replica`` (per-GPU sharding and counter reduction),
``weights_io`` (trained net with batch-norm -> canonical blob -> NSBW file for ``infer::B200::load``),
``onnx_io`` (the reference's ONNX model files <-> canonical blob, no `onnx` package needed).  There is no
CPU fallback anywhere in this package.
"""
from . import binding, infer, onnx_io, replica, synth, teacher_io, weights_io  # noqa: F401
