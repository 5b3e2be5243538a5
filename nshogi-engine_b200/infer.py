"""Python twin of the reference's executor plug-in interface, for parity tests and benchmarks.

Mirrors reference src/infer/infer.h:19-32 (``class Infer``: computeNonBlocking / computeBlocking /
await / isComputing) and src/evaluate/evaluator.{h,cc} (``Evaluator`` owns the pinned host batch
buffers and forwards to the Infer).  Method names and argument meaning follow the reference so the
tests read like its own; the production host mirror is the C++ one in host/infer_b200.h.
"""
from __future__ import annotations

import numpy as np

from . import binding as nb


class Infer:
    """reference src/infer/infer.h:19-32."""

    def computeNonBlocking(self, Features, BatchSize, DstPolicy, DstWinRate, DstDrawRate):
        raise NotImplementedError

    def computeBlocking(self, Features, BatchSize, DstPolicy, DstWinRate, DstDrawRate):
        raise NotImplementedError

    def await_(self):
        raise NotImplementedError

    def isComputing(self) -> bool:
        raise NotImplementedError


class B200(Infer):
    """Drop-in beside the reference's Zero / Nothing / Random / TensorRT executors.  Constructor
    follows TensorRT(GPUId, BatchSizeMax, NumChannels) + load() (src/infer/trt.cc:52-80,109-232)."""

    def __init__(self, GPUId: int, BatchSizeMax: int, NumChannels: int, desc: nb.NetDesc, blob=None,
                 seed=None, slots: int = 1):
        if NumChannels != nb.FEATURE_CHANNELS:
            raise nb.NsbError(f"NumChannels must be {nb.FEATURE_CHANNELS}")
        self.ctx = nb.Context(desc, BatchSizeMax, slots=slots, gpu=GPUId, blob=blob, seed=seed)
        self.BatchSizeM = BatchSizeMax
        self._slot = 0

    def load(self, blob):
        self.ctx.load_weights(blob)

    def resetGPU(self):  # src/infer/trt.cc:289-291
        self.ctx.bind_thread()

    def computeNonBlocking(self, Features, BatchSize, DstPolicy, DstWinRate, DstDrawRate):
        assert BatchSize <= self.BatchSizeM and not self.isComputing()  # trt.cc:237-238
        self.ctx.eval_async(self._slot, Features, BatchSize, DstPolicy, DstWinRate, DstDrawRate)

    def computeBlocking(self, Features, BatchSize, DstPolicy, DstWinRate, DstDrawRate):
        self.computeNonBlocking(Features, BatchSize, DstPolicy, DstWinRate, DstDrawRate)
        self.await_()

    def await_(self):
        self.ctx.await_(self._slot)

    def isComputing(self) -> bool:
        return self.ctx.is_computing(self._slot)

    def close(self):
        self.ctx.close()


class Evaluator:
    """reference src/evaluate/evaluator.{h,cc}: owns FeatureBitboards[B*86], Policy[B*2187],
    WinRate[B], DrawRate[B], page-locked (evaluator.cc:85-106), forwards to the Infer."""

    def __init__(self, BatchSize: int, PInfer: Infer):
        self.BatchSizeMax = BatchSize
        self.PInfer = PInfer
        self._fb = nb.PinnedArray((BatchSize * nb.FEATURE_CHANNELS,), nb.FEATURE_BITBOARD)
        self._policy = nb.PinnedArray((BatchSize * nb.POLICY_SIZE,), np.float32)
        self._win = nb.PinnedArray((BatchSize,), np.float32)
        self._draw = nb.PinnedArray((BatchSize,), np.float32)

    def getFeatureBitboards(self):
        return self._fb.array

    def getPolicy(self):
        return self._policy.array

    def getWinRate(self):
        return self._win.array

    def getDrawRate(self):
        return self._draw.array

    def computeNonBlocking(self, BatchSize):
        self.PInfer.computeNonBlocking(self._fb.array, BatchSize, self._policy.array, self._win.array,
                                       self._draw.array)

    def computeBlocking(self, BatchSize):
        self.PInfer.computeBlocking(self._fb.array, BatchSize, self._policy.array, self._win.array,
                                    self._draw.array)

    def await_(self):
        self.PInfer.await_()

    def isComputing(self):
        return self.PInfer.isComputing()

    def close(self):
        for a in (self._fb, self._policy, self._win, self._draw):
            a.free()
