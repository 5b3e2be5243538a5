"""Reader of the teacher-record files the self-play harness writes (host/teacher_io.h, "NSBT"): the positions of
finished games at which a full search was conducted, with the move played and the game's winner - what the reference's
SaveWorker::save emits (reference src/selfplay/saveworker.cc:160-182).  The byte format is this repo's (libnshogi's
simple_teacher format is not available); test and tooling plumbing only."""
from __future__ import annotations

import struct

import numpy as np

from .binding import POSITION

RECORD = np.dtype({"names": ["position", "from", "to", "promote", "piece", "winner"],
                   "formats": [POSITION, "u1", "u1", "u1", "u1", "u1"],
                   "offsets": [0, 108, 109, 110, 111, 112], "itemsize": 116})
WINNER_BLACK, WINNER_WHITE, WINNER_NONE = 0, 1, 2


def read_nsbt(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        head = f.read(12)
        if len(head) != 12 or head[:4] != b"NSBT":
            raise ValueError(f"{path}: not an NSBT teacher file")
        version, size = struct.unpack("<II", head[4:])
        if version != 1 or size != RECORD.itemsize:
            raise ValueError(f"{path}: NSBT version {version} / record size {size} not understood")
        body = f.read()
    if len(body) % RECORD.itemsize:
        raise ValueError(f"{path}: truncated record")
    return np.frombuffer(body, dtype=RECORD).copy()
