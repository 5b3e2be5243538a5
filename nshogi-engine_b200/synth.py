"""Seeded synthetic inputs for the leaf-evaluation path (SURVEY.md §8d "Synthetic inputs").

Pure numpy, deterministic for a given seed; used by tests/, bench.py and the golden-vector script.
Nothing here computes any part of the hot path.
"""
from __future__ import annotations

import numpy as np

from .binding import FEATURE_BITBOARD, FEATURE_CHANNELS, MAX_LEGAL_MOVES, POLICY_SIZE, POSITION

SEED = 20240203  # echoes the reference's only hot-path test (src/test/test_extractbit.cc:72)

# piece types in channel order of reference src/evaluate/preset.h:20-33
P, L, N, S, G, K, B, R = range(8)
PROMOTE = {P: 8, L: 9, N: 10, S: 11, B: 12, R: 13}
COUNTS = [(P, 18), (L, 4), (N, 4), (S, 4), (G, 4), (B, 2), (R, 2), (K, 2)]
HAND_SLOT = {P: 0, L: 1, N: 2, S: 3, G: 4, B: 5, R: 6}


def random_feature_bitboards(n_planes: int, seed: int = SEED, garbage: bool = True) -> np.ndarray:
    """FeatureBitboard-level fuzz: random 81-bit masks, rotate in {0,1}, value in {1.0f, random
    fp32 in (0,1)}, garbage in the unused bits 18..23 / 25..31 of the high word (must be ignored,
    reference src/cuda/extractbit.cu:20-37)."""
    rng = np.random.default_rng(seed)
    fb = np.zeros(n_planes, dtype=FEATURE_BITBOARD)
    lo = rng.integers(0, 1 << 63, size=n_planes, dtype=np.uint64)
    hi18 = rng.integers(0, 1 << 18, size=n_planes, dtype=np.uint64)
    density = rng.random(n_planes)
    lo = np.where(density < 0.15, np.uint64(0), np.where(density > 0.85, np.uint64((1 << 63) - 1), lo))
    hi18 = np.where(density < 0.15, np.uint64(0), np.where(density > 0.85, np.uint64((1 << 18) - 1), hi18))
    rot = rng.integers(0, 2, size=n_planes, dtype=np.uint64)
    val = np.where(rng.random(n_planes) < 0.5, np.float32(1.0), rng.random(n_planes, dtype=np.float32))
    val = val.astype(np.float32)
    bits = val.view(np.uint32).astype(np.uint64)
    hi = hi18 | (rot << np.uint64(24)) | (bits << np.uint64(32))
    if garbage:
        g1 = rng.integers(0, 1 << 6, size=n_planes, dtype=np.uint64) << np.uint64(18)
        g2 = rng.integers(0, 1 << 7, size=n_planes, dtype=np.uint64) << np.uint64(25)
        hi = hi | g1 | g2
    # bit 63 of lo is not a square either (squares 0..62 live in lo)
    lo = lo | (rng.integers(0, 2, size=n_planes, dtype=np.uint64) << np.uint64(63) if garbage else np.uint64(0))
    fb["lo"], fb["hi"] = lo, hi
    return fb


def random_positions(n: int, seed: int = SEED) -> np.ndarray:
    """Random (not necessarily legal) shogi positions holding the 40 pieces: owner uniform,
    p(hand)=0.25 (kings always on board), promoted with p=0.2, distinct squares, side uniform,
    MaxPly~U[224,640], ply~U[0,MaxPly), draw values per reference src/selfplay/worker.cc:135-150."""
    rng = np.random.default_rng(seed)
    pos = np.zeros(n, dtype=POSITION)
    for i in range(n):
        squares = rng.permutation(81)
        k = 0
        board = np.zeros(81, dtype=np.uint8)
        hands = np.zeros((2, 7), dtype=np.uint8)
        for pt, cnt in COUNTS:
            for j in range(cnt):
                owner = j if pt == K else int(rng.integers(0, 2))
                if pt != K and rng.random() < 0.25:
                    hands[owner, HAND_SLOT[pt]] += 1
                    continue
                t = pt
                if pt in PROMOTE and rng.random() < 0.2:
                    t = PROMOTE[pt]
                board[squares[k]] = 1 + t + 14 * owner
                k += 1
        pos["board"][i] = board
        pos["hands"][i] = hands
        pos["side"][i] = rng.integers(0, 2)
        mp = int(rng.integers(224, 641))
        pos["max_ply"][i] = mp
        pos["ply"][i] = rng.integers(0, mp)
        pos["black_draw_value"][i] = np.float32(rng.random())
        pos["white_draw_value"][i] = np.float32(1.0) - pos["black_draw_value"][i]
    return pos


def startpos(n: int = 1) -> np.ndarray:
    """Hirate start position x n, as reference src/bench/batchsize.cc:47-59 fills its batch.
    Square numbering s = 9*(file-1) + (rank-1) (builder-defined)."""
    pos = np.zeros(n, dtype=POSITION)
    board = np.zeros(81, dtype=np.uint8)

    def put(file, rank, pt, colour):
        board[9 * (file - 1) + (rank - 1)] = 1 + pt + 14 * colour

    back = [L, N, S, G, K, G, S, N, L]
    for f in range(1, 10):
        put(f, 9, back[f - 1], 0)
        put(f, 1, back[9 - f], 1)
        put(f, 7, P, 0)
        put(f, 3, P, 1)
    put(8, 8, B, 0)
    put(2, 8, R, 0)
    put(2, 2, B, 1)
    put(8, 2, R, 1)
    pos["board"][:] = board
    pos["max_ply"][:] = 320
    pos["black_draw_value"][:] = 0.5
    pos["white_draw_value"][:] = 0.5
    return pos


def random_legal_moves(n: int, seed: int = SEED, edge_rows: bool = True):
    """CSR legal-move policy indices: n_i ~ clamp(round(N(80, 35^2)), 1, 593), indices distinct
    uniform draws from [0, 2187).  With edge_rows the first rows are n = 1, 164, 165, 593, 2."""
    rng = np.random.default_rng(seed + 1)
    counts = np.clip(np.rint(rng.normal(80.0, 35.0, size=n)), 1, MAX_LEGAL_MOVES).astype(np.int64)
    if edge_rows:
        for j, v in enumerate([1, 164, 165, MAX_LEGAL_MOVES, 2]):
            if j < n:
                counts[j] = v
    off = np.zeros(n + 1, dtype=np.uint32)
    off[1:] = np.cumsum(counts)
    idx = np.empty(int(off[-1]), dtype=np.uint16)
    for i in range(n):
        idx[off[i]:off[i + 1]] = rng.choice(POLICY_SIZE, size=int(counts[i]), replace=False)
    return off, idx


def random_logits(n: int, seed: int = SEED, special: bool = True):
    """N(0, 3^2) fp32 logits + win/draw in (0,1); with `special`, rows 5.. carry +-inf / NaN."""
    rng = np.random.default_rng(seed + 2)
    policy = (rng.standard_normal((n, POLICY_SIZE)) * 3.0).astype(np.float32)
    win = rng.random(n, dtype=np.float32)
    draw = rng.random(n, dtype=np.float32)
    if special and n > 9:
        policy[5, :] = np.nan
        policy[6, ::7] = np.nan
        policy[7, ::5] = -np.inf
        win[8] = np.nan
        draw[9] = np.nan
    return policy, win, draw


def feature_stack_bytes(n: int) -> int:
    return n * FEATURE_CHANNELS * FEATURE_BITBOARD.itemsize
