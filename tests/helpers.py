"""Independent restatements used to pin the oracle (numpy / torch fp32), plus small utilities."""
import numpy as np


def expand_numpy(fb, n, channels, channels_first=True):
    """Vectorised numpy restatement of reference src/cuda/extractbit.cu:15-39 / :41-68, written
    independently of oracle/nsb_oracle.c (python ints -> no shift pitfalls)."""
    lo = fb["lo"].astype(np.uint64).reshape(n, channels)
    hi = fb["hi"].astype(np.uint64).reshape(n, channels)
    rot = ((hi >> np.uint64(24)) & np.uint64(1)).astype(np.int64)
    val = (hi >> np.uint64(32)).astype(np.uint32)
    t = np.arange(81, dtype=np.int64)
    sq = np.where(rot[..., None] == 1, 80 - t, t)  # [n, c, 81]
    use_hi = sq >= 63
    sh = np.where(use_hi, sq - 63, sq).astype(np.uint64)
    word = np.where(use_hi, hi[..., None], lo[..., None])
    bit = ((word >> sh) & np.uint64(1)).astype(np.uint32)
    out = (bit * val[..., None]).astype(np.uint32)  # [n, c, 81]
    if not channels_first:
        out = np.ascontiguousarray(out.transpose(0, 2, 1))
    return out.view(np.float32)


def split_blob(desc, blob):
    """Canonical blob layout (DESIGN.md §5 / csrc/weights.cc header)."""
    C, IN, H, NB = desc.channels, desc.in_channels, desc.value_hidden, desc.blocks
    o = 0

    def take(shape):
        nonlocal o
        n = int(np.prod(shape))
        a = blob[o:o + n].reshape(shape)
        o += n
        return a

    w = {"stem_w": take((C, IN, 3, 3)), "stem_b": take((C,)), "blocks": []}
    for _ in range(NB):
        w["blocks"].append((take((C, C, 3, 3)), take((C,)), take((C, C, 3, 3)), take((C,))))
    w["pol_w"], w["pol_b"] = take((27, C)), take((27,))
    w["val_w"], w["val_b"] = take((C,)), take((1,))
    w["fc1_w"], w["fc1_b"] = take((H, 81)), take((H,))
    w["fc2_w"], w["fc2_b"] = take((2, H)), take((2,))
    assert o == blob.size
    return w


def forward_torch(desc, blob, planes):
    """fp32 PyTorch forward of the canonical net (TF32 off, CPU): the independent check of
    oracle.forward(emulate_bf16=0).  planes: [n][in_channels][81] fp32."""
    import torch
    import torch.nn.functional as F

    torch.backends.cudnn.allow_tf32 = False
    w = split_blob(desc, blob)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    x = t(planes).reshape(-1, desc.in_channels, 9, 9)
    x = F.relu(F.conv2d(x, t(w["stem_w"]), t(w["stem_b"]), padding=1))
    for (w1, b1, w2, b2) in w["blocks"]:
        y = F.relu(F.conv2d(x, t(w1), t(b1), padding=1))
        x = F.relu(F.conv2d(y, t(w2), t(b2), padding=1) + x)
    pol = F.conv2d(x, t(w["pol_w"]).reshape(27, -1, 1, 1), t(w["pol_b"])).reshape(-1, 2187)
    v = F.relu(F.conv2d(x, t(w["val_w"]).reshape(1, -1, 1, 1), t(w["val_b"]))).reshape(-1, 81)
    h = F.relu(F.linear(v, t(w["fc1_w"]), t(w["fc1_b"])))
    o = torch.sigmoid(F.linear(h, t(w["fc2_w"]), t(w["fc2_b"])))
    return pol.numpy(), o[:, 0].numpy().copy(), o[:, 1].numpy().copy()


def softmax_rows(policy, off, idx):
    out = np.zeros(int(off[-1]), dtype=np.float64)
    for i in range(len(off) - 1):
        g = policy[i, idx[off[i]:off[i + 1]]].astype(np.float64)
        e = np.exp(g - g.max())
        out[off[i]:off[i + 1]] = e / e.sum()
    return out


def drift_metrics(nb, orc, desc, blob, fb, n, got, off, idx):
    """Drift of an executor result `got` = (logits [n,2187], win, draw) against the oracle at both precisions, on the
    bitboards `fb`: the numbers of DESIGN.md §4's drift table and of tests/test_depth_parity.py.
      *_vs_bf16 : against oracle.forward(emulate_bf16=True) - same rounding points as the kernels (bf16 inputs, weights
                  and activations, fp32 accumulation); what is left is accumulation order
      *_vs_fp32 : against oracle.forward(emulate_bf16=False) - the precision bar of the north star (the reference's
                  TensorRT path has fp32 I/O with TF32 allowed, src/infer/trt.cc:144-161): decoded probabilities,
                  KL(fp32 || got) per row, win / draw rate
      oracle_*  : the bf16-emulating oracle against the fp32 oracle, i.e. the share of the drift that is bf16 itself"""
    planes = orc.expand(fb, n)
    pb, wb, db = orc.forward(desc, blob, planes, emulate_bf16=True)
    pf, wf, df = orc.forward(desc, blob, planes, emulate_bf16=False)
    gp, gw, gd = got
    qg, _ = orc.decode(gp, gw, gd, off, idx, nb.DECODE_PROBS)
    qb, _ = orc.decode(pb, wb, db, off, idx, nb.DECODE_PROBS)
    qf, _ = orc.decode(pf, wf, df, off, idx, nb.DECODE_PROBS)
    kl = 0.0
    for i in range(n):
        a, b = qf[off[i]:off[i + 1]].astype(np.float64), qg[off[i]:off[i + 1]].astype(np.float64)
        m = a > 0
        kl = max(kl, float(np.sum(a[m] * np.log(a[m] / np.maximum(b[m], 1e-300)))))
    f = lambda x: float(np.max(np.abs(x))) if x.size else 0.0
    return {
        "layers": 2 * desc.blocks + 2, "n": n, "logit_rms_fp32": float(np.sqrt(np.mean(pf.astype(np.float64) ** 2))),
        "logit_vs_bf16": f(gp - pb), "logit_vs_bf16_mean": float(np.mean(np.abs(gp - pb))),
        "value_vs_bf16": max(f(gw - wb), f(gd - db)),
        "logit_vs_fp32": f(gp - pf), "prob_vs_fp32": f(qg - qf), "kl_vs_fp32": kl,
        "win_vs_fp32": f(gw - wf), "draw_vs_fp32": f(gd - df),
        "oracle_prob_bf16_vs_fp32": f(qb - qf), "oracle_value_bf16_vs_fp32": max(f(wb - wf), f(db - df)),
        "win_spread_fp32": float(np.ptp(wf)) if n > 1 else 0.0, "max_prob_fp32": float(qf.max()) if qf.size else 0.0,
    }
