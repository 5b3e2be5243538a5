/*
 * nsb_oracle.c — CPU oracle for the leaf-evaluation hot path.  TEST INFRASTRUCTURE ONLY:
 * loaded by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs; never by the
 * product library.  See nsb_oracle.h for what is pinned against the reference and what is not.
 */
#define _GNU_SOURCE
#include "nsb_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------ */
/* expand: reference src/cuda/extractbit.cu:15-39 (NCHW) and :41-68 (NHWC)                     */
/* ------------------------------------------------------------------------------------------ */

/* One (plane, output position t) element, exactly the kernel's integer arithmetic:
 *   Rotate = (hi >> 24) & 1; Value = hi >> 32            (extractbit.cu:20-21)
 *   TargetSquare = t*(1-2*Rotate) + 80*Rotate            (:26)
 *   word/shift pick at square 63                         (:30-34)
 *   out = ((word & mask) >> shift) * Value  as int32     (:36-37) */
static inline uint32_t expand_one(uint64_t lo, uint64_t hi, int t) {
    const int rotate = (int)((hi >> 24) & 1u);
    const uint32_t value = (uint32_t)(hi >> 32);
    const int sq = t * (1 - 2 * rotate) + 80 * rotate;
    const int use_hi = sq >= 63;
    const int sh = sq - 63 * use_hi;
    const uint64_t word = use_hi ? hi : lo;
    return (uint32_t)((word >> sh) & 1u) * value;
}

void nsb_oracle_expand(const nsb_feature_bitboard* fb, size_t n, int channels, int channels_first,
                       float* planes) {
    uint32_t* out = (uint32_t*)planes; /* fp32 bit patterns are stored as integers (:80-85) */
    for (size_t b = 0; b < n; ++b) {
        for (int c = 0; c < channels; ++c) {
            const uint64_t lo = fb[b * channels + c].lo, hi = fb[b * channels + c].hi;
            for (int t = 0; t < 81; ++t) {
                const uint32_t v = expand_one(lo, hi, t);
                if (channels_first)
                    out[(b * channels + c) * 81 + t] = v; /* Dest[Index*81 + BitIndex] (:37) */
                else
                    out[(b * 81 + t) * channels + c] = v; /* (:65-66) */
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* pack: position -> 86 feature bitboards (channel order: reference src/evaluate/preset.h:20-66;
 * plane semantics: SURVEY.md App. A.2, builder-defined — libnshogi is not available)          */
/* ------------------------------------------------------------------------------------------ */

static const int kStandMax[7] = {6, 4, 4, 4, 4, 2, 2}; /* P L N S G B R: 26 planes per side */

static inline uint32_t f32_bits(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}

static inline void fb_set(nsb_feature_bitboard* f, uint64_t lo, uint64_t hi18, int rotate,
                          float value) {
    f->lo = lo;
    f->hi = hi18 | ((uint64_t)(rotate & 1) << 24) | ((uint64_t)f32_bits(value) << 32);
}

#define ALL_LO ((1ULL << 63) - 1ULL)
#define ALL_HI 0x3FFFFULL

void nsb_oracle_pack(const nsb_position* pos, size_t n, nsb_feature_bitboard* fb) {
    for (size_t b = 0; b < n; ++b) {
        const nsb_position* p = &pos[b];
        nsb_feature_bitboard* f = &fb[b * NSB_FEATURE_CHANNELS];
        const int me = p->side & 1, op = me ^ 1, rot = me; /* white to move => rotate */
        uint64_t lo[2][14], hi[2][14];
        memset(lo, 0, sizeof lo);
        memset(hi, 0, sizeof hi);
        for (int s = 0; s < 81; ++s) {
            const int code = p->board[s];
            if (!code) continue;
            const int colour = (code - 1) / 14, pt = (code - 1) % 14;
            if (s < 63)
                lo[colour][pt] |= 1ULL << s;
            else
                hi[colour][pt] |= 1ULL << (s - 63);
        }
        int c = 0;
        for (int pt = 0; pt < 14; ++pt) fb_set(&f[c++], lo[me][pt], hi[me][pt], rot, 1.0f);
        for (int pt = 0; pt < 14; ++pt) fb_set(&f[c++], lo[op][pt], hi[op][pt], rot, 1.0f);
        for (int side = 0; side < 2; ++side) {
            const int col = side == 0 ? me : op;
            for (int k = 0; k < 7; ++k)
                for (int j = 1; j <= kStandMax[k]; ++j) {
                    const int on = p->hands[col][k] >= j;
                    fb_set(&f[c++], on ? ALL_LO : 0, on ? ALL_HI : 0, rot, 1.0f);
                }
        }
        fb_set(&f[c++], me == 0 ? ALL_LO : 0, me == 0 ? ALL_HI : 0, rot, 1.0f); /* Black */
        fb_set(&f[c++], me == 1 ? ALL_LO : 0, me == 1 ? ALL_HI : 0, rot, 1.0f); /* White */
        const float maxply = (float)(p->max_ply ? p->max_ply : 1);
        fb_set(&f[c++], ALL_LO, ALL_HI, rot, (float)p->ply / maxply); /* Progress      */
        fb_set(&f[c++], ALL_LO, ALL_HI, rot, 1.0f / maxply);          /* ProgressUnit  */
        const float bd = p->black_draw_value, wd = p->white_draw_value;
        fb_set(&f[c++], ALL_LO, ALL_HI, rot, me == 0 ? bd : wd); /* MyDrawValue */
        fb_set(&f[c++], ALL_LO, ALL_HI, rot, me == 0 ? wd : bd); /* OpDrawValue */
    }
}

/* ------------------------------------------------------------------------------------------ */
/* decode: reference src/mcts/feedworker.cc:56-136, src/selfplay/frame.cc:93-136                */
/* ------------------------------------------------------------------------------------------ */

/* reference src/math/math.h:23-39: bit-pattern NaN test (robust under -ffast-math). */
static inline int isnan_bits(float x) {
    const uint32_t u = f32_bits(x);
    return (u & 0x7F800000u) == 0x7F800000u && (u & 0x007FFFFFu) != 0;
}

/* ml::math::softmax_(x, n, T = 1) (libnshogi, UNPINNED: assumed max-subtracted exp / sum).  Called at
 * feedworker.cc:127 and frame.cc:117. */
static void softmax_inplace(float* x, uint32_t m) {
    if (m == 0) return;
    float mx = x[0];
    for (uint32_t j = 1; j < m; ++j) mx = x[j] > mx ? x[j] : mx;
    float sum = 0.0f;
    for (uint32_t j = 0; j < m; ++j) {
        x[j] = expf(x[j] - mx);
        sum += x[j];
    }
    const float inv = 1.0f / sum;
    for (uint32_t j = 0; j < m; ++j) x[j] *= inv;
}

/* One leaf, MCTS flavour = FeedWorker::feedResult<NaNFallbackEnabled> (feedworker.cc:56-137) as far as the
 * policy row, NaNFound and the cache decision go.  Three separate pieces of the reference:
 *   :58-85   win / draw NaN  -> NaNFound = true; the VALUE is replaced from the parent (restated in
 *            nsb_oracle_value_fallback below); the policy row is NOT touched
 *   :100-103 one child       -> LegalPolicy[0] = 1, no gather, no NaN test of the logit
 *   :105-118 fallback build  -> gather; the first NaN logit sets NaNFound and makes EVERY legal logit 1
 *   :119-126 default build   -> plain gather (NaNs flow into the softmax; NaNFound stays false)
 *   :127     softmax_        :134 cache store iff !NaNFound
 * Returns NaNFound. */
static int feed_result_row(const float* row, const uint16_t* idx, uint32_t m, int fallback, float win, float draw,
                           float* out) {
    int nan_found = 0;
    if (fallback && (isnan_bits(win) || isnan_bits(draw))) nan_found = 1; /* :61-62, :73-74 */
    if (m == 1) {                                                          /* :100-103 */
        out[0] = 1.0f;
        return nan_found;
    }
    if (fallback) { /* :105-118 */
        for (uint32_t j = 0; j < m; ++j) {
            out[j] = row[idx[j]];
            if (isnan_bits(out[j])) {
                nan_found = 1;
                for (uint32_t k = 0; k < m; ++k) out[k] = 1.0f;
                break;
            }
        }
    } else { /* :119-126 */
        for (uint32_t j = 0; j < m; ++j) out[j] = row[idx[j]];
    }
    softmax_inplace(out, m); /* :127 */
    return nan_found;
}

/* One leaf, self-play flavour = Frame::setEvaluation<false> (frame.cc:93-136) up to the Dirichlet mix:
 *   :96-107  gather the raw logits      :110-114 EvalCache->store(raw logits), unconditional
 *   :116-118 softmax_ unless (Gumbel && root): the caller passes that as NSB_ROW_SKIP_SOFTMAX
 * There is no NaN handling in this function of the reference; with the NSB_DECODE_NAN_FALLBACK bit the
 * executor (and this restatement) only REPORT a NaN among logits / win / draw and keep such a row out of the
 * cache - an extension, off by default like the reference's own switch (src/context.h:103). */
static int set_evaluation_row(const float* row, const uint16_t* idx, uint32_t m, int fallback, int row_flags,
                              float win, float draw, float* out, float* logits_out) {
    int any_nan = isnan_bits(win) || isnan_bits(draw);
    for (uint32_t j = 0; j < m; ++j) { /* :96-107 */
        out[j] = row[idx[j]];
        any_nan |= isnan_bits(out[j]);
        if (logits_out) logits_out[j] = out[j];
    }
    if (!(row_flags & NSB_ROW_SKIP_SOFTMAX)) softmax_inplace(out, m); /* :116-118 */
    return fallback && any_nan;
}

void nsb_oracle_decode_ex(const float* policy, const float* win, const float* draw, size_t n,
                          const uint32_t* move_off, const uint16_t* move_idx, int mode, const uint8_t* row_flags,
                          float* legal_out, float* logits_out, uint8_t* nan_flag) {
    const int fallback = (mode & NSB_DECODE_NAN_FALLBACK) != 0, kind = mode & NSB_DECODE_MODE_MASK;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t b = move_off[i], m = move_off[i + 1] - b;
        const float* row = policy + i * NSB_POLICY_SIZE;
        int flag;
        if (kind == NSB_DECODE_PROBS)
            flag = feed_result_row(row, move_idx + b, m, fallback, win[i], draw[i], legal_out + b);
        else if (kind == NSB_DECODE_LOGITS) /* the gather alone: frame.cc:96-107 */
            flag = set_evaluation_row(row, move_idx + b, m, fallback, NSB_ROW_SKIP_SOFTMAX, win[i], draw[i],
                                      legal_out + b, NULL);
        else
            flag = set_evaluation_row(row, move_idx + b, m, fallback, row_flags ? row_flags[i] : 0, win[i], draw[i],
                                      legal_out + b, logits_out ? logits_out + b : NULL);
        if (nan_flag) nan_flag[i] = (uint8_t)flag;
    }
}

void nsb_oracle_decode(const float* policy, const float* win, const float* draw, size_t n,
                       const uint32_t* move_off, const uint16_t* move_idx, int mode,
                       float* legal_out, uint8_t* nan_flag) {
    nsb_oracle_decode_ex(policy, win, draw, n, move_off, move_idx, mode, NULL, legal_out, NULL, nan_flag);
}

/* feedworker.cc:58-85: the NaN fallback of the leaf's VALUE from its parent's running statistics
 * (has_parent == 0: the root, :64-65, :76-77).  Returns NaNFound. */
int nsb_oracle_value_fallback(float* win, float* draw, int has_parent, double parent_win_acc,
                              double parent_draw_acc, uint64_t parent_visits) {
    int nan_found = 0;
    if (isnan_bits(*win)) {
        nan_found = 1;
        *win = has_parent ? (float)(1.0 - parent_win_acc / (double)parent_visits) : 0.5f;
    }
    if (isnan_bits(*draw)) {
        nan_found = 1;
        *draw = has_parent ? (float)(parent_draw_acc / (double)parent_visits) : 0.0f;
    }
    return nan_found;
}

/* frame.cc:121-133: Dirichlet noise at the AlphaZero root of a full search, in double, rounded once. */
void nsb_oracle_dirichlet_mix(float* probs, const double* noise, uint32_t m) {
    const double EPS = 0.25;
    for (uint32_t j = 0; j < m; ++j) probs[j] = (float)((1 - EPS) * (double)probs[j] + EPS * noise[j]);
}

/* ------------------------------------------------------------------------------------------ */
/* forward: fp32 definition of the canonical net                                */
/* ------------------------------------------------------------------------------------------ */

float nsb_oracle_bf16_round(float x) {
    uint32_t u = f32_bits(x);
    if ((u & 0x7F800000u) == 0x7F800000u) return x; /* inf / nan unchanged */
    u += 0x7FFFu + ((u >> 16) & 1u);                /* round to nearest even */
    u &= 0xFFFF0000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}

#define PW 11            /* padded board width  */
#define PSZ (PW * PW + 8) /* padded plane floats (tail slack for the 99-wide window) */

/* 3x3 conv, pad 1, on [cin][9][9] -> [cout][9][9]; h = t / 9, w = t % 9 (NCHW of
 * reference src/infer/trt.cc:144-150).  in_pad holds zero-bordered 11x11 planes. */
static void conv3x3(const float* in_pad, int cin, const float* w, const float* bias, int cout,
                    float* out /* [cout][81] */) {
    for (int co = 0; co < cout; ++co) {
        float acc[9 * PW];
        for (int o = 0; o < 9 * PW; ++o) acc[o] = bias[co];
        const float* wc = w + (size_t)co * cin * 9;
        for (int ci = 0; ci < cin; ++ci) {
            const float* p = in_pad + (size_t)ci * PSZ;
            for (int kh = 0; kh < 3; ++kh)
                for (int kw = 0; kw < 3; ++kw) {
                    const float wv = wc[ci * 9 + kh * 3 + kw];
                    const float* q = p + kh * PW + kw;
                    for (int o = 0; o < 9 * PW; ++o) acc[o] += wv * q[o];
                }
        }
        for (int h = 0; h < 9; ++h)
            for (int x = 0; x < 9; ++x) out[co * 81 + h * 9 + x] = acc[h * PW + x];
    }
}

static void pad_planes(const float* in /* [c][81] */, int c, float* in_pad) {
    memset(in_pad, 0, sizeof(float) * (size_t)c * PSZ);
    for (int ch = 0; ch < c; ++ch)
        for (int h = 0; h < 9; ++h)
            for (int x = 0; x < 9; ++x)
                in_pad[(size_t)ch * PSZ + (h + 1) * PW + (x + 1)] = in[ch * 81 + h * 9 + x];
}

/* What the device does to its INPUT in bf16 mode (csrc/nsb_internal.h kFirstScalarChannel): the planes before channel
 * 82 are 0/1 and exact in bf16; every plane from 82 on (Progress, ProgressUnit, draw values, scores - arbitrary fp32
 * fill values) enters the stem twice: rounded to bf16, and - in a TWIN channel appended behind the real ones, with the
 * same (bf16-exact) weights - as the bf16 rounding of the remainder.  The restatement does exactly that: `stem_w_ext`
 * is the stem weight tensor [C][IN + T][3][3] with the twins' weights repeated. */
#define ORACLE_FIRST_SCALAR_CHANNEL 82
static int stem_twins(int in_channels) {
    return in_channels > ORACLE_FIRST_SCALAR_CHANNEL ? in_channels - ORACLE_FIRST_SCALAR_CHANNEL : 0;
}

typedef struct {
    const nsb_net_desc* net;
    const float *blob, *planes;
    size_t n;
    int emulate_bf16, tid, nthreads;
    float *policy, *win, *draw;
    const float* stem_w_ext; /* bf16 mode with twins only, else NULL */
} fwd_job;

static void* forward_worker(void* arg) {
    const fwd_job* J = (const fwd_job*)arg;
    const nsb_net_desc* net = J->net;
    const float *blob = J->blob, *planes = J->planes;
    const int emulate_bf16 = J->emulate_bf16;
    float *policy = J->policy, *win = J->win, *draw = J->draw;
    const int IN = net->in_channels, C = net->channels, NB = net->blocks, H = net->value_hidden;
    const int T = J->stem_w_ext ? stem_twins(IN) : 0, IN_EXT = IN + T;
    const int maxc = C > IN_EXT ? C : IN_EXT;
    for (size_t b = (size_t)J->tid; b < J->n; b += (size_t)J->nthreads) {
        float* x = (float*)malloc(sizeof(float) * (size_t)maxc * 81);
        float* t = (float*)malloc(sizeof(float) * (size_t)maxc * 81);
        float* y = (float*)malloc(sizeof(float) * (size_t)maxc * 81);
        float* pad = (float*)malloc(sizeof(float) * (size_t)maxc * PSZ);
        const float* w = blob;
        for (int i = 0; i < IN * 81; ++i) {
            const float v = planes[(size_t)b * IN * 81 + i];
            x[i] = emulate_bf16 ? nsb_oracle_bf16_round(v) : v;
        }
        for (int k = 0; k < T; ++k) /* twin of channel 82 + k: bf16 of what the rounding above lost */
            for (int i = 0; i < 81; ++i) {
                const int src = (ORACLE_FIRST_SCALAR_CHANNEL + k) * 81 + i;
                const float v = planes[(size_t)b * IN * 81 + src];
                x[(IN + k) * 81 + i] = isfinite(v) ? nsb_oracle_bf16_round(v - x[src]) : 0.0f;
            }
        /* stem */
        pad_planes(x, IN_EXT, pad);
        conv3x3(pad, IN_EXT, T ? J->stem_w_ext : w, w + (size_t)C * IN * 9, C, t);
        w += (size_t)C * IN * 9 + C;
        for (int i = 0; i < C * 81; ++i) {
            float v = t[i] > 0.f ? t[i] : 0.f;
            x[i] = emulate_bf16 ? nsb_oracle_bf16_round(v) : v;
        }
        /* residual blocks */
        for (int blk = 0; blk < NB; ++blk) {
            pad_planes(x, C, pad);
            conv3x3(pad, C, w, w + (size_t)C * C * 9, C, t);
            w += (size_t)C * C * 9 + C;
            for (int i = 0; i < C * 81; ++i) {
                float v = t[i] > 0.f ? t[i] : 0.f;
                t[i] = emulate_bf16 ? nsb_oracle_bf16_round(v) : v;
            }
            pad_planes(t, C, pad);
            conv3x3(pad, C, w, w + (size_t)C * C * 9, C, y);
            w += (size_t)C * C * 9 + C;
            for (int i = 0; i < C * 81; ++i) {
                float v = y[i] + x[i];
                v = v > 0.f ? v : 0.f;
                x[i] = emulate_bf16 ? nsb_oracle_bf16_round(v) : v;
            }
        }
        /* policy head: conv1x1 C -> 27, plane-major logits (ChannelsFirst, globalconfig.h:20) */
        const float* pw = w;
        const float* pb = w + (size_t)NSB_POLICY_PLANES * C;
        w += (size_t)NSB_POLICY_PLANES * C + NSB_POLICY_PLANES;
        for (int p = 0; p < NSB_POLICY_PLANES; ++p)
            for (int s = 0; s < 81; ++s) {
                float acc = pb[p];
                for (int c = 0; c < C; ++c) acc += pw[p * C + c] * x[c * 81 + s];
                policy[(size_t)b * NSB_POLICY_SIZE + p * 81 + s] = acc;
            }
        /* value head: conv1x1 C -> 1, ReLU, FC 81 -> H, ReLU, FC H -> 2, sigmoid */
        const float* vw = w;
        const float vb = w[C];
        w += C + 1;
        float v81[81];
        for (int s = 0; s < 81; ++s) {
            float acc = vb;
            for (int c = 0; c < C; ++c) acc += vw[c] * x[c * 81 + s];
            v81[s] = acc > 0.f ? acc : 0.f;
        }
        const float* f1w = w;
        const float* f1b = w + (size_t)H * 81;
        w += (size_t)H * 81 + H;
        const float* f2w = w;
        const float* f2b = w + 2 * (size_t)H;
        float o0 = f2b[0], o1 = f2b[1];
        for (int h = 0; h < H; ++h) {
            float acc = f1b[h];
            for (int s = 0; s < 81; ++s) acc += f1w[h * 81 + s] * v81[s];
            acc = acc > 0.f ? acc : 0.f;
            o0 += f2w[h] * acc;
            o1 += f2w[H + h] * acc;
        }
        win[b] = 1.0f / (1.0f + expf(-o0));
        draw[b] = 1.0f / (1.0f + expf(-o1));
        free(x);
        free(t);
        free(y);
        free(pad);
    }
    return NULL;
}

void nsb_oracle_forward(const nsb_net_desc* net, const float* blob, const float* planes, size_t n,
                        int emulate_bf16, float* policy, float* win, float* draw) {
    long nt = sysconf(_SC_NPROCESSORS_ONLN);
    if (nt < 1) nt = 1;
    if (nt > 64) nt = 64;
    if ((size_t)nt > n) nt = (long)(n ? n : 1);
    pthread_t th[64];
    fwd_job jobs[64];
    float* stem_ext = NULL;
    const int IN = net->in_channels, C = net->channels, T = stem_twins(IN);
    if (emulate_bf16 && T > 0) {
        stem_ext = (float*)malloc(sizeof(float) * (size_t)C * (IN + T) * 9);
        for (int co = 0; co < C; ++co)
            for (int ci = 0; ci < IN + T; ++ci) {
                const int src = ci < IN ? ci : ORACLE_FIRST_SCALAR_CHANNEL + (ci - IN);
                memcpy(stem_ext + ((size_t)co * (IN + T) + ci) * 9, blob + ((size_t)co * IN + src) * 9, 9 * sizeof(float));
            }
    }
    for (long t = 0; t < nt; ++t) {
        jobs[t] = (fwd_job){net, blob, planes, n, emulate_bf16, (int)t, (int)nt, policy, win, draw, stem_ext};
        pthread_create(&th[t], NULL, forward_worker, &jobs[t]);
    }
    for (long t = 0; t < nt; ++t) pthread_join(th[t], NULL);
    free(stem_ext);
}

/* ------------------------------------------------------------------------------------------ */
/* Random executor: reference src/infer/random.cc:28-42                                        */
/*   std::mt19937_64 Rng(Seed); static std::uniform_real_distribution<float> Distribution(0,1)  */
/*   (libstdc++ 13 generate_canonical<float,24>: one 64-bit draw, float(u) / 2^64, clamped to   */
/*   nextafter(1,0) when it rounds to 1).                                                        */
/* ------------------------------------------------------------------------------------------ */

struct nsb_oracle_rng {
    uint64_t mt[312];
    int idx;
};

nsb_oracle_rng* nsb_oracle_rng_create(uint64_t seed) {
    nsb_oracle_rng* r = (nsb_oracle_rng*)malloc(sizeof *r);
    r->mt[0] = seed;
    for (int i = 1; i < 312; ++i)
        r->mt[i] = 6364136223846793005ULL * (r->mt[i - 1] ^ (r->mt[i - 1] >> 62)) + (uint64_t)i;
    r->idx = 312;
    return r;
}

void nsb_oracle_rng_destroy(nsb_oracle_rng* r) { free(r); }

static inline uint64_t mt_next(nsb_oracle_rng* r) {
    if (r->idx >= 312) {
        const uint64_t UM = 0xFFFFFFFF80000000ULL, LM = 0x7FFFFFFFULL, A = 0xB5026F5AA96619E9ULL;
        for (int i = 0; i < 312; ++i) {
            const uint64_t x = (r->mt[i] & UM) | (r->mt[(i + 1) % 312] & LM);
            r->mt[i] = r->mt[(i + 156) % 312] ^ (x >> 1) ^ ((x & 1ULL) ? A : 0ULL);
        }
        r->idx = 0;
    }
    uint64_t y = r->mt[r->idx++];
    y ^= (y >> 29) & 0x5555555555555555ULL;
    y ^= (y << 17) & 0x71D67FFFEDA60000ULL;
    y ^= (y << 37) & 0xFFF7EEE000000000ULL;
    y ^= y >> 43;
    return y;
}

static inline float canonical_f32(nsb_oracle_rng* r) {
    const float sum = (float)mt_next(r);
    float ret = sum / 18446744073709551616.0f;
    if (ret >= 1.0f) ret = nextafterf(1.0f, 0.0f);
    return ret;
}

void nsb_oracle_random_fill(nsb_oracle_rng* r, size_t n, float* policy, float* win, float* draw) {
    for (size_t i = 0; i < n; ++i) { /* random.cc:33-41 */
        for (int j = 0; j < NSB_POLICY_SIZE; ++j) policy[i * NSB_POLICY_SIZE + j] = canonical_f32(r);
        win[i] = canonical_f32(r);
        draw[i] = canonical_f32(r);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* CPU path of config 1: pack [+ expand] + Random fill + decode, N threads.    */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    const nsb_position* pos;
    size_t n_pos;
    const uint32_t* move_off;
    const uint16_t* move_idx;
    int batch, batches, with_expand, tid;
    nsb_fill_fn fill;
    nsb_fill_make_fn mk;
    double checksum;
} cpu_job;

static void port_fill(void* h, size_t n, float* p, float* w, float* d) {
    nsb_oracle_random_fill((nsb_oracle_rng*)h, n, p, w, d);
}
static void* port_make(uint64_t seed) { return nsb_oracle_rng_create(seed); }

static void* cpu_worker(void* arg) {
    cpu_job* j = (cpu_job*)arg;
    const int B = j->batch;
    nsb_feature_bitboard* fb = (nsb_feature_bitboard*)malloc(sizeof(*fb) * B * NSB_FEATURE_CHANNELS);
    float* planes = j->with_expand ? (float*)malloc(sizeof(float) * B * NSB_FEATURE_CHANNELS * 81) : NULL;
    float* policy = (float*)malloc(sizeof(float) * B * NSB_POLICY_SIZE);
    float* win = (float*)malloc(sizeof(float) * B);
    float* draw = (float*)malloc(sizeof(float) * B);
    float* legal = (float*)malloc(sizeof(float) * B * NSB_MAX_LEGAL_MOVES);
    uint32_t* off = (uint32_t*)malloc(sizeof(uint32_t) * (B + 1));
    uint8_t* flag = (uint8_t*)malloc(B);
    void* h = j->mk((uint64_t)j->tid); /* reference: Random(Seed) per executor */
    double cs = 0.0;
    size_t cursor = ((size_t)j->tid * 7919u) % j->n_pos;
    for (int it = 0; it < j->batches; ++it) {
        if (cursor + B > j->n_pos) cursor = 0;
        const size_t s = cursor;
        cursor += B;
        nsb_oracle_pack(j->pos + s, B, fb);
        if (planes) nsb_oracle_expand(fb, B, NSB_FEATURE_CHANNELS, 1, planes);
        j->fill(h, B, policy, win, draw);
        const uint32_t base = j->move_off[s];
        for (int i = 0; i <= B; ++i) off[i] = j->move_off[s + i] - base;
        nsb_oracle_decode(policy, win, draw, B, off, j->move_idx + base, NSB_DECODE_PROBS, legal, flag);
        cs += legal[0] + win[B - 1] + (planes ? planes[80] : 0.f);
    }
    j->checksum = cs;
    free(fb);
    free(planes);
    free(policy);
    free(win);
    free(draw);
    free(legal);
    free(off);
    free(flag);
    return NULL;
}

double nsb_oracle_cpu_path(const nsb_position* pos, size_t n_pos, const uint32_t* move_off,
                           const uint16_t* move_idx, int batch, int threads,
                           int batches_per_thread, int with_expand, nsb_fill_fn fill,
                           nsb_fill_make_fn mk, double* seconds_out) {
    if (threads < 1) threads = 1;
    if ((size_t)batch > n_pos) return -1.0;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    cpu_job* jobs = (cpu_job*)calloc(threads, sizeof(cpu_job));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < threads; ++t) {
        jobs[t] = (cpu_job){pos, n_pos, move_off, move_idx, batch, batches_per_thread, with_expand,
                            t, fill ? fill : port_fill, (fill && mk) ? mk : port_make, 0.0};
        pthread_create(&th[t], NULL, cpu_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double sec = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    if (seconds_out) *seconds_out = sec;
    free(th);
    free(jobs);
    return (double)threads * batches_per_thread * batch / sec;
}


/* ---- evaluation cache: restatement of reference src/mcts/evalcache.{h,cc} ---------------------- */
#define NSB_ORACLE_CACHE_ROW 164 /* evalcache.h:26 MAX_CACHE_MOVES_COUNT */
#define NSB_ORACLE_BUNDLE 3      /* evalcache.h:44 CACHE_BUNDLE_SIZE     */

typedef struct cache_elem {
    int used;
    uint64_t hash;
    uint32_t n;
    float row[NSB_ORACLE_CACHE_ROW];
    float win, draw;
    struct cache_elem *next, *prev;
} cache_elem;

struct nsb_oracle_cache {
    uint64_t num_bundles;
    cache_elem* mem;
    cache_elem** head;
};

nsb_oracle_cache* nsb_oracle_cache_create(uint64_t num_bundles) { /* evalcache.cc:17-47 */
    nsb_oracle_cache* c = (nsb_oracle_cache*)calloc(1, sizeof *c);
    c->num_bundles = num_bundles;
    c->mem = (cache_elem*)calloc((size_t)num_bundles * NSB_ORACLE_BUNDLE, sizeof(cache_elem));
    c->head = (cache_elem**)calloc((size_t)num_bundles, sizeof(cache_elem*));
    for (uint64_t b = 0; b < num_bundles; ++b) {
        cache_elem* e = c->mem + b * NSB_ORACLE_BUNDLE;
        c->head[b] = e;
        for (int j = 0; j < NSB_ORACLE_BUNDLE; ++j) {
            e[j].used = 0;
            e[j].prev = j ? &e[j - 1] : NULL;
            e[j].next = j + 1 < NSB_ORACLE_BUNDLE ? &e[j + 1] : NULL;
        }
    }
    return c;
}

void nsb_oracle_cache_destroy(nsb_oracle_cache* c) {
    if (!c) return;
    free(c->mem);
    free(c->head);
    free(c);
}

/* The "Reorder" blocks of evalcache.cc:75-86,96-107,146-157, statement for statement.  Note what they
 * do NOT do: the old head's Prev is left NULL.  An element that has been the head therefore looks
 * like the head forever after (its Prev is only repaired when its predecessor is moved away), and
 * `Prev != nullptr` - the guard of every reorder - keeps it where it is.  This is the reference's
 * observable replacement policy, so it is restated, not repaired. */
static void cache_to_front(cache_elem** head, cache_elem* e) {
    if (e->prev == NULL) return;
    e->prev->next = e->next;
    if (e->next) e->next->prev = e->prev;
    e->next = *head;
    e->prev = NULL;
    *head = e;
}

int nsb_oracle_cache_store(nsb_oracle_cache* c, uint64_t hash, uint32_t n, const float* row, float win, float draw) {
    if (n > NSB_ORACLE_CACHE_ROW) return 0; /* :51-53 */
    cache_elem** head = &c->head[hash % c->num_bundles];
    cache_elem* e = *head;
    for (;;) {
        if (!e->used) break;                    /* :66-68 */
        if (e->hash == hash && e->n == n) {     /* :70-88: already there, refresh recency only */
            cache_to_front(head, e);
            return 1;
        }
        if (e->next == NULL) break;             /* :90-92: overwrite the least recent entry */
        e = e->next;
    }
    cache_to_front(head, e);
    e->used = 1;
    e->hash = hash;
    e->n = n;
    memcpy(e->row, row, sizeof(float) * n);
    e->win = win;
    e->draw = draw;
    return 1;
}

int nsb_oracle_cache_load(nsb_oracle_cache* c, uint64_t hash, uint32_t expected_n, float* row, float* win, float* draw) {
    cache_elem** head = &c->head[hash % c->num_bundles];
    for (cache_elem* e = *head; e != NULL; e = e->next) {
        if (!e->used) break;                    /* :136-138 */
        if (e->hash == hash) {                  /* :140-160 */
            const int usable = e->n == expected_n; /* searchworker.cc:546 */
            if (usable) {
                memcpy(row, e->row, sizeof(float) * e->n);
                *win = e->win;
                *draw = e->draw;
            }
            cache_to_front(head, e);
            return usable;
        }
    }
    return 0;
}
