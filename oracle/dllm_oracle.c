/*
 * dllm_oracle.c — CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY
 * (see dllm_oracle.h for the rules and the parity status).
 *
 * Build: gcc -O3 -mavx2 -ffp-contract=off -fno-fast-math (oracle/Makefile).  No FMA contraction,
 * no reassociation: Rust never contracts a*b+c and the f32 operation order below is the
 * reference's.  Rust semantics restated explicitly:
 *   f32::max/min      -> fmaxf/fminf (NaN-ignoring)
 *   f32::clamp        -> rs_clampf   (NaN-propagating)
 *   f32::round        -> roundf      (half away from zero)
 *   `as u8` / `as i32`-> saturating, NaN -> 0, toward zero
 */
#include "dllm_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---- Rust cast / clamp semantics ---- */
static inline uint8_t rs_f32_as_u8(float v) {
    if (v != v) return 0;
    if (v <= 0.0f) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)v; /* in range: C truncates toward zero like Rust */
}
static inline int32_t rs_f32_as_i32(float v) {
    if (v != v) return 0;
    if (v <= -2147483648.0f) return INT32_MIN;
    if (v >= 2147483648.0f) return INT32_MAX;
    return (int32_t)v;
}
static inline float rs_clampf(float x, float lo, float hi) {
    /* core::f32::clamp: `if x < min {min} else if x > max {max} else {x}` — NaN falls through */
    if (x < lo) x = lo;
    if (x > hi) x = hi;
    return x;
}
static inline int32_t rs_clampi(int32_t x, int32_t lo, int32_t hi) {
    return x < lo ? lo : (x > hi ? hi : x);
}

/* ===================== Quantizer B ===================== */
/* diffuse-llm-rs/src/quantization.rs:38-68 */
static void b_params(const float *x, size_t n, uint8_t bits, float *scale_out, uint8_t *zp_out) {
    float max_val = -INFINITY, min_val = INFINITY;           /* :41-46 */
    for (size_t i = 0; i < n; ++i) max_val = fmaxf(max_val, x[i]);
    for (size_t i = 0; i < n; ++i) min_val = fminf(min_val, x[i]);
    const float q_min = 0.0f;
    const float q_max = (float)(1u << bits) - 1.0f;          /* :50 */
    float scale = (max_val - min_val) / (q_max - q_min);     /* :52 */
    if (scale == 0.0f) scale = 1.0f;                         /* :53 */
    float zp = q_min - min_val / scale;                      /* :55 */
    *zp_out = rs_f32_as_u8(roundf(rs_clampf(zp, q_min, q_max))); /* :56 */
    *scale_out = scale;
}

void orc_quantize_codes_b(const float *x, size_t n, uint8_t bits, float scale, float zp,
                          uint8_t *codes) {
    const int32_t hi = (int32_t)((1u << bits) - 1u);
    for (size_t i = 0; i < n; ++i) {                         /* :59-65 */
        float v = roundf((x[i] / scale) + zp);
        codes[i] = (uint8_t)rs_clampi(rs_f32_as_i32(v), 0, hi);
    }
}

int32_t orc_quantize_tensor(const float *x, size_t n, uint8_t bits,
                            uint8_t *codes, float *scale, float *zero_point) {
    if (bits < 1 || bits > 8) return ORC_ERR_INVALID_PARAMS; /* :39 assert! */
    uint8_t zp;
    b_params(x, n, bits, scale, &zp);
    orc_quantize_codes_b(x, n, bits, *scale, (float)zp, codes);
    *zero_point = (float)zp;                                 /* :67 */
    return ORC_OK;
}

void orc_dequantize_tensor(const uint8_t *codes, size_t n, float scale, float zero_point,
                           float *out) {
    for (size_t i = 0; i < n; ++i) out[i] = ((float)codes[i] - zero_point) * scale; /* :81-85 */
}

float orc_compression_ratio(size_t numel, size_t data_len, uint8_t bits) {
    size_t original = numel * 4;                              /* :121 */
    size_t compressed = (data_len * (size_t)bits + 7) / 8;    /* :122 */
    return (float)original / (float)compressed;               /* :123 */
}

int32_t orc_quantize_weight_grouped(const float *w, size_t K, size_t N, uint8_t bits,
                                    size_t group, uint8_t *codes, float *scales, float *zps) {
    if (bits < 1 || bits > 8) return ORC_ERR_INVALID_PARAMS;
    if (group == 0 || K % group != 0) return ORC_ERR_SHAPE;
    float *col = (float *)malloc(group * sizeof(float));
    uint8_t *cc = (uint8_t *)malloc(group);
    for (size_t g = 0; g < K / group; ++g)
        for (size_t nn = 0; nn < N; ++nn) {
            for (size_t j = 0; j < group; ++j) col[j] = w[(g * group + j) * N + nn];
            float s; uint8_t zp;
            b_params(col, group, bits, &s, &zp);
            orc_quantize_codes_b(col, group, bits, s, (float)zp, cc);
            for (size_t j = 0; j < group; ++j) codes[(g * group + j) * N + nn] = cc[j];
            scales[g * N + nn] = s;
            zps[g * N + nn] = (float)zp;
        }
    free(col); free(cc);
    return ORC_OK;
}

void orc_dequantize_weight_grouped(const uint8_t *codes, size_t K, size_t N, size_t group,
                                   const float *scales, const float *zps, float *w) {
    for (size_t k = 0; k < K; ++k) {
        const size_t g = k / group;
        for (size_t nn = 0; nn < N; ++nn)
            w[k * N + nn] = ((float)codes[k * N + nn] - zps[g * N + nn]) * scales[g * N + nn];
    }
}

/* ===================== Quantizer A ===================== */
/* quantization/src/quantize.rs:111-124 (quantize_value), :127-154 (quantize_tensor) */
int32_t orc_quantize_a(const float *x, size_t n, int32_t qtype, float scale, int32_t zero_point,
                       uint8_t *codes) {
    float lo, hi;
    switch (qtype) {                                          /* :139-144 */
        case ORC_QT_INT8:   lo = -128.0f; hi = 127.0f; break;
        case ORC_QT_INT4:   lo = -8.0f;   hi = 7.0f;   break;
        case ORC_QT_BINARY: lo = 0.0f;    hi = 1.0f;   break;
        case ORC_QT_FLOAT8: lo = -127.0f; hi = 127.0f; break;
        default: return ORC_ERR_INVALID_PARAMS;
    }
    const float zp = (float)zero_point;                       /* :136 */
    for (size_t i = 0; i < n; ++i) {
        float q = roundf(fminf(fmaxf(x[i] / scale + zp, lo), hi)); /* :119-122 */
        codes[i] = rs_f32_as_u8(q);                           /* :150 `q as u8` */
    }
    return ORC_OK;
}

void orc_dequantize_a(const uint8_t *codes, size_t n, float scale, int32_t zero_point, float *out) {
    const float zp = (float)zero_point;
    for (size_t i = 0; i < n; ++i) out[i] = ((float)codes[i] - zp) * scale; /* :179-181 */
}

/* quantization/src/calibrate.rs:72-110 */
int32_t orc_calibrate_params(float min, float max, size_t total_samples, uint8_t bits,
                             int32_t symmetric, float *scale, int32_t *zero_point) {
    if (total_samples == 0) return ORC_ERR_INVALID_PARAMS;    /* :73-75 */
    const float num_levels = (float)(1u << bits);             /* 2u32.pow(bits) as f32, :77 */
    const float range = max - min;                            /* :78 */
    if (range <= 1.1920929e-07f) { *scale = 1.0f; *zero_point = 0; return ORC_OK; } /* :80-88 */
    if (symmetric) {
        float max_abs = fmaxf(fabsf(max), fabsf(min));        /* :91 */
        *scale = max_abs * 2.0f / (num_levels - 1.0f);        /* :92 */
        *zero_point = rs_f32_as_i32(num_levels / 2.0f - 1.0f);/* :98 */
    } else {
        *scale = range / (num_levels - 1.0f);                 /* :94 */
        *zero_point = rs_f32_as_i32(roundf(-min / *scale));   /* :100 */
    }
    return ORC_OK;
}

/* ===================== Quantizer C ===================== */
float orc_bitquantizer_scale_c(uint8_t bits) {
    return 1.0f / (float)(int32_t)((1u << bits) - 1u);        /* prefill-kvquant-rs/lib.rs:105 */
}

int32_t orc_quantize_c(const float *x, size_t n, uint8_t bits, float scale, float zp,
                       uint8_t *codes) {
    if (bits > 30) return ORC_ERR_INVALID_PARAMS;             /* `1 << bits` overflows i32 */
    const float max_val = (float)(int32_t)((1u << bits) - 1u);/* :40 */
    for (size_t i = 0; i < n; ++i) {
        float scaled = (x[i] - zp) / scale;                   /* :42 */
        codes[i] = rs_f32_as_u8(rs_clampf(scaled, 0.0f, max_val)); /* :43 */
    }
    return ORC_OK;
}

void orc_dequantize_cd(const uint8_t *codes, size_t n, float scale, float zp, float *out) {
    for (size_t i = 0; i < n; ++i) out[i] = (float)codes[i] * scale + zp; /* :49-51; prefill_kv.rs:62-66 */
}

int32_t orc_kvquant_quantize_vectors(const float *emb, size_t nvec, size_t elems_per_vec,
                                     const uint8_t *cfg_bits, size_t ncfg,
                                     const uint8_t *bits, size_t nbits, uint8_t *codes) {
    if (nbits == 0) return ORC_OK; /* zip with an empty cycle yields nothing (:132) */
    for (size_t v = 0; v < nvec; ++v) {
        const uint8_t b = bits[v % nbits];                    /* :132 cycle */
        const size_t qi = (size_t)b / 2;                      /* :133 */
        if (qi >= ncfg) return ORC_ERR_INDEX;
        const float scale = orc_bitquantizer_scale_c(cfg_bits[qi]);
        int32_t rc = orc_quantize_c(emb + v * elems_per_vec, elems_per_vec, b, scale, 0.0f,
                                    codes + v * elems_per_vec);
        if (rc) return rc;
    }
    return ORC_OK;
}

/* ===================== Quantizer D ===================== */
/* diffusion_prefill/src/prefill_kv.rs:104-121 with BitQuantizer::quantize :53-59 */
int32_t orc_quantize_d_rows(const float *x, size_t rows, size_t dim,
                            const uint8_t *bits, size_t nbits,
                            uint8_t *codes, float *scales, float *zps) {
    if (nbits == 0) return ORC_ERR_INVALID_PARAMS;            /* i % 0 panics, fusion_ann.rs:58 */
    for (size_t r = 0; r < rows; ++r) {
        const uint8_t b = bits[r % nbits];
        if (b > 31) return ORC_ERR_INVALID_PARAMS;            /* 1u32 << bits overflow */
        const float *v = x + r * dim;
        float min_val = INFINITY, max_val = -INFINITY;        /* :105-106 */
        for (size_t i = 0; i < dim; ++i) min_val = fminf(min_val, v[i]);
        for (size_t i = 0; i < dim; ++i) max_val = fmaxf(max_val, v[i]);
        const float levels = (float)((1u << b) - 1u);
        const float scale = (max_val - min_val) / levels;     /* :107 */
        const float zp = min_val;                             /* :108 */
        for (size_t i = 0; i < dim; ++i) {                    /* :55-58 */
            float scaled = rs_clampf((v[i] - zp) / scale, 0.0f, levels);
            codes[r * dim + i] = rs_f32_as_u8(scaled);
        }
        scales[r] = scale;
        zps[r] = zp;
    }
    return ORC_OK;
}

void orc_dequantize_d_rows(const uint8_t *codes, size_t rows, size_t dim,
                           const float *scales, const float *zps, float *out) {
    for (size_t r = 0; r < rows; ++r)
        orc_dequantize_cd(codes + r * dim, dim, scales[r], zps[r], out + r * dim);
}

/* ===================== pack / unpack ===================== */
size_t orc_packed_len(size_t n, uint8_t bits) { return (n * (size_t)bits + 7) / 8; }

int32_t orc_pack(const uint8_t *codes, size_t n, uint8_t bits, uint8_t *packed) {
    if (!(bits == 1 || bits == 2 || bits == 4 || bits == 8)) return ORC_ERR_INVALID_PARAMS;
    const size_t per = 8 / bits;
    const uint8_t mask = (uint8_t)((1u << bits) - 1u);
    memset(packed, 0, orc_packed_len(n, bits));
    for (size_t i = 0; i < n; ++i)
        packed[i / per] |= (uint8_t)((codes[i] & mask) << ((i % per) * bits));
    return ORC_OK;
}

int32_t orc_unpack(const uint8_t *packed, size_t n, uint8_t bits, uint8_t *codes) {
    if (!(bits == 1 || bits == 2 || bits == 4 || bits == 8)) return ORC_ERR_INVALID_PARAMS;
    const size_t per = 8 / bits;
    const uint8_t mask = (uint8_t)((1u << bits) - 1u);
    for (size_t i = 0; i < n; ++i)
        codes[i] = (uint8_t)((packed[i / per] >> ((i % per) * bits)) & mask);
    return ORC_OK;
}

/* ===================== Linear ===================== */
/* diffuse-llm-rs/src/lib.rs:812  x.dot(&self.weights) + &self.bias.  The reference's sgemm
 * (matrixmultiply) is un-vendored: its accumulation order is unpinned.  This is the plain
 * sequential-k f32 form; orc_linear_f64 is the truth tolerances are stated against. */
static void linear_f32_cols(const float *x, const float *w, const float *bias,
                            size_t M, size_t K, size_t N, size_t n_begin, size_t n_end, float *y) {
    /* Cache-blocked, but every y[m][n] is still the plain sequential-k f32 sum
     * ((0 + x0*w0) + x1*w1) + ... with separate multiply and add, then + bias: blocking only
     * changes which (m, n) are in flight, not the order of any single accumulation. */
    enum { MB = 8, NB = 128 };
    float acc[MB][NB];
    for (size_t n0 = n_begin; n0 < n_end; n0 += NB) {
        const size_t nb = n_end - n0 < NB ? n_end - n0 : NB;
        for (size_t m0 = 0; m0 < M; m0 += MB) {
            const size_t mb = M - m0 < MB ? M - m0 : MB;
            for (size_t mi = 0; mi < mb; ++mi)
                for (size_t j = 0; j < nb; ++j) acc[mi][j] = 0.0f;
            for (size_t k = 0; k < K; ++k) {
                const float *wr = w + k * N + n0;
                for (size_t mi = 0; mi < mb; ++mi) {
                    const float xv = x[(m0 + mi) * K + k];
                    float *a = acc[mi];
                    for (size_t j = 0; j < nb; ++j) a[j] = a[j] + xv * wr[j];
                }
            }
            for (size_t mi = 0; mi < mb; ++mi) {
                float *yr = y + (m0 + mi) * N + n0;
                for (size_t j = 0; j < nb; ++j) yr[j] = bias ? acc[mi][j] + bias[n0 + j] : acc[mi][j];
            }
        }
    }
}

void orc_linear_f32(const float *x, const float *w, const float *bias,
                    size_t M, size_t K, size_t N, float *y) {
    linear_f32_cols(x, w, bias, M, K, N, 0, N, y);
}

void orc_linear_f64(const float *x, const float *w, const float *bias,
                    size_t M, size_t K, size_t N, double *y) {
    for (size_t m = 0; m < M; ++m) {
        double *yr = y + m * N;
        for (size_t nn = 0; nn < N; ++nn) yr[nn] = 0.0;
        for (size_t k = 0; k < K; ++k) {
            const double xv = (double)x[m * K + k];
            const float *wr = w + k * N;
            for (size_t nn = 0; nn < N; ++nn) yr[nn] += xv * (double)wr[nn];
        }
        if (bias) for (size_t nn = 0; nn < N; ++nn) yr[nn] += (double)bias[nn];
    }
}

struct lin_job { const float *x, *w, *bias; size_t M, K, N, n0, n1; float *y; };
static void *lin_worker(void *p) {
    struct lin_job *j = (struct lin_job *)p;
    if (j->n1 > j->n0) linear_f32_cols(j->x, j->w, j->bias, j->M, j->K, j->N, j->n0, j->n1, j->y);
    return NULL;
}
/* threads split the OUTPUT COLUMNS (each streams its own panel of W); every y[m][n] keeps the
 * same sequential-k sum, so the result is bit-identical to orc_linear_f32 */
void orc_linear_f32_mt(const float *x, const float *w, const float *bias,
                       size_t M, size_t K, size_t N, float *y, int threads) {
    if (threads < 1) threads = 1;
    size_t chunks = (N + 127) / 128;
    if ((size_t)threads > chunks) threads = (int)(chunks ? chunks : 1);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    struct lin_job *jobs = (struct lin_job *)malloc(sizeof(struct lin_job) * (size_t)threads);
    for (int t = 0; t < threads; ++t) {
        size_t c0 = chunks * (size_t)t / (size_t)threads, c1 = chunks * (size_t)(t + 1) / (size_t)threads;
        size_t n0 = c0 * 128, n1 = c1 * 128 < N ? c1 * 128 : N;
        jobs[t] = (struct lin_job){x, w, bias, M, K, N, n0, n1, y};
        pthread_create(&th[t], NULL, lin_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

void orc_linear_i8_exact(const uint8_t *qx, int32_t zx, const uint8_t *qw, int32_t zw,
                         size_t M, size_t K, size_t N, int64_t *acc) {
    for (size_t m = 0; m < M; ++m)
        for (size_t nn = 0; nn < N; ++nn) {
            int64_t s = 0;
            for (size_t k = 0; k < K; ++k)
                s += (int64_t)((int32_t)qx[m * K + k] - zx) * (int64_t)((int32_t)qw[k * N + nn] - zw);
            acc[m * N + nn] = s;
        }
}

/* ===================== schedules, p_sample ===================== */
/* diffuse-llm-rs/src/lib.rs:554-593 */
int32_t orc_beta_schedule(int32_t kind, size_t T, float beta_start, float beta_end, float *betas) {
    if (T == 0) return ORC_ERR_INVALID_PARAMS;
    const float PI_F = 3.14159265358979323846f;               /* std::f32::consts::PI */
    for (size_t t = 0; t < T; ++t) {
        switch (kind) {
            case ORC_BETA_LINEAR:                             /* :560-563 */
                betas[t] = beta_start + (beta_end - beta_start) * (float)t / (float)(T - 1);
                break;
            case ORC_BETA_QUADRATIC: {                        /* :569-573 */
                float t_norm = (float)t / (float)(T - 1);
                betas[t] = beta_start + (beta_end - beta_start) * t_norm * t_norm;
                break;
            }
            case ORC_BETA_COSINE: {                           /* :578-587 */
                const float s = 0.008f;
                float t_norm = (float)t / (float)T;
                float c_t = cosf((t_norm + s) / (1.0f + s) * PI_F / 2.0f);
                float f_t = c_t * c_t;                        /* powi(2) */
                float c_0 = cosf(s / (1.0f + s) * PI_F / 2.0f);
                float f_0 = c_0 * c_0;
                betas[t] = fminf(1.0f - (f_t / f_0), 0.999f);
                break;
            }
            default: return ORC_ERR_INVALID_PARAMS;
        }
    }
    return ORC_OK;
}

static float alpha_bar_at(const float *betas, size_t i) {
    /* alpha_bars[0] = 1; alpha_bars[i] = alpha_bars[i-1] * alphas[i-1]  (lib.rs:1162-1165) */
    float ab = 1.0f;
    for (size_t j = 1; j <= i; ++j) ab = ab * (1.0f - betas[j - 1]);
    return ab;
}

void orc_p_sample_coeffs(const float *betas, size_t T, size_t t, float *c1, float *c2, float *std) {
    const size_t ti = t < T - 1 ? t : T - 1;                  /* :1174 */
    const float ab_t = alpha_bar_at(betas, ti);
    const float beta_t = betas[ti];
    const float alpha_t = 1.0f - betas[ti];
    const float ab_prev = ti > 0 ? alpha_bar_at(betas, ti - 1) : 1.0f; /* :1180-1184 */
    *c1 = (sqrtf(ab_prev) * beta_t) / (1.0f - ab_t);          /* :1189-1190 */
    *c2 = (sqrtf(alpha_t) * (1.0f - ab_prev)) / (1.0f - ab_t);/* :1191-1192, alphas read as alpha_t */
    const float variance = ((1.0f - ab_prev) / (1.0f - ab_t)) * beta_t; /* :1208 */
    *std = sqrtf(variance);                                   /* :1209 */
}

void orc_p_sample(const float *x_t, const float *noise_pred, const float *z,
                  const size_t *t, size_t batch, size_t feat,
                  const float *betas, size_t T, int32_t guard_t0, float *x_prev) {
    const int add_noise = (z != NULL) && batch > 0 && t[0] > 0; /* :1199-1205 */
    for (size_t b = 0; b < batch; ++b) {
        float c1, c2, sd;
        orc_p_sample_coeffs(betas, T, t[b], &c1, &c2, &sd);
        const size_t ti = t[b] < T - 1 ? t[b] : T - 1;
        const int degenerate = guard_t0 && (1.0f - alpha_bar_at(betas, ti)) == 0.0f;
        for (size_t i = 0; i < feat; ++i) {
            const size_t o = b * feat + i;
            if (degenerate) { x_prev[o] = x_t[o]; continue; }
            float mean = c1 * x_t[o] + c2 * noise_pred[o];    /* :1195-1196 */
            float nz = add_noise ? z[o] : 0.0f;
            x_prev[o] = mean + sd * nz;                       /* :1212 */
        }
    }
}

/* diffuse-llm-rs/src/lib.rs:1100-1137 (noise supplied by the caller: the reference's own draw is an unseeded thread_rng) */
void orc_add_noise(const float *x_start, const float *noise, const size_t *t, size_t batch, size_t feat,
                   const float *betas, size_t T, float *noisy) {
    for (size_t b = 0; b < batch; ++b) {
        const size_t ti = t[b] < T - 1 ? t[b] : T - 1;        /* :1123 */
        const float ab = alpha_bar_at(betas, ti);             /* :1116-1124 */
        const float sa = sqrtf(ab);                           /* :1131 */
        const float sd = sqrtf(1.0f - ab);                    /* :1132 */
        for (size_t i = 0; i < feat; ++i) {
            const size_t o = b * feat + i;
            const float mean = x_start[o] * sa;               /* :1131 */
            noisy[o] = mean + noise[o] * sd;                  /* :1133 */
        }
    }
}

/* ===================== "dllm_noise v1": the counter-based normal generator of the seeded sampling loop =====================
 * Defined by this build (the reference's noise is an unseeded thread_rng, lib.rs:875-878/:1201, and cannot be reproduced);
 * specification in diffusion-llm-rs_b200/csrc/noise.cuh.  Every step is exact or ONE correctly rounded f32 operation in a
 * fixed order, polynomials are evaluated with fmaf: this restatement and the CUDA kernels agree bit for bit. */
static uint64_t dn_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static void dn_sincos_poly(float phi, float *s, float *c) {
    const float p2 = phi * phi;
    float ps = fmaf(p2, 2.75573192e-6f, -1.98412698e-4f);
    ps = fmaf(p2, ps, 8.33333333e-3f);
    ps = fmaf(p2, ps, -1.66666667e-1f);
    ps = fmaf(p2, ps, 1.0f);
    *s = phi * ps;
    float pc = fmaf(p2, -2.75573192e-7f, 2.48015873e-5f);
    pc = fmaf(p2, pc, -1.38888889e-3f);
    pc = fmaf(p2, pc, 4.16666667e-2f);
    pc = fmaf(p2, pc, -0.5f);
    *c = fmaf(p2, pc, 1.0f);
}
static void dn_normal_pair(uint64_t key, uint64_t p, float *z0, float *z1) {
    const uint64_t r = dn_splitmix64(key + p);
    const uint32_t a = (uint32_t)(r >> 41), b = (uint32_t)(r >> 17) & 0xFFFFFFu;
    const uint32_t v = 2u * a + 1u;
    int e = 31 - __builtin_clz(v);
    union { uint32_t u; float f; } pw;
    pw.u = (uint32_t)(127 - e) << 23;
    float m = (float)v * pw.f;
    if (m > 1.41421356f) { m = m * 0.5f; e += 1; }
    const float num = m - 1.0f, den = m + 1.0f;
    const float t = num / den;
    const float t2 = t * t;
    float pl = fmaf(t2, 1.11111111e-1f, 1.42857143e-1f);
    pl = fmaf(t2, pl, 0.2f);
    pl = fmaf(t2, pl, 3.33333343e-1f);
    pl = fmaf(t2, pl, 1.0f);
    const float tt = 2.0f * t;
    const float ln_m = tt * pl;
    const float ln_u = fmaf((float)(e - 24), 6.93147181e-1f, ln_m);
    const float arg = -2.0f * ln_u;
    const float radius = sqrtf(arg);
    const uint32_t quad = b >> 22, f = b & 0x3FFFFFu;
    const int refl = f >= 0x200000u;
    const uint32_t g = refl ? 0x3FFFFFu - f : f;
    const float phi = (float)(2u * g + 1u) * 1.87253514e-7f;
    float sp, cp;
    dn_sincos_poly(phi, &sp, &cp);
    const float sq = refl ? cp : sp, cq = refl ? sp : cp;
    float cs, sn;
    switch (quad) {
        case 0: cs = cq; sn = sq; break;
        case 1: cs = -sq; sn = cq; break;
        case 2: cs = -cq; sn = -sq; break;
        default: cs = sq; sn = -cq; break;
    }
    *z0 = radius * cs;
    *z1 = radius * sn;
}
void orc_noise_normal(uint64_t seed, uint64_t stream, uint64_t i0, size_t n, float *out) {
    const uint64_t key = dn_splitmix64(dn_splitmix64(seed) + stream);
    for (size_t j = 0; j < n; ++j) {
        const uint64_t i = i0 + j;
        float z0, z1;
        dn_normal_pair(key, i >> 1, &z0, &z1);
        out[j] = (i & 1) ? z1 : z0;
    }
}

uint8_t orc_progressive_bits(size_t num_steps, size_t t, uint8_t decode_bits, uint8_t min_bits,
                             int32_t *is_prefill) {
    *is_prefill = t > num_steps / 2;                          /* :886 */
    float progress = (float)(num_steps - t) / (float)(num_steps / 2);      /* :895 */
    float target = (float)decode_bits * (1.0f - progress) + (float)min_bits * progress; /* :896-897 */
    return rs_f32_as_u8(target);
}
