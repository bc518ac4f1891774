/*
 * mcl_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the per-particle arithmetic of
 * gustavorvillela/mcmh_localization, function by function, in the reference's
 * own operation order and precision (all fp64 unless the reference rounds to
 * fp32), so that its results are bit-comparable with the numba-compiled
 * reference.  Every function cites the reference file:line it follows
 * (paths relative to the reference tree; "pu" = app/scripts/parallel_utils.py,
 * "node" = app/scripts/amcmh_localizer.py).
 *
 * Pinning: oracle/gen_golden.py runs the UNMODIFIED reference (numba) in the
 * build container and stores its inputs/outputs under tests/golden/; the
 * CPU test-suite checks this file against those vectors bit-for-bit
 * (tests/test_oracle_golden.py).  numba's scalar libm calls (sin, cos, exp,
 * log, sqrt, fmod, atan2) resolve to glibc, which is what this file links.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product
 * (mcmh_localization_b200/) never does.
 *
 * Build: see oracle/Makefile  (-O2 -ffp-contract=off -fopenmp; no fast-math).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PI 3.141592653589793  /* == np.pi */

#ifdef _OPENMP
#include <omp.h>
#endif

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* pu:62-67  normalize_angle: (theta + pi) % (2 pi) - pi with Python's
 * sign-of-divisor modulo (numba lowers float % to fmod + sign fix-up). */
static inline double py_fmod(double a, double b) {
    double r = fmod(a, b);
    if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b;
    return r;
}
double orc_normalize_angle(double theta) {
    return py_fmod(theta + ORC_PI, 2 * ORC_PI) - ORC_PI;
}

/* pu:69-83 normalize_angle_array: result is float32. */
void orc_normalize_angle_array(const double *angles, double mean_angle, int64_t n, float *out) {
    for (int64_t i = 0; i < n; ++i) out[i] = (float)orc_normalize_angle(angles[i] - mean_angle);
}

/* pu:31-33 gaussian_prob */
static inline double gaussian_prob(double diff, double sigma) {
    double q = diff / sigma;
    return exp(-0.5 * (q * q)) / sqrt(2 * ORC_PI * (sigma * sigma));
}

/* pu:85-149 compute_likelihoods.  particles (N,3) f64 C-order; scan/angles
 * f32; distance_map f32 flattened my*W+mx; every expression promoted to f64
 * except dist**2 (f32), and the final store rounds to f32. */
void orc_compute_likelihoods(const float *scan, const float *angles, int M,
                             const double *particles, int64_t N,
                             const float *dist_map, double res, double ox, double oy,
                             int W, int H, double sigma_hit, double z_hit, double z_rand,
                             double max_range, int step, float *scores) {
    if (step < 1) step = 1;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        const double x = particles[3 * i], y = particles[3 * i + 1], theta = particles[3 * i + 2];
        double log_score = 0.0;
        int64_t valid_count = 0;
        for (int j = 0; j < M; j += step) {
            const double r = (double)scan[j];
            if (isfinite(r) && r < max_range) {                     /* pu:123 */
                valid_count += 1;                                   /* pu:124 (before bounds) */
                const double a = theta + (double)angles[j];
                const double lx = x + r * cos(a);                   /* pu:126 */
                const double ly = y + r * sin(a);                   /* pu:127 */
                const int64_t mx = (int64_t)((lx - ox) / res);      /* pu:128 trunc toward 0 */
                const int64_t my = (int64_t)((ly - oy) / res);      /* pu:129 */
                if (mx < 0 || mx >= W || my < 0 || my >= H) continue; /* pu:131-132 */
                const float dist_f = dist_map[my * (int64_t)W + mx];
                const double dist = (double)dist_f;
                /* numba types float32 ** 2 as float32: the square is rounded to f32 before
                 * promotion (pinned by tests/golden/likelihood_*.npz; SURVEY A.2 says f64 -- wrong). */
                const double dist_sq = (double)(dist_f * dist_f);
                double p_hit;
                if (dist <= max_range)                              /* pu:135-138 */
                    p_hit = exp(-0.5 * dist_sq / (sigma_hit * sigma_hit)) /
                            sqrt(2 * ORC_PI * (sigma_hit * sigma_hit));
                else
                    p_hit = 0.0;
                const double p_rand = (0 <= r && r <= max_range) ? 1.0 / max_range : 0.0; /* pu:139 */
                double p = z_hit * p_hit + z_rand * p_rand;         /* pu:140 */
                p = p > 1e-6 ? p : 1e-6;                            /* pu:141 max(p,1e-6) */
                log_score += log(p);                                /* pu:142 */
            }
        }
        if (valid_count > 0) scores[i] = (float)(log_score / (double)valid_count); /* pu:144-145 */
        else scores[i] = -50.0f;                                    /* pu:147 */
    }
}

/* pu:4-29 raycast on a (H,W) float64 grid (row = y).  limits[0]=x_min, limits[2]=y_min. */
double orc_raycast(double x, double y, double angle, double max_range, const double *limits,
                   double resolution, const double *grid, int W, int H) {
    const double dx = cos(angle), dy = sin(angle);
    const double step_size = 0.1;
    const int max_steps = (int)(max_range / step_size);
    for (int i = 1; i <= max_steps; ++i) {
        const double cx = x + i * step_size * dx;   /* (i*step)*dx */
        const double cy = y + i * step_size * dy;
        const int64_t gx = (int64_t)((cx - limits[0]) / resolution);
        const int64_t gy = (int64_t)((cy - limits[2]) / resolution);
        if (!(0 <= gx && gx < W && 0 <= gy && gy < H)) return max_range;
        if (grid[gy * (int64_t)W + gx] > 0.5) return i * step_size;
    }
    return max_range;
}

/* pu:388-396 is_valid_position (cell == 0 only). */
static inline int is_valid_position(double x, double y, const int8_t *map, int W, int H,
                                    double res, double ox, double oy) {
    const int64_t mx = (int64_t)((x - ox) / res);
    const int64_t my = (int64_t)((y - oy) / res);
    if (0 <= mx && mx < W && 0 <= my && my < H) return map[my * (int64_t)W + mx] == 0;
    return 0;
}

/* pu:398-413 compute_valid_mask */
void orc_compute_valid_mask(const double *particles, int64_t N, const int8_t *map, int W, int H,
                            double res, double ox, double oy, uint8_t *mask) {
    for (int64_t i = 0; i < N; ++i)
        mask[i] = (uint8_t)is_valid_position(particles[3 * i], particles[3 * i + 1], map, W, H, res, ox, oy);
}

/* ---------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon et al., SC'11; Random123 reference constants).     */
/* Not part of the reference (which uses numba's MT19937); it is the        */
/* product's counter-based generator, restated here so that the GPU         */
/* production-RNG mode can be checked draw-for-draw on the CPU.             */
/* ---------------------------------------------------------------------- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Stream ids (ctr[3] low byte) -- must match mcmh_localization_b200/csrc/philox.cuh */
enum { ORC_STREAM_MOTION = 1, ORC_STREAM_MH = 2, ORC_STREAM_RESAMPLE = 3, ORC_STREAM_INIT = 4,
       ORC_STREAM_KLD = 5, ORC_STREAM_MOTION_R = 6, ORC_STREAM_MOTION_R2 = 7 };

static inline void draw4(uint64_t seed, uint64_t step, uint64_t item, uint32_t sub, uint32_t stream,
                         uint32_t out[4]) {
    uint32_t ctr[4], key[2];
    ctr[0] = (uint32_t)item;
    ctr[1] = (uint32_t)step;
    ctr[2] = sub;
    ctr[3] = (stream & 0xffu) | ((uint32_t)(item >> 32) << 8) | ((uint32_t)((step >> 32) & 0xffu) << 24);
    key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
    orc_philox4x32_10(ctr, key, out);
}
void orc_draw4(uint64_t seed, uint64_t step, uint64_t item, uint32_t sub, uint32_t stream, uint32_t out[4]) {
    draw4(seed, step, item, sub, stream, out);
}
/* 53-bit uniform in [0,1) from two words, same bit recipe as numpy's legacy random_sample. */
static inline double u53(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}
double orc_uniform53(uint64_t seed, uint64_t step, uint64_t item, uint32_t sub, uint32_t stream) {
    uint32_t o[4]; draw4(seed, step, item, sub, stream, o); return u53(o[0], o[1]);
}
/* Three standard normals per (particle, attempt t): Box-Muller on 32-bit uniforms, u1 in (0,1],
 * u2 in [0,1):  z0 = R1 cos(2 pi u2), z1 = R1 sin(2 pi u2), z2 = R2 cos(2 pi u4), R = sqrt(-2 ln u).
 * u2, u3, u4 come from words 0, 1, 2 of the MOTION block (particle, step, t).  The RADIUS WORD (u1) is dealt so that
 * the GPU rejection loop can screen many attempts per Philox call (|z0|, |z1| <= R1: a small R1 cannot leave an
 * all-blocked neighbourhood), every bit being used exactly once:
 *   t = 0 : word 3 of the MOTION block.
 *   t >= 1: top nibble = nibble (t & 31) of the MOTION_R block (particle, step, t >> 5);
 *           nibble != 0: low 28 bits = top 28 bits of word 3 of the MOTION block;
 *           nibble == 0: next 16 bits = half-word (c & 7) of the MOTION_R2 block (particle, step, c >> 3), where
 *           c = number of attempts 1 <= s < t whose nibble is zero; low 12 bits = low 12 bits of word 3. */
static inline uint32_t motion_nibble(const uint32_t a[4], uint32_t t) {
    const uint32_t k = t & 31u;
    return (a[k >> 3] >> (4u * (k & 7u))) & 15u;
}
static inline uint32_t motion_half(const uint32_t b[4], uint32_t c) {
    const uint32_t k = c & 7u, w = b[k >> 1];
    return (k & 1u) ? (w >> 16) : (w & 0xffffu);
}
static inline void normals3_words(uint32_t wr, const uint32_t o[4], double z[3]) {
    const double u1 = ((double)wr + 1.0) * 2.3283064365386963e-10;
    const double u2 = (double)o[0] * 2.3283064365386963e-10;
    const double u3 = ((double)o[1] + 1.0) * 2.3283064365386963e-10;
    const double u4 = (double)o[2] * 2.3283064365386963e-10;
    const double r1 = sqrt(-2.0 * log(u1)), r2 = sqrt(-2.0 * log(u3));
    const double a1 = 6.283185307179586 * u2, a2 = 6.283185307179586 * u4;
    z[0] = r1 * cos(a1); z[1] = r1 * sin(a1); z[2] = r2 * cos(a2);
}
/* Sequential reader of one particle's attempts 0, 1, 2, ... (what the rejection loop needs). */
typedef struct {
    uint64_t seed, step, item;
    uint32_t next, zeros;          /* next attempt to be read; zero nibbles among attempts 1 .. next - 1 */
    uint32_t a[4];                 /* MOTION_R block of attempt `next` (valid when next & 31 or next was loaded) */
    int a_block;
} MotionDraws;
static inline void motion_draws_init(MotionDraws *d, uint64_t seed, uint64_t step, uint64_t item) {
    d->seed = seed; d->step = step; d->item = item; d->next = 0; d->zeros = 0; d->a_block = -1;
}
static inline void motion_draws_next(MotionDraws *d, double z[3]) {
    uint32_t o[4], wr;
    const uint32_t t = d->next++;
    draw4(d->seed, d->step, d->item, t, ORC_STREAM_MOTION, o);
    if (t == 0) {
        wr = o[3];
    } else {
        if (d->a_block != (int)(t >> 5)) {
            draw4(d->seed, d->step, d->item, t >> 5, ORC_STREAM_MOTION_R, d->a);
            d->a_block = (int)(t >> 5);
        }
        const uint32_t nib = motion_nibble(d->a, t);
        if (nib) {
            wr = (nib << 28) | (o[3] >> 4);
        } else {
            uint32_t b[4];
            const uint32_t c = d->zeros++;
            draw4(d->seed, d->step, d->item, c >> 3, ORC_STREAM_MOTION_R2, b);
            wr = (motion_half(b, c) << 12) | (o[3] & 0xfffu);
        }
    }
    normals3_words(wr, o, z);
}
/* Random access (tests): reads the attempts before it. */
static inline void normals3(uint64_t seed, uint64_t step, uint64_t item, uint32_t attempt, double z[3]) {
    uint32_t o[4], wr;
    draw4(seed, step, item, attempt, ORC_STREAM_MOTION, o);
    if (attempt == 0) {
        wr = o[3];
    } else {
        uint32_t a[4], c = 0, nib = 0;
        for (uint32_t g = 0; g <= (attempt >> 5); ++g) {
            draw4(seed, step, item, g, ORC_STREAM_MOTION_R, a);
            for (uint32_t t = 32u * g; t < 32u * g + 32u && t <= attempt; ++t) {
                if (t == 0) continue;
                nib = motion_nibble(a, t);
                if (t < attempt && nib == 0) ++c;
            }
        }
        if (nib) {
            wr = (nib << 28) | (o[3] >> 4);
        } else {
            uint32_t b[4];
            draw4(seed, step, item, c >> 3, ORC_STREAM_MOTION_R2, b);
            wr = (motion_half(b, c) << 12) | (o[3] & 0xfffu);
        }
    }
    normals3_words(wr, o, z);
}
void orc_normals3(uint64_t seed, uint64_t step, uint64_t item, uint32_t attempt, double z[3]) {
    normals3(seed, step, item, attempt, z);
}
/* attempts 0 .. n-1 of one particle through the sequential reader (what orc_apply_motion_model uses) */
void orc_normals3_seq(uint64_t seed, uint64_t step, uint64_t item, int n, double *z_out) {
    MotionDraws d;
    motion_draws_init(&d, seed, step, item);
    for (int t = 0; t < n; ++t) motion_draws_next(&d, z_out + 3 * (size_t)t);
}

/* pu:332-363 apply_motion_model_parallel.
 * normals != NULL : injected draws, layout (N, A, 3) f64; attempt t uses row t % A
 *                   (tests size A so that no wrap occurs when comparing with the reference).
 * normals == NULL : Philox draws keyed (seed, step, first_index + i, attempt).
 * alpha is the float32[4] of node:28-33 promoted to f64 inside the njit code.
 * np.random.normal(0, s) == 0.0 + s * z.
 * attempts_out (nullable): 1-based index of the accepted attempt, 0 = fallback (pu:360-361). */
void orc_apply_motion_model(const double *particles, int64_t N, const double delta[3],
                            const float alpha_f32[4], const int8_t *map, double res, double ox,
                            double oy, int W, int H, const double *normals, int A, uint64_t seed,
                            uint64_t step, uint64_t first_index, int max_attempts,
                            double *out, int32_t *attempts_out) {
    const double rot1 = delta[0], trans = delta[1], rot2 = delta[2];
    const double a1 = (double)alpha_f32[0], a2 = (double)alpha_f32[1], a3 = (double)alpha_f32[2],
                 a4 = (double)alpha_f32[3];
    const double s1 = a1 * fabs(rot1) + a2 * fabs(trans);                  /* pu:345 */
    const double s2 = a3 * fabs(trans) + a4 * (fabs(rot1) + fabs(rot2));   /* pu:346 */
    const double s3 = a1 * fabs(rot2) + a2 * fabs(trans);                  /* pu:347 */
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < N; ++i) {
        const double x = particles[3 * i], y = particles[3 * i + 1], theta = particles[3 * i + 2];
        int success = 0;
        MotionDraws draws;
        motion_draws_init(&draws, seed, step, first_index + (uint64_t)i);
        for (int t = 0; t < max_attempts; ++t) {
            double z[3];
            if (normals) {
                const double *zz = normals + ((size_t)i * A + (size_t)(t % A)) * 3;
                z[0] = zz[0]; z[1] = zz[1]; z[2] = zz[2];
            } else {
                motion_draws_next(&draws, z);
            }
            const double r1_hat = rot1 + (0.0 + s1 * z[0]);
            const double t_hat = trans + (0.0 + s2 * z[1]);
            const double r2_hat = rot2 + (0.0 + s3 * z[2]);
            const double x_new = x + t_hat * cos(theta + r1_hat);          /* pu:351 */
            const double y_new = y + t_hat * sin(theta + r1_hat);          /* pu:352 */
            const double th_new = orc_normalize_angle(theta + r1_hat + r2_hat); /* pu:353 */
            if (is_valid_position(x_new, y_new, map, W, H, res, ox, oy)) { /* pu:355 */
                out[3 * i] = x_new; out[3 * i + 1] = y_new; out[3 * i + 2] = th_new;
                if (attempts_out) attempts_out[i] = t + 1;
                success = 1;
                break;
            }
        }
        if (!success) {                                                    /* pu:360-361 */
            out[3 * i] = x; out[3 * i + 1] = y; out[3 * i + 2] = theta;
            if (attempts_out) attempts_out[i] = 0;
        }
    }
}

/* pu:208-236 mh_resampling.  p_new / p_old is an fp32 IEEE divide promoted to f64 for
 * min(1.0, .) and the compare with the f64 uniform.
 * uniforms != NULL: injected; else Philox u53 keyed (seed, step, first_index+i). */
void orc_mh_resampling(const double *particles, const double *proposed, const float *likelihoods,
                       const float *old_weights, int64_t N, const double *uniforms, uint64_t seed,
                       uint64_t step, uint64_t first_index, double *new_particles,
                       float *new_weights, uint8_t *accept) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        const float p_old = old_weights[i], p_new = likelihoods[i];
        double alpha = 1.0;
        if (p_old > 0) {
            const float q = p_new / p_old;
            const double qd = (double)q;
            alpha = (qd < 1.0) ? qd : 1.0;      /* min(1.0, q): NaN -> 1.0 like Python's min(1.0, nan) */
        }
        const double u = uniforms ? uniforms[i]
                                  : orc_uniform53(seed, step, first_index + (uint64_t)i, 0, ORC_STREAM_MH);
        const int acc = u < alpha;
        const double *src = acc ? proposed : particles;
        new_particles[3 * i] = src[3 * i]; new_particles[3 * i + 1] = src[3 * i + 1];
        new_particles[3 * i + 2] = src[3 * i + 2];
        new_weights[i] = acc ? p_new : p_old;
        if (accept) accept[i] = (uint8_t)acc;
    }
}

/* pu:238-276 assym_mh_resampling.  np.log(f32 array + 1e-10) is computed in f64
 * (SURVEY A.2); alpha = min(1, exp(log_alpha)) only if log_den > 0, else 1 (pu:269). */
void orc_assym_mh_resampling(const double *particles, const double *proposed,
                             const float *likelihoods, const float *old_weights,
                             const double *trans_forward, const double *trans_backward, int64_t N,
                             const double *uniforms, double *new_particles, float *new_weights,
                             uint8_t *accept) {
    for (int64_t i = 0; i < N; ++i) {
        const double log_pre = log((double)old_weights[i] + 1e-10);
        const double log_post = log((double)likelihoods[i] + 1e-10);
        const double log_tf = log(trans_forward[i] + 1e-10);
        const double log_tb = log(trans_backward[i] + 1e-10);
        const double log_num = log_post + log_tb;
        const double log_den = log_pre + log_tf;
        const double log_alpha = log_num - log_den;
        double alpha = 1.0;
        if (log_den > 0) { const double e = exp(log_alpha); alpha = (e < 1.0) ? e : 1.0; }
        const int acc = uniforms[i] < alpha;
        const double *src = acc ? proposed : particles;
        new_particles[3 * i] = src[3 * i]; new_particles[3 * i + 1] = src[3 * i + 1];
        new_particles[3 * i + 2] = src[3 * i + 2];
        new_weights[i] = acc ? likelihoods[i] : old_weights[i];
        if (accept) accept[i] = (uint8_t)acc;
    }
}

/* pu:282-330 motion_model_odometry_parallel (density, normalised by the population sum).
 * alpha is float32[4] promoted to f64.  The sum is a plain sequential f64 loop here
 * (numba may tree-reduce it under parallel=True; compare with a tolerance). */
void orc_motion_model_odometry(const double *prev, const double *curr, int64_t N,
                               const double delta[3], const float alpha_f32[4], double *probs) {
    const double rot1 = delta[0], trans = delta[1], rot2 = delta[2];
    const double a1 = (double)alpha_f32[0], a2 = (double)alpha_f32[1], a3 = (double)alpha_f32[2],
                 a4 = (double)alpha_f32[3];
    for (int64_t i = 0; i < N; ++i) {
        const double dx = curr[3 * i] - prev[3 * i];
        const double dy = curr[3 * i + 1] - prev[3 * i + 1];
        const double th_prev = prev[3 * i + 2], th_curr = curr[3 * i + 2];
        const double trans_hat = sqrt(dx * dx + dy * dy);
        const double rot1_hat = orc_normalize_angle(atan2(dy, dx) - th_prev);
        const double rot2_hat = orc_normalize_angle(th_curr - th_prev - rot1_hat);
        const double s_rot1 = a1 * fabs(rot1) + a2 * fabs(trans);
        const double s_trans = a3 * fabs(trans) + a4 * (fabs(rot1) + fabs(rot2));
        const double s_rot2 = a1 * fabs(rot2) + a2 * fabs(trans);
        const double p1 = gaussian_prob(orc_normalize_angle(rot1 - rot1_hat), s_rot1);
        const double p2 = gaussian_prob(trans - trans_hat, s_trans);
        const double p3 = gaussian_prob(orc_normalize_angle(rot2 - rot2_hat), s_rot2);
        probs[i] = p1 * p2 * p3;
    }
    double s = 0.0;
    for (int64_t i = 0; i < N; ++i) s += probs[i];
    if (s > 0) for (int64_t i = 0; i < N; ++i) probs[i] /= s;
}

/* pu:416-446 low_variance_resample_numba -> indices.
 * weights / np.sum(weights): numba's np.sum on an f32 array accumulates sequentially in f32
 * (SURVEY A.3); the divide is f32; c accumulates in f32; U = r + m*step in f64; the compare
 * U > c promotes c to f64.  r is passed in (== np.random.uniform(0, 1/N) = 0 + (1/N - 0)*u). */
void orc_low_variance_resample(const float *weights, int64_t n_in, int64_t N, double r, int32_t *idx) {
    float sum = 0.0f;
    for (int64_t i = 0; i < n_in; ++i) sum += weights[i];
    const double step = 1.0 / (double)N;
    float c = weights[0] / sum;
    int64_t i = 0;
    for (int64_t m = 0; m < N; ++m) {
        const double U = r + (double)m * step;
        while (U > (double)c && i < N - 1) {
            i += 1;
            c += weights[i] / sum;
        }
        idx[m] = (int32_t)i;
    }
}

/* PRODUCTION resampling semantics of the GPU build (DESIGN.md "resampling modes"): systematic
 * resampling on weights quantised to 64-bit fixed point, so that the cumulative sum is exact
 * integer arithmetic -- associative, hence identical for any block/rank decomposition.
 * Not in the reference; restated here so the GPU scan+search can be checked index-for-index.
 *   q_i   = (uint64) trunc( (double)w_i * scale ),  scale = 2^(K-e), 2^e >= max w, K = 62-ceil(log2 n)
 *   C_i   = sum_{k<=i} q_k                      (uint64, exact)
 *   T_m   = (uint64) ceil( (r + m*step) * (double)C_{n-1} )   (f64 add, mul, no FMA)
 *   idx_m = min{ i : C_i >= T_m }, clamped to n-1           (== the walk of pu:439-444)
 */
double orc_resample_scale(float wmax, int64_t n_global) {
    int e = 0; if (wmax > 0) { frexp((double)wmax, &e); }   /* wmax = f * 2^e, f in [0.5,1) => 2^e > wmax */
    int lg = 0; while (((int64_t)1 << lg) < n_global) ++lg;
    return ldexp(1.0, 62 - lg - e);
}
void orc_systematic_resample_q(const float *weights, int64_t n_in, int64_t N, double r, double scale,
                               int32_t *idx) {
    uint64_t total = 0;
    for (int64_t i = 0; i < n_in; ++i) total += (uint64_t)((double)weights[i] * scale);
    const double step = 1.0 / (double)N;
    const double totd = (double)total;
    uint64_t c = (uint64_t)((double)weights[0] * scale);
    int64_t i = 0;
    for (int64_t m = 0; m < N; ++m) {
        const double U = r + (double)m * step;
        const double t = ceil(U * totd);
        const uint64_t T = t >= 18446744073709551616.0 ? UINT64_MAX : (uint64_t)t;
        while (T > c && i < n_in - 1) {
            i += 1;
            c += (uint64_t)((double)weights[i] * scale);
        }
        idx[m] = (int32_t)i;
    }
}

/* pu:529-591 kld_sampling_amcl with injected draws: r (the one uniform) and normals (max_samples, 3)
 * standard normals (np.random.normal(0, s) = 0 + s*z, three per sample, in order).
 * weights are used as given (node:276-278 normalises them before); c accumulates in f32.
 * The Python set of bin triples is restated as an open-addressing hash set.
 * Returns the number of samples; out is (max_samples, 3) f32 like the reference's buffer. */
int64_t orc_kld_sampling(const double *particles, const float *weights, int64_t n, double bin_xy,
                         double bin_theta, double epsilon, double z, int64_t max_samples,
                         int64_t min_particles, double r, const double *normals, float *out) {
    const double noise_std[3] = {0.001, 0.001, 0.02};                  /* pu:552 */
    int64_t cap = 16;
    while (cap < 4 * (max_samples + 1)) cap <<= 1;
    int64_t *keys = (int64_t *)malloc((size_t)cap * 3 * sizeof(int64_t));
    unsigned char *used = (unsigned char *)calloc((size_t)cap, 1);
    int64_t nbins = 0, count = 0, i = 0;
    float c = weights[0];                                               /* pu:555 */
    while (count < max_samples) {
        const double u = r + (double)count / (double)max_samples;       /* pu:560 */
        while (u > (double)c && i < n - 1) { i += 1; c += weights[i]; } /* pu:561-563 */
        double noisy[3];
        for (int j = 0; j < 3; ++j)
            noisy[j] = particles[3 * i + j] + (0.0 + noise_std[j] * normals[3 * count + j]);   /* pu:570 */
        const int64_t xb = (int64_t)(noisy[0] / bin_xy), yb = (int64_t)(noisy[1] / bin_xy),
                      tb = (int64_t)(noisy[2] / bin_theta);             /* pu:573-575 int() truncation */
        uint64_t hsh = (uint64_t)xb * 0x9E3779B97F4A7C15ull ^ (uint64_t)yb * 0xC2B2AE3D27D4EB4Full ^
                       (uint64_t)tb * 0x165667B19E3779F9ull;
        int64_t slot = (int64_t)(hsh & (uint64_t)(cap - 1));
        int found = 0;
        while (used[slot]) {
            if (keys[3 * slot] == xb && keys[3 * slot + 1] == yb && keys[3 * slot + 2] == tb) { found = 1; break; }
            slot = (slot + 1) & (cap - 1);
        }
        if (!found) {                                                   /* pu:578-586 */
            used[slot] = 1; keys[3 * slot] = xb; keys[3 * slot + 1] = yb; keys[3 * slot + 2] = tb;
            nbins += 1;
            const int64_t k = nbins;
            if (k > 1 && count >= min_particles) {
                const double km1 = (double)(k - 1);
                const double a = 1.0 - 2.0 / (9.0 * km1) + sqrt(2.0 / (9.0 * km1)) * z;
                const double chi2 = km1 * (a * a * a);
                if ((double)count > chi2 / (2.0 * epsilon)) break;      /* breaks BEFORE storing this sample */
            }
        }
        out[3 * count] = (float)noisy[0]; out[3 * count + 1] = (float)noisy[1]; out[3 * count + 2] = (float)noisy[2];
        count += 1;
    }
    free(keys); free(used);
    return count;
}

/* pu:151-201 compute_likelihoods_raycast (+ pu:36-58 p_hit / p_rand): beam model by ray marching, hard-coded
 * sigma_hit 0.05, z_hit 0.8, z_rand 0.1, max_range 10.  grid is (H, W) f64 (row = y), occupied where > 0.5;
 * limits = [x_min, x_max, y_min, y_max].  No valid beam -> -inf (pu:199). */
void orc_compute_likelihoods_raycast(const float *scan, const float *angles, int M, const double *particles,
                                     int64_t N, const double *grid, int W, int H, double res,
                                     const double *limits, float *scores) {
    const double sigma_hit = 0.05, z_hit = 0.8, z_rand = 0.1, max_range = 10.0;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < N; ++i) {
        const double x = particles[3 * i], y = particles[3 * i + 1], theta = particles[3 * i + 2];
        double log_score = 0.0;
        int64_t valid_count = 0;
        for (int j = 0; j < M; ++j) {
            const double r_meas = (double)scan[j];
            if (isfinite(r_meas) && r_meas < max_range) {
                valid_count += 1;
                const double r_pred = orc_raycast(x, y, theta + (double)angles[j], max_range, limits, res, grid, W, H);
                double prob_hit = 0.0;                                         /* pu:36-41 */
                if (0 <= r_meas && r_meas <= max_range) {
                    const double q = (r_meas - r_pred) / sigma_hit;
                    prob_hit = (1 / (sqrt(2 * ORC_PI) * sigma_hit)) * exp(-0.5 * (q * q));
                }
                const double prob_rand = (0 <= r_meas && r_meas <= max_range) ? 1.0 / max_range : 0.0;   /* pu:55-58 */
                double p = z_hit * prob_hit + z_rand * prob_rand;
                p = p > 1e-6 ? p : 1e-6;
                log_score += log(p);
            }
        }
        scores[i] = valid_count > 0 ? (float)(log_score / (double)valid_count) : -INFINITY;
    }
}

/* ---------------------------------------------------------------------- */
/* Functions the node imports (node:13) but its callbacks never reach      */
/* (SURVEY 8(a) row a14).                                                  */
/* ---------------------------------------------------------------------- */

/* pu:369-386 compute_valid_indices: int() truncation, cell in the map, map_data[cell] <= 10 */
int64_t orc_compute_valid_indices(const double *particles, int64_t N, const int8_t *map, int W, int H, double res,
                                  double ox, double oy, int32_t *out) {
    int64_t k = 0;
    for (int64_t i = 0; i < N; ++i) {
        const int64_t mx = (int64_t)((particles[3 * i] - ox) / res);
        const int64_t my = (int64_t)((particles[3 * i + 1] - oy) / res);
        if (0 <= mx && mx < W && 0 <= my && my < H && map[my * (int64_t)W + mx] <= 10) out[k++] = (int32_t)i;
    }
    return k;
}

/* pu:467-477 parallel_resample_simple with the uniforms passed in: cum = np.cumsum(weights) (numba: sequential,
 * float32 accumulator), idx = np.searchsorted(cum, u) (left: first i with cum[i] >= u, compared in f64).
 * The reference reads particles[N] when u > cum[-1]; here that case takes N - 1. */
void orc_parallel_resample_simple(const float *weights, int64_t n, const double *u, int64_t N, int32_t *idx) {
    float *cum = (float *)malloc(sizeof(float) * (size_t)n);
    float c = 0.0f;
    for (int64_t i = 0; i < n; ++i) { c += weights[i]; cum[i] = c; }
    for (int64_t m = 0; m < N; ++m) {
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if ((double)cum[mid] < u[m]) lo = mid + 1; else hi = mid;
        }
        idx[m] = (int32_t)(lo < n ? lo : n - 1);
    }
    free(cum);
}

/* pu:486-502 low_variance_resample_amcl: weights as given, c a sequential f32 sum starting at weights[0],
 * U = r + m / target_size (int / int true division), walk bounded by len(particles) - 1 */
void orc_low_variance_resample_amcl(const float *weights, int64_t n_in, int64_t target, double r, int32_t *idx) {
    float c = weights[0];
    int64_t i = 0;
    for (int64_t m = 0; m < target; ++m) {
        const double U = r + (double)m / (double)target;
        while (U > (double)c && i < n_in - 1) {
            i += 1;
            c += weights[i];
        }
        idx[m] = (int32_t)(i % n_in);
    }
}

/* pu:504-526 reinitialize_particles_numba with the draws passed in: choice[i] indexes np.argwhere(map == 0)
 * (row-major), pose = lower-left corner of that cell, float32 output */
int64_t orc_reinitialize_particles(int64_t n, const int8_t *occ, int W, int H, double res, double ox, double oy,
                                   const int64_t *choice, const double *theta, float *out) {
    const int64_t cells = (int64_t)W * H;
    int32_t *free_cells = (int32_t *)malloc(sizeof(int32_t) * (size_t)(cells > 0 ? cells : 1));
    int64_t nf = 0;
    for (int64_t c = 0; c < cells; ++c) if (occ[c] == 0) free_cells[nf++] = (int32_t)c;
    for (int64_t i = 0; i < n; ++i) {
        if (nf == 0) { out[3 * i] = (float)ox; out[3 * i + 1] = (float)oy; out[3 * i + 2] = (float)theta[i]; continue; }
        const int cell = free_cells[choice[i]];
        const int my = cell / W, mx = cell % W;
        out[3 * i] = (float)((double)mx * res + ox);
        out[3 * i + 1] = (float)((double)my * res + oy);
        out[3 * i + 2] = (float)theta[i];
    }
    free(free_cells);
    return nf;
}

/* pu:600-614 validate_samples: in place; a sample outside the map or with distance_map >= 1.0 becomes (0, 0, 0) */
void orc_validate_samples(double *samples, int64_t n, const float *dist, int W, int H, double res, double ox, double oy) {
    for (int64_t i = 0; i < n; ++i) {
        const int64_t mx = (int64_t)((samples[3 * i] - ox) / res);
        const int64_t my = (int64_t)((samples[3 * i + 1] - oy) / res);
        if (!(0 <= mx && mx < W && 0 <= my && my < H && dist[my * (int64_t)W + mx] < 1.0f))
            samples[3 * i] = samples[3 * i + 1] = samples[3 * i + 2] = 0.0;
    }
}
