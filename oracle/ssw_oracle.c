/* TEST INFRASTRUCTURE (oracle / CPU baseline) -- never linked into libssw, never on the product path.
 *
 * Single-threaded FP32 C restatement of the reference's embed / extract / similarity path that keeps
 * the reference's *structure* (so its run time is a fair stand-in for the reference's CPU path, which
 * cannot be built here: no cargo/rustc, no crates registry, no network):
 *   - u8 -> f32 / 255, three separate planes            image::into_rgb32f + src/yiq.rs:58-62,177-186
 *   - per-line gather -> 1-D DCT -> scaled scatter,      src/dct2d.rs:83-219 (longest dimension first,
 *     columns strided by width                            x2 forward, x0.5 inverse, x4/(W*H) at the end)
 *   - FULL stable sort of all W*H-1 (index,&coef) pairs  src/algorithm.rs:200-221 (comparator through a
 *     through an indirect comparator                      function pointer like the boxed dyn Fn)
 *   - scatter embed / gather extract / sequential sim    src/algorithm.rs:382-432,543-593,696-714
 *   - clamped YIQ->RGB, round-half-away to u8            src/yiq.rs:139-147,187-197 + image::into_rgb8
 * The 1-D DCT arithmetic of the reference lives in rustdct 0.7.0 / rustfft 6.0.1 (Cargo.lock:584-603,
 * un-vendored); its definition is restated here as the classic N-point FFT algorithm (Makhoul
 * reordering + mixed-radix Cooley-Tukey + quarter-sample twiddle), O(N log N) like rustdct's.
 *
 * Pinned by tests/test_oracle_c.py against the numpy oracle (itself pinned to the reference's golden
 * PNG) and against the reference's scipy known answers (src/dct2d.rs:229-524).
 * Paths are relative to /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct { float re, im; } cpx;

/* ---------------------------------------------------------------- 1-D complex FFT, any length */
typedef struct fft_plan {
    int n;
    int nfac;
    int fac[32];
    cpx* tw;      /* exp(-2 pi i j / n), j < n */
    cpx* scratch; /* n */
} fft_plan;

static fft_plan* fft_plan_new(int n) {
    fft_plan* p = (fft_plan*)calloc(1, sizeof(fft_plan));
    p->n = n;
    int r = n;
    while (r % 4 == 0) { p->fac[p->nfac++] = 4; r /= 4; }
    while (r % 2 == 0) { p->fac[p->nfac++] = 2; r /= 2; }
    for (int f = 3; f * f <= r; f += 2) while (r % f == 0) { p->fac[p->nfac++] = f; r /= f; }
    if (r > 1) p->fac[p->nfac++] = r;
    p->tw = (cpx*)malloc(sizeof(cpx) * (size_t)(n > 0 ? n : 1));
    p->scratch = (cpx*)malloc(sizeof(cpx) * (size_t)(n > 0 ? n : 1));
    for (int j = 0; j < n; ++j) {
        double a = -2.0 * M_PI * (double)j / (double)n;
        p->tw[j].re = (float)cos(a);
        p->tw[j].im = (float)sin(a);
    }
    return p;
}

static void fft_plan_free(fft_plan* p) {
    if (!p) return;
    free(p->tw); free(p->scratch); free(p);
}

/* recursive decimation in time: out[0..n) = DFT of in[0], in[stride], ...  (n = product of fac[level..]) */
static void fft_rec(const fft_plan* p, int level, int n, const cpx* in, int stride, cpx* out) {
    if (n == 1) { out[0] = in[0]; return; }
    const int r = p->fac[level];
    const int m = n / r;
    for (int q = 0; q < r; ++q) fft_rec(p, level + 1, m, in + (size_t)q * stride, stride * r, out + (size_t)q * m);
    const int tws = p->n / n; /* twiddle stride: exp(-2 pi i k q / n) = tw[k q tws] */
    cpx tmp[64];
    cpx* t = r <= 64 ? tmp : (cpx*)malloc(sizeof(cpx) * (size_t)r);
    for (int k = 0; k < m; ++k) {
        for (int q = 0; q < r; ++q) {
            const cpx w = p->tw[(size_t)k * q * tws]; /* k*q*tws < n */
            const cpx v = out[(size_t)q * m + k];
            t[q].re = v.re * w.re - v.im * w.im;
            t[q].im = v.re * w.im + v.im * w.re;
        }
        if (r == 2) {
            out[k].re = t[0].re + t[1].re; out[k].im = t[0].im + t[1].im;
            out[k + m].re = t[0].re - t[1].re; out[k + m].im = t[0].im - t[1].im;
        } else if (r == 4) {
            const float ar = t[0].re + t[2].re, ai = t[0].im + t[2].im, br = t[0].re - t[2].re, bi = t[0].im - t[2].im;
            const float cr = t[1].re + t[3].re, ci = t[1].im + t[3].im, dr = t[1].re - t[3].re, di = t[1].im - t[3].im;
            out[k].re = ar + cr; out[k].im = ai + ci;
            out[k + m].re = br + di; out[k + m].im = bi - dr;
            out[k + 2 * m].re = ar - cr; out[k + 2 * m].im = ai - ci;
            out[k + 3 * m].re = br - di; out[k + 3 * m].im = bi + dr;
        } else {
            const int rs = p->n / r; /* exp(-2 pi i q j / r) = tw[(q j mod r) rs] */
            for (int j = 0; j < r; ++j) {
                float sr = 0.f, si = 0.f;
                for (int q = 0; q < r; ++q) {
                    const cpx w = p->tw[(size_t)((q * j) % r) * rs];
                    sr += t[q].re * w.re - t[q].im * w.im;
                    si += t[q].re * w.im + t[q].im * w.re;
                }
                out[(size_t)j * m + k].re = sr; out[(size_t)j * m + k].im = si;
            }
        }
    }
    if (t != tmp) free(t);
}

/* ---------------------------------------------------------------- 1-D DCT-II / DCT-III (rustdct scaling) */
typedef struct dct_plan {
    int n;
    fft_plan* fft;
    cpx* q;   /* exp(-i pi k / (2n)) */
    cpx* buf; /* n */
    cpx* out; /* n */
} dct_plan;

static dct_plan* dct_plan_new(int n) {
    dct_plan* d = (dct_plan*)calloc(1, sizeof(dct_plan));
    d->n = n;
    d->fft = fft_plan_new(n);
    d->q = (cpx*)malloc(sizeof(cpx) * (size_t)n);
    d->buf = (cpx*)malloc(sizeof(cpx) * (size_t)n);
    d->out = (cpx*)malloc(sizeof(cpx) * (size_t)n);
    for (int k = 0; k < n; ++k) {
        double a = -M_PI * (double)k / (2.0 * (double)n);
        d->q[k].re = (float)cos(a);
        d->q[k].im = (float)sin(a);
    }
    return d;
}

static void dct_plan_free(dct_plan* d) {
    if (!d) return;
    fft_plan_free(d->fft); free(d->q); free(d->buf); free(d->out); free(d);
}

/* rustdct process_dct2: X_k = sum_n x_n cos(pi k (2n+1) / 2N)   (no factor 2; src/dct2d.rs:229-248) */
static void dct2_1d(dct_plan* d, float* x) {
    const int n = d->n;
    for (int i = 0; i < (n + 1) / 2; ++i) { d->buf[i].re = x[2 * i]; d->buf[i].im = 0.f; }
    for (int i = 0; i < n / 2; ++i) { d->buf[n - 1 - i].re = x[2 * i + 1]; d->buf[n - 1 - i].im = 0.f; }
    fft_rec(d->fft, 0, n, d->buf, 1, d->out);
    for (int k = 0; k < n; ++k) x[k] = d->out[k].re * d->q[k].re - d->out[k].im * d->q[k].im;
}

/* rustdct process_dct3: x_k = X_0/2 + sum_{n>=1} X_n cos(pi n (2k+1) / 2N) */
static void dct3_1d(dct_plan* d, float* x) {
    const int n = d->n;
    /* v = Re( IFFT-unnormalised( conj-twiddled Hermitian spectrum ) ): V_k = 1/2 e^{+i pi k/2N} (X_k - i X_{N-k}) */
    for (int k = 0; k < n; ++k) {
        const float a = x[k], b = (k == 0) ? 0.f : x[n - k];
        /* conj(V_k) = 1/2 conj(e^{+i pi k/2N}) (a + i b) = 1/2 q_k (a + i b) */
        d->buf[k].re = 0.5f * (d->q[k].re * a - d->q[k].im * b);
        d->buf[k].im = 0.5f * (d->q[k].re * b + d->q[k].im * a);
    }
    fft_rec(d->fft, 0, n, d->buf, 1, d->out); /* conj(IFFT(V)) = FFT(conj V) */
    for (int i = 0; i < (n + 1) / 2; ++i) x[2 * i] = d->out[i].re;
    for (int i = 0; i < n / 2; ++i) x[2 * i + 1] = d->out[n - 1 - i].re;
}

/* ---------------------------------------------------------------- dct2d::dct2_2d, src/dct2d.rs:83-219 */
enum { T_DCT2 = 0, T_DCT2_ORTHO = 1, T_DCT3 = 2 };

void oracle_dct2_2d(int type, int width, int height, float* data) {
    const int first_is_row = width >= height; /* :93-97 */
    const float scaling = (type == T_DCT3) ? 0.5f : 2.0f; /* :107-111 */
    for (int pass = 0; pass < 2; ++pass) {
        const int is_row = (pass == 0) ? first_is_row : !first_is_row;
        const int length = is_row ? width : height;
        dct_plan* plan = dct_plan_new(length); /* planner.plan_dct2/3(length) :119-123 */
        float* tmp = (float*)malloc(sizeof(float) * (size_t)length);
        const float s0 = sqrtf(1.0f / (4.0f * (float)length)), sn = sqrtf(1.0f / (2.0f * (float)length));
        const int lines = is_row ? height : width;
        for (int l = 0; l < lines; ++l) {
            float* base = is_row ? data + (size_t)l * width : data + l;
            const size_t step = is_row ? 1 : (size_t)width;
            for (int i = 0; i < length; ++i) tmp[i] = base[i * step]; /* gather :132-136,174-178 */
            if (type == T_DCT3) dct3_1d(plan, tmp); else dct2_1d(plan, tmp);
            if (type == T_DCT2_ORTHO) { /* :153-162 */
                for (int i = 0; i < length; ++i) base[i * step] = (i == 0 ? s0 : sn) * scaling * tmp[i];
            } else {
                for (int i = 0; i < length; ++i) base[i * step] = scaling * tmp[i]; /* :163-168 */
            }
        }
        free(tmp);
        dct_plan_free(plan);
    }
    if (type == T_DCT3) { /* :213-217 */
        const float sc = 4.0f / (float)((size_t)width * (size_t)height);
        const size_t n = (size_t)width * height;
        for (size_t i = 0; i < n; ++i) data[i] = data[i] * sc;
    }
}

/* ---------------------------------------------------------------- colour, src/yiq.rs */
static void rgb8_to_yiq(const uint8_t* rgb, size_t npix, float* y, float* i, float* q) {
    for (size_t p = 0; p < npix; ++p) {
        const float r = (float)rgb[3 * p] / 255.0f, g = (float)rgb[3 * p + 1] / 255.0f, b = (float)rgb[3 * p + 2] / 255.0f;
        y[p] = 0.30f * r + 0.59f * g + 0.11f * b;
        i[p] = 0.60f * r + -0.28f * g + -0.32f * b;
        q[p] = 0.21f * r + -0.52f * g + 0.31f * b;
    }
}

static float clamp01(float v) { return v < 0.f ? 0.f : (v > 1.f ? 1.f : v); }

static void yiq_to_rgb8(const float* y, const float* i, const float* q, size_t npix, uint8_t* rgb) {
    for (size_t p = 0; p < npix; ++p) {
        const float r = clamp01(1.0f * y[p] + 0.948262f * i[p] + 0.624013f * q[p]);
        const float g = clamp01(1.0f * y[p] + -0.276066f * i[p] + -0.639810f * q[p]);
        const float b = clamp01(1.0f * y[p] + -1.105450f * i[p] + 1.729860f * q[p]);
        rgb[3 * p] = (uint8_t)roundf(clamp01(r) * 255.0f);
        rgb[3 * p + 1] = (uint8_t)roundf(clamp01(g) * 255.0f);
        rgb[3 * p + 2] = (uint8_t)roundf(clamp01(b) * 255.0f);
    }
}

/* ---------------------------------------------------------------- ordering, src/algorithm.rs:200-280 */
typedef struct { size_t index; const float* coef; } pair_t; /* Vec<(usize, &f32)> :204 */

typedef int (*order_fn)(size_t li, float l, size_t ri, float r); /* <0, 0, >0 like Ordering */

static int32_t total_key(float v) {
    int32_t b;
    memcpy(&b, &v, 4);
    return b ^ (int32_t)(((uint32_t)(b >> 31)) >> 1); /* f32::total_cmp */
}
static int total_cmp(float a, float b) {
    const int32_t x = total_key(a), y = total_key(b);
    return (x > y) - (x < y);
}

static int g_w, g_h;
static int ordering_by_energy(size_t li, float l, size_t ri, float r) { (void)li; (void)ri; return total_cmp(l * l, r * r); }
static int ordering_by_largest(size_t li, float l, size_t ri, float r) { (void)li; (void)ri; return total_cmp(l, r); }
static float ortho_scaling(size_t index, float value) { /* :240-266 */
    const float s_k0_w = sqrtf(1.0f / (4.0f * (float)g_w)), s_k0_h = sqrtf(1.0f / (4.0f * (float)g_h));
    const float s_w = sqrtf(1.0f / (2.0f * (float)g_w)), s_h = sqrtf(1.0f / (2.0f * (float)g_h));
    float scaling = 1.0f;
    scaling *= (index < (size_t)g_w) ? s_k0_w : s_w;
    scaling *= (index % (size_t)g_w == 0) ? s_k0_h : s_h;
    return scaling * value;
}
static int ordering_energy_ortho(size_t li, float l, size_t ri, float r) {
    return ordering_by_energy(li, ortho_scaling(li, l), ri, ortho_scaling(ri, r));
}
static int ordering_legacy(size_t li, float l, size_t ri, float r) {
    return ordering_by_largest(li, ortho_scaling(li, l), ri, ortho_scaling(ri, r));
}

/* out: all w*h-1 indices in order */
void oracle_obtain_indices(const float* coeff, int width, int height, int ordering, uint64_t* out) {
    const size_t n = (size_t)width * height;
    if (n <= 1) return;
    g_w = width; g_h = height;
    order_fn f = ordering == 0 ? ordering_by_energy : (ordering == 1 ? ordering_energy_ortho : ordering_legacy);
    const size_t m = n - 1;
    pair_t* a = (pair_t*)malloc(sizeof(pair_t) * m);
    pair_t* b = (pair_t*)malloc(sizeof(pair_t) * m);
    for (size_t i = 0; i < m; ++i) { a[i].index = i + 1; a[i].coef = coeff + i + 1; } /* enumerate().skip(1) */
    /* stable bottom-up merge sort with `sort_by(|a, b| f(b, a))` semantics (:205): on a tie the left
     * (lower index) element stays first; take the right one only when f(left, right) < 0 */
    pair_t *src = a, *dst = b;
    for (size_t width_ = 1; width_ < m; width_ *= 2) {
        for (size_t lo = 0; lo < m; lo += 2 * width_) {
            size_t mid = lo + width_ < m ? lo + width_ : m, hi = lo + 2 * width_ < m ? lo + 2 * width_ : m;
            size_t i = lo, j = mid, k = lo;
            while (i < mid && j < hi) {
                if (f(src[i].index, *src[i].coef, src[j].index, *src[j].coef) < 0) dst[k++] = src[j++];
                else dst[k++] = src[i++];
            }
            while (i < mid) dst[k++] = src[i++];
            while (j < hi) dst[k++] = src[j++];
        }
        pair_t* t = src; src = dst; dst = t;
    }
    for (size_t i = 0; i < m; ++i) out[i] = (uint64_t)src[i].index;
    free(a); free(b);
}

/* ---------------------------------------------------------------- embed / extract / similarity */
static float insert_fn(int method, float alpha, float orig, float w) { /* :414-432 */
    if (method == 1) return orig + alpha * w;
    if (method == 2) return orig * (1.0f + alpha * w);
    return orig * expf(alpha * w);
}
static float extract_fn(int method, float alpha, float b, float d) { /* :566-593 */
    if (method == 1) return (d - b) / alpha;
    if (method == 2) return (d - b) / (b * alpha);
    return logf(d / b) / alpha;
}

void oracle_embed_watermark(float* coeff, const uint64_t* indices, size_t n_indices, const float* const* marks,
                            const size_t* lens, size_t n_marks, int method, float alpha, size_t n_coeff) { /* :382-410 */
    if (n_marks == 1) {
        const size_t n = lens[0] < n_indices ? lens[0] : n_indices;
        for (size_t i = 0; i < n; ++i) coeff[indices[i]] = insert_fn(method, alpha, coeff[indices[i]], marks[0][i]);
    } else {
        float* orig = (float*)malloc(sizeof(float) * n_coeff);
        memcpy(orig, coeff, sizeof(float) * n_coeff);
        for (size_t m = 0; m < n_marks; ++m) {
            const size_t n = lens[m] < n_indices ? lens[m] : n_indices;
            for (size_t i = 0; i < n; ++i) {
                const float updated = insert_fn(method, alpha, orig[indices[i]], marks[m][i]);
                const float change = updated - orig[indices[i]];
                coeff[indices[i]] += change;
            }
        }
        free(orig);
    }
}

float oracle_similarity(const float* extracted, const float* mark, size_t n) { /* :696-714 */
    float nominator = 0.0f, denominator = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        nominator += extracted[i] * mark[i];
        denominator += extracted[i] * extracted[i];
    }
    return nominator / sqrtf(denominator);
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* Writer::new(img, cfg).mark(&[mark]).into_rgb8().  timings (optional, 6 doubles, seconds):
 * rgb->yiq, forward dct, full sort, embed, inverse dct, yiq->rgb8.  coeff_out/idx_out optional. */
int oracle_embed_rgb8(const uint8_t* rgb, int width, int height, const float* mark, size_t n, int method, float alpha,
                      int ordering, uint8_t* out_rgb, float* coeff_out, uint64_t* idx_out, double* timings) {
    const size_t np = (size_t)width * height;
    float* y = (float*)malloc(sizeof(float) * np);
    float* i = (float*)malloc(sizeof(float) * np);
    float* q = (float*)malloc(sizeof(float) * np);
    uint64_t* idx = (uint64_t*)malloc(sizeof(uint64_t) * (np > 1 ? np - 1 : 1));
    double t0 = now_s();
    rgb8_to_yiq(rgb, np, y, i, q);
    double t1 = now_s();
    oracle_dct2_2d(T_DCT2, width, height, y);
    double t2 = now_s();
    oracle_obtain_indices(y, width, height, ordering, idx);
    double t3 = now_s();
    if (coeff_out) memcpy(coeff_out, y, sizeof(float) * np);
    if (idx_out) memcpy(idx_out, idx, sizeof(uint64_t) * (n < np - 1 ? n : np - 1));
    const float* marks[1] = {mark};
    size_t lens[1] = {n};
    double t3b = now_s();
    oracle_embed_watermark(y, idx, np - 1, marks, lens, 1, method, alpha, np);
    double t4 = now_s();
    oracle_dct2_2d(T_DCT3, width, height, y);
    double t5 = now_s();
    yiq_to_rgb8(y, i, q, np, out_rgb);
    double t6 = now_s();
    if (timings) {
        timings[0] = t1 - t0; timings[1] = t2 - t1; timings[2] = t3 - t2;
        timings[3] = t4 - t3b; timings[4] = t5 - t4; timings[5] = t6 - t5;
    }
    free(y); free(i); free(q); free(idx);
    return 0;
}

/* Reader::base + Reader::derived + extract (+ Tester::similarity when mark != NULL).
 * timings (optional, 4 doubles): 2x rgb->yiq, 2x forward dct, full sort, gather+similarity. */
int oracle_extract_rgb8(const uint8_t* base_rgb, const uint8_t* derived_rgb, int width, int height, size_t n,
                        int method, float alpha, int ordering, float* extracted, const float* mark, float* sim,
                        double* timings) {
    const size_t np = (size_t)width * height;
    if (n >= np) return -1; /* :553-555 */
    float* yb = (float*)malloc(sizeof(float) * np);
    float* yd = (float*)malloc(sizeof(float) * np);
    float* i = (float*)malloc(sizeof(float) * np);
    float* q = (float*)malloc(sizeof(float) * np);
    uint64_t* idx = (uint64_t*)malloc(sizeof(uint64_t) * (np > 1 ? np - 1 : 1));
    double t0 = now_s();
    rgb8_to_yiq(base_rgb, np, yb, i, q);
    rgb8_to_yiq(derived_rgb, np, yd, i, q);
    double t1 = now_s();
    oracle_dct2_2d(T_DCT2, width, height, yb);
    oracle_dct2_2d(T_DCT2, width, height, yd);
    double t2 = now_s();
    oracle_obtain_indices(yb, width, height, ordering, idx);
    double t3 = now_s();
    for (size_t k = 0; k < n; ++k) extracted[k] = extract_fn(method, alpha, yb[idx[k]], yd[idx[k]]);
    if (mark && sim) *sim = oracle_similarity(extracted, mark, n);
    double t4 = now_s();
    if (timings) { timings[0] = t1 - t0; timings[1] = t2 - t1; timings[2] = t3 - t2; timings[3] = t4 - t3; }
    free(yb); free(yd); free(i); free(q); free(idx);
    return 0;
}

int oracle_forward_rgb8(const uint8_t* rgb, int width, int height, float* coeff) {
    const size_t np = (size_t)width * height;
    float* i = (float*)malloc(sizeof(float) * np);
    float* q = (float*)malloc(sizeof(float) * np);
    rgb8_to_yiq(rgb, np, coeff, i, q);
    oracle_dct2_2d(T_DCT2, width, height, coeff);
    free(i); free(q);
    return 0;
}

/* Synthetic natural-image-like RGB8 frame of the benchmark (value-noise octaves 2..8 +
 * dither, integer only).  Not part of the reference: the C form of ssw_oracle.py:synth_frame, so that the reference
 * arm of bench.py can build its inputs without a GPU and without libssw.  rows [row0, row0 + n_rows) of frame `img`. */
static uint64_t sm64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void oracle_synth_rows(int width, uint64_t seed, uint32_t img, uint32_t row0, uint32_t n_rows, uint8_t* out) {
    const uint64_t base = seed ^ ((uint64_t)img * 0x9E3779B97F4A7C15ull);
    for (uint32_t yl = 0; yl < n_rows; ++yl) {
        const uint64_t y = (uint64_t)row0 + yl;
        for (uint64_t x = 0; x < (uint64_t)width; ++x) {
            uint8_t* o = out + ((size_t)yl * (size_t)width + x) * 3;
            for (uint64_t c = 0; c < 3; ++c) {
                uint64_t acc = 0;
                for (uint64_t oct = 2; oct <= 8; ++oct) {
                    const uint64_t s = 1ull << oct;
                    const uint64_t X = x >> oct, Y = y >> oct, fx = x & (s - 1), fy = y & (s - 1);
                    const uint64_t tag = base ^ (oct << 58) ^ (c << 56);
                    const uint64_t l00 = sm64(tag ^ (Y << 28) ^ X) >> 56, l10 = sm64(tag ^ (Y << 28) ^ (X + 1)) >> 56;
                    const uint64_t l01 = sm64(tag ^ ((Y + 1) << 28) ^ X) >> 56, l11 = sm64(tag ^ ((Y + 1) << 28) ^ (X + 1)) >> 56;
                    const uint64_t top = l00 * (s - fx) + l10 * fx, bot = l01 * (s - fx) + l11 * fx;
                    acc += ((top * (s - fy) + bot * fy) >> (2 * oct)) << oct;
                }
                const int64_t noise = (int64_t)(sm64(base ^ 0xABCDEFull ^ (c << 56) ^ (y << 28) ^ x) >> 61);
                int64_t v = (int64_t)(acc / 508ull) + noise - 4;
                o[c] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
            }
        }
    }
}
