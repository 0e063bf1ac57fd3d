/*
 * libdf_oracle.c -- CPU restatement of the reference's Rust DSP core (libDF).
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA hot
 * path in deepfilternet_b200/csrc.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference leg may load the library built
 * from it.  The product path never calls into oracle/.
 *
 * Each function cites the reference lines it follows (paths relative to
 * /root/reference).  The reference is single threaded Rust (f32 arithmetic).
 *
 * Third-party arithmetic: the reference's FFT is `realfft 3.3.0` on top of
 * `rustfft 6.2.0` (Cargo.lock:3696,3870; libDF/Cargo.toml:97-98); neither crate
 * is vendored under /root/reference.  Call sites: libDF/src/lib.rs:117-118
 * (plans), :385-388 (forward), :398-405 (inverse).  Both compute the
 * unnormalised DFT / inverse DFT of a real length-N sequence.  The oracle
 * evaluates that definition with a mixed-radix FFT in f64 and rounds the result
 * to f32, so it differs from the f32 reference FFT only by f32 round-off
 * (~1e-7 relative); no reference test pins FFT bits.
 *
 * Parity pinning: tests/test_oracle_golden.py checks this file against
 *   (1) the ERB widths stored in the shipped DeepFilterNet{2,3} checkpoints
 *       (`erb_fb` buffer, bit exact),
 *   (2) the STFT->ISTFT reconstruction test of libDF/src/transforms.rs:618-638,
 *   (3) the band-gain equality test of libDF/src/lib.rs:626-652,
 *   (4) the SI-SDR known answers of DeepFilterNet/df/scripts/test_df.py:44-78
 *       (through the reference PyTorch modules, fixtures in tests/golden).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ ERB -- */

/* libDF/src/lib.rs:42-47 (f32 ln_1p / exp) */
static float freq2erb(float f) { return 9.265f * log1pf(f / (24.7f * 9.265f)); }
static float erb2freq(float e) { return 24.7f * 9.265f * (expf(e / 9.265f) - 1.f); }

/* libDF/src/lib.rs:68-100.  Integer result, must be bit exact. */
int dfo_erb_widths(int sr, int fft_size, int nb_bands, int min_nb_freqs, int64_t *out) {
    int nyq = sr / 2;
    float freq_width = (float)sr / (float)fft_size;
    float erb_low = freq2erb(0.f);
    float erb_high = freq2erb((float)nyq);
    float step = (erb_high - erb_low) / (float)nb_bands;
    int prev_freq = 0, freq_over = 0;
    for (int i = 1; i <= nb_bands; i++) {
        float f = erb2freq(erb_low + (float)i * step);
        int fb = (int)roundf(f / freq_width);
        int nb_freqs = fb - prev_freq - freq_over;
        if (nb_freqs < min_nb_freqs) {
            freq_over = min_nb_freqs - nb_freqs;
            nb_freqs = min_nb_freqs;
        } else {
            freq_over = 0;
        }
        out[i - 1] = nb_freqs;
        prev_freq = fb;
    }
    out[nb_bands - 1] += 1;
    int64_t sum = 0;
    for (int i = 0; i < nb_bands; i++) sum += out[i];
    int64_t too_large = sum - (fft_size / 2 + 1);
    if (too_large > 0) out[nb_bands - 1] -= too_large;
    return 0;
}

/* ---------------------------------------------------------------- state -- */

typedef struct {
    int sr, fft_size, hop_size, freq_size, nb_erb;
    float *window;        /* [fft_size]  lib.rs:126-132 */
    float wnorm;          /* lib.rs:133 */
    int64_t *erb;         /* [nb_erb]    lib.rs:124 */
    float *analysis_mem;  /* [fft-hop]   lib.rs:119 */
    float *synthesis_mem; /* [fft-hop]   lib.rs:120 */
    /* f64 FFT plan */
    int nfac, fac[32];
    double *tw_re, *tw_im; /* e^{-2 pi i k / N}, k in [0,N) */
    double *wr, *wi, *sr_, *si_;
} dfo_state;

static void factorize(int n, int *nfac, int *fac) {
    *nfac = 0;
    int p[] = {4, 2, 3, 5};
    for (int i = 0; i < 4; i++)
        while (n % p[i] == 0) { fac[(*nfac)++] = p[i]; n /= p[i]; }
    for (int q = 7; n > 1; q += 2)
        while (n % q == 0) { fac[(*nfac)++] = q; n /= q; }
}

/* lib.rs:104-154 */
dfo_state *dfo_create(int sr, int fft_size, int hop_size, int nb_bands, int min_nb_freqs) {
    if (hop_size * 2 > fft_size) return NULL; /* assert at lib.rs:111 */
    dfo_state *s = (dfo_state *)calloc(1, sizeof(dfo_state));
    s->sr = sr; s->fft_size = fft_size; s->hop_size = hop_size;
    s->freq_size = fft_size / 2 + 1; s->nb_erb = nb_bands;
    s->window = (float *)malloc(sizeof(float) * fft_size);
    int wh = fft_size / 2;
    for (int i = 0; i < fft_size; i++) {
        double sn = sin(0.5 * M_PI * ((double)i + 0.5) / (double)wh);
        s->window[i] = (float)sin(0.5 * M_PI * sn * sn);
    }
    s->wnorm = 1.f / ((float)((int64_t)fft_size * fft_size) / (float)(2 * hop_size));
    s->erb = (int64_t *)malloc(sizeof(int64_t) * nb_bands);
    dfo_erb_widths(sr, fft_size, nb_bands, min_nb_freqs, s->erb);
    s->analysis_mem = (float *)calloc(fft_size - hop_size, sizeof(float));
    s->synthesis_mem = (float *)calloc(fft_size - hop_size, sizeof(float));
    factorize(fft_size, &s->nfac, s->fac);
    s->tw_re = (double *)malloc(sizeof(double) * fft_size);
    s->tw_im = (double *)malloc(sizeof(double) * fft_size);
    for (int k = 0; k < fft_size; k++) {
        s->tw_re[k] = cos(2.0 * M_PI * k / fft_size);
        s->tw_im[k] = -sin(2.0 * M_PI * k / fft_size);
    }
    s->wr = (double *)malloc(sizeof(double) * fft_size);
    s->wi = (double *)malloc(sizeof(double) * fft_size);
    s->sr_ = (double *)malloc(sizeof(double) * fft_size);
    s->si_ = (double *)malloc(sizeof(double) * fft_size);
    return s;
}

void dfo_free(dfo_state *s) {
    if (!s) return;
    free(s->window); free(s->erb); free(s->analysis_mem); free(s->synthesis_mem);
    free(s->tw_re); free(s->tw_im); free(s->wr); free(s->wi); free(s->sr_); free(s->si_);
    free(s);
}

/* lib.rs:156-159 */
void dfo_reset(dfo_state *s) {
    memset(s->analysis_mem, 0, sizeof(float) * (s->fft_size - s->hop_size));
    memset(s->synthesis_mem, 0, sizeof(float) * (s->fft_size - s->hop_size));
}

int dfo_get_erb_widths(const dfo_state *s, int64_t *out) {
    memcpy(out, s->erb, sizeof(int64_t) * s->nb_erb); return s->nb_erb;
}
int dfo_get_window(const dfo_state *s, float *out) {
    memcpy(out, s->window, sizeof(float) * s->fft_size); return s->fft_size;
}
float dfo_get_wnorm(const dfo_state *s) { return s->wnorm; }

/* --------------------------------------------------------------- f64 FFT -- */

/* Recursive decimation-in-time mixed radix complex DFT of length n
 * (sign -1 when inv == 0, +1 when inv != 0), unnormalised.
 * in: stride `is`; out: contiguous.  Twiddles from the size-N table with
 * stride N/n. */
static void fft_rec(const dfo_state *s, int n, int fi, const double *xr, const double *xi, int is,
                    double *yr, double *yi, double *tr, double *ti, int inv) {
    const int N = s->fft_size;
    if (n == 1) { yr[0] = xr[0]; yi[0] = xi[0]; return; }
    int p = s->fac[fi];
    int m = n / p;
    /* p sub-transforms of length m over the decimated inputs */
    for (int r = 0; r < p; r++)
        fft_rec(s, m, fi + 1, xr + (size_t)r * is, xi + (size_t)r * is, is * p, tr + r * m, ti + r * m,
                yr + r * m, yi + r * m, inv);
    int tstride = N / n;
    for (int k = 0; k < m; k++) {
        for (int q = 0; q < p; q++) {
            /* Y[k + q m] = sum_r W_n^{r (k + q m)} T_r[k] */
            double ar = 0, ai = 0;
            for (int r = 0; r < p; r++) {
                int e = (int)(((int64_t)r * (k + q * m)) % n) * tstride;
                double wr = s->tw_re[e], wi = inv ? -s->tw_im[e] : s->tw_im[e];
                double vr = tr[r * m + k], vi = ti[r * m + k];
                ar += vr * wr - vi * wi;
                ai += vr * wi + vi * wr;
            }
            yr[k + q * m] = ar; yi[k + q * m] = ai;
        }
    }
}

static void cfft(dfo_state *s, double *xr, double *xi, double *yr, double *yi, int inv) {
    fft_rec(s, s->fft_size, 0, xr, xi, 1, yr, yi, s->sr_, s->si_, inv);
}

/* ------------------------------------------------------ frame analysis -- */

/* lib.rs:356-394.  input [hop], output [freq_size] interleaved re/im. */
void dfo_frame_analysis(dfo_state *s, const float *input, float *output) {
    const int N = s->fft_size, H = s->hop_size, M = N - H;
    double *br = s->wr, *bi = s->wi;
    /* First part of the window on the previous frame(s) (:365-371); products in f32 */
    for (int i = 0; i < M; i++) { br[i] = (double)(s->analysis_mem[i] * s->window[i]); bi[i] = 0; }
    /* Second part of the window on the new input frame (:373-375) */
    for (int i = 0; i < H; i++) { br[M + i] = (double)(input[i] * s->window[M + i]); bi[M + i] = 0; }
    /* Shift analysis_mem (:376-384) */
    int split = M - H;
    if (split > 0) memmove(s->analysis_mem, s->analysis_mem + H, sizeof(float) * split);
    memcpy(s->analysis_mem + split, input, sizeof(float) * H);
    /* forward real FFT (:385-388), then wnorm in f32 (:390-393) */
    double *yr = (double *)malloc(sizeof(double) * N * 2), *yi = yr + N;
    cfft(s, br, bi, yr, yi, 0);
    for (int k = 0; k < s->freq_size; k++) {
        output[2 * k] = (float)yr[k] * s->wnorm;
        output[2 * k + 1] = (float)yi[k] * s->wnorm;
    }
    free(yr);
}

/* lib.rs:396-427.  input [freq_size] interleaved, output [hop]. */
void dfo_frame_synthesis(dfo_state *s, const float *input, float *output) {
    const int N = s->fft_size, H = s->hop_size, M = N - H, F = s->freq_size;
    double *xr = s->wr, *xi = s->wi;
    /* Hermitian extension; imag of DC and Nyquist are discarded (realfft
     * ComplexToReal ignores them and only reports InputValues, which lib.rs:402
     * swallows). */
    for (int k = 0; k < F; k++) { xr[k] = input[2 * k]; xi[k] = input[2 * k + 1]; }
    xi[0] = 0;
    if (N % 2 == 0) xi[N / 2] = 0;
    for (int k = F; k < N; k++) { xr[k] = xr[N - k]; xi[k] = -xi[N - k]; }
    double *yr = (double *)malloc(sizeof(double) * N * 2), *yi = yr + N;
    cfft(s, xr, xi, yr, yi, 1);
    float *x = (float *)malloc(sizeof(float) * N);
    for (int i = 0; i < N; i++) x[i] = (float)yr[i] * s->window[i]; /* :406 */
    for (int i = 0; i < H; i++) output[i] = x[i] + s->synthesis_mem[i]; /* :407-411 */
    int split = M - H;
    if (split > 0) memmove(s->synthesis_mem, s->synthesis_mem + H, sizeof(float) * split); /* rotate_left :415 */
    for (int i = 0; i < split; i++) s->synthesis_mem[i] += x[H + i];        /* :419-422 */
    for (int i = split; i < M; i++) s->synthesis_mem[i] = x[H + i];         /* :423-426 */
    free(x); free(yr);
}

/* --------------------------------------------- batched wrappers (pyDF) -- */

/* pyDF/src/lib.rs:41-72: input f32[C,T] -> c64[C, T/hop, F]; state reset per channel */
int dfo_analysis(dfo_state *s, const float *input, int64_t C, int64_t T, int reset, float *out) {
    int64_t Tf = T / s->hop_size;
    for (int64_t c = 0; c < C; c++) {
        if (reset) dfo_reset(s);
        for (int64_t t = 0; t < Tf; t++)
            dfo_frame_analysis(s, input + c * T + t * s->hop_size,
                               out + ((c * Tf + t) * s->freq_size) * 2);
    }
    return 0;
}

/* pyDF/src/lib.rs:74-107: c64[C,Tf,F] -> f32[C, Tf*hop] */
int dfo_synthesis(dfo_state *s, const float *input, int64_t C, int64_t Tf, int reset, float *out) {
    for (int64_t c = 0; c < C; c++) {
        if (reset) dfo_reset(s);
        for (int64_t t = 0; t < Tf; t++)
            dfo_frame_synthesis(s, input + ((c * Tf + t) * s->freq_size) * 2,
                                out + (c * Tf + t) * s->hop_size);
    }
    return 0;
}

/* lib.rs:280-295 (compute_band_corr with x == p) + dB lib.rs:207-210 /
 * transforms.rs:236-253.  input c64[n_frames, F] -> f32[n_frames, E]. */
int dfo_erb(const float *spec, int64_t n_frames, int64_t F, const int64_t *erb_fb, int E, int db,
            float *out) {
    int64_t sum = 0;
    for (int b = 0; b < E; b++) sum += erb_fb[b];
    if (sum != F) return -1;
    for (int64_t t = 0; t < n_frames; t++) {
        const float *x = spec + t * F * 2;
        int64_t bc = 0;
        for (int b = 0; b < E; b++) {
            float k = 1.f / (float)erb_fb[b];
            float acc = 0.f;
            for (int64_t j = 0; j < erb_fb[b]; j++) {
                int64_t i = bc + j;
                acc += (x[2 * i] * x[2 * i] + x[2 * i + 1] * x[2 * i + 1]) * k;
            }
            bc += erb_fb[b];
            out[t * E + b] = db ? log10f(acc + 1e-10f) * 10.f : acc;
        }
    }
    return 0;
}

/* lib.rs:328-337 / transforms.rs:285-299.  f32[n,E] -> f32[n,F] */
int dfo_erb_inv(const float *gains, int64_t n_frames, const int64_t *erb_fb, int E, float *out) {
    int64_t F = 0;
    for (int b = 0; b < E; b++) F += erb_fb[b];
    for (int64_t t = 0; t < n_frames; t++) {
        int64_t bc = 0;
        for (int b = 0; b < E; b++) {
            for (int64_t j = 0; j < erb_fb[b]; j++) out[t * F + bc + j] = gains[t * E + b];
            bc += erb_fb[b];
        }
    }
    return 0;
}

/* lib.rs:314-326 / transforms.rs:255-273: spec c64[n,F] *= gains[n,E] (in place) */
int dfo_apply_erb_gains(float *spec, const float *gains, int64_t n_frames, const int64_t *erb_fb,
                        int E) {
    int64_t F = 0;
    for (int b = 0; b < E; b++) F += erb_fb[b];
    for (int64_t t = 0; t < n_frames; t++) {
        int64_t bc = 0;
        for (int b = 0; b < E; b++) {
            float g = gains[t * E + b];
            for (int64_t j = 0; j < erb_fb[b]; j++) {
                spec[(t * F + bc + j) * 2] *= g;
                spec[(t * F + bc + j) * 2 + 1] *= g;
            }
            bc += erb_fb[b];
        }
    }
    return 0;
}

/* ndarray linspace(a, b, n): a + i * (b - a) / (n - 1)   (transforms.rs:310,341) */
static void linspace_f32(float a, float b, int64_t n, float *out) {
    if (n == 1) { out[0] = a; return; }
    float step = (b - a) / (float)(n - 1);
    for (int64_t i = 0; i < n; i++) out[i] = a + (float)i * step;
}

void dfo_unit_norm_init(int64_t n, float *out) { linspace_f32(0.001f, 0.0001f, n, out); }
void dfo_mean_norm_init(int64_t n, float *out) { linspace_f32(-60.f, -90.f, n, out); }

/* transforms.rs:301-330 + lib.rs:244-251.  erb f32[C,T,E] in place; state f32[C,E] or NULL */
int dfo_erb_norm(float *erb, int64_t C, int64_t T, int64_t E, float alpha, const float *state_in) {
    float *st = (float *)malloc(sizeof(float) * E);
    for (int64_t c = 0; c < C; c++) {
        if (state_in) memcpy(st, state_in + c * E, sizeof(float) * E);
        else dfo_mean_norm_init(E, st);
        for (int64_t t = 0; t < T; t++) {
            float *x = erb + (c * T + t) * E;
            for (int64_t b = 0; b < E; b++) {
                st[b] = x[b] * (1.f - alpha) + st[b] * alpha;
                x[b] -= st[b];
                x[b] /= 40.f;
            }
        }
    }
    free(st);
    return 0;
}

/* transforms.rs:332-361 + lib.rs:253-259.  spec c64[C,T,F] in place. */
int dfo_unit_norm(float *spec, int64_t C, int64_t T, int64_t F, float alpha, const float *state_in) {
    float *st = (float *)malloc(sizeof(float) * F);
    for (int64_t c = 0; c < C; c++) {
        if (state_in) memcpy(st, state_in + c * F, sizeof(float) * F);
        else dfo_unit_norm_init(F, st);
        for (int64_t t = 0; t < T; t++) {
            float *x = spec + (c * T + t) * F * 2;
            for (int64_t k = 0; k < F; k++) {
                float nrm = hypotf(x[2 * k], x[2 * k + 1]); /* Complex32::norm */
                st[k] = nrm * (1.f - alpha) + st[k] * alpha;
                float d = sqrtf(st[k]);
                x[2 * k] /= d;
                x[2 * k + 1] /= d;
            }
        }
    }
    free(st);
    return 0;
}
