// enc_analysis.cuh -- the float analysis front-end of the CELT encoder, ONE WARP PER STREAM.
//
// Stages (reference lines in each function): pre-emphasis, the pitch pre-filter (2:1 down-sampling + LPC whitening, coarse / fine
// cross-correlation with one lag per lane, the sub-harmonic check with one candidate period per lane, the comb filter as a plain
// FIR over lanes), transient detection (the high-pass and the two masking envelopes are linear recurrences evaluated as
// chunk-per-lane scans), the forward MDCT (window/fold fused into the pre-rotation, every butterfly of an FFT stage a work item,
// all short blocks at once), band energies / normalisation as warp reductions, and the decision heuristics that turn floats into
// the integers the bitstream carries (dynalloc, tf, spread, stereo, trim, VBR target).
//
// Execution model: warp-uniform control.  Scalars live in registers of every lane (all lanes compute the same value), vectors are
// strided over the lanes, everything that crosses lanes goes through ob_coop.cuh.  Small per-band arrays live in shared memory and
// are written with the same value by every lane.  The same source compiles for the host with one lane, either in the reference's
// summation order (ObSolo: bit-identical to the reference's C build) or in the warp's order (ObSoloW: what the GPU must produce).
#pragma once
#include "enc_range.cuh"
#include "ob_coop.cuh"
#include "dec_synth.cuh"     // ob_fft_stage, tables

#ifndef __CUDACC__
#include <math.h>
#include <string.h>
#endif

#define OB_MAXPERIOD 1024           // COMBFILTER_MAXPERIOD (celt.h:218)
#define OB_MINPERIOD 15

OB_DEV float ob_fmax(float a, float b) { return a > b ? a : b; }     // MAX16/MAX32 (arch.h): (a) > (b) ? (a) : (b)
OB_DEV float ob_fmin(float a, float b) { return a < b ? a : b; }
OB_DEV float ob_log2(float x) { return (float)(1.442695040888963387 * log((double)x)); }     // celt_log2 (mathops.h:168)
OB_DEV float ob_exp2(float x) { return (float)exp(0.6931471805599453094 * (double)x); }      // celt_exp2 (mathops.h:169)

// celt_preemphasis (celt_encoder.c:507-578): inp[i] = 32768*x[i] - coef0 * 32768*x[i-1] -- a two-tap FIR, one sample per lane
template <class G>
OB_STAGE void ob_preemphasis(const G &g, const float *__restrict__ pcm, float *__restrict__ inp, int N, int CC, float *mem, int clip)
{
    const float coef0 = OB_PREEMPH[0], m_in = *mem;
    for (int i = g.lane; i < N; i += g.n) {
        float x = pcm[CC * i] * 32768.f, xp = i > 0 ? pcm[CC * (i - 1)] * 32768.f : 0.f;
        if (clip) { x = ob_fmax(-65536.f, ob_fmin(65536.f, x)); xp = ob_fmax(-65536.f, ob_fmin(65536.f, xp)); }
        inp[i] = x - (i > 0 ? coef0 * xp : m_in);
    }
    float xl = pcm[CC * (N - 1)] * 32768.f;
    if (clip) xl = ob_fmax(-65536.f, ob_fmin(65536.f, xl));
    *mem = coef0 * xl;
    g.sync();
}

// comb_filter (celt.c:190-256) with y != x: the encoder's pre-filter reads only the unfiltered signal, so every output sample is independent
template <class G>
OB_STAGE void ob_comb_filter_xy(const G &g, float *y, const float *x, int T0, int T1, int N, float g0, float g1, int tapset0, int tapset1, int overlap)
{
    const float gains[3][3] = {{0.3066406250f, 0.2170410156f, 0.1296386719f}, {0.4638671875f, 0.2680664062f, 0.f}, {0.7998046875f, 0.1000976562f, 0.f}};
    if (g0 == 0 && g1 == 0) { for (int i = g.lane; i < N; i += g.n) y[i] = x[i]; g.sync(); return; }
    T0 = ob_imax(T0, OB_MINPERIOD); T1 = ob_imax(T1, OB_MINPERIOD);
    const float g00 = g0 * gains[tapset0][0], g01 = g0 * gains[tapset0][1], g02 = g0 * gains[tapset0][2];
    const float g10 = g1 * gains[tapset1][0], g11 = g1 * gains[tapset1][1], g12 = g1 * gains[tapset1][2];
    if (g0 == g1 && T0 == T1 && tapset0 == tapset1) overlap = 0;
    for (int i = g.lane; i < N; i += g.n) {
        const float x0 = x[i - T1 + 2], x1 = x[i - T1 + 1], x2 = x[i - T1], x3 = x[i - T1 - 1], x4 = x[i - T1 - 2];
        if (i < overlap) {
            const float f = OB_WINDOW[i] * OB_WINDOW[i];
            y[i] = x[i] + ((1.0f - f) * g00) * x[i - T0] + ((1.0f - f) * g01) * (x[i - T0 + 1] + x[i - T0 - 1])
                        + ((1.0f - f) * g02) * (x[i - T0 + 2] + x[i - T0 - 2])
                        + (f * g10) * x2 + (f * g11) * (x1 + x3) + (f * g12) * (x0 + x4);
        } else if (g1 == 0) y[i] = x[i];
        else y[i] = x[i] + g10 * x2 + g11 * (x1 + x3) + g12 * (x0 + x4);          // comb_filter_const_c (celt.c:162-185)
    }
    g.sync();
}

template <class G>
OB_DEV float ob_maxabs(const G &g, const float *x, int len)            // celt_maxabs16 (mathops.h:79-91)
{
    return ob_pmax(g, len, 0.f, [&](int i) { return fabsf(x[i]); });
}

// ---- pitch analysis (opus/celt/pitch.c) ----------------------------------------------------------------------------
// serial in-order inner product (one lane): the lane-per-lag / lane-per-candidate stages below keep the reference's summation order
OB_DEV float ob_inner_prod(const float *x, const float *y, int N) { float s = 0; for (int i = 0; i < N; i++) s = s + x[i] * y[i]; return s; }

// celt_pitch_xcorr_c (pitch.c:225-300): one lag per lane (four per lane and pass, sharing every x[j]); each lag is the reference's in-order sum
template <class G>
OB_DEV void ob_pitch_xcorr(const G &g, const float *x, const float *y, float *xcorr, int len, int max_pitch)
{
    for (int base = 0; base < max_pitch; base += 4 * g.n) {
        const int i0 = base + g.lane, i1 = i0 + g.n, i2 = i1 + g.n, i3 = i2 + g.n;
        const float *y0 = y + ob_imin(i0, max_pitch - 1), *y1 = y + ob_imin(i1, max_pitch - 1), *y2 = y + ob_imin(i2, max_pitch - 1), *y3 = y + ob_imin(i3, max_pitch - 1);
        float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        for (int j = 0; j < len; j++) {
            const float xj = x[j];
            s0 = s0 + xj * y0[j]; s1 = s1 + xj * y1[j]; s2 = s2 + xj * y2[j]; s3 = s3 + xj * y3[j];
        }
        if (i0 < max_pitch) xcorr[i0] = s0;
        if (i1 < max_pitch) xcorr[i1] = s1;
        if (i2 < max_pitch) xcorr[i2] = s2;
        if (i3 < max_pitch) xcorr[i3] = s3;
    }
    g.sync();
}

// pitch_downsample (pitch.c:140-217) incl. _celt_autocorr (celt_lpc.c:277-351, lag 4, no window), _celt_lpc (:37-91), celt_fir5 (pitch.c:105-137).
// raw: n floats of scratch (the decimated signal before whitening); x_lp: n floats out.
template <class G>
OB_STAGE void ob_pitch_downsample(const G &g, const float *x0, const float *x1, float *x_lp, float *raw, int len, int C)
{
    const int n = len >> 1;
    for (int i = g.lane; i < n; i += g.n) {
        float v = i > 0 ? .25f * x0[2 * i - 1] + .25f * x0[2 * i + 1] + .5f * x0[2 * i] : .25f * x0[1] + .5f * x0[0];
        if (C == 2) v += i > 0 ? .25f * x1[2 * i - 1] + .25f * x1[2 * i + 1] + .5f * x1[2 * i] : .25f * x1[1] + .5f * x1[0];
        raw[i] = v;
    }
    g.sync();
    float ac[5];
    {
        const int lag = 4, fastN = n - lag;
        for (int k = 0; k <= lag; k++) {
            ac[k] = ob_psum(g, fastN, 0.f, [&](int j) { return raw[j] * raw[j + k]; });
            float d = 0;
            for (int i = k + fastN; i < n; i++) d = d + raw[i] * raw[i - k];
            ac[k] += d;
        }
    }
    ac[0] *= 1.0001f;
    for (int i = 1; i <= 4; i++) ac[i] -= ac[i] * (.008f * i) * (.008f * i);
    float lpc[4] = {0, 0, 0, 0};
    {
        float error = ac[0];
        if (ac[0] > 1e-10f) {
            for (int i = 0; i < 4; i++) {
                float rr = 0;
                for (int j = 0; j < i; j++) rr += lpc[j] * ac[i - j];
                rr += ac[i + 1];
                const float r = -(rr / error);
                lpc[i] = r;
                for (int j = 0; j < (i + 1) >> 1; j++) {
                    const float t1 = lpc[j], t2 = lpc[i - 1 - j];
                    lpc[j] = t1 + r * t2;
                    lpc[i - 1 - j] = t2 + r * t1;
                }
                error = error - (r * r) * error;
                if (error <= .001f * ac[0]) break;
            }
        }
    }
    float tmp = 1.0f;
    for (int i = 0; i < 4; i++) { tmp = .9f * tmp; lpc[i] = lpc[i] * tmp; }
    const float c1 = .8f;
    const float num0 = lpc[0] + .8f, num1 = lpc[1] + c1 * lpc[0], num2 = lpc[2] + c1 * lpc[1], num3 = lpc[3] + c1 * lpc[2], num4 = c1 * lpc[3];
    for (int i = g.lane; i < n; i += g.n) {                            // celt_fir5: the memories are the five previous INPUT samples
        float sum = raw[i];
        sum = sum + num0 * (i >= 1 ? raw[i - 1] : 0.f); sum = sum + num1 * (i >= 2 ? raw[i - 2] : 0.f); sum = sum + num2 * (i >= 3 ? raw[i - 3] : 0.f);
        sum = sum + num3 * (i >= 4 ? raw[i - 4] : 0.f); sum = sum + num4 * (i >= 5 ? raw[i - 5] : 0.f);
        x_lp[i] = sum;
    }
    g.sync();
}

OB_DEV void ob_find_best_pitch(const float *xcorr, const float *y, int len, int max_pitch, int *best_pitch, float Syy)   // pitch.c:45-103 (Syy = 1 + sum y^2)
{
    float best_num[2] = {-1, -1}, best_den[2] = {0, 0};
    best_pitch[0] = 0; best_pitch[1] = 1;
    for (int i = 0; i < max_pitch; i++) {
        if (xcorr[i] > 0) {
            float xcorr16 = xcorr[i];
            xcorr16 *= 1e-12f;
            const float num = xcorr16 * xcorr16;
            if (num * best_den[1] > best_num[1] * Syy) {
                if (num * best_den[0] > best_num[0] * Syy) {
                    best_num[1] = best_num[0]; best_den[1] = best_den[0]; best_pitch[1] = best_pitch[0];
                    best_num[0] = num; best_den[0] = Syy; best_pitch[0] = i;
                } else { best_num[1] = num; best_den[1] = Syy; best_pitch[1] = i; }
            }
        }
        Syy += y[i + len] * y[i + len] - y[i] * y[i];
        Syy = ob_fmax(1, Syy);
    }
}

// pitch_search (pitch.c:302-411).  scratch: >= (len>>2) + ((len+max_pitch)>>2) + (max_pitch>>1) floats.
template <class G>
OB_STAGE void ob_pitch_search(const G &g, const float *x_lp, const float *y, int len, int max_pitch, int *pitch, float *scratch)
{
    const int lag = len + max_pitch;
    float *x_lp4 = scratch, *y_lp4 = x_lp4 + (len >> 2), *xcorr = y_lp4 + (lag >> 2);
    int best_pitch[2] = {0, 0}, offset;
    for (int j = g.lane; j < len >> 2; j += g.n) x_lp4[j] = x_lp[2 * j];
    for (int j = g.lane; j < lag >> 2; j += g.n) y_lp4[j] = y[2 * j];
    g.sync();
    ob_pitch_xcorr(g, x_lp4, y_lp4, xcorr, len >> 2, max_pitch >> 2);
    float Syy = ob_psum(g, len >> 2, 1.f, [&](int j) { return y_lp4[j] * y_lp4[j]; });
    ob_find_best_pitch(xcorr, y_lp4, len >> 2, max_pitch >> 2, best_pitch, Syy);
    g.sync();
    for (int i = g.lane; i < max_pitch >> 1; i += g.n) xcorr[i] = 0;
    g.sync();
    for (int q = g.lane; q < 10; q += g.n) {                           // the (at most ten) lags within 2 of either candidate: one lane each
        const int i = q < 5 ? 2 * best_pitch[0] - 2 + q : 2 * best_pitch[1] - 2 + (q - 5);
        if (i < 0 || i >= max_pitch >> 1) continue;
        const float sum = ob_inner_prod(x_lp, y + i, len >> 1);
        xcorr[i] = ob_fmax(-1, sum);
    }
    g.sync();
    Syy = ob_psum(g, len >> 1, 1.f, [&](int j) { return y[j] * y[j]; });
    ob_find_best_pitch(xcorr, y, len >> 1, max_pitch >> 1, best_pitch, Syy);
    if (best_pitch[0] > 0 && best_pitch[0] < (max_pitch >> 1) - 1) {
        const float a = xcorr[best_pitch[0] - 1], b = xcorr[best_pitch[0]], c = xcorr[best_pitch[0] + 1];
        if ((c - a) > .7f * (b - a)) offset = 1;
        else if ((a - c) > .7f * (b - c)) offset = -1;
        else offset = 0;
    } else offset = 0;
    *pitch = 2 * best_pitch[0] - offset;
    g.sync();
}

OB_DEV float ob_pitch_gain(float xy, float xx, float yy) { return xy / sqrtf(1 + xx * yy); }       // pitch.c:441-444

// remove_doubling (pitch.c:449-555).  scratch: >= maxperiod/2 + 1 + 32 floats.  The fourteen sub-harmonic candidates are independent of
// each other (their thresholds depend on T0 and g0 only): all 28 correlations run at once, one lane each, then one uniform pass decides.
template <class G>
OB_STAGE float ob_remove_doubling(const G &g, const float *x, int maxperiod, int minperiod, int N, int *T0_, int prev_period, float prev_gain, float *scratch)
{
    const int second_check[16] = {0, 0, 3, 2, 3, 2, 5, 2, 3, 2, 3, 2, 5, 2, 3, 2};
    int k, T, T0, offset;
    float gg, g0, pg, xy, xx, yy, xcorr[3], best_xy, best_yy;
    const int minperiod0 = minperiod;
    maxperiod /= 2; minperiod /= 2; *T0_ /= 2; prev_period /= 2; N /= 2;
    float *yy_lookup = scratch, *cand = scratch + maxperiod + 1;
    x += maxperiod;
    if (*T0_ >= maxperiod) *T0_ = maxperiod - 1;
    T = T0 = *T0_;
    xx = 0; xy = 0;
    ob_psum2(g, N, xx, xy, [&](int i, float &a, float &b) { a = a + x[i] * x[i]; b = b + x[i] * x[i - T0]; });      // dual_inner_prod(x, x, x-T0)
    if (ObOrder<G>::value == 1) {
        if (g.lane == 0) {
            yy_lookup[0] = xx;
            float y2 = xx;
            for (int i = 1; i <= maxperiod; i++) { y2 = y2 + x[-i] * x[-i] - x[N - i] * x[N - i]; yy_lookup[i] = ob_fmax(0, y2); }
        }
    } else {
        if (g.lane == 0) yy_lookup[0] = xx;
        ob_prefix_sum(g, maxperiod, xx, [&](int i) { return x[-(i + 1)] * x[-(i + 1)] - x[N - (i + 1)] * x[N - (i + 1)]; },
                      [&](int i, float v) { yy_lookup[i + 1] = ob_fmax(0, v); });
    }
    g.sync();
    yy = yy_lookup[T0];
    best_xy = xy; best_yy = yy;
    gg = g0 = ob_pitch_gain(xy, xx, yy);
    for (int w = g.lane; w < 28; w += g.n) {
        k = 2 + (w >> 1);
        const int T1 = (int)((uint32_t)(2 * T0 + k) / (uint32_t)(2 * k));
        int Tq = T1;
        if (w & 1) {
            if (k == 2) Tq = T1 + T0 > maxperiod ? T0 : T0 + T1;
            else Tq = (int)((uint32_t)(2 * second_check[k] * T0 + k) / (uint32_t)(2 * k));
        }
        cand[w] = T1 < minperiod ? 0.f : ob_inner_prod(x, x - Tq, N);
    }
    g.sync();
    for (k = 2; k <= 15; k++) {
        int T1, T1b;
        float g1, cont, thresh;
        T1 = (int)((uint32_t)(2 * T0 + k) / (uint32_t)(2 * k));
        if (T1 < minperiod) break;
        if (k == 2) { if (T1 + T0 > maxperiod) T1b = T0; else T1b = T0 + T1; }
        else T1b = (int)((uint32_t)(2 * second_check[k] * T0 + k) / (uint32_t)(2 * k));
        xy = .5f * (cand[2 * (k - 2)] + cand[2 * (k - 2) + 1]);
        yy = .5f * (yy_lookup[T1] + yy_lookup[T1b]);
        g1 = ob_pitch_gain(xy, xx, yy);
        int dT = T1 - prev_period; if (dT < 0) dT = -dT;
        if (dT <= 1) cont = prev_gain;
        else if (dT <= 2 && 5 * k * k < T0) cont = .5f * prev_gain;
        else cont = 0;
        thresh = ob_fmax(.3f, .7f * g0 - cont);
        if (T1 < 3 * minperiod) thresh = ob_fmax(.4f, .85f * g0 - cont);
        else if (T1 < 2 * minperiod) thresh = ob_fmax(.5f, .9f * g0 - cont);
        if (g1 > thresh) { best_xy = xy; best_yy = yy; T = T1; gg = g1; }
    }
    best_xy = ob_fmax(0, best_xy);
    if (best_yy <= best_xy) pg = 1.0f;
    else pg = best_xy / (best_yy + 1);
    for (k = 0; k < 3; k++) { const float *xs = x - (T + k - 1); xcorr[k] = ob_psum(g, N, 0.f, [&](int i) { return x[i] * xs[i]; }); }
    if ((xcorr[2] - xcorr[0]) > .7f * (xcorr[1] - xcorr[0])) offset = 1;
    else if ((xcorr[0] - xcorr[2]) > .7f * (xcorr[1] - xcorr[2])) offset = -1;
    else offset = 0;
    if (pg > gg) pg = gg;
    *T0_ = 2 * T + offset;
    if (*T0_ < minperiod0) *T0_ = minperiod0;
    g.sync();
    return pg;
}

// transient_analysis (celt_encoder.c:227-419), allow_weak_transients = 0.  tmp: >= len floats, env: >= len/2 floats (both scratch).
// The high-pass filter is a second-order and the two masking envelopes are first-order linear recurrences: chunk-per-lane scans.
OB_TABLE(uint8_t, OB_INV_TABLE, 128) = {                               // inv_table (celt_encoder.c:246-255)
    255, 255, 156, 110, 86, 70, 59, 51, 45, 40, 37, 33, 31, 28, 26, 25, 23, 22, 21, 20, 19, 18, 17, 16, 16, 15, 15, 14, 13, 13, 12, 12,
    12, 12, 11, 11, 11, 10, 10, 10, 9, 9, 9, 9, 9, 9, 8, 8, 8, 8, 8, 7, 7, 7, 7, 7, 7, 6, 6, 6, 6, 6, 6, 6,
    6, 6, 6, 6, 6, 6, 6, 6, 6, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 2};
template <class G>
OB_STAGE int ob_transient_analysis(const G &g, const float *in, int len, int C, float *tf_estimate, int *tf_chan, float *tmp, float *env)
{
    const float forward_decay = .0625f;
    int mask_metric = 0;
    const int len2 = len / 2;
    for (int c = 0; c < C; c++) {
        const float *xin = in + c * len;
        ObState2 z; z.s0 = 0.f; z.s1 = 0.f;
        ob_scan2(g, len, 1.f, .5f, -1.f, 0.f, z, [&](int i, float &mem0, float &mem1, bool emit) {
            const float x = xin[i];
            const float y = mem0 + x;
            const float mem00 = mem0;
            mem0 = mem0 - x + .5f * mem1;
            mem1 = x - mem00;
            if (emit) tmp[i] = i < 12 ? 0.f : y;
        });
        g.sync();
        float mean = ob_psum(g, len2, 0.f, [&](int i) { return tmp[2 * i] * tmp[2 * i] + tmp[2 * i + 1] * tmp[2 * i + 1]; });
        ob_scan1(g, len2, 1.f - forward_decay, 0.f, false, [&](int i) { return tmp[2 * i] * tmp[2 * i] + tmp[2 * i + 1] * tmp[2 * i + 1]; },
                 [&](int i, float m) { env[i] = forward_decay * m; });
        g.sync();
        ob_scan1(g, len2, 0.875f, 0.f, true, [&](int i) { return env[i]; }, [&](int i, float m) { env[i] = 0.125f * m; });
        g.sync();
        const float maxE = ob_pmax(g, len2, 0.f, [&](int i) { return env[i]; });
        mean = (float)sqrt((double)(mean * maxE) * .5 * (double)len2);          // celt_sqrt(mean * maxE*.5*len2): ".5" makes it double
        const float norm = (float)len2 / (1e-15f + mean);                         // SHL32/SHR32 are identities in the float build
        int unmask = (int)ob_psum_u32(g, (len2 - 5 - 12 + 3) / 4, [&](int q) {
            const int i = 12 + 4 * q;
            double v = floor((double)(64 * norm * (env[i] + 1e-15f)));
            if (v > 127) v = 127;
            if (v < 0) v = 0;
            return (uint32_t)OB_INV_TABLE[(int)v];
        });
        unmask = 64 * unmask * 4 / (6 * (len2 - 17));
        if (unmask > mask_metric) { *tf_chan = c; mask_metric = unmask; }
        g.sync();
    }
    const int is_transient = mask_metric > 200;
    const float tf_max = ob_fmax(0, (float)sqrt((double)(27 * mask_metric)) - 42);
    {   // celt_sqrt(MAX32(0, MULT16_16(0.0069, MIN16(163, tf_max)) - 0.139)): the subtraction and max are in double
        double v = (double)((float)0.0069 * ob_fmin(163, tf_max)) - 0.139;
        if (!(v > 0)) v = 0;
        *tf_estimate = (float)sqrt(v);
    }
    return is_transient;
}

// ---- forward MDCT (mdct.c:119-238) of nblk blocks of one channel at once.  in: the channel's [overlap | nblk*N2] samples; out: coefficient k
// of block b at out[b + nblk*k]; f2: nblk*N2 floats of scratch (N4 complex per block).  Window + fold feed the pre-rotation directly. --------------
template <class G>
OB_STAGE void ob_mdct_forward(const G &g, const float *__restrict__ in, float *__restrict__ out, int shift, int nblk, float *__restrict__ f2)
{
    const int N2 = 1920 >> (shift + 1), N4 = N2 >> 1, overlap = OB_OVERLAP, q = (overlap + 3) >> 2;
    const float *trig = OB_MDCT_TRIG + (shift == 0 ? 0 : shift == 1 ? 960 : shift == 2 ? 1440 : 1680);
    const float scale = shift == 0 ? 0.002083333f : shift == 1 ? 0.004166667f : shift == 2 ? 0.008333333f : 0.016666667f;   // kiss_fft_state.scale
    const int16_t *br = ob_fft_bitrev(shift);
    const ObDiv dN4 = ob_div_make(N4);
    for (int t = g.lane; t < nblk * N4; t += g.n) {
        const int b = nblk > 1 ? ob_div(t, dN4) : 0, i = t - b * N4;
        const float *xp1 = in + b * N2 + (overlap >> 1) + 2 * i, *xp2 = in + b * N2 + N2 - 1 + (overlap >> 1) - 2 * i;
        float re, im;
        if (i < q) {
            const float w1 = OB_WINDOW[(overlap >> 1) + 2 * i], w2 = OB_WINDOW[(overlap >> 1) - 1 - 2 * i];
            re = w2 * xp1[N2] + w1 * *xp2;
            im = w1 * *xp1 - w2 * xp2[-N2];
        } else if (i < N4 - q) { re = *xp2; im = *xp1; }
        else {
            const int r = i - (N4 - q);
            const float w1 = OB_WINDOW[2 * r], w2 = OB_WINDOW[overlap - 1 - 2 * r];
            re = -(w1 * xp1[-N2]) + w2 * *xp2;
            im = w2 * *xp1 + w1 * xp2[N2];
        }
        const float t0 = trig[i], t1 = trig[N4 + i];
        const float yr = re * t0 - im * t1, yi = im * t0 + re * t1;
        float *dst = f2 + b * N2 + 2 * br[i];
        dst[0] = scale * yr;
        dst[1] = scale * yi;
    }
    g.sync();
    {
        const int16_t *fac = ob_fft_factors(shift);
        int fstride[9], L = 0, m, m2, p;
        fstride[0] = 1;
        do { p = fac[2 * L]; m = fac[2 * L + 1]; fstride[L + 1] = fstride[L] * p; L++; } while (m != 1);
        m = fac[2 * L - 1];
        for (int i = L - 1; i >= 0; i--) {
            m2 = i != 0 ? fac[2 * i - 1] : 1;
            ob_fft_stage(g, f2, nblk, N2, fac[2 * i], fstride[i] << shift, m, fstride[i], m2);
            m = m2;
        }
    }
    for (int t = g.lane; t < nblk * N4; t += g.n) {
        const int b = nblk > 1 ? ob_div(t, dN4) : 0, i = t - b * N4;
        const float *fp = f2 + b * N2 + 2 * i;
        const float yr = fp[1] * trig[N4 + i] - fp[0] * trig[i];
        const float yi = fp[0] * trig[N4 + i] + fp[1] * trig[i];
        out[b + nblk * (2 * i)] = yr;
        out[b + nblk * (N2 - 1 - 2 * i)] = yi;
    }
    g.sync();
}

// compute_mdcts (celt_encoder.c:461-504), upsample == 1
template <class G>
OB_DEV void ob_compute_mdcts(const G &g, int shortBlocks, const float *in, float *out, int C, int CC, int LM, float *f2)
{
    int B, N, shift;
    if (shortBlocks) { B = shortBlocks; N = OB_SHORT; shift = 3; } else { B = 1; N = OB_SHORT << LM; shift = 3 - LM; }
    for (int c = 0; c < CC; c++) ob_mdct_forward(g, in + c * (B * N + OB_OVERLAP), out + c * N * B, shift, B, f2);
    if (CC == 2 && C == 1) { for (int i = g.lane; i < B * N; i += g.n) out[i] = .5f * out[i] + .5f * out[B * N + i]; g.sync(); }
}

// compute_band_energies (bands.c:159-174, float), amp2Log2 (quant_bands.c:544-563), normalise_bands (bands.c:177-191)
template <class G>
OB_STAGE void ob_band_energies(const G &g, const float *X, float *bandE, int end, int C, int LM)
{
    const int N = OB_SHORT << LM;
    for (int c = 0; c < C; c++) for (int i = 0; i < end; i++) {
        const float *x = &X[c * N + (OB_EBANDS[i] << LM)];
        const int n = (OB_EBANDS[i + 1] - OB_EBANDS[i]) << LM;
        const float sum = 1e-27f + ob_psum(g, n, 0.f, [&](int j) { return x[j] * x[j]; });
        bandE[i + c * OB_NB] = sqrtf(sum);
    }
    g.sync();
}
template <class G>
OB_STAGE void ob_amp2log2(const G &g, int effEnd, int end, const float *bandE, float *bandLogE, int C)
{
    for (int k = g.lane; k < C * OB_NB; k += g.n) {
        const int i = k % OB_NB;
        if (i < effEnd) bandLogE[k] = ob_log2(bandE[k]) - OB_EMEANS[i];
        else if (i < end) bandLogE[k] = -14.f;
    }
    g.sync();
}
template <class G>
OB_STAGE void ob_normalise_bands(const G &g, const float *__restrict__ freq, float *__restrict__ X, const float *__restrict__ bandE, int end, int C, int M)
{
    const int N = M * OB_SHORT;
    for (int c = 0; c < C; c++) for (int i = 0; i < end; i++) {
        const float gn = 1.f / (1e-27f + bandE[i + c * OB_NB]);
        for (int j = M * OB_EBANDS[i] + g.lane; j < M * OB_EBANDS[i + 1]; j += g.n) X[j + c * N] = freq[j + c * N] * gn;
    }
    g.sync();
}

// patch_transient_decision (celt_encoder.c:423-457), start = 0
OB_DEV int ob_patch_transient(const float *newE, const float *oldE, int end, int C, float *spread_old)
{
    float mean_diff = 0;
    if (C == 1) {
        spread_old[0] = oldE[0];
        for (int i = 1; i < end; i++) spread_old[i] = ob_fmax(spread_old[i - 1] - 1.0f, oldE[i]);
    } else {
        spread_old[0] = ob_fmax(oldE[0], oldE[OB_NB]);
        for (int i = 1; i < end; i++) spread_old[i] = ob_fmax(spread_old[i - 1] - 1.0f, ob_fmax(oldE[i], oldE[i + OB_NB]));
    }
    for (int i = end - 2; i >= 0; i--) spread_old[i] = ob_fmax(spread_old[i], spread_old[i + 1] - 1.0f);
    for (int c = 0; c < C; c++) for (int i = 2; i < end - 1; i++) {
        const float x1 = ob_fmax(0, newE[i + c * OB_NB]), x2 = ob_fmax(0, spread_old[i]);
        mean_diff = mean_diff + ob_fmax(0, x1 - x2);
    }
    mean_diff = mean_diff / (float)(C * (end - 1 - 2));
    return mean_diff > 1.f;
}

OB_DEV float ob_median_of_5(const float *x)                          // celt_encoder.c:921-959
{
    float t0, t1, t2 = x[2], t3, t4;
    if (x[0] > x[1]) { t0 = x[1]; t1 = x[0]; } else { t0 = x[0]; t1 = x[1]; }
    if (x[3] > x[4]) { t3 = x[4]; t4 = x[3]; } else { t3 = x[3]; t4 = x[4]; }
    if (t0 > t3) { float t = t0; t0 = t3; t3 = t; t = t1; t1 = t4; t4 = t; }
    if (t2 > t1) { if (t1 < t3) return ob_fmin(t2, t3); else return ob_fmin(t4, t1); }
    else { if (t2 < t3) return ob_fmin(t1, t3); else return ob_fmin(t2, t4); }
}
OB_DEV float ob_median_of_3(const float *x)                          // celt_encoder.c:961-979
{
    float t0, t1, t2;
    if (x[0] > x[1]) { t0 = x[1]; t1 = x[0]; } else { t0 = x[0]; t1 = x[1]; }
    t2 = x[2];
    if (t1 < t2) return t1; else if (t0 < t2) return t2; else return t0;
}

// dynalloc_analysis (celt_encoder.c:981-1185), start = 0, lfe = 0, no surround mask.  21-element recurrences: warp-uniform.
// scr: >= 6 * 21 floats of scratch.
OB_STAGE float ob_dynalloc_analysis(const float *bandLogE, const float *bandLogE2, const float *oldBandE, int end, int C, int *offsets,
        int lsb_depth, int isTransient, int vbr, int constrained_vbr, int LM, int effectiveBytes, int32_t *tot_boost_,
        int *importance, int *spread_weight, const uint8_t *leak_boost, float *scr)
{
    int32_t tot_boost = 0;
    float maxDepth = -31.9f;
    float *follower = scr, *noise_floor = scr + 2 * OB_NB, *bandLogE3 = scr + 3 * OB_NB, *mask = scr + 4 * OB_NB, *sig = scr + 5 * OB_NB;
    for (int i = 0; i < OB_NB; i++) offsets[i] = 0;
    for (int i = 0; i < end; i++)
        noise_floor[i] = 0.0625f * (float)OB_LOGN[i] + .5f + (float)(9 - lsb_depth) - OB_EMEANS[i] + .0062f * (float)((i + 5) * (i + 5));
    for (int c = 0; c < C; c++) for (int i = 0; i < end; i++) maxDepth = ob_fmax(maxDepth, bandLogE[c * OB_NB + i] - noise_floor[i]);
    {
        for (int i = 0; i < end; i++) mask[i] = bandLogE[i] - noise_floor[i];
        if (C == 2) for (int i = 0; i < end; i++) mask[i] = ob_fmax(mask[i], bandLogE[OB_NB + i] - noise_floor[i]);
        for (int i = 0; i < end; i++) sig[i] = mask[i];
        for (int i = 1; i < end; i++) mask[i] = ob_fmax(mask[i], mask[i - 1] - 2.f);
        for (int i = end - 2; i >= 0; i--) mask[i] = ob_fmax(mask[i], mask[i + 1] - 3.f);
        for (int i = 0; i < end; i++) {
            const float smr = sig[i] - ob_fmax(ob_fmax(0, maxDepth - 12.f), mask[i]);
            const int shift = ob_imin(5, ob_imax(0, -(int)floor((double)(.5f + smr))));
            spread_weight[i] = 32 >> shift;
        }
    }
    if (effectiveBytes >= (30 + 5 * LM)) {
        int last = 0;
        for (int c = 0; c < C; c++) {
            float *f = &follower[c * OB_NB];
            for (int i = 0; i < end; i++) bandLogE3[i] = bandLogE2[c * OB_NB + i];
            if (LM == 0) for (int i = 0; i < ob_imin(8, end); i++) bandLogE3[i] = ob_fmax(bandLogE2[c * OB_NB + i], oldBandE[c * OB_NB + i]);
            f[0] = bandLogE3[0];
            for (int i = 1; i < end; i++) {
                if (bandLogE3[i] > bandLogE3[i - 1] + .5f) last = i;
                f[i] = ob_fmin(f[i - 1] + 1.5f, bandLogE3[i]);
            }
            for (int i = last - 1; i >= 0; i--) f[i] = ob_fmin(f[i], ob_fmin(f[i + 1] + 2.f, bandLogE3[i]));
            const float offset = 1.f;
            for (int i = 2; i < end - 2; i++) f[i] = ob_fmax(f[i], ob_median_of_5(&bandLogE3[i - 2]) - offset);
            float t = ob_median_of_3(&bandLogE3[0]) - offset;
            f[0] = ob_fmax(f[0], t); f[1] = ob_fmax(f[1], t);
            t = ob_median_of_3(&bandLogE3[end - 3]) - offset;
            f[end - 2] = ob_fmax(f[end - 2], t); f[end - 1] = ob_fmax(f[end - 1], t);
            for (int i = 0; i < end; i++) f[i] = ob_fmax(f[i], noise_floor[i]);
        }
        if (C == 2) {
            for (int i = 0; i < end; i++) {
                follower[OB_NB + i] = ob_fmax(follower[OB_NB + i], follower[i] - 4.f);
                follower[i] = ob_fmax(follower[i], follower[OB_NB + i] - 4.f);
                follower[i] = .5f * (ob_fmax(0, bandLogE[i] - follower[i]) + ob_fmax(0, bandLogE[OB_NB + i] - follower[OB_NB + i]));
            }
        } else for (int i = 0; i < end; i++) follower[i] = ob_fmax(0, bandLogE[i] - follower[i]);
        // surround_dynalloc is all zero on this path: follower = MAX16(follower, 0) is a no-op (follower >= 0)
        for (int i = 0; i < end; i++) importance[i] = (int)floor((double)(.5f + 13 * ob_exp2(ob_fmin(follower[i], 4.f))));
        if ((!vbr || constrained_vbr) && !isTransient) for (int i = 0; i < end; i++) follower[i] = .5f * follower[i];
        for (int i = 0; i < end; i++) {
            if (i < 8) follower[i] *= 2;
            if (i >= 12) follower[i] = .5f * follower[i];
        }
        if (leak_boost) for (int i = 0; i < ob_imin(19, end); i++) follower[i] = follower[i] + (1.f / 64.f) * leak_boost[i];   // celt_encoder.c:1139-1143
        for (int i = 0; i < end; i++) {
            int boost, boost_bits;
            follower[i] = ob_fmin(follower[i], 4.f);
            const int width = C * (OB_EBANDS[i + 1] - OB_EBANDS[i]) << LM;
            if (width < 6) { boost = (int)follower[i]; boost_bits = boost * width << OB_BITRES; }
            else if (width > 48) { boost = (int)(follower[i] * 8); boost_bits = (boost * width << OB_BITRES) / 8; }
            else { boost = (int)(follower[i] * width / 6); boost_bits = boost * 6 << OB_BITRES; }
            if ((!vbr || (constrained_vbr && !isTransient)) && (tot_boost + boost_bits) >> OB_BITRES >> 3 > 2 * effectiveBytes / 3) {
                const int32_t cap = ((2 * effectiveBytes / 3) << OB_BITRES << 3);
                offsets[i] = cap - tot_boost;
                tot_boost = cap;
                break;
            } else { offsets[i] = boost; tot_boost += boost_bits; }
        }
    } else for (int i = 0; i < end; i++) importance[i] = 13;
    *tot_boost_ = tot_boost;
    return maxDepth;
}

// tf_analysis (celt_encoder.c:595-754): per band, the L1 norm after each Haar level (reductions over the band), then the 21-step Viterbi
// search (warp-uniform).  tmp, tmp_1: OB_MAX_BAND floats each; iscr: >= 3 * 21 ints.
template <class G>
OB_DEV float ob_l1_metric(const G &g, const float *tmp, int N, int LM, float bias)      // celt_encoder.c:582-593
{
    float L1 = ob_psum(g, N, 0.f, [&](int i) { return fabsf(tmp[i]); });
    L1 = L1 + (LM * bias) * L1;
    return L1;
}
template <class G>
OB_STAGE int ob_tf_analysis(const G &g, int len, int isTransient, int *tf_res, int lambda, const float *X, int N0, int LM, float tf_estimate, int tf_chan,
        const int *importance, float *tmp, float *tmp_1, int *iscr)
{
    int *metric = iscr, *path0 = iscr + OB_NB, *path1 = iscr + 2 * OB_NB;
    int cost0, cost1, selcost[2], tf_select = 0;
    const float bias = .04f * ob_fmax(-.25f, .5f - tf_estimate);
    for (int i = 0; i < len; i++) {
        const int N = (OB_EBANDS[i + 1] - OB_EBANDS[i]) << LM, narrow = (OB_EBANDS[i + 1] - OB_EBANDS[i]) == 1;
        float L1, best_L1;
        int best_level = 0;
        const float *src = X + tf_chan * N0 + (OB_EBANDS[i] << LM);
        for (int j = g.lane; j < N; j += g.n) { const float v = src[j]; tmp[j] = v; tmp_1[j] = v; }
        g.sync();
        L1 = ob_l1_metric(g, tmp, N, isTransient ? LM : 0, bias);
        best_L1 = L1;
        if (isTransient && !narrow) {
            ob_haar1(g, tmp_1, N >> LM, 1 << LM);
            L1 = ob_l1_metric(g, tmp_1, N, LM + 1, bias);
            if (L1 < best_L1) { best_L1 = L1; best_level = -1; }
        }
        for (int k = 0; k < LM + !(isTransient || narrow); k++) {
            const int B = isTransient ? (LM - k - 1) : k + 1;
            ob_haar1(g, tmp, N >> k, 1 << k);
            L1 = ob_l1_metric(g, tmp, N, B, bias);
            if (L1 < best_L1) { best_L1 = L1; best_level = k + 1; }
        }
        int mt = isTransient ? 2 * best_level : -2 * best_level;
        if (narrow && (mt == 0 || mt == -2 * LM)) mt -= 1;
        metric[i] = mt;
        g.sync();
    }
#define OB_TFS(sel, k) (2 * OB_TF_SELECT[LM * 8 + 4 * isTransient + 2 * (sel) + (k)])
#define OB_IABS(v) ((v) < 0 ? -(v) : (v))
    for (int sel = 0; sel < 2; sel++) {
        cost0 = importance[0] * OB_IABS(metric[0] - OB_TFS(sel, 0));
        cost1 = importance[0] * OB_IABS(metric[0] - OB_TFS(sel, 1)) + (isTransient ? 0 : lambda);
        for (int i = 1; i < len; i++) {
            const int curr0 = ob_imin(cost0, cost1 + lambda), curr1 = ob_imin(cost0 + lambda, cost1);
            cost0 = curr0 + importance[i] * OB_IABS(metric[i] - OB_TFS(sel, 0));
            cost1 = curr1 + importance[i] * OB_IABS(metric[i] - OB_TFS(sel, 1));
        }
        cost0 = ob_imin(cost0, cost1);
        selcost[sel] = cost0;
    }
    if (selcost[1] < selcost[0] && isTransient) tf_select = 1;
    cost0 = importance[0] * OB_IABS(metric[0] - OB_TFS(tf_select, 0));
    cost1 = importance[0] * OB_IABS(metric[0] - OB_TFS(tf_select, 1)) + (isTransient ? 0 : lambda);
    for (int i = 1; i < len; i++) {
        int curr0, curr1, from0 = cost0, from1 = cost1 + lambda;
        if (from0 < from1) { curr0 = from0; path0[i] = 0; } else { curr0 = from1; path0[i] = 1; }
        from0 = cost0 + lambda; from1 = cost1;
        if (from0 < from1) { curr1 = from0; path1[i] = 0; } else { curr1 = from1; path1[i] = 1; }
        cost0 = curr0 + importance[i] * OB_IABS(metric[i] - OB_TFS(tf_select, 0));
        cost1 = curr1 + importance[i] * OB_IABS(metric[i] - OB_TFS(tf_select, 1));
    }
    tf_res[len - 1] = cost0 < cost1 ? 0 : 1;
    for (int i = len - 2; i >= 0; i--) tf_res[i] = tf_res[i + 1] == 1 ? path1[i + 1] : path0[i + 1];
#undef OB_TFS
    g.sync();
    return tf_select;
}

// spreading_decision (bands.c:479-570): the three threshold counts of a band travel through one integer reduction (10 bits each)
template <class G>
OB_STAGE int ob_spreading_decision(const G &g, const float *X, int *average, int last_decision, int *hf_average, int *tapset_decision, int update_hf,
        int end, int C, int M, const int *spread_weight)
{
    int sum = 0, nbBands = 0, hf_sum = 0, decision;
    const int N0 = M * OB_SHORT;
    if (M * (OB_EBANDS[end] - OB_EBANDS[end - 1]) <= 8) return 0;
    for (int c = 0; c < C; c++) for (int i = 0; i < end; i++) {
        const float *x = X + M * OB_EBANDS[i] + c * N0;
        const int N = M * (OB_EBANDS[i + 1] - OB_EBANDS[i]);
        if (N <= 8) continue;
        const uint32_t packed = ob_psum_u32(g, N, [&](int j) {
            const float x2N = (x[j] * x[j]) * (float)N;
            return (uint32_t)(x2N < 0.25f) | (uint32_t)(x2N < 0.0625f) << 10 | (uint32_t)(x2N < 0.015625f) << 20;
        });
        const int t0 = (int)(packed & 1023u), t1 = (int)((packed >> 10) & 1023u), t2 = (int)(packed >> 20);
        if (i > OB_NB - 4) hf_sum += (int)((uint32_t)(32 * (t1 + t0)) / (uint32_t)N);
        const int tmp = (2 * t2 >= N) + (2 * t1 >= N) + (2 * t0 >= N);
        sum += tmp * spread_weight[i];
        nbBands += spread_weight[i];
    }
    if (update_hf) {
        if (hf_sum) hf_sum = (int)((uint32_t)hf_sum / (uint32_t)(C * (4 - OB_NB + end)));
        *hf_average = (*hf_average + hf_sum) >> 1;
        hf_sum = *hf_average;
        if (*tapset_decision == 2) hf_sum += 4; else if (*tapset_decision == 0) hf_sum -= 4;
        if (hf_sum > 22) *tapset_decision = 2; else if (hf_sum > 18) *tapset_decision = 1; else *tapset_decision = 0;
    }
    sum = (int)((uint32_t)(sum << 8) / (uint32_t)nbBands);
    sum = (sum + *average) >> 1;
    *average = sum;
    sum = (3 * sum + (((3 - last_decision) << 7) + 64) + 2) >> 2;
    if (sum < 80) decision = 3; else if (sum < 256) decision = 2; else if (sum < 384) decision = 1; else decision = 0;
    return decision;
}

// stereo_analysis (celt_encoder.c:889-919): bands 0..12 are contiguous bins
template <class G>
OB_DEV int ob_stereo_analysis(const G &g, const float *X, int LM, int N0)
{
    float sumLR = 1e-15f, sumMS = 1e-15f;
    ob_psum2(g, OB_EBANDS[13] << LM, sumLR, sumMS, [&](int j, float &a, float &b) {
        const float L = X[j], R = X[N0 + j], M = L + R, S = L - R;
        a = a + (fabsf(L) + fabsf(R));
        b = b + (fabsf(M) + fabsf(S));
    });
    sumMS = 0.707107f * sumMS;
    int thetas = 13;
    if (LM <= 1) thetas -= 8;
    return (float)((OB_EBANDS[13] << (LM + 1)) + thetas) * sumMS > (float)(OB_EBANDS[13] << (LM + 1)) * sumLR;
}

// hysteresis_decision (bands.c:46-59)
OB_DEV int ob_hysteresis_decision(float val, const float *thresholds, const float *hysteresis, int N, int prev)
{
    int i;
    for (i = 0; i < N; i++) if (val < thresholds[i]) break;
    if (i > prev && val < thresholds[prev] + hysteresis[prev]) i = prev;
    if (i < prev && val > thresholds[prev - 1] - hysteresis[prev - 1]) i = prev;
    return i;
}

// alloc_trim_analysis (celt_encoder.c:797-887), surround_trim = 0
template <class G>
OB_STAGE int ob_alloc_trim_analysis(const G &g, const float *X, const float *bandLogE, int end, int LM, int C, int N0, float *stereo_saving, float tf_estimate,
        int intensity, int32_t equiv_rate, int an_valid, float an_tonality_slope)
{
    float diff = 0, trim = 5.f;
    if (equiv_rate < 64000) trim = 4.f;
    else if (equiv_rate < 80000) { const int32_t frac = (equiv_rate - 64000) >> 10; trim = 4.f + (1.f / 16.f) * frac; }
    if (C == 2) {
        float sum = 0, minXC, logXC, logXC2;
        for (int i = 0; i < 8; i++) {
            const float *a = &X[OB_EBANDS[i] << LM], *b = &X[N0 + (OB_EBANDS[i] << LM)];
            sum = sum + ob_psum(g, (OB_EBANDS[i + 1] - OB_EBANDS[i]) << LM, 0.f, [&](int j) { return a[j] * b[j]; });
        }
        sum = (1.f / 8) * sum;
        sum = ob_fmin(1.f, fabsf(sum));
        minXC = sum;
        for (int i = 8; i < intensity; i++) {
            const float *a = &X[OB_EBANDS[i] << LM], *b = &X[N0 + (OB_EBANDS[i] << LM)];
            const float partial = ob_psum(g, (OB_EBANDS[i + 1] - OB_EBANDS[i]) << LM, 0.f, [&](int j) { return a[j] * b[j]; });
            minXC = ob_fmin(minXC, fabsf(partial));
        }
        minXC = ob_fmin(1.f, fabsf(minXC));
        logXC = ob_log2(1.001f - sum * sum);
        logXC2 = ob_fmax(.5f * logXC, ob_log2(1.001f - minXC * minXC));
        trim += ob_fmax(-4.f, .75f * logXC);
        *stereo_saving = ob_fmin(*stereo_saving + 0.25f, -(.5f * logXC2));
    }
    for (int c = 0; c < C; c++) for (int i = 0; i < end - 1; i++) diff += bandLogE[i + c * OB_NB] * (int32_t)(2 + 2 * i - end);
    diff /= C * (end - 1);
    trim -= ob_fmax(-2.f, ob_fmin(2.f, (diff + 1.f) / 6));
    trim -= 2 * tf_estimate;
    if (an_valid) trim -= ob_fmax(-2.f, ob_fmin(2.f, 2.f * (an_tonality_slope + .05f)));
    int trim_index = (int)floor((double)(.5f + trim));
    trim_index = ob_imax(0, ob_imin(10, trim_index));
    return trim_index;
}

// compute_vbr (celt_encoder.c:1320-1429), no surround mask, lfe = 0
OB_DEV int32_t ob_compute_vbr(int32_t base_target, int LM, int32_t bitrate, int lastCodedBands, int C, int intensity, int constrained_vbr,
        float stereo_saving, int tot_boost, float tf_estimate, float maxDepth, float temporal_vbr, int an_valid, float an_activity, float an_tonality,
        int pitch_change)
{
    int32_t target = base_target;
    const int coded_bands = lastCodedBands ? lastCodedBands : OB_NB;
    int coded_bins = OB_EBANDS[coded_bands] << LM;
    if (C == 2) coded_bins += OB_EBANDS[ob_imin(intensity, coded_bands)] << LM;
    if (an_valid && (double)an_activity < .4) target -= (int32_t)((float)(coded_bins << OB_BITRES) * (.4f - an_activity));
    if (C == 2) {
        const int coded_stereo_bands = ob_imin(intensity, coded_bands);
        const int coded_stereo_dof = (OB_EBANDS[coded_stereo_bands] << LM) - coded_stereo_bands;
        const float max_frac = (0.8f * (float)coded_stereo_dof) / (float)coded_bins;
        stereo_saving = ob_fmin(stereo_saving, 1.f);
        target -= (int32_t)ob_fmin(max_frac * (float)target, (stereo_saving - 0.1f) * (float)(coded_stereo_dof << OB_BITRES));
    }
    target += tot_boost - (19 << LM);
    const float tf_calibration = 0.044f;
    target += (int32_t)((tf_estimate - tf_calibration) * (float)target);
    if (an_valid) {                                                  // tonality boost (celt_encoder.c:1373-1386)
        const float tonal = ob_fmax(0.f, an_tonality - .15f) - 0.12f;
        int32_t tonal_target = target + (int32_t)((float)(coded_bins << OB_BITRES) * 1.2f * tonal);
        if (pitch_change) tonal_target += (int32_t)((float)(coded_bins << OB_BITRES) * .8f);
        target = tonal_target;
    }
    {
        const int bins = OB_EBANDS[OB_NB - 2] << LM;
        int32_t floor_depth = (int32_t)((float)(C * bins << OB_BITRES) * maxDepth);
        floor_depth = ob_imax(floor_depth, target >> 2);
        target = ob_imin(target, floor_depth);
    }
    if (constrained_vbr) target = base_target + (int32_t)(0.67f * (float)(target - base_target));
    if (tf_estimate < .2f) {
        const float amount = .0000031f * (float)ob_imax(0, ob_imin(32000, 96000 - bitrate));
        const float tvbr_factor = temporal_vbr * amount;
        target += (int32_t)(tvbr_factor * (float)target);
    }
    target = ob_imin(2 * base_target, target);
    return target;
}
