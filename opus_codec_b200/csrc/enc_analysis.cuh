// enc_analysis.cuh -- the float analysis front-end of the CELT encoder, per stream:
// pre-emphasis, pitch pre-filter (down-sampling, LPC whitening, cross-correlation search, sub-harmonic check, comb
// filter), transient detection, forward MDCT, band energies / normalisation, and the decision heuristics that turn
// floats into the integers the bitstream carries (dynalloc, tf, spread, stereo, trim, VBR target).
// Every function restates the float build of the reference in the same operation order (opus_val16/32 = float,
// shifts are identities: opus/celt/arch.h:225-290) so that decisions match the reference's pure-C paths.
#pragma once
#include "enc_range.cuh"
#include "ob_group.cuh"
#include "dec_synth.cuh"     // ob_fft_stage, tables

#ifndef __CUDACC__
#include <math.h>
#endif

#define OB_MAXPERIOD 1024           // COMBFILTER_MAXPERIOD (celt.h:218)
#define OB_MINPERIOD 15

OB_DEV float ob_fmax(float a, float b) { return a > b ? a : b; }     // MAX16/MAX32 (arch.h): (a) > (b) ? (a) : (b)
OB_DEV float ob_fmin(float a, float b) { return a < b ? a : b; }
OB_DEV float ob_log2(float x) { return (float)(1.442695040888963387 * log((double)x)); }     // celt_log2 (mathops.h:168)
OB_DEV float ob_exp2(float x) { return (float)exp(0.6931471805599453094 * (double)x); }      // celt_exp2 (mathops.h:169)

// celt_preemphasis fast path (celt_encoder.c:507-531): inp = 32768*x - m ; m = coef0 * 32768*x
OB_DEV void ob_preemphasis(const float *__restrict__ pcm, float *__restrict__ inp, int N, int CC, float *mem, int clip)
{
    const float coef0 = OB_PREEMPH[0];
    float m = *mem;
    if (!clip) {
#pragma unroll 8
        for (int i = 0; i < N; i++) { const float x = pcm[CC * i] * 32768.f; inp[i] = x - m; m = coef0 * x; }
    } else {
        for (int i = 0; i < N; i++) {
            float x = pcm[CC * i] * 32768.f;
            x = ob_fmax(-65536.f, ob_fmin(65536.f, x));
            inp[i] = x - m; m = coef0 * x;
        }
    }
    *mem = m;
}

// comb_filter (celt.c:190-256), general form y != x, window may be absent (overlap = 0).
OB_DEV void ob_comb_filter_xy(float *y, const float *x, int T0, int T1, int N, float g0, float g1, int tapset0, int tapset1, int overlap)
{
    const float gains[3][3] = {{0.3066406250f, 0.2170410156f, 0.1296386719f}, {0.4638671875f, 0.2680664062f, 0.f}, {0.7998046875f, 0.1000976562f, 0.f}};
    if (g0 == 0 && g1 == 0) { if (x != y) for (int i = 0; i < N; i++) y[i] = x[i]; return; }
    T0 = ob_imax(T0, OB_MINPERIOD); T1 = ob_imax(T1, OB_MINPERIOD);
    const float g00 = g0 * gains[tapset0][0], g01 = g0 * gains[tapset0][1], g02 = g0 * gains[tapset0][2];
    const float g10 = g1 * gains[tapset1][0], g11 = g1 * gains[tapset1][1], g12 = g1 * gains[tapset1][2];
    float x1 = x[-T1 + 1], x2 = x[-T1], x3 = x[-T1 - 1], x4 = x[-T1 - 2], x0;
    if (g0 == g1 && T0 == T1 && tapset0 == tapset1) overlap = 0;
    int i;
    for (i = 0; i < overlap; i++) {
        const float f = OB_WINDOW[i] * OB_WINDOW[i];
        x0 = x[i - T1 + 2];
        y[i] = x[i] + ((1.0f - f) * g00) * x[i - T0] + ((1.0f - f) * g01) * (x[i - T0 + 1] + x[i - T0 - 1])
                    + ((1.0f - f) * g02) * (x[i - T0 + 2] + x[i - T0 - 2])
                    + (f * g10) * x2 + (f * g11) * (x1 + x3) + (f * g12) * (x0 + x4);
        x4 = x3; x3 = x2; x2 = x1; x1 = x0;
    }
    if (g1 == 0) { if (x != y) for (; i < N; i++) y[i] = x[i]; return; }
    x4 = x[i - T1 - 2]; x3 = x[i - T1 - 1]; x2 = x[i - T1]; x1 = x[i - T1 + 1];
    for (; i < N; i++) {                                           // comb_filter_const_c (celt.c:162-185)
        x0 = x[i - T1 + 2];
        y[i] = x[i] + g10 * x2 + g11 * (x1 + x3) + g12 * (x0 + x4);
        x4 = x3; x3 = x2; x2 = x1; x1 = x0;
    }
}

OB_DEV float ob_maxabs(const float *x, int len)                      // celt_maxabs16 (mathops.h:79-91)
{
    float maxval = 0, minval = 0;
    for (int i = 0; i < len; i++) { maxval = ob_fmax(maxval, x[i]); minval = ob_fmin(minval, x[i]); }
    return ob_fmax(maxval, -minval);
}

// ---- pitch analysis (opus/celt/pitch.c) ----------------------------------------------------------------------------
OB_DEV float ob_inner_prod(const float *x, const float *y, int N) { float s = 0; for (int i = 0; i < N; i++) s = s + x[i] * y[i]; return s; }

// celt_pitch_xcorr_c (pitch.c:225-300): every lag is a plain in-order sum (xcorr_kernel_c accumulates lag by lag in j order)
OB_DEV void ob_pitch_xcorr(const float *x, const float *y, float *xcorr, int len, int max_pitch)
{
    int i = 0;
    for (; i + 4 <= max_pitch; i += 4) {                   // four lags share every x[j] and a sliding window of y: 2 loads per 4 MACs instead of 8
        float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        float y0 = y[i], y1 = y[i + 1], y2 = y[i + 2];
        for (int j = 0; j < len; j++) {
            const float xj = x[j], y3 = y[i + j + 3];
            s0 = s0 + xj * y0; s1 = s1 + xj * y1; s2 = s2 + xj * y2; s3 = s3 + xj * y3;
            y0 = y1; y1 = y2; y2 = y3;
        }
        xcorr[i] = s0; xcorr[i + 1] = s1; xcorr[i + 2] = s2; xcorr[i + 3] = s3;
    }
    for (; i < max_pitch; i++) xcorr[i] = ob_inner_prod(x, y + i, len);
}

// pitch_downsample (pitch.c:140-217) incl. _celt_autocorr (celt_lpc.c:277-351, lag 4, no window), _celt_lpc (:37-91), celt_fir5 (pitch.c:105-137)
OB_DEV void ob_pitch_downsample(const float *x0, const float *x1, float *x_lp, int len, int C)
{
    const int n = len >> 1;
    for (int i = 1; i < n; i++) x_lp[i] = .25f * x0[2 * i - 1] + .25f * x0[2 * i + 1] + .5f * x0[2 * i];
    x_lp[0] = .25f * x0[1] + .5f * x0[0];
    if (C == 2) {
        for (int i = 1; i < n; i++) x_lp[i] += .25f * x1[2 * i - 1] + .25f * x1[2 * i + 1] + .5f * x1[2 * i];
        x_lp[0] += .25f * x1[1] + .5f * x1[0];
    }
    float ac[5];
    {
        const int lag = 4, fastN = n - lag;
        ob_pitch_xcorr(x_lp, x_lp, ac, fastN, lag + 1);
        for (int k = 0; k <= lag; k++) {
            float d = 0;
            for (int i = k + fastN; i < n; i++) d = d + x_lp[i] * x_lp[i - k];
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
    float num[5];
    num[0] = lpc[0] + .8f;
    num[1] = lpc[1] + c1 * lpc[0];
    num[2] = lpc[2] + c1 * lpc[1];
    num[3] = lpc[3] + c1 * lpc[2];
    num[4] = c1 * lpc[3];
    float m0 = 0, m1 = 0, m2 = 0, m3 = 0, m4 = 0;
    for (int i = 0; i < n; i++) {
        float sum = x_lp[i];
        sum = sum + num[0] * m0; sum = sum + num[1] * m1; sum = sum + num[2] * m2; sum = sum + num[3] * m3; sum = sum + num[4] * m4;
        m4 = m3; m3 = m2; m2 = m1; m1 = m0; m0 = x_lp[i];
        x_lp[i] = sum;
    }
}

OB_DEV void ob_find_best_pitch(const float *xcorr, const float *y, int len, int max_pitch, int *best_pitch)   // pitch.c:45-103
{
    float Syy = 1, best_num[2] = {-1, -1}, best_den[2] = {0, 0};
    best_pitch[0] = 0; best_pitch[1] = 1;
    for (int j = 0; j < len; j++) Syy = Syy + y[j] * y[j];
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
OB_DEV void ob_pitch_search(const float *x_lp, const float *y, int len, int max_pitch, int *pitch, float *scratch)
{
    const int lag = len + max_pitch;
    float *x_lp4 = scratch, *y_lp4 = x_lp4 + (len >> 2), *xcorr = y_lp4 + (lag >> 2);
    int best_pitch[2] = {0, 0}, offset;
    for (int j = 0; j < len >> 2; j++) x_lp4[j] = x_lp[2 * j];
    for (int j = 0; j < lag >> 2; j++) y_lp4[j] = y[2 * j];
    ob_pitch_xcorr(x_lp4, y_lp4, xcorr, len >> 2, max_pitch >> 2);
    ob_find_best_pitch(xcorr, y_lp4, len >> 2, max_pitch >> 2, best_pitch);
    for (int i = 0; i < max_pitch >> 1; i++) {
        xcorr[i] = 0;
        int d0 = i - 2 * best_pitch[0], d1 = i - 2 * best_pitch[1];
        if (d0 < 0) d0 = -d0;
        if (d1 < 0) d1 = -d1;
        if (d0 > 2 && d1 > 2) continue;
        const float sum = ob_inner_prod(x_lp, y + i, len >> 1);
        xcorr[i] = ob_fmax(-1, sum);
    }
    ob_find_best_pitch(xcorr, y, len >> 1, max_pitch >> 1, best_pitch);
    if (best_pitch[0] > 0 && best_pitch[0] < (max_pitch >> 1) - 1) {
        const float a = xcorr[best_pitch[0] - 1], b = xcorr[best_pitch[0]], c = xcorr[best_pitch[0] + 1];
        if ((c - a) > .7f * (b - a)) offset = 1;
        else if ((a - c) > .7f * (b - c)) offset = -1;
        else offset = 0;
    } else offset = 0;
    *pitch = 2 * best_pitch[0] - offset;
}

OB_DEV float ob_pitch_gain(float xy, float xx, float yy) { return xy / sqrtf(1 + xx * yy); }       // pitch.c:441-444

// remove_doubling (pitch.c:449-555).  yy_lookup: >= maxperiod/2 + 1 floats.
OB_DEV float ob_remove_doubling(const float *x, int maxperiod, int minperiod, int N, int *T0_, int prev_period, float prev_gain, float *yy_lookup)
{
    const int second_check[16] = {0, 0, 3, 2, 3, 2, 5, 2, 3, 2, 3, 2, 5, 2, 3, 2};
    int k, i, T, T0, offset;
    float g, g0, pg, xy, xx, yy, xy2, xcorr[3], best_xy, best_yy;
    const int minperiod0 = minperiod;
    maxperiod /= 2; minperiod /= 2; *T0_ /= 2; prev_period /= 2; N /= 2;
    x += maxperiod;
    if (*T0_ >= maxperiod) *T0_ = maxperiod - 1;
    T = T0 = *T0_;
    xx = 0; xy = 0;
    for (i = 0; i < N; i++) { xx = xx + x[i] * x[i]; xy = xy + x[i] * x[i - T0]; }      // dual_inner_prod(x, x, x-T0)
    yy_lookup[0] = xx;
    yy = xx;
    for (i = 1; i <= maxperiod; i++) {
        yy = yy + x[-i] * x[-i] - x[N - i] * x[N - i];
        yy_lookup[i] = ob_fmax(0, yy);
    }
    yy = yy_lookup[T0];
    best_xy = xy; best_yy = yy;
    g = g0 = ob_pitch_gain(xy, xx, yy);
    for (k = 2; k <= 15; k++) {
        int T1, T1b;
        float g1, cont, thresh;
        T1 = (int)((uint32_t)(2 * T0 + k) / (uint32_t)(2 * k));
        if (T1 < minperiod) break;
        if (k == 2) { if (T1 + T0 > maxperiod) T1b = T0; else T1b = T0 + T1; }
        else T1b = (int)((uint32_t)(2 * second_check[k] * T0 + k) / (uint32_t)(2 * k));
        xy = 0; xy2 = 0;
        for (i = 0; i < N; i++) { xy = xy + x[i] * x[i - T1]; xy2 = xy2 + x[i] * x[i - T1b]; }
        xy = .5f * (xy + xy2);
        yy = .5f * (yy_lookup[T1] + yy_lookup[T1b]);
        g1 = ob_pitch_gain(xy, xx, yy);
        int dT = T1 - prev_period; if (dT < 0) dT = -dT;
        if (dT <= 1) cont = prev_gain;
        else if (dT <= 2 && 5 * k * k < T0) cont = .5f * prev_gain;
        else cont = 0;
        thresh = ob_fmax(.3f, .7f * g0 - cont);
        if (T1 < 3 * minperiod) thresh = ob_fmax(.4f, .85f * g0 - cont);
        else if (T1 < 2 * minperiod) thresh = ob_fmax(.5f, .9f * g0 - cont);
        if (g1 > thresh) { best_xy = xy; best_yy = yy; T = T1; g = g1; }
    }
    best_xy = ob_fmax(0, best_xy);
    if (best_yy <= best_xy) pg = 1.0f;
    else pg = best_xy / (best_yy + 1);
    for (k = 0; k < 3; k++) xcorr[k] = ob_inner_prod(x, x - (T + k - 1), N);
    if ((xcorr[2] - xcorr[0]) > .7f * (xcorr[1] - xcorr[0])) offset = 1;
    else if ((xcorr[0] - xcorr[2]) > .7f * (xcorr[1] - xcorr[2])) offset = -1;
    else offset = 0;
    if (pg > g) pg = g;
    *T0_ = 2 * T + offset;
    if (*T0_ < minperiod0) *T0_ = minperiod0;
    return pg;
}

// transient_analysis (celt_encoder.c:227-419), allow_weak_transients = 0.  tmp: >= len floats.
OB_DEV int ob_transient_analysis(const float *in, int len, int C, float *tf_estimate, int *tf_chan, float *tmp)
{
    // inv_table (celt_encoder.c:246-255)
    const uint8_t inv_table[128] = {
        255, 255, 156, 110, 86, 70, 59, 51, 45, 40, 37, 33, 31, 28, 26, 25, 23, 22, 21, 20, 19, 18, 17, 16, 16, 15, 15, 14, 13, 13, 12, 12,
        12, 12, 11, 11, 11, 10, 10, 10, 9, 9, 9, 9, 9, 9, 8, 8, 8, 8, 8, 7, 7, 7, 7, 7, 7, 6, 6, 6, 6, 6, 6, 6,
        6, 6, 6, 6, 6, 6, 6, 6, 6, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
        4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 2};
    const float forward_decay = .0625f;
    int mask_metric = 0;
    const int len2 = len / 2;
    for (int c = 0; c < C; c++) {
        float mean, mem0 = 0, mem1 = 0, maxE, norm;
        int unmask = 0;
#pragma unroll 8
        for (int i = 0; i < len; i++) {
            const float x = in[i + c * len];
            const float y = mem0 + x;
            const float mem00 = mem0;
            mem0 = mem0 - x + .5f * mem1;
            mem1 = x - mem00;
            tmp[i] = y;
        }
        for (int i = 0; i < 12; i++) tmp[i] = 0;
        mean = 0; mem0 = 0;
        for (int i = 0; i < len2; i++) {
            const float x2 = tmp[2 * i] * tmp[2 * i] + tmp[2 * i + 1] * tmp[2 * i + 1];
            mean += x2;
            mem0 = x2 + (1.f - forward_decay) * mem0;
            tmp[i] = forward_decay * mem0;
        }
        mem0 = 0; maxE = 0;
        for (int i = len2 - 1; i >= 0; i--) {
            mem0 = tmp[i] + 0.875f * mem0;
            tmp[i] = 0.125f * mem0;
            maxE = ob_fmax(maxE, 0.125f * mem0);
        }
        mean = (float)sqrt((double)(mean * maxE) * .5 * (double)len2);          // celt_sqrt(mean * maxE*.5*len2): ".5" makes it double
        norm = (float)len2 / (1e-15f + mean);                                    // SHL32/SHR32 are identities in the float build
        for (int i = 12; i < len2 - 5; i += 4) {
            double v = floor((double)(64 * norm * (tmp[i] + 1e-15f)));
            if (v > 127) v = 127;
            if (v < 0) v = 0;
            unmask += inv_table[(int)v];
        }
        unmask = 64 * unmask * 4 / (6 * (len2 - 17));
        if (unmask > mask_metric) { *tf_chan = c; mask_metric = unmask; }
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

// ---- forward MDCT (mdct.c:119-238); f: N2 floats scratch, f2: N4 complex scratch --------------------------------------
OB_DEV void ob_mdct_forward(const float *__restrict__ in, float *__restrict__ out, int shift, int stride, float *__restrict__ f, float *__restrict__ f2)
{
    int N = 1920;
    const float *trig = OB_MDCT_TRIG;
    for (int i = 0; i < shift; i++) { N >>= 1; trig += N; }
    const int N2 = N >> 1, N4 = N >> 2, overlap = OB_OVERLAP;
    const float scale = shift == 0 ? 0.002083333f : shift == 1 ? 0.004166667f : shift == 2 ? 0.008333333f : 0.016666667f;   // kiss_fft_state.scale
    {
        const float *__restrict__ xp1 = in + (overlap >> 1), *__restrict__ xp2 = in + N2 - 1 + (overlap >> 1);
        float *__restrict__ yp = f;
        const float *wp1 = OB_WINDOW + (overlap >> 1), *wp2 = OB_WINDOW + (overlap >> 1) - 1;
        int i;
        for (i = 0; i < ((overlap + 3) >> 2); i++) {
            *yp++ = *wp2 * xp1[N2] + *wp1 * *xp2;
            *yp++ = *wp1 * *xp1 - *wp2 * xp2[-N2];
            xp1 += 2; xp2 -= 2; wp1 += 2; wp2 -= 2;
        }
        wp1 = OB_WINDOW; wp2 = OB_WINDOW + overlap - 1;
#pragma unroll 8
        for (; i < N4 - ((overlap + 3) >> 2); i++) { *yp++ = *xp2; *yp++ = *xp1; xp1 += 2; xp2 -= 2; }
        for (; i < N4; i++) {
            *yp++ = -(*wp1 * xp1[-N2]) + *wp2 * *xp2;
            *yp++ = *wp2 * *xp1 + *wp1 * xp2[N2];
            xp1 += 2; xp2 -= 2; wp1 += 2; wp2 -= 2;
        }
    }
    {
        const int16_t *br = ob_fft_bitrev(shift);
        const float *__restrict__ yp = f;
#pragma unroll 4
        for (int i = 0; i < N4; i++) {
            const float t0 = trig[i], t1 = trig[N4 + i], re = *yp++, im = *yp++;
            const float yr = re * t0 - im * t1, yi = im * t0 + re * t1;
            f2[2 * br[i]] = scale * yr;
            f2[2 * br[i] + 1] = scale * yi;
        }
    }
    {
        ObSolo g;
        const int16_t *fac = ob_fft_factors(shift);
        int fstride[9], L = 0, m, m2, p;
        fstride[0] = 1;
        do { p = fac[2 * L]; m = fac[2 * L + 1]; fstride[L + 1] = fstride[L] * p; L++; } while (m != 1);
        m = fac[2 * L - 1];
        for (int i = L - 1; i >= 0; i--) {
            m2 = i != 0 ? fac[2 * i - 1] : 1;
            ob_fft_stage(g, f2, 1, 0, fac[2 * i], fstride[i] << shift, m, fstride[i], m2);
            m = m2;
        }
    }
    {
        const float *__restrict__ fp = f2;
        float *__restrict__ yp1 = out, *__restrict__ yp2 = out + stride * (N2 - 1);
#pragma unroll 4
        for (int i = 0; i < N4; i++) {
            const float yr = fp[1] * trig[N4 + i] - fp[0] * trig[i];
            const float yi = fp[0] * trig[N4 + i] + fp[1] * trig[i];
            *yp1 = yr; *yp2 = yi;
            fp += 2; yp1 += 2 * stride; yp2 -= 2 * stride;
        }
    }
}

// compute_mdcts (celt_encoder.c:461-504), upsample == 1
OB_DEV void ob_compute_mdcts(int shortBlocks, const float *in, float *out, int C, int CC, int LM, float *f, float *f2)
{
    int B, N, shift;
    if (shortBlocks) { B = shortBlocks; N = OB_SHORT; shift = 3; } else { B = 1; N = OB_SHORT << LM; shift = 3 - LM; }
    for (int c = 0; c < CC; c++)
        for (int b = 0; b < B; b++)
            ob_mdct_forward(in + c * (B * N + OB_OVERLAP) + b * N, &out[b + c * N * B], shift, B, f, f2);
    if (CC == 2 && C == 1) for (int i = 0; i < B * N; i++) out[i] = .5f * out[i] + .5f * out[B * N + i];
}

// compute_band_energies (bands.c:159-174, float), amp2Log2 (quant_bands.c:544-563), normalise_bands (bands.c:177-191)
OB_DEV void ob_band_energies(const float *X, float *bandE, int end, int C, int LM)
{
    const int N = OB_SHORT << LM;
    for (int c = 0; c < C; c++) for (int i = 0; i < end; i++) {
        const float *x = &X[c * N + (OB_EBANDS[i] << LM)];
        const int n = (OB_EBANDS[i + 1] - OB_EBANDS[i]) << LM;
        const float sum = 1e-27f + ob_inner_prod(x, x, n);
        bandE[i + c * OB_NB] = sqrtf(sum);
    }
}
OB_DEV void ob_amp2log2(int effEnd, int end, const float *bandE, float *bandLogE, int C)
{
    for (int c = 0; c < C; c++) {
        for (int i = 0; i < effEnd; i++) bandLogE[i + c * OB_NB] = ob_log2(bandE[i + c * OB_NB]) - OB_EMEANS[i];
        for (int i = effEnd; i < end; i++) bandLogE[c * OB_NB + i] = -14.f;
    }
}
OB_DEV void ob_normalise_bands(const float *__restrict__ freq, float *__restrict__ X, const float *__restrict__ bandE, int end, int C, int M)
{
    const int N = M * OB_SHORT;
    for (int c = 0; c < C; c++) for (int i = 0; i < end; i++) {
        const float g = 1.f / (1e-27f + bandE[i + c * OB_NB]);
#pragma unroll 8
        for (int j = M * OB_EBANDS[i]; j < M * OB_EBANDS[i + 1]; j++) X[j + c * N] = freq[j + c * N] * g;
    }
}

// patch_transient_decision (celt_encoder.c:423-457), start = 0
OB_DEV int ob_patch_transient(const float *newE, const float *oldE, int end, int C)
{
    float mean_diff = 0, spread_old[26];
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

// dynalloc_analysis (celt_encoder.c:981-1185), start = 0, lfe = 0, no surround mask, analysis invalid
OB_DEV float ob_dynalloc_analysis(const float *bandLogE, const float *bandLogE2, const float *oldBandE, int end, int C, int *offsets,
        int lsb_depth, int isTransient, int vbr, int constrained_vbr, int LM, int effectiveBytes, int32_t *tot_boost_,
        int *importance, int *spread_weight, const uint8_t *leak_boost)
{
    int32_t tot_boost = 0;
    float maxDepth = -31.9f, follower[2 * OB_NB], noise_floor[OB_NB], bandLogE3[OB_NB];
    for (int i = 0; i < OB_NB; i++) offsets[i] = 0;
    for (int i = 0; i < end; i++)
        noise_floor[i] = 0.0625f * (float)OB_LOGN[i] + .5f + (float)(9 - lsb_depth) - OB_EMEANS[i] + .0062f * (float)((i + 5) * (i + 5));
    for (int c = 0; c < C; c++) for (int i = 0; i < end; i++) maxDepth = ob_fmax(maxDepth, bandLogE[c * OB_NB + i] - noise_floor[i]);
    {
        float mask[OB_NB], sig[OB_NB];
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

// haar1 on a private copy (bands.c:632-645), scalar
OB_DEV void ob_haar1_s(float *X, int N0, int stride)
{
    N0 >>= 1;
    for (int i = 0; i < stride; i++) for (int j = 0; j < N0; j++) {
        const float t1 = .70710678f * X[stride * 2 * j + i], t2 = .70710678f * X[stride * (2 * j + 1) + i];
        X[stride * 2 * j + i] = t1 + t2;
        X[stride * (2 * j + 1) + i] = t1 - t2;
    }
}
OB_DEV float ob_l1_metric(const float *tmp, int N, int LM, float bias)      // celt_encoder.c:582-593
{
    float L1 = 0;
    for (int i = 0; i < N; i++) L1 += fabsf(tmp[i]);
    L1 = L1 + (LM * bias) * L1;
    return L1;
}

// tf_analysis (celt_encoder.c:595-754)
OB_DEV int ob_tf_analysis(int len, int isTransient, int *tf_res, int lambda, const float *X, int N0, int LM, float tf_estimate, int tf_chan,
        const int *importance)
{
    int metric[OB_NB], path0[OB_NB], path1[OB_NB], cost0, cost1, selcost[2], tf_select = 0;
    float tmp[OB_MAX_BAND], tmp_1[OB_MAX_BAND];
    const float bias = .04f * ob_fmax(-.25f, .5f - tf_estimate);
    for (int i = 0; i < len; i++) {
        const int N = (OB_EBANDS[i + 1] - OB_EBANDS[i]) << LM, narrow = (OB_EBANDS[i + 1] - OB_EBANDS[i]) == 1;
        float L1, best_L1;
        int best_level = 0;
        for (int j = 0; j < N; j++) tmp[j] = X[tf_chan * N0 + (OB_EBANDS[i] << LM) + j];
        L1 = ob_l1_metric(tmp, N, isTransient ? LM : 0, bias);
        best_L1 = L1;
        if (isTransient && !narrow) {
            for (int j = 0; j < N; j++) tmp_1[j] = tmp[j];
            ob_haar1_s(tmp_1, N >> LM, 1 << LM);
            L1 = ob_l1_metric(tmp_1, N, LM + 1, bias);
            if (L1 < best_L1) { best_L1 = L1; best_level = -1; }
        }
        for (int k = 0; k < LM + !(isTransient || narrow); k++) {
            const int B = isTransient ? (LM - k - 1) : k + 1;
            ob_haar1_s(tmp, N >> k, 1 << k);
            L1 = ob_l1_metric(tmp, N, B, bias);
            if (L1 < best_L1) { best_L1 = L1; best_level = k + 1; }
        }
        metric[i] = isTransient ? 2 * best_level : -2 * best_level;
        if (narrow && (metric[i] == 0 || metric[i] == -2 * LM)) metric[i] -= 1;
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
    return tf_select;
}

// spreading_decision (bands.c:479-570)
OB_DEV int ob_spreading_decision(const float *X, int *average, int last_decision, int *hf_average, int *tapset_decision, int update_hf,
        int end, int C, int M, const int *spread_weight)
{
    int sum = 0, nbBands = 0, hf_sum = 0, decision;
    const int N0 = M * OB_SHORT;
    if (M * (OB_EBANDS[end] - OB_EBANDS[end - 1]) <= 8) return 0;
    for (int c = 0; c < C; c++) for (int i = 0; i < end; i++) {
        int tcount[3] = {0, 0, 0};
        const float *x = X + M * OB_EBANDS[i] + c * N0;
        const int N = M * (OB_EBANDS[i + 1] - OB_EBANDS[i]);
        if (N <= 8) continue;
        for (int j = 0; j < N; j++) {
            const float x2N = (x[j] * x[j]) * (float)N;
            if (x2N < 0.25f) tcount[0]++;
            if (x2N < 0.0625f) tcount[1]++;
            if (x2N < 0.015625f) tcount[2]++;
        }
        if (i > OB_NB - 4) hf_sum += (int)((uint32_t)(32 * (tcount[1] + tcount[0])) / (uint32_t)N);
        const int tmp = (2 * tcount[2] >= N) + (2 * tcount[1] >= N) + (2 * tcount[0] >= N);
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

// stereo_analysis (celt_encoder.c:889-919)
OB_DEV int ob_stereo_analysis(const float *X, int LM, int N0)
{
    float sumLR = 1e-15f, sumMS = 1e-15f;
    for (int i = 0; i < 13; i++) for (int j = OB_EBANDS[i] << LM; j < OB_EBANDS[i + 1] << LM; j++) {
        const float L = X[j], R = X[N0 + j], M = L + R, S = L - R;
        sumLR = sumLR + (fabsf(L) + fabsf(R));
        sumMS = sumMS + (fabsf(M) + fabsf(S));
    }
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
OB_DEV int ob_alloc_trim_analysis(const float *X, const float *bandLogE, int end, int LM, int C, int N0, float *stereo_saving, float tf_estimate,
        int intensity, int32_t equiv_rate, int an_valid, float an_tonality_slope)
{
    float diff = 0, trim = 5.f;
    if (equiv_rate < 64000) trim = 4.f;
    else if (equiv_rate < 80000) { const int32_t frac = (equiv_rate - 64000) >> 10; trim = 4.f + (1.f / 16.f) * frac; }
    if (C == 2) {
        float sum = 0, minXC, logXC, logXC2;
        for (int i = 0; i < 8; i++)
            sum = sum + ob_inner_prod(&X[OB_EBANDS[i] << LM], &X[N0 + (OB_EBANDS[i] << LM)], (OB_EBANDS[i + 1] - OB_EBANDS[i]) << LM);
        sum = (1.f / 8) * sum;
        sum = ob_fmin(1.f, fabsf(sum));
        minXC = sum;
        for (int i = 8; i < intensity; i++) {
            const float partial = ob_inner_prod(&X[OB_EBANDS[i] << LM], &X[N0 + (OB_EBANDS[i] << LM)], (OB_EBANDS[i + 1] - OB_EBANDS[i]) << LM);
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
