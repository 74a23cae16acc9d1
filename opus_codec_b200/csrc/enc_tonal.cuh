// enc_tonal.cuh -- the Opus-layer signal analysis the reference runs at complexity >= 7 (opus/src/analysis.c, opus/src/mlp.c):
// 24 kHz down-mix through a half-band all-pass pair, 480-point FFT of overlapping 20 ms windows, per-bin phase-stability
// tonality, 18-band energies / stationarity / leakage, bandwidth detection, a 25-32-24-2 dense/GRU/dense network for the
// speech-music and activity probabilities, and the 100-entry look-ahead ring tonality_get_info reads from.  Per stream, one thread,
// float arithmetic in the reference's order and width (its double-precision promotions are kept: they decide comparisons).
#pragma once
#include "enc_quant.cuh"      // ob_fast_atan2f
#include "dec_synth.cuh"      // ob_fft_stage, bit-reversal and factor tables of the 480-point FFT

#include "analysis_tables.inc"

#define OB_AN_NB_FRAMES 8
#define OB_AN_NB_TBANDS 18
#define OB_AN_BUF 720                       // ANALYSIS_BUF_SIZE: 30 ms at 24 kHz
#define OB_AN_COUNT_MAX 10000
#define OB_AN_DETECT 100                    // DETECT_SIZE
#define OB_AN_LEAK_BANDS 19
#define OB_AN_SKIP_BANDS 9                  // NB_TONAL_SKIP_BANDS

struct ObAnalysisInfo {                     // AnalysisInfo (celt.h:59-73)
    int32_t valid;
    float tonality, tonality_slope, noisiness, activity, music_prob, music_prob_min, music_prob_max;
    int32_t bandwidth;
    float activity_probability, max_pitch_ratio;
    uint8_t leak_boost[OB_AN_LEAK_BANDS];
    uint8_t pad;
};

struct ObTonalState {                       // TonalityAnalysisState (analysis.h:47-79), Fs = 48000
    float angle[240], d_angle[240], d2_angle[240];
    float inmem[OB_AN_BUF];
    int32_t mem_fill;
    float prev_band_tonality[OB_AN_NB_TBANDS];
    float prev_tonality;
    int32_t prev_bandwidth;
    float E[OB_AN_NB_FRAMES][OB_AN_NB_TBANDS], logE[OB_AN_NB_FRAMES][OB_AN_NB_TBANDS];
    float lowE[OB_AN_NB_TBANDS], highE[OB_AN_NB_TBANDS], meanE[OB_AN_NB_TBANDS + 1];
    float mem[32], cmean[8], std[9];
    float Etracker, lowECount;
    int32_t E_count, count, analysis_offset, write_pos, read_pos, read_subframe;
    float hp_ener_accum;
    int32_t initialized;
    float rnn_state[32];
    float downmix_state[3];
    ObAnalysisInfo info[OB_AN_DETECT];
};

OB_DEV void ob_tonal_reset(ObTonalState &t)                        // tonality_analysis_reset (analysis.c:226-231)
{
    uint32_t *z = reinterpret_cast<uint32_t *>(&t);
    for (int i = 0; i < (int)(sizeof(ObTonalState) / 4); i++) z[i] = 0;
}

// ---- the small network (mlp.c:36-132); weights are int8 scaled by 1/128 ----
OB_DEV float ob_tansig(float x)
{
    const float N0 = 952.52801514f, N1 = 96.39235687f, N2 = 0.60863042f, D0 = 952.72399902f, D1 = 413.36801147f, D2 = 11.88600922f;
    const float X2 = x * x;
    float num = (N2 * X2 + N1) * X2 + N0;
    const float den = (D2 * X2 + D1) * X2 + D0;
    num = num * x / den;
    return ob_fmax(-1.f, ob_fmin(1.f, num));
}
OB_DEV float ob_sigmoid(float x) { return .5f + .5f * ob_tansig(.5f * x); }
// The network of analysis.c (mlp.c: analysis_compute_dense / analysis_compute_gru).  Output neuron i is one lane's work: its sum runs over the
// inputs in the reference's order, so the result does not depend on how many lanes share the layer.  Callers sync before (inputs complete).
template <class G>
OB_DEV void ob_gemm_accum(const G &g, float *out, const int8_t *w, int rows, int cols, int col_stride, const float *x)
{
    for (int i = g.lane; i < rows; i += g.n) {
        float a = out[i];
        for (int j = 0; j < cols; j++) a += w[j * col_stride + i] * x[j];
        out[i] = a;
    }
}
template <class G>
OB_DEV void ob_dense(const G &g, const int8_t *bias, const int8_t *w, int M, int N, int sigmoid, float *output, const float *input)
{
    for (int i = g.lane; i < N; i += g.n) output[i] = bias[i];
    ob_gemm_accum(g, output, w, N, M, N, input);
    for (int i = g.lane; i < N; i += g.n) { const float v = output[i] * (1.f / 128); output[i] = sigmoid ? ob_sigmoid(v) : ob_tansig(v); }
    g.sync();
}
template <class G>
OB_DEV void ob_gru(const G &g, float *state, const float *input, float *scr)               // layer1: 32 inputs, 24 neurons; scr: 4 * 32 floats
{
    const int M = 32, N = 24, stride = 3 * N;
    float *tmp = scr, *z = scr + 32, *r = scr + 64, *h = scr + 96;
    for (int i = g.lane; i < N; i += g.n) { z[i] = OB_AN_L1_B[i]; r[i] = OB_AN_L1_B[N + i]; h[i] = OB_AN_L1_B[2 * N + i]; }
    ob_gemm_accum(g, z, OB_AN_L1_W, N, M, stride, input);
    ob_gemm_accum(g, z, OB_AN_L1_R, N, N, stride, state);
    ob_gemm_accum(g, r, OB_AN_L1_W + N, N, M, stride, input);
    ob_gemm_accum(g, r, OB_AN_L1_R + N, N, N, stride, state);
    for (int i = g.lane; i < N; i += g.n) {
        z[i] = ob_sigmoid((1.f / 128) * z[i]);
        r[i] = ob_sigmoid((1.f / 128) * r[i]);
        tmp[i] = state[i] * r[i];
    }
    g.sync();                                                        // every tmp[j] is read by every neuron below
    ob_gemm_accum(g, h, OB_AN_L1_W + 2 * N, N, M, stride, input);
    ob_gemm_accum(g, h, OB_AN_L1_R + 2 * N, N, N, stride, tmp);
    for (int i = g.lane; i < N; i += g.n) state[i] = z[i] * state[i] + (1 - z[i]) * ob_tansig((1.f / 128) * h[i]);       // nobody reads state[j != i] any more
    g.sync();
}

// silk_resampler_down2_hp (analysis.c:115-161), float build; returns the high-pass energy.  The three all-pass memories are independent
// first-order recurrences (S0 over the even samples, S1 and S2 over the odd ones): three exact-step scans write their contributions to
// five temporaries, one element-wise pass combines them in the reference's order.  tmp: >= 5 * 240 floats; runs in chunks of 240 outputs.
template <class G>
OB_DEV float ob_down2_hp(const G &g, float *S, float *out, const float *in, int inLen, float *tmp)
{
    const int len2 = inLen / 2;
    float hp_ener = 0, S0 = S[0], S1 = S[1], S2 = S[2];
    if (ObOrder<G>::value == 1) {                                    // one lane, reference order: the reference's fused loop (same arithmetic, no temporaries)
        for (int k = 0; k < len2; k++) {
            float in32 = in[2 * k];
            float Y = in32 - S0;
            float X = 0.6074371f * Y;
            float out32 = S0 + X;
            S0 = in32 + X;
            float out32_hp = out32;
            in32 = in[2 * k + 1];
            Y = in32 - S1;
            X = 0.15063f * Y;
            out32 = out32 + S1;
            out32 = out32 + X;
            S1 = in32 + X;
            Y = -in32 - S2;
            X = 0.15063f * Y;
            out32_hp = out32_hp + S2;
            out32_hp = out32_hp + X;
            S2 = -in32 + X;
            hp_ener += out32_hp * out32_hp;
            out[k] = .5f * out32;
        }
        S[0] = S0; S[1] = S1; S[2] = S2;
        return hp_ener;
    }
    for (int base = 0; base < len2; base += 240) {
        const int n = ob_imin(240, len2 - base);
        const float *x = in + 2 * base;
        float *a0 = tmp, *s1 = tmp + 240, *x1 = tmp + 480, *s2 = tmp + 720, *x2 = tmp + 960;
        S0 = ob_scan1s(g, n, -0.6074371f, S0, [&](int k, float &m, bool emit) {
            const float in32 = x[2 * k], Y = in32 - m, X = 0.6074371f * Y;
            if (emit) a0[k] = m + X;
            m = in32 + X;
        });
        S1 = ob_scan1s(g, n, -0.15063f, S1, [&](int k, float &m, bool emit) {
            const float in32 = x[2 * k + 1], Y = in32 - m, X = 0.15063f * Y;
            if (emit) { s1[k] = m; x1[k] = X; }
            m = in32 + X;
        });
        S2 = ob_scan1s(g, n, -0.15063f, S2, [&](int k, float &m, bool emit) {
            const float in32 = x[2 * k + 1], Y = -in32 - m, X = 0.15063f * Y;
            if (emit) { s2[k] = m; x2[k] = X; }
            m = -in32 + X;
        });
        g.sync();
        for (int k = g.lane; k < n; k += g.n) {
            float out32 = a0[k];
            out32 = out32 + s1[k];
            out32 = out32 + x1[k];
            out[base + k] = .5f * out32;
        }
        hp_ener = ob_psum(g, n, hp_ener, [&](int k) {
            float hp = a0[k];
            hp = hp + s2[k];
            hp = hp + x2[k];
            return hp * hp;
        });
        g.sync();
    }
    if (g.lane == 0) { S[0] = S0; S[1] = S1; S[2] = S2; }
    g.sync();
    return hp_ener;
}

// downmix_and_resample (analysis.c:163-215) with downmix_float (opus_encoder.c:657-678), c1 = 0, c2 = -2, Fs = 48000.
// x: the caller's interleaved float PCM; dm: >= 960 floats of scratch; tmp: >= 1200 floats.
template <class G>
OB_DEV float ob_downmix_and_resample(const G &g, const float *x, float *y, float *S, int subframe, int offset, int C, float *dm, float *tmp)
{
    if (subframe == 0) return 0;
    subframe *= 2;
    offset *= 2;
    float scale = 1.f / 32768;
    scale /= C;
    for (int j = g.lane; j < subframe; j += g.n) {
        float v = x[(j + offset) * C] * 32768.f;
        for (int c = 1; c < C; c++) v += x[(j + offset) * C + c] * 32768.f;
        dm[j] = v * scale;
    }
    g.sync();
    return ob_down2_hp(g, S, y, dm, subframe, tmp);
}

// opus_fft (kiss_fft.c:569-589) of 480 complex points: scale, bit-reverse, then the decoder's radix stages (every butterfly a work item).
template <class G>
OB_DEV void ob_fft480(const G &g, const float *fin, float *fout)
{
    const float scale = 0.002083333f;
    const int16_t *br = ob_fft_bitrev(0);
    for (int i = g.lane; i < 480; i += g.n) { fout[2 * br[i]] = scale * fin[2 * i]; fout[2 * br[i] + 1] = scale * fin[2 * i + 1]; }
    g.sync();
    const int16_t *fac = ob_fft_factors(0);
    int fstride[9], L = 0, m, m2, p;
    fstride[0] = 1;
    do { p = fac[2 * L]; m = fac[2 * L + 1]; fstride[L + 1] = fstride[L] * p; L++; } while (m != 1);
    m = fac[2 * L - 1];
    for (int i = L - 1; i >= 0; i--) {
        m2 = i != 0 ? fac[2 * i - 1] : 1;
        ob_fft_stage(g, fout, 1, 0, fac[2 * i], fstride[i], m, fstride[i], m2);
        m = m2;
    }
}

// tonality_analysis (analysis.c:446-953).  x: interleaved float PCM of the frame being encoded; len/offset in 48 kHz samples.
// Cooperative part (all lanes): down-mix + 2:1 resampling (three recurrence scans), windowing, the 480-point FFT, the per-bin phase statistics
// and the per-band energy sums (one band per lane, each the reference's in-order sum).  The per-band statistics, the bandwidth detector and the
// 25-32-24-2 network that follow are ~7 k instructions of scalar code on 18-element arrays: lane 0 runs them on the sums the warp left behind.
// work: >= 2 720 floats (FFT input 960 | down-mix 960 | 3 x 240 per-bin values | 80 band sums); fbuf: >= 1 200 floats (resampler temporaries, then
// the FFT's 960-float output) -- shared memory in the warp-per-stream kernel.
template <class G>
OB_DEV_NOINLINE void ob_tonality_analysis(const G &g, ObTonalState &tonal, const float *x, int len, int offset, int C, int lsb_depth, float *work, float *fbuf)
{
    const int N = 480, N2 = 240, NB = OB_AN_NB_TBANDS;
    float *A = tonal.angle, *dA = tonal.d_angle, *d2A = tonal.d2_angle;
    float *in = work, *out = fbuf, *dm = work + 960, *tonality = work + 1920, *noisiness = work + 2160, *tonality2 = work + 2400, *bsum = work + 2640;
    float band_tonality[18], logE[18], BFCC[8], features[25], midE[8], band_log2[19], leakage_from[19], leakage_to[19];
    int is_masked[19];
    const float pi4 = (float)(3.14159265358979323846 * 3.14159265358979323846 * 3.14159265358979323846 * 3.14159265358979323846);
    float slope = 0, frame_tonality, max_frame_tonality, frame_noisiness, frame_stationarity, relativeE, frame_loudness, bandwidth_mask, maxE, noise_floor;
    float spec_variability = 0, below_max_pitch, above_max_pitch, hp_ener;
    int bandwidth = 0, b, i;

    // the scalar state is read by every lane (uniform) and written back by lane 0 alone
    int mem_fill = tonal.mem_fill, write_pos = tonal.write_pos;
    const int count = tonal.count;
    float hp_accum = tonal.hp_ener_accum;
    g.sync();
    if (!tonal.initialized) mem_fill = 240;
    const float alpha = 1.f / ob_imin(10, 1 + count), alphaE = 1.f / ob_imin(25, 1 + count);
    float alphaE2 = 1.f / ob_imin(100, 1 + count);
    if (count <= 1) alphaE2 = 1;
    len /= 2; offset /= 2;                                           // now at 24 kHz
    hp_accum += ob_downmix_and_resample(g, x, &tonal.inmem[mem_fill], tonal.downmix_state, ob_imin(len, OB_AN_BUF - mem_fill), offset, C, dm, fbuf);
    if (mem_fill + len < OB_AN_BUF) {
        if (g.lane == 0) { tonal.initialized = 1; tonal.mem_fill = mem_fill + len; tonal.hp_ener_accum = hp_accum; }
        g.sync();
        return;
    }
    hp_ener = hp_accum;
    ObAnalysisInfo *info = &tonal.info[write_pos++];
    if (write_pos >= OB_AN_DETECT) write_pos -= OB_AN_DETECT;
    const int is_silence = ob_maxabs(g, tonal.inmem, OB_AN_BUF) <= (float)1 / (1 << lsb_depth);
    for (i = g.lane; i < N2; i += g.n) {
        const float w = OB_AN_WINDOW[i];
        in[2 * i] = w * tonal.inmem[i];
        in[2 * i + 1] = w * tonal.inmem[N2 + i];
        in[2 * (N - i - 1)] = w * tonal.inmem[N - i - 1];
        in[2 * (N - i - 1) + 1] = w * tonal.inmem[N + N2 - i - 1];
    }
    g.sync();
    for (i = g.lane; i < 240; i += g.n) tonal.inmem[i] = tonal.inmem[OB_AN_BUF - 240 + i];
    g.sync();
    const int remaining = len - (OB_AN_BUF - mem_fill);
    hp_accum = ob_downmix_and_resample(g, x, &tonal.inmem[240], tonal.downmix_state, remaining, offset + OB_AN_BUF - mem_fill, C, dm, fbuf);
    if (g.lane == 0) { tonal.initialized = 1; tonal.mem_fill = 240 + remaining; tonal.hp_ener_accum = hp_accum; tonal.write_pos = write_pos; }
    g.sync();
    if (is_silence) {                                                // copy the previous analysis
        int prev_pos = write_pos - 2;
        if (prev_pos < 0) prev_pos += OB_AN_DETECT;
        if (g.lane == 0) *info = tonal.info[prev_pos];
        g.sync();
        return;
    }
    ob_fft480(g, in, out);
    if (out[0] != out[0]) { if (g.lane == 0) info->valid = 0; g.sync(); return; }
#define RE(k) out[2 * (k)]
#define IM(k) out[2 * (k) + 1]
    for (i = 1 + g.lane; i < N2; i += g.n) {
        const float X1r = RE(i) + RE(N - i), X1i = IM(i) - IM(N - i), X2r = IM(i) + IM(N - i), X2i = RE(N - i) - RE(i);
        const float angle = (float)(.5f / 3.14159265358979323846) * ob_fast_atan2f(X1i, X1r);
        const float d_angle = angle - A[i], d2_angle = d_angle - dA[i];
        const float angle2 = (float)(.5f / 3.14159265358979323846) * ob_fast_atan2f(X2i, X2r);
        const float d_angle2 = angle2 - angle, d2_angle2 = d_angle2 - d_angle;
        float mod1 = d2_angle - (float)OB_F2I_RN(d2_angle);
        float nz = fabsf(mod1);
        mod1 *= mod1; mod1 *= mod1;
        float mod2 = d2_angle2 - (float)OB_F2I_RN(d2_angle2);
        nz += fabsf(mod2);
        noisiness[i] = nz;
        mod2 *= mod2; mod2 *= mod2;
        const float avg_mod = .25f * (d2A[i] + mod1 + 2 * mod2);
        tonality[i] = 1.f / (1.f + 40.f * 16.f * pi4 * avg_mod) - .015f;
        tonality2[i] = 1.f / (1.f + 40.f * 16.f * pi4 * mod2) - .015f;
        A[i] = angle2; dA[i] = d_angle2; d2A[i] = mod2;
    }
    g.sync();
    for (i = 2 + g.lane; i < N2 - 1; i += g.n) {
        const float tt = ob_fmin(tonality2[i], ob_fmax(tonality2[i - 1], tonality2[i + 1]));
        tonality[i] = .9f * ob_fmax(tonality[i], tt - .1f);
    }
    g.sync();
#define BINE(k) (RE(k) * RE(k) + RE(N - (k)) * RE(N - (k)) + IM(k) * IM(k) + IM(N - (k)) * IM(N - (k)))
    // per-band sums, one band per lane, each in the reference's order: bsum[b] = E, bsum[20 + b] = tonal energy, bsum[40 + b] = noisy energy,
    // bsum[60] = the DC band's energy (bins 0..3)
    for (b = g.lane; b <= NB; b += g.n) {
        if (b == NB) {
            const float X1r = 2 * RE(0), X2r = 2 * IM(0);
            float E = X1r * X1r + X2r * X2r;
            for (i = 1; i < 4; i++) { const float binE = BINE(i); E += binE; }
            bsum[60] = E;
        } else {
            float E = 0, tE = 0, nE = 0;
            for (i = OB_AN_TBANDS[b]; i < OB_AN_TBANDS[b + 1]; i++) {
                const float binE = BINE(i);
                E += binE;
                tE += binE * ob_fmax(0, tonality[i]);
                nE += binE * 2.f * (.5f - noisiness[i]);
            }
            bsum[b] = E; bsum[20 + b] = tE; bsum[40 + b] = nE;
        }
    }
    g.sync();
    // per-band transcendental work of the statistics below, one band per lane (each value exactly what the one-lane loop computed: the double-precision
    // log / sqrt of the band energy, and the eight-frame sums L1 / L2 in the reference's order), into the spent FFT buffer
    float *pre = out + 320;                                          // [b] log, [32 + b] sqrt, [64 + b] L1, [96 + b] L2
    {
        const int Ec = tonal.E_count;
        for (b = g.lane; b < NB; b += g.n) {
            const float E = bsum[b];
            float L1 = 0, L2 = 0;
            pre[b] = (float)log((double)(E + 1e-10f));
            pre[32 + b] = (float)sqrt((double)(E + 1e-10f));
            for (i = 0; i < OB_AN_NB_FRAMES; i++) { const float Ei = i == Ec ? E : tonal.E[i][b]; L1 += (float)sqrt((double)Ei); L2 += Ei; }
            pre[64 + b] = L1; pre[96 + b] = L2;
        }
    }
    g.sync();
    if (g.lane == 0) {                                               // ---- from here on: scalar statistics, one lane ----
    int E_count = tonal.E_count;
    frame_tonality = 0; max_frame_tonality = 0; info->activity = 0; frame_noisiness = 0; frame_stationarity = 0;
    if (!count) for (b = 0; b < NB; b++) { tonal.lowE[b] = 1e10; tonal.highE[b] = -1e10; }
    relativeE = 0; frame_loudness = 0;
#define BINE(k) (RE(k) * RE(k) + RE(N - (k)) * RE(N - (k)) + IM(k) * IM(k) + IM(N - (k)) * IM(N - (k)))
    {   // the very first band is special because of DC
        const float E = bsum[60];
        band_log2[0] = .5f * 1.442695f * (float)log((double)(E + 1e-10f));
    }
    int bad = 0;
    for (b = 0; b < NB; b++) {
        const float E = bsum[b], tE = bsum[20 + b], nE = bsum[40 + b];
        float L1, L2, stationarity;
        if (!(E < 1e9f) || E != E) { info->valid = 0; bad = 1; break; }
        tonal.E[E_count][b] = E;
        frame_noisiness += nE / (1e-15f + E);
        frame_loudness += pre[32 + b];
        logE[b] = pre[b];
        band_log2[b + 1] = .5f * 1.442695f * pre[b];
        tonal.logE[E_count][b] = logE[b];
        if (count == 0) tonal.highE[b] = tonal.lowE[b] = logE[b];
        if ((double)tonal.highE[b] > (double)tonal.lowE[b] + 7.5) {
            if (tonal.highE[b] - logE[b] > logE[b] - tonal.lowE[b]) tonal.highE[b] -= .01f;
            else tonal.lowE[b] += .01f;
        }
        if (logE[b] > tonal.highE[b]) {
            tonal.highE[b] = logE[b];
            tonal.lowE[b] = ob_fmax(tonal.highE[b] - 15, tonal.lowE[b]);
        } else if (logE[b] < tonal.lowE[b]) {
            tonal.lowE[b] = logE[b];
            tonal.highE[b] = ob_fmin(tonal.lowE[b] + 15, tonal.highE[b]);
        }
        relativeE += (logE[b] - tonal.lowE[b]) / (1e-5f + (tonal.highE[b] - tonal.lowE[b]));
        L1 = pre[64 + b]; L2 = pre[96 + b];
        stationarity = ob_fmin(0.99f, L1 / (float)sqrt(1e-15 + (double)(OB_AN_NB_FRAMES * L2)));
        stationarity *= stationarity;
        stationarity *= stationarity;
        frame_stationarity += stationarity;
        band_tonality[b] = ob_fmax(tE / (1e-15f + E), stationarity * tonal.prev_band_tonality[b]);
        frame_tonality += band_tonality[b];
        if (b >= NB - OB_AN_SKIP_BANDS) frame_tonality -= band_tonality[b - NB + OB_AN_SKIP_BANDS];
        max_frame_tonality = ob_fmax(max_frame_tonality, (1.f + .03f * (b - NB)) * frame_tonality);
        slope += band_tonality[b] * (b - 8);
        tonal.prev_band_tonality[b] = band_tonality[b];
    }
    if (!bad) {
    leakage_from[0] = band_log2[0];
    leakage_to[0] = band_log2[0] - 2.5f;
    for (b = 1; b < NB + 1; b++) {
        const float leak_slope = 2.f * (OB_AN_TBANDS[b] - OB_AN_TBANDS[b - 1]) / 4;
        leakage_from[b] = ob_fmin(leakage_from[b - 1] + leak_slope, band_log2[b]);
        leakage_to[b] = ob_fmax(leakage_to[b - 1] - leak_slope, band_log2[b] - 2.5f);
    }
    for (b = NB - 2; b >= 0; b--) {
        const float leak_slope = 2.f * (OB_AN_TBANDS[b + 1] - OB_AN_TBANDS[b]) / 4;
        leakage_from[b] = ob_fmin(leakage_from[b + 1] + leak_slope, leakage_from[b]);
        leakage_to[b] = ob_fmax(leakage_to[b + 1] - leak_slope, leakage_to[b]);
    }
    for (b = 0; b < NB + 1; b++) {
        const float boost = ob_fmax(0, leakage_to[b] - band_log2[b]) + ob_fmax(0, band_log2[b] - (leakage_from[b] + 2.5f));
        info->leak_boost[b] = (uint8_t)ob_imin(255, (int)floor(.5 + (double)(64.f * boost)));
    }
    for (; b < OB_AN_LEAK_BANDS; b++) info->leak_boost[b] = 0;
    // spec_variability (analysis.c:803-820) only feeds the network: it is computed by all lanes after this block
    bandwidth_mask = 0; bandwidth = 0; maxE = 0;
    noise_floor = 5.7e-4f / (1 << ob_imax(0, lsb_depth - 8));
    noise_floor *= noise_floor;
    below_max_pitch = 0; above_max_pitch = 0;
    for (b = 0; b < NB; b++) {
        const float E = bsum[b];
        float Em;
        const int band_start = OB_AN_TBANDS[b], band_end = OB_AN_TBANDS[b + 1];
        maxE = ob_fmax(maxE, E);
        if (band_start < 64) below_max_pitch += E; else above_max_pitch += E;
        tonal.meanE[b] = ob_fmax((1 - alphaE2) * tonal.meanE[b], E);
        Em = ob_fmax(E, tonal.meanE[b]);
        if (E * 1e9f > maxE && (Em > 3 * noise_floor * (band_end - band_start) || E > noise_floor * (band_end - band_start))) bandwidth = b + 1;
        is_masked[b] = E < (tonal.prev_bandwidth >= b + 1 ? .01f : .05f) * bandwidth_mask;
        bandwidth_mask = ob_fmax(.05f * bandwidth_mask, E);
    }
    {   // the last two bands: only the energy above 12 kHz from the down-sampler's high-pass branch
        float Em, E = hp_ener * (1.f / (60 * 60));
        const float noise_ratio = tonal.prev_bandwidth == 20 ? 10.f : 30.f;
        above_max_pitch += E;
        tonal.meanE[b] = ob_fmax((1 - alphaE2) * tonal.meanE[b], E);
        Em = ob_fmax(E, tonal.meanE[b]);
        if (Em > 3 * noise_ratio * noise_floor * 160 || E > noise_ratio * noise_floor * 160) bandwidth = 20;
        is_masked[b] = E < (tonal.prev_bandwidth == 20 ? .01f : .05f) * bandwidth_mask;
    }
    if (above_max_pitch > below_max_pitch) info->max_pitch_ratio = below_max_pitch / above_max_pitch;
    else info->max_pitch_ratio = 1;
    if (bandwidth == 20 && is_masked[NB]) bandwidth -= 2;
    else if (bandwidth > 0 && bandwidth <= NB && is_masked[bandwidth - 1]) bandwidth--;
    if (count <= 2) bandwidth = 20;
    frame_loudness = 20 * (float)log10((double)frame_loudness);
    tonal.Etracker = ob_fmax(tonal.Etracker - .003f, frame_loudness);
    tonal.lowECount *= (1 - alphaE);
    if (frame_loudness < tonal.Etracker - 30) tonal.lowECount += alphaE;
    for (i = 0; i < 8; i++) {
        float sum = 0;
        for (b = 0; b < 16; b++) sum += OB_AN_DCT[i * 16 + b] * logE[b];
        BFCC[i] = sum;
    }
    for (i = 0; i < 8; i++) {
        float sum = 0;
        for (b = 0; b < 16; b++) sum += OB_AN_DCT[i * 16 + b] * .5f * (tonal.highE[b] + tonal.lowE[b]);
        midE[i] = sum;
    }
    frame_stationarity /= NB;
    relativeE /= NB;
    if (count < 10) relativeE = .5f;
    frame_noisiness /= NB;
    info->activity = frame_noisiness + (1 - frame_noisiness) * relativeE;
    frame_tonality = (max_frame_tonality / (NB - OB_AN_SKIP_BANDS));
    frame_tonality = ob_fmax(frame_tonality, tonal.prev_tonality * .8f);
    tonal.prev_tonality = frame_tonality;
    slope /= 8 * 8;
    info->tonality_slope = slope;
    tonal.E_count = (E_count + 1) % OB_AN_NB_FRAMES;
    const int count1 = ob_imin(count + 1, OB_AN_COUNT_MAX);
    tonal.count = count1;
    info->tonality = frame_tonality;
    for (i = 0; i < 4; i++)
        features[i] = -0.12299f * (BFCC[i] + tonal.mem[i + 24]) + 0.49195f * (tonal.mem[i] + tonal.mem[i + 16]) + 0.69693f * tonal.mem[i + 8] - 1.4349f * tonal.cmean[i];
    for (i = 0; i < 4; i++) tonal.cmean[i] = (1 - alpha) * tonal.cmean[i] + alpha * BFCC[i];
    for (i = 0; i < 4; i++) features[4 + i] = 0.63246f * (BFCC[i] - tonal.mem[i + 24]) + 0.31623f * (tonal.mem[i] - tonal.mem[i + 16]);
    for (i = 0; i < 3; i++)
        features[8 + i] = 0.53452f * (BFCC[i] + tonal.mem[i + 24]) - 0.26726f * (tonal.mem[i] + tonal.mem[i + 16]) - 0.53452f * tonal.mem[i + 8];
    if (count1 > 5) for (i = 0; i < 9; i++) tonal.std[i] = (1 - alpha) * tonal.std[i] + alpha * features[i] * features[i];
    for (i = 0; i < 4; i++) features[i] = BFCC[i] - midE[i];
    for (i = 0; i < 8; i++) {
        tonal.mem[i + 24] = tonal.mem[i + 16];
        tonal.mem[i + 16] = tonal.mem[i + 8];
        tonal.mem[i + 8] = tonal.mem[i];
        tonal.mem[i] = BFCC[i];
    }
    for (i = 0; i < 9; i++) features[11 + i] = (float)sqrt((double)tonal.std[i]) - OB_AN_STD_BIAS[i];
    features[18] = spec_variability - 0.78f;
    features[20] = info->tonality - 0.154723f;
    features[21] = info->activity - 0.724643f;
    features[22] = frame_stationarity - 0.743717f;
    features[23] = info->tonality_slope + 0.069216f;
    features[24] = tonal.lowECount - 0.067930f;
    for (i = 0; i < 25; i++) out[i] = features[i];                   // the FFT output is spent: its buffer carries the network's vectors
    info->bandwidth = bandwidth;
    tonal.prev_bandwidth = bandwidth;
    info->noisiness = frame_noisiness;
    info->valid = 1;
    }                                                                // !bad
    out[255] = bad ? 0.f : 1.f;
    }                                                                // lane 0
    g.sync();
    if (out[255] != 0.f) {                                           // the network, all lanes: one output neuron per lane (mlp.c)
        {   // spec_variability: the 64 frame-pair distances one per lane (18 terms each, in the reference's order), then min / sum as the reference does
            float *dm = out + 448;
            for (int p = g.lane; p < OB_AN_NB_FRAMES * OB_AN_NB_FRAMES; p += g.n) {
                const int fi = p / OB_AN_NB_FRAMES, fj = p % OB_AN_NB_FRAMES;
                float dist = 0;
                for (int k = 0; k < NB; k++) { const float tmp = tonal.logE[fi][k] - tonal.logE[fj][k]; dist += tmp * tmp; }
                dm[p] = dist;
            }
            g.sync();
            if (g.lane == 0) {
                float sv = 0;
                for (i = 0; i < OB_AN_NB_FRAMES; i++) {
                    float mindist = 1e15f;
                    for (int j = 0; j < OB_AN_NB_FRAMES; j++) if (j != i) mindist = ob_fmin(mindist, dm[i * OB_AN_NB_FRAMES + j]);
                    sv += mindist;
                }
                out[18] = (float)sqrt((double)(sv / OB_AN_NB_FRAMES / NB)) - 0.78f;
            }
            g.sync();
        }
        float *layer_out = out + 32, *frame_probs = out + 64;
        ob_dense(g, OB_AN_L0_B, OB_AN_L0_W, 25, 32, 0, layer_out, out);
        ob_gru(g, tonal.rnn_state, layer_out, out + 96);
        ob_dense(g, OB_AN_L2_B, OB_AN_L2_W, 24, 2, 1, frame_probs, tonal.rnn_state);
        if (g.lane == 0) { info->activity_probability = frame_probs[1]; info->music_prob = frame_probs[0]; }
    }
    g.sync();
#undef RE
#undef IM
#undef BINE
}

// tonality_get_info (analysis.c:233-409), Fs = 48000
OB_DEV_NOINLINE void ob_tonality_get_info(ObTonalState &tonal, ObAnalysisInfo &info_out, int len)
{
    int pos = tonal.read_pos, i;
    int curr_lookahead = tonal.write_pos - tonal.read_pos;
    if (curr_lookahead < 0) curr_lookahead += OB_AN_DETECT;
    tonal.read_subframe += len / (48000 / 400);
    while (tonal.read_subframe >= 8) { tonal.read_subframe -= 8; tonal.read_pos++; }
    if (tonal.read_pos >= OB_AN_DETECT) tonal.read_pos -= OB_AN_DETECT;
    if (len > 48000 / 50 && pos != tonal.write_pos) { pos++; if (pos == OB_AN_DETECT) pos = 0; }
    if (pos == tonal.write_pos) pos--;
    if (pos < 0) pos = OB_AN_DETECT - 1;
    const int pos0 = pos;
    info_out = tonal.info[pos];
    if (!info_out.valid) return;
    float tonality_max = info_out.tonality, tonality_avg = info_out.tonality;
    int tonality_count = 1, bandwidth_span = 6;
    for (i = 0; i < 3; i++) {
        pos++;
        if (pos == OB_AN_DETECT) pos = 0;
        if (pos == tonal.write_pos) break;
        tonality_max = ob_fmax(tonality_max, tonal.info[pos].tonality);
        tonality_avg += tonal.info[pos].tonality;
        tonality_count++;
        info_out.bandwidth = ob_imax(info_out.bandwidth, tonal.info[pos].bandwidth);
        bandwidth_span--;
    }
    pos = pos0;
    for (i = 0; i < bandwidth_span; i++) {
        pos--;
        if (pos < 0) pos = OB_AN_DETECT - 1;
        if (pos == tonal.write_pos) break;
        info_out.bandwidth = ob_imax(info_out.bandwidth, tonal.info[pos].bandwidth);
    }
    info_out.tonality = ob_fmax(tonality_avg / tonality_count, tonality_max - .2f);
    int mpos = pos0, vpos = pos0;
    if (curr_lookahead > 15) {
        mpos += 5; if (mpos >= OB_AN_DETECT) mpos -= OB_AN_DETECT;
        vpos += 1; if (vpos >= OB_AN_DETECT) vpos -= OB_AN_DETECT;
    }
    float prob_min = 1.f, prob_max = 0.f;
    const float vad_prob = tonal.info[vpos].activity_probability;
    float prob_count = ob_fmax(.1f, vad_prob);
    float prob_avg = ob_fmax(.1f, vad_prob) * tonal.info[mpos].music_prob;
    while (1) {
        mpos++; if (mpos == OB_AN_DETECT) mpos = 0;
        if (mpos == tonal.write_pos) break;
        vpos++; if (vpos == OB_AN_DETECT) vpos = 0;
        if (vpos == tonal.write_pos) break;
        const float pos_vad = tonal.info[vpos].activity_probability;
        prob_min = ob_fmin((prob_avg - 10 * (vad_prob - pos_vad)) / prob_count, prob_min);
        prob_max = ob_fmax((prob_avg + 10 * (vad_prob - pos_vad)) / prob_count, prob_max);
        prob_count += ob_fmax(.1f, pos_vad);
        prob_avg += ob_fmax(.1f, pos_vad) * tonal.info[mpos].music_prob;
    }
    info_out.music_prob = prob_avg / prob_count;
    prob_min = ob_fmin(prob_avg / prob_count, prob_min);
    prob_max = ob_fmax(prob_avg / prob_count, prob_max);
    prob_min = ob_fmax(prob_min, 0.f);
    prob_max = ob_fmin(prob_max, 1.f);
    if (curr_lookahead < 10) {
        float pmin = prob_min, pmax = prob_max;
        pos = pos0;
        for (i = 0; i < ob_imin(tonal.count - 1, 15); i++) {
            pos--;
            if (pos < 0) pos = OB_AN_DETECT - 1;
            pmin = ob_fmin(pmin, tonal.info[pos].music_prob);
            pmax = ob_fmax(pmax, tonal.info[pos].music_prob);
        }
        pmin = ob_fmax(0.f, pmin - .1f * vad_prob);
        pmax = ob_fmin(1.f, pmax + .1f * vad_prob);
        prob_min += (1.f - .1f * curr_lookahead) * (pmin - prob_min);
        prob_max += (1.f - .1f * curr_lookahead) * (pmax - prob_max);
    }
    info_out.music_prob_min = prob_min;
    info_out.music_prob_max = prob_max;
}

// run_analysis (analysis.c:955-981) for one frame of frame_size samples at 48 kHz (analysis_frame_size == frame_size)
template <class G>
OB_DEV void ob_run_analysis(const G &g, ObTonalState &an, const float *pcm, int frame_size, int C, int lsb_depth, ObAnalysisInfo &info, float *work, float *fbuf)
{
    int analysis_frame_size = frame_size - (frame_size & 1);
    analysis_frame_size = ob_imin((OB_AN_DETECT - 5) * 48000 / 50, analysis_frame_size);
    const int offset0 = an.analysis_offset;
    g.sync();
    int pcm_len = analysis_frame_size - offset0, offset = offset0;
    while (pcm_len > 0) {
        ob_tonality_analysis(g, an, pcm, ob_imin(48000 / 50, pcm_len), offset, C, lsb_depth, work, fbuf);
        offset += 48000 / 50;
        pcm_len -= 48000 / 50;
    }
    if (g.lane == 0) {
        an.analysis_offset = analysis_frame_size - frame_size;
        ob_tonality_get_info(an, info, frame_size);
    }
    g.sync();
}
