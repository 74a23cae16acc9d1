// dec_synth.cuh -- per-stream synthesis, cooperatively by one thread block per stream, frames in order:
//   energy de-quantisation (quant_bands.c:428-542) -> anti-collapse (bands.c:268-362) -> denormalise
//   (bands.c:196-265) -> inverse MDCT = pre-rotation, mixed-radix FFT, post-rotation, TDAC window
//   (mdct.c:242-342, kiss_fft.c:48-567) -> pitch post-filter (celt.c:162-256, celt_decoder.c:1301-1325)
//   -> energy-history update (celt_decoder.c:1327-1357) -> de-emphasis (celt_decoder.c:249-377).
// This is the only stage that carries per-stream float state from frame to frame; it keeps that state in
// shared memory across all frames of a launch.
#pragma once
#include "ob_ir.h"
#include "ob_group.cuh"
#include "dec_bands.cuh"     // ObLcg, OB_SQRTF, tables via dec_symbols.cuh

#define OB_HISTK 1056                       // history kept per channel in SHARED memory: the post-filter reads back COMBFILTER_MAXPERIOD + 2
                                            // samples (celt.h:218), the concealment's LPC analysis MAX_PERIOD + CELT_LPC_ORDER = 1048
#define OB_RING 2048                        // DECODE_BUFFER_SIZE (celt_decoder.c:72): the full history, kept as a ring in GLOBAL memory;
                                            // only the concealment's pitch search (first lost frame of a burst) reads beyond OB_HISTK
#define OB_BUF_LEN (OB_HISTK + OB_MAX_N + OB_OVERLAP)

// Persistent per-stream decoder state (struct OpusCustomDecoder, celt_decoder.c:80-123, minus the configuration).
struct ObDecState {
    // owned by the plan pass (ObPlanState, dec_plc.cuh): the integer loss state machine
    uint32_t rng;                 // st->rng: range-coder state left by the last decoded frame, advanced by noise PLC (noise seed)
    int32_t loss_duration, skip_plc, plc_end, last_frame_size;
    // owned by the synthesis kernel
    uint32_t final_range;         // OpusDecoder.rangeFinal of the last call (opus_decoder.c:651-654)
    int32_t pf_period, pf_period_old, pf_tapset, pf_tapset_old;
    float pf_gain, pf_gain_old;
    float preemph_mem[2];
    int32_t last_packet_duration;
    int32_t prefilter_and_fold, last_pitch_index, ring_pos;
    int32_t pad;
    float softclip_mem[2];        // OpusDecoder.softclip_mem: the int16 path's soft clipper carries its last gain (opus_decoder.c:96)
    int32_t pkt_samples;          // samples (or error) accumulated over the frames of a multi-frame packet still being decoded
    float lpc[2][24];             // CELT_LPC_ORDER coefficients per channel, kept across consecutive losses (celt_decoder.c:637)
    float oldBandE[2 * OB_NB], oldLogE[2 * OB_NB], oldLogE2[2 * OB_NB], backgroundLogE[2 * OB_NB];
};

struct ObCpx { float r, i; };
#define OB_TW(k) (((const ObCpx *)OB_FFT_TWIDDLES)[k])
#define OB_CMUL(m, a, b) do { (m).r = (a).r * (b).r - (a).i * (b).i; (m).i = (a).r * (b).i + (a).i * (b).r; } while (0)

OB_DEV const int16_t *ob_fft_bitrev(int shift) { return shift == 0 ? OB_FFT_BITREV480 : shift == 1 ? OB_FFT_BITREV240 : shift == 2 ? OB_FFT_BITREV120 : OB_FFT_BITREV60; }
OB_DEV const int16_t *ob_fft_factors(int shift) { return shift == 0 ? OB_FFT_FACTORS480 : shift == 1 ? OB_FFT_FACTORS240 : shift == 2 ? OB_FFT_FACTORS120 : OB_FFT_FACTORS60; }

// n / ds for the output down-sampling factor ds (1, 2, 3, 4 or 6): 48 kHz output (ds == 1) skips the division
OB_DEV int ob_div_ds(int n, int ds) { return ds == 1 ? n : n / ds; }

// One decimation-in-time stage of opus_fft_impl (kiss_fft.c:521-567) over `nblk` independent transforms whose data
// start blk_stride floats apart; every butterfly of the stage is an independent work item.
template <class G>
OB_DEV void ob_fft_stage(const G &g, float *base, int nblk, int blk_stride, int p, int fstride, int m, int Nst, int mm)
{
    if (p == 2) {                                                   // kf_bfly2, m == 4 (kiss_fft.c:48-100)
        const float tw = 0.7071067812f;
        const int per = Nst * 4, total = nblk * per;
        const ObDiv dper = ob_div_make_blocks(per, nblk);
        for (int t = g.lane; t < total; t += g.n) {
            const int blk = ob_div_q(t, dper);
            ObCpx *F = (ObCpx *)(base + blk * blk_stride) + ((t - blk * per) >> 2) * 8;
            const int k = t & 3;
            ObCpx *F2 = F + 4, x = F2[k], tt;
            if (k == 0) tt = x;
            else if (k == 1) { tt.r = (x.r + x.i) * tw; tt.i = (x.i - x.r) * tw; }
            else if (k == 2) { tt.r = x.i; tt.i = -x.r; }
            else { tt.r = (x.i - x.r) * tw; tt.i = (-(x.i + x.r)) * tw; }
            F2[k].r = F[k].r - tt.r; F2[k].i = F[k].i - tt.i;
            F[k].r += tt.r; F[k].i += tt.i;
        }
    } else if (p == 4) {                                            // kf_bfly4 (kiss_fft.c:102-171)
        const int per = Nst * m, total = nblk * per, lm = ob_log2i(m);       // m is a power of two in radix-4 stages
        const ObDiv dper = ob_div_make_blocks(per, nblk);
        for (int t = g.lane; t < total; t += g.n) {
            const int blk = ob_div_q(t, dper);
            const int w = t - blk * per, i = w >> lm, j = w & (m - 1);
            ObCpx *F = (ObCpx *)(base + blk * blk_stride) + i * mm + j;
            ObCpx s0, s1, s2, s3, s4, s5;
            if (m == 1) { s0 = F[1]; s1 = F[2]; s2 = F[3]; }
            else { OB_CMUL(s0, F[m], OB_TW(j * fstride)); OB_CMUL(s1, F[2 * m], OB_TW(j * fstride * 2)); OB_CMUL(s2, F[3 * m], OB_TW(j * fstride * 3)); }
            s5.r = F->r - s1.r; s5.i = F->i - s1.i;
            F->r += s1.r; F->i += s1.i;
            s3.r = s0.r + s2.r; s3.i = s0.i + s2.i;
            s4.r = s0.r - s2.r; s4.i = s0.i - s2.i;
            F[2 * m].r = F->r - s3.r; F[2 * m].i = F->i - s3.i;
            F->r += s3.r; F->i += s3.i;
            F[m].r = s5.r + s4.i; F[m].i = s5.i - s4.r;
            F[3 * m].r = s5.r - s4.i; F[3 * m].i = s5.i + s4.r;
        }
    } else if (p == 3) {                                            // kf_bfly3 (kiss_fft.c:176-236)
        const int per = Nst * m, total = nblk * per, lm = ob_log2i(m);       // m is a power of two in radix-3 stages
        const ObDiv dper = ob_div_make_blocks(per, nblk);
        const float epi3i = OB_TW(fstride * m).i;
        for (int t = g.lane; t < total; t += g.n) {
            const int blk = ob_div_q(t, dper);
            const int w = t - blk * per, i = w >> lm, j = w & (m - 1);
            ObCpx *F = (ObCpx *)(base + blk * blk_stride) + i * mm + j;
            ObCpx s0, s1, s2, s3;
            OB_CMUL(s1, F[m], OB_TW(j * fstride)); OB_CMUL(s2, F[2 * m], OB_TW(j * fstride * 2));
            s3.r = s1.r + s2.r; s3.i = s1.i + s2.i;
            s0.r = s1.r - s2.r; s0.i = s1.i - s2.i;
            F[m].r = F->r - s3.r * .5f; F[m].i = F->i - s3.i * .5f;
            s0.r *= epi3i; s0.i *= epi3i;
            F->r += s3.r; F->i += s3.i;
            F[2 * m].r = F[m].r + s0.i; F[2 * m].i = F[m].i - s0.r;
            F[m].r = F[m].r - s0.i; F[m].i = F[m].i + s0.r;
        }
    } else {                                                        // kf_bfly5 (kiss_fft.c:240-308)
        const int per = Nst * m, total = nblk * per;                          // radix-5 is always the last stage: Nst == 1
        const ObDiv dper = ob_div_make_blocks(per, nblk);
        const ObCpx ya = OB_TW(fstride * m), yb = OB_TW(fstride * 2 * m);
        for (int t = g.lane; t < total; t += g.n) {
            const int blk = ob_div_q(t, dper);
            const int i = 0, u = t - blk * per;
            ObCpx *F0 = (ObCpx *)(base + blk * blk_stride) + i * mm + u;
            ObCpx *F1 = F0 + m, *F2 = F0 + 2 * m, *F3 = F0 + 3 * m, *F4 = F0 + 4 * m;
            ObCpx s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12;
            s0 = *F0;
            OB_CMUL(s1, *F1, OB_TW(u * fstride)); OB_CMUL(s2, *F2, OB_TW(2 * u * fstride));
            OB_CMUL(s3, *F3, OB_TW(3 * u * fstride)); OB_CMUL(s4, *F4, OB_TW(4 * u * fstride));
            s7.r = s1.r + s4.r; s7.i = s1.i + s4.i; s10.r = s1.r - s4.r; s10.i = s1.i - s4.i;
            s8.r = s2.r + s3.r; s8.i = s2.i + s3.i; s9.r = s2.r - s3.r; s9.i = s2.i - s3.i;
            F0->r = s0.r + (s7.r + s8.r); F0->i = s0.i + (s7.i + s8.i);
            s5.r = s0.r + (s7.r * ya.r + s8.r * yb.r); s5.i = s0.i + (s7.i * ya.r + s8.i * yb.r);
            s6.r = s10.i * ya.i + s9.i * yb.i; s6.i = -(s10.r * ya.i + s9.r * yb.i);
            F1->r = s5.r - s6.r; F1->i = s5.i - s6.i; F4->r = s5.r + s6.r; F4->i = s5.i + s6.i;
            s11.r = s0.r + (s7.r * yb.r + s8.r * ya.r); s11.i = s0.i + (s7.i * yb.r + s8.i * ya.r);
            s12.r = s9.i * ya.i - s10.i * yb.i; s12.i = s10.r * yb.i - s9.r * ya.i;
            F2->r = s11.r + s12.r; F2->i = s11.i + s12.i; F3->r = s11.r - s12.r; F3->i = s11.i - s12.i;
        }
    }
    g.sync();
}

// nblk inverse MDCTs of N2 = 1920>>(shift+1) coefficients each, inputs interleaved with stride nblk in `in`,
// output block b at out + b*N2 .. (clt_mdct_backward, mdct.c:242-342; the B short blocks of celt_synthesis
// celt_decoder.c:445-447 are independent up to the TDAC step, so all of them run together).
template <class G>
OB_DEV void ob_imdct(const G &g, const float *in, float *out, int shift, int nblk)
{
    const int N2 = 1920 >> (shift + 1), N4 = N2 >> 1;
    const float *trig = OB_MDCT_TRIG + (shift == 0 ? 0 : shift == 1 ? 960 : shift == 2 ? 1440 : 1680);
    const int16_t *br = ob_fft_bitrev(shift);
    const ObDiv dN4 = ob_div_make_blocks(N4, nblk);
    for (int t = g.lane; t < nblk * N4; t += g.n) {                  // pre-rotation into bit-reversed order
        const int b = ob_div_q(t, dN4), i = t - b * N4;
        const float x1 = in[b + nblk * (2 * i)], x2 = in[b + nblk * (N2 - 1 - 2 * i)];
        float *yp = out + b * N2 + (OB_OVERLAP >> 1);
        const int rev = br[i];
        yp[2 * rev + 1] = x2 * trig[i] + x1 * trig[N4 + i];
        yp[2 * rev] = x1 * trig[i] - x2 * trig[N4 + i];
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
            ob_fft_stage(g, out + (OB_OVERLAP >> 1), nblk, N2, fac[2 * i], fstride[i] << shift, m, fstride[i], m2);
            m = m2;
        }
    }
    const int half = (N4 + 1) >> 1;
    const ObDiv dhalf = ob_div_make_blocks(half, nblk);
    for (int t = g.lane; t < nblk * half; t += g.n) {                // post-rotation, both ends at once
        const int b = ob_div_q(t, dhalf), i = t - b * half;
        float *yp0 = out + b * N2 + (OB_OVERLAP >> 1) + 2 * i, *yp1 = out + b * N2 + (OB_OVERLAP >> 1) + N2 - 2 - 2 * i;
        float re = yp0[1], im = yp0[0], t0 = trig[i], t1 = trig[N4 + i];
        const float yr0 = re * t0 + im * t1, yi0 = re * t1 - im * t0;
        re = yp1[1]; im = yp1[0];
        t0 = trig[N4 - i - 1]; t1 = trig[N2 - i - 1];
        const float yr1 = re * t0 + im * t1, yi1 = re * t1 - im * t0;
        yp0[0] = yr0; yp1[1] = yi0;
        yp1[0] = yr1; yp0[1] = yi1;
    }
    g.sync();
    for (int t = g.lane; t < nblk * (OB_OVERLAP / 2); t += g.n) {    // TDAC mirror
        const int b = t / (OB_OVERLAP / 2), i = t - b * (OB_OVERLAP / 2);
        float *o = out + b * N2;
        const float x1 = o[OB_OVERLAP - 1 - i], x2 = o[i], w1 = OB_WINDOW[i], w2 = OB_WINDOW[OB_OVERLAP - 1 - i];
        o[i] = w2 * x2 - w1 * x1;
        o[OB_OVERLAP - 1 - i] = w1 * x2 + w2 * x1;
    }
    g.sync();
}

// comb_filter with y == x (celt.c:190-256).  y[i] only reads outputs at least min(T)-2 samples back, so samples are
// produced in order in chunks of that many, each chunk in parallel.
template <class G>
OB_DEV void ob_comb_filter(const G &g, float *x, int T0, int T1, int N, float g0, float g1, int tapset0, int tapset1)
{
    if (g0 == 0 && g1 == 0) return;
    const float gains[3][3] = {{0.3066406250f, 0.2170410156f, 0.1296386719f}, {0.4638671875f, 0.2680664062f, 0.f}, {0.7998046875f, 0.1000976562f, 0.f}};
    T0 = ob_imax(T0, 15); T1 = ob_imax(T1, 15);
    const float g00 = g0 * gains[tapset0][0], g01 = g0 * gains[tapset0][1], g02 = g0 * gains[tapset0][2];
    const float g10 = g1 * gains[tapset1][0], g11 = g1 * gains[tapset1][1], g12 = g1 * gains[tapset1][2];
    int overlap = OB_OVERLAP;
    if (g0 == g1 && T0 == T1 && tapset0 == tapset1) overlap = 0;
    const int chunk0 = ob_imin(T0, T1) - 2, chunk1 = T1 - 2;
    for (int base = 0; base < overlap; base += chunk0) {
        const int lim = ob_imin(overlap, base + chunk0);
        for (int i = base + g.lane; i < lim; i += g.n) {
            const float f = OB_WINDOW[i] * OB_WINDOW[i];
            x[i] = x[i] + ((1.0f - f) * g00) * x[i - T0] + ((1.0f - f) * g01) * (x[i - T0 + 1] + x[i - T0 - 1])
                        + ((1.0f - f) * g02) * (x[i - T0 + 2] + x[i - T0 - 2])
                        + (f * g10) * x[i - T1] + (f * g11) * (x[i - T1 + 1] + x[i - T1 - 1])
                        + (f * g12) * (x[i - T1 + 2] + x[i - T1 - 2]);
        }
        g.sync();
    }
    if (g1 == 0) return;
    for (int base = overlap; base < N; base += chunk1) {             // comb_filter_const (celt.c:162-185)
        const int lim = ob_imin(N, base + chunk1);
        for (int i = base + g.lane; i < lim; i += g.n)
            x[i] = x[i] + g10 * x[i - T1] + g11 * (x[i - T1 + 1] + x[i - T1 - 1]) + g12 * (x[i - T1 + 2] + x[i - T1 - 2]);
        g.sync();
    }
}

// Shared-memory working set of one stream.  CH = 1: a mono decoder that meets mono frames only (20 KB instead of 29 KB: 8 blocks per SM).
template <int CH>
struct ObSynthSharedT {
    // Threads of the synthesis block and resident blocks per SM (opus_b200.cu ob_k_synth).  Measured, ms per 819 200 mono frames / per 163 840 stereo
    // frames: 128 threads x 8 / 7 blocks 18.34 / 7.05, 96 x 10 / 9: 16.76 / 7.08, 96 x 9 / 7: 16.91 / 6.72, 96 x 8 / 6: 17.17 / 6.86, 64 x 10 / 7: 16.53 / 6.96,
    // 160 x 6: 18.79, 192 x 5: 20.47 -- a frame is 960 samples and 120-240 butterflies per FFT stage: small blocks waste fewer lanes per barrier.
    static constexpr int synth_threads = CH == 1 ? 64 : 96, synth_blocks = CH == 1 ? 10 : 7;
    float buf[CH][OB_BUF_LEN];       // [history | current frame | overlap tail] per channel (the tail of decode_mem)
    float freq[2][OB_MAX_N];         // X tile -> MDCT coefficients -> PCM staging; both halves are the pitch concealment's work area, mono or not
    float oldBandE[2 * OB_NB], oldLogE[2 * OB_NB], oldLogE2[2 * OB_NB], backgroundLogE[2 * OB_NB];
    float gain[2 * OB_NB];
    float scanA[256], scanB[256];
    float red[32];
    ObFrameHdr hdr;
    uint32_t lcg_a[33], lcg_c[33];   // celt_lcg_rand^k as affine maps, k = 0..32 (anti-collapse noise, bands.c:335-340)
    uint8_t band_of_bin[104];        // 2.5 ms bin -> band index (21 = above the last band)
    int32_t pf_period, pf_period_old, pf_tapset, pf_tapset_old;
    float pf_gain, pf_gain_old, preemph_mem[2];
    float lpc[2][24];
    int32_t last_pitch_index, paf;   // paf = st->prefilter_and_fold
    float decode_gain;               // linear gain of OPUS_SET_GAIN (1 = none)
    float softclip_mem[2];           // OpusDecoder.softclip_mem (int16 API)
    int32_t ds;                      // st->downsample = 48000 / output rate (1, 2, 3, 4 or 6): spectrum cut at N/ds, every ds-th sample kept
    int32_t ring_pos;                // next write position (= oldest sample) of the global history ring
    float *ring;                     // this stream's ring: [CC][OB_RING]
};
typedef ObSynthSharedT<2> ObSynthShared;

// Once per block: lookup tables in shared memory.
template <class G, class SH>
OB_DEV void ob_synth_init(const G &g, SH &sh)
{
    for (int k = g.lane; k < 33; k += g.n) { const ObLcg p = ob_lcg_pow((uint32_t)k); sh.lcg_a[k] = p.a; sh.lcg_c[k] = p.c; }
    for (int b = g.lane; b < 104; b += g.n) {
        int band = 0;
        while (band < OB_NB && OB_EBANDS[band + 1] <= b) band++;
        sh.band_of_bin[b] = (uint8_t)band;
    }
    g.sync();
}

// anti_collapse (bands.c:268-362) on the X tile.
template <class G, class SH>
OB_DEV void ob_anti_collapse(const G &g, SH &sh, float *X, int N, uint32_t seed0)
{
    const ObFrameHdr &h = sh.hdr;
    const int LM = h.LM, C = h.C, end = h.end;
    uint32_t seed = seed0;                                           // uniform across lanes: state before the next block
    for (int i = 0; i < end; i++) {
        const int N0 = OB_EBANDS[i + 1] - OB_EBANDS[i];              // <= 22
        const int depth = (int)((uint32_t)(1 + h.pulses[i]) / (uint32_t)N0) >> LM;
        const float thresh = .5f * exp2f(-.125f * (float)depth);
        const float sqrt_1 = 1.f / OB_SQRTF((float)(N0 << LM));
        for (int c = 0; c < C; c++) {
            const int mask = h.collapse_masks[i * C + c];
            if (mask == (1 << (1 << LM)) - 1) continue;
            float prev1 = sh.oldLogE[c * OB_NB + i], prev2 = sh.oldLogE2[c * OB_NB + i];
            if (C == 1) { prev1 = fmaxf(prev1, sh.oldLogE[OB_NB + i]); prev2 = fmaxf(prev2, sh.oldLogE2[OB_NB + i]); }
            float Ediff = sh.oldBandE[c * OB_NB + i] - fminf(prev1, prev2);
            Ediff = fmaxf(0.f, Ediff);
            float r = 2.f * exp2f(-Ediff);
            if (LM == 3) r *= 1.41421356f;
            r = fminf(thresh, r);
            r = r * sqrt_1;
            float *Xb = X + c * N + (OB_EBANDS[i] << LM);
            for (int k = 0; k < 1 << LM; k++) {
                if (!(mask & (1 << k))) {
                    for (int j = g.lane; j < N0; j += g.n) {
                        const uint32_t sj = sh.lcg_a[j + 1] * seed + sh.lcg_c[j + 1];
                        Xb[(j << LM) + k] = (sj & 0x8000u) ? r : -r;
                    }
                    seed = sh.lcg_a[N0] * seed + sh.lcg_c[N0];
                }
            }
            g.sync();
            float e = 0.f;
            for (int j = g.lane; j < N0 << LM; j += g.n) e += Xb[j] * Xb[j];
            e = 1e-15f + g.sum(e);
            const float gg = 1.f / OB_SQRTF(e);
            for (int j = g.lane; j < N0 << LM; j += g.n) Xb[j] = gg * Xb[j];
            g.sync();
        }
    }
}

#include "dec_plc.cuh"

#ifdef __CUDA_ARCH__
#define OB_F2I_RN(v) __float2int_rn(v)
#else
#include <math.h>
#define OB_F2I_RN(v) ((int)lrintf(v))
#endif

// denormalise_bands (bands.c:196-265) on the X tile in sh.freq (C coded channels, energies in sh.oldBandE), the mono<->stereo
// cases of celt_synthesis (celt_decoder.c:415-441) and the inverse MDCTs into buf[c] + HISTK.
template <class G, class SH>
OB_DEV void ob_denorm_imdct(const G &g, SH &sh, int C, int CC, int N, int LM, int end, int transient, int silence)
{
    const int M = 1 << LM;
    for (int t = g.lane; t < C * OB_NB; t += g.n) {
        const int i = t % OB_NB;
        const float lg = sh.oldBandE[t] + OB_EMEANS[i];
        sh.gain[t] = (i < end && !silence) ? (float)exp(0.6931471805599453094 * (double)(lg < 32.f ? lg : 32.f)) : 0.f;
    }
    g.sync();
    const int lim = ob_div_ds(N, sh.ds);                                      // bound: bands.c:206-208
    for (int c = 0; c < C; c++)
        for (int j = g.lane; j < N; j += g.n) {
            const int bin = j >> LM, band = bin < 100 ? sh.band_of_bin[bin] : OB_NB;
            sh.freq[c][j] = (band < OB_NB && j < lim) ? sh.freq[c][j] * sh.gain[c * OB_NB + band] : 0.f;
        }
    g.sync();
    if (CC == 2 && C == 1) { for (int j = g.lane; j < N; j += g.n) sh.freq[1][j] = sh.freq[0][j]; g.sync(); }
    if (CC == 1 && C == 2) { for (int j = g.lane; j < N; j += g.n) sh.freq[0][j] = .5f * sh.freq[0][j] + .5f * sh.freq[1][j]; g.sync(); }
    for (int c = 0; c < CC; c++)
        ob_imdct(g, sh.freq[c], sh.buf[c] + OB_HISTK, transient ? 3 : 3 - LM, transient ? M : 1);
}

// de-emphasis (celt_decoder.c:249-377): y[n] = x[n] + coef*y[n-1] as a two-level scan, interleaved PCM out, then the history
// slides by N: buf[j] <- buf[j+N] for j < HISTK + overlap (celt_decoder.c:1265-1267, done after the frame instead of before).
template <class G, class SH>
OB_DEV void ob_synth_tail(const G &g, SH &sh, float *pcm, int N, int CC)
{
    {
        const float coef = OB_PREEMPH[0];
        const int per = (N + g.n - 1) / g.n;                         // samples per lane
        for (int c = 0; c < CC; c++) {
            const float *x = sh.buf[c] + OB_HISTK;
            float *y = sh.freq[c];
            const int lo = ob_imin(N, g.lane * per), hi = ob_imin(N, lo + per);
            float m = 0.f;
            for (int j = lo; j < hi; j++) { const float t = x[j] + 1e-30f + m; m = coef * t; }
            // m = coef * (local response at the chunk end); affine map of the carry: m_out = a * m_in + m
            float a = 1.f;
            for (int j = lo; j < hi; j++) a *= coef;
            // carry-in of this lane = the composition of the lanes before it applied to the stream's memory
            float ea, eb, ta, tb;
            g.affine_scan(a, m, ea, eb, ta, tb);
            const float carry = ea * sh.preemph_mem[c] + eb, last_all = ta * sh.preemph_mem[c] + tb;
            m = carry;
            for (int j = lo; j < hi; j++) { const float t = x[j] + 1e-30f + m; m = coef * t; y[j] = t * (1.f / 32768.f); }
            g.sync();
            if (g.lane == 0) sh.preemph_mem[c] = last_all;         // next read: the next frame, many barriers away
        }
        const float dg = sh.decode_gain;                             // OPUS_SET_GAIN, applied by the Opus layer (opus_decoder.c:639-649)
        const int ds = sh.ds, Nd = ob_div_ds(N, ds);                           // output rates below 48 kHz keep every ds-th sample (celt_decoder.c:326-373)
        if (CC == 1) { for (int t = g.lane; t < Nd; t += g.n) { const float v = sh.freq[0][t * ds]; pcm[t] = dg == 1.f ? v : v * dg; } }
        else { for (int t = g.lane; t < 2 * Nd; t += g.n) { const float v = sh.freq[t & 1][(t >> 1) * ds]; pcm[t] = dg == 1.f ? v : v * dg; } }
        g.sync();
    }
    for (int c = 0; c < CC; c++) {                                   // the frame joins the full-length history ring
        float *r = sh.ring + c * OB_RING;
        const float *x = sh.buf[c] + OB_HISTK;
        for (int j = g.lane; j < N; j += g.n) r[(sh.ring_pos + j) & (OB_RING - 1)] = x[j];
    }
    g.sync();
    if (g.lane == 0) sh.ring_pos = (sh.ring_pos + N) & (OB_RING - 1);
    for (int c = 0; c < CC; c++) {
        float *b = sh.buf[c];
        const int total = OB_HISTK + OB_OVERLAP;
        for (int base = 0; base < total; base += N) {               // move in blocks of N: source block lies fully ahead of dest
            const int lim = ob_imin(total, base + N);
            for (int j = base + g.lane; j < lim; j += g.n) b[j] = b[j + N];
            g.sync();
        }
    }
}

// A lost packet or DTX payload: conceal h.status samples as the reference does, frame by frame (opus_decoder.c:313-335,
// celt_decode_lost celt_decoder.c:604-968).  The integer side of the state (loss duration, skip_plc, noise seed) was stamped
// into the header by the plan pass and is re-derived here per concealment frame with the same functions.
template <class G, class SH>
OB_DEV_NOINLINE int ob_conceal(const G &g, SH &sh, float *pcm, int CC)
{
    const ObFrameHdr &h = sh.hdr;
    const int total = h.status;
    if (h.end_in == 0) {                                             // nothing decoded yet: zeros (opus_decoder.c:302-309)
        for (int t = g.lane; t < ob_div_ds(total, sh.ds) * CC; t += g.n) pcm[t] = 0.f;
        g.sync();
        return ob_div_ds(total, sh.ds);
    }
    ObPlanState p;
    p.rng = h.seed_in; p.loss_duration = h.loss_in; p.skip_plc = h.skip_in; p.plc_end = h.end_in;
    p.last_fs = (h.flags & OB_F_DTX) ? OB_SHORT << h.LM : h.lastfs_in;
    for (int done = 0; done < total;) {
        const int N = ob_plc_chunk(total - done, p.last_fs), LM = N == 960 ? 3 : N == 480 ? 2 : N == 240 ? 1 : 0;
        if (ob_plc_noise_based(p.loss_duration, p.skip_plc)) {
            if (sh.paf) ob_prefilter_and_fold(g, sh, CC);
            ob_plc_noise_fill(g, sh, N, LM, p.loss_duration, p.rng, p.plc_end, CC);
            ob_denorm_imdct(g, sh, CC, CC, N, LM, ob_imin(p.plc_end, OB_NB), 0, 0);
            g.sync();
            if (g.lane == 0) sh.paf = 0;
        } else {
            ob_plc_pitch(g, sh, N, p.loss_duration, CC);
            if (g.lane == 0) sh.paf = 1;
        }
        g.sync();
        ob_plc_advance(p, N, CC);
        ob_synth_tail(g, sh, pcm + (size_t)ob_div_ds(done, sh.ds) * CC, N, CC);
        done += N;
    }
    return ob_div_ds(total, sh.ds);
}

// opus_pcm_soft_clip (opus/src/opus.c:39-144) for ONE channel of an interleaved packet that has already been limited to +-2:
// x[i*C], i < N.  Serial by nature (zero-crossing searches, a gain carried from packet to packet); run by one lane, and only for
// packets that actually exceed +-1 or follow one that did.
OB_DEV float ob_soft_clip_channel(float *x, int N, int C, float a)
{
    int i;
    for (i = 0; i < N; i++) {                                        // continue the previous packet's non-linearity
        if (x[i * C] * a >= 0) break;
        x[i * C] = OB_MACS(x[i * C], OB_FMUL(a, x[i * C]), x[i * C]);      // x + (a*x)*x, un-fused like the reference's C
    }
    int curr = 0;
    const float x0 = x[0];
    while (1) {
        int start, end, peak_pos, special;
        float maxval;
        for (i = curr; i < N; i++) if (x[i * C] > 1 || x[i * C] < -1) break;
        if (i == N) { a = 0; break; }
        peak_pos = i;
        start = end = i;
        maxval = fabsf(x[i * C]);
        while (start > 0 && x[i * C] * x[(start - 1) * C] >= 0) start--;
        while (end < N && x[i * C] * x[end * C] >= 0) {
            if (fabsf(x[end * C]) > maxval) { maxval = fabsf(x[end * C]); peak_pos = end; }
            end++;
        }
        special = (start == 0 && x[i * C] * x[0] >= 0);
        a = (maxval - 1) / (maxval * maxval);
        a = OB_MACS(a, a, 2.4e-7f);
        if (x[i * C] > 0) a = -a;
        for (i = start; i < end; i++) x[i * C] = OB_MACS(x[i * C], OB_FMUL(a, x[i * C]), x[i * C]);
        if (special && peak_pos >= 2) {
            float offset = x0 - x[0];
            const float delta = offset / peak_pos;
            for (i = curr; i < peak_pos; i++) {
                offset -= delta;
                x[i * C] += offset;
                x[i * C] = fmaxf(-1.f, fminf(1.f, x[i * C]));
            }
        }
        curr = end;
        if (curr == N) break;
    }
    return a;
}

// The int16 API (opus_decode, opus_decoder.c:861-894): soft clip over the whole packet, then FLOAT2INT16 (round to nearest even,
// saturating).  x: the packet's float PCM (global memory, n samples per channel, written by this block); out: its int16 slot.
template <class G>
OB_DEV void ob_packet_to_int16(const G &g, float *x, int16_t *out, int n, int CC, float *softclip_mem, int soft = 1)
{
    g.sync();                                                        // the frames' PCM stores are visible to the whole block
    if (!soft) {                                                     // a concealed packet: opus_decode_native returns before its soft clip (opus_decoder.c:714-729)
        for (int t = g.lane; t < n * CC; t += g.n) {
            float v = x[t] * 32768.f;
            v = fminf(32767.f, fmaxf(-32768.f, v));
            out[t] = (int16_t)OB_F2I_RN(v);
        }
        g.sync();
        return;
    }
    float over = 0.f;
    for (int t = g.lane; t < n * CC; t += g.n) {
        const float v = fmaxf(-2.f, fminf(2.f, x[t]));
        x[t] = v;
        if (v > 1.f || v < -1.f) over = 1.f;
    }
    over = g.sum(over);
    const bool carry = softclip_mem[0] != 0.f || (CC == 2 && softclip_mem[1] != 0.f);
    g.sync();
    if (over > 0.f || carry) {
        for (int c = g.lane; c < CC; c += g.n) softclip_mem[c] = ob_soft_clip_channel(x + c, n, CC, softclip_mem[c]);
        g.sync();
    }
    for (int t = g.lane; t < n * CC; t += g.n) {
        float v = x[t] * 32768.f;
        v = fminf(32767.f, fmaxf(-32768.f, v));
        out[t] = (int16_t)OB_F2I_RN(v);
    }
    g.sync();
}

// Decodes frame `ir` (already reconstructed normalised spectrum Xg: C*N floats in global memory) into pcm
// (interleaved, CC channels) and advances the shared-memory state.  Returns samples per channel or an error.
template <class G, class SH>
OB_DEV int ob_synth_frame(const G &g, SH &sh, const ObFrameIR *ir, const float *Xg, float *pcm, int CC)
{
    // ---- header to shared memory ----
    {
        const uint32_t *src = (const uint32_t *)&ir->hdr;
        uint32_t *dst = (uint32_t *)&sh.hdr;
        for (int j = g.lane; j < (int)(sizeof(ObFrameHdr) / 4); j += g.n) dst[j] = src[j];
    }
    g.sync();
    const ObFrameHdr &h = sh.hdr;
    if (h.status <= 0) return h.status;
    if (h.flags & OB_F_LOST) return ob_conceal(g, sh, pcm, CC);
    const int LM = h.LM, M = 1 << LM, N = OB_SHORT << LM, C = h.C, end = h.end;
    const int transient = (h.flags & OB_F_TRANSIENT) != 0, silence = (h.flags & OB_F_SILENCE) != 0;
    const uint32_t seed_in = h.seed_in;

    // ---- X tile, first half: the global loads are issued here into registers and land in shared memory after the energy recurrence below,
    // which does not depend on them (the synthesis kernel's largest single stall was the wait for this tile, profiles/r02_hot_lines_synth.txt) ----
    const int bound = OB_EBANDS[end] << LM;                          // bands >= end were not written by the bands stage
#ifdef __CUDA_ARCH__
    constexpr int XR = ((int)(sizeof(sh.buf) / sizeof(sh.buf[0])) * OB_MAX_N + SH::synth_threads - 1) / SH::synth_threads;     // tile elements per thread of the synthesis block
    float xr[XR];
    const bool x_early = g.n * XR >= C * N;
    if (x_early) {
#pragma unroll
        for (int k = 0; k < XR; k++) {
            const int t = g.lane + k * g.n, j = t >= N ? t - N : t;
            xr[k] = (t < C * N && j < bound) ? Xg[t] : 0.f;
        }
    }
#endif

    // ---- energies: coarse recurrence + fine + finalise (quant_bands.c:428-542); one lane per channel ----
    if (C == 1) for (int i = g.lane; i < OB_NB; i += g.n) sh.oldBandE[i] = fmaxf(sh.oldBandE[i], sh.oldBandE[OB_NB + i]);
    g.sync();
    if (!(h.flags & OB_F_INTRA) && h.loss_in != 0) ob_post_loss_energy(g, sh, LM, end, h.loss_in);
    {
        // Only the inter-band predictor `prev` is a recurrence (quant_bands.c:473): one lane per channel runs it (two operations per band) and
        // leaves prev-before-band-i in sh.gain; the rest of the update is independent per (channel, band) and runs one per lane, with the
        // reference's operations in the reference's order.
        const int intra = (h.flags & OB_F_INTRA) != 0;
        const float coef = intra ? 0.f : OB_PRED_COEF[LM], beta = intra ? OB_BETA_INTRA[0] : OB_BETA_COEF[LM];
        for (int c = g.lane; c < C; c += g.n) {
            float prev = 0.f;
            for (int i = 0; i < end; i++) {
                const float q = (float)h.coarse_qi[c * OB_NB + i];
                sh.gain[c * OB_NB + i] = prev;
                prev = prev + q - beta * q;
            }
        }
        g.sync();
        for (int t = g.lane; t < C * OB_NB; t += g.n) {
            const int i = t >= OB_NB ? t - OB_NB : t;
            if (i >= end) continue;
            const float q = (float)h.coarse_qi[t];
            float e = fmaxf(-9.f, sh.oldBandE[t]);
            e = coef * e + sh.gain[t] + q;
            const int fq = h.fine_quant[i];
            if (fq > 0) e += ((float)h.fine_q2[t] + .5f) * (float)(1 << (14 - fq)) * (1.f / 16384) - .5f;
            const int fb = h.final_bit[t];
            if (fb >= 0) e += ((float)fb - .5f) * (float)(1 << (14 - fq - 1)) * (1.f / 16384);
            sh.oldBandE[t] = e;
        }
    }
    g.sync();

    // ---- X tile -> shared, anti-collapse ----
#ifdef __CUDA_ARCH__
    if (x_early) {
#pragma unroll
        for (int k = 0; k < XR; k++) {
            const int t = g.lane + k * g.n;
            if (t < C * N) { if (t >= N) sh.freq[1][t - N] = xr[k]; else sh.freq[0][t] = xr[k]; }
        }
    } else
#endif
    {
        for (int c = 0; c < C; c++)
            for (int j = g.lane; j < N; j += g.n) sh.freq[c][j] = j < bound ? Xg[c * N + j] : 0.f;
    }
    g.sync();
    if (h.flags & OB_F_ANTICOLLAPSE) {
        const ObLcg jump = ob_lcg_pow(h.lcg_total);
        ob_anti_collapse(g, sh, &sh.freq[0][0], OB_MAX_N, jump.a * seed_in + jump.c);   // note: channel stride is OB_MAX_N in the tile
    }
    if (silence) { for (int i = g.lane; i < C * OB_NB; i += g.n) sh.oldBandE[i] = -28.f; }
    g.sync();
    if (sh.paf) ob_prefilter_and_fold(g, sh, CC);                    // first frame after a pitch-based concealment (celt_decoder.c:1295-1297)

    // ---- denormalise + inverse MDCT into buf[c] + HISTK (celt_synthesis, celt_decoder.c:382-458) ----
    ob_denorm_imdct(g, sh, C, CC, N, LM, end, transient, silence);

    // ---- pitch post-filter (celt_decoder.c:1301-1325) ----
    const float pf_gain_new = (h.flags & OB_F_POSTFILTER) ? .09375f * (float)(h.pf_qg + 1) : 0.f;
    const int pf_pitch_new = (h.flags & OB_F_POSTFILTER) ? h.pf_pitch : 0, pf_tapset_new = (h.flags & OB_F_POSTFILTER) ? h.pf_tapset : 0;
    {
        const int p = ob_imax(sh.pf_period, 15), po = ob_imax(sh.pf_period_old, 15);
        for (int c = 0; c < CC; c++) {
            float *x = sh.buf[c] + OB_HISTK;
            ob_comb_filter(g, x, po, p, OB_SHORT, sh.pf_gain_old, sh.pf_gain, sh.pf_tapset_old, sh.pf_tapset);
            if (LM != 0) ob_comb_filter(g, x + OB_SHORT, p, pf_pitch_new, N - OB_SHORT, sh.pf_gain, pf_gain_new, sh.pf_tapset, pf_tapset_new);
        }
        g.sync();
        if (g.lane == 0) {
            sh.pf_period_old = p; sh.pf_gain_old = sh.pf_gain; sh.pf_tapset_old = sh.pf_tapset;
            sh.pf_period = pf_pitch_new; sh.pf_gain = pf_gain_new; sh.pf_tapset = pf_tapset_new;
            if (LM != 0) { sh.pf_period_old = sh.pf_period; sh.pf_gain_old = sh.pf_gain; sh.pf_tapset_old = sh.pf_tapset; }
            sh.paf = 0;
        }
    }

    // ---- energy history (celt_decoder.c:1327-1357) ----
    g.sync();
    if (C == 1) for (int i = g.lane; i < OB_NB; i += g.n) sh.oldBandE[OB_NB + i] = sh.oldBandE[i];
    g.sync();
    for (int i = g.lane; i < 2 * OB_NB; i += g.n) {
        const float e = sh.oldBandE[i];
        if (!transient) { sh.oldLogE2[i] = sh.oldLogE[i]; sh.oldLogE[i] = e; }
        else sh.oldLogE[i] = fminf(sh.oldLogE[i], e);
        sh.backgroundLogE[i] = fminf(sh.backgroundLogE[i] + (float)ob_imin(160, h.loss_in + M) * 0.001f, e);
        if ((i % OB_NB) >= end) { sh.oldBandE[i] = 0.f; sh.oldLogE[i] = sh.oldLogE2[i] = -28.f; }
    }
    g.sync();

    ob_synth_tail(g, sh, pcm, N, CC);
    return ob_div_ds(N, sh.ds);
}
