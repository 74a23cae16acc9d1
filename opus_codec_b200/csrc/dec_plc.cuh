// dec_plc.cuh -- packet-loss concealment for lost packets and DTX payloads, cooperatively by the synthesis block of the stream:
//   the loss state machine (celt_decoder.c:604-968 bookkeeping; opus_decoder.c:284-334, :715-729 for the frame sizes),
//   noise-based PLC / CNG (celt_decoder.c:648-699), pitch-based PLC (:700-905) with celt_plc_pitch_search (:499-513),
//   _celt_autocorr + _celt_lpc (celt_lpc.c:277-351, :37-91), celt_fir / celt_iir (celt_lpc.c:93-275),
//   prefilter_and_fold (:515-550) and the post-loss energy safety of the next good frame (:1171-1198).
// Included by dec_synth.cuh after ObSynthShared is defined.
#pragma once
// In-place FIR passes of the concealment compute every output of a lane's contiguous chunk into registers before any is written back: 1024 samples over
// the synthesis block's lanes, so a chunk is <= 16 samples for blocks of >= 64 lanes (dec_synth.cuh ObSynthSharedT::synth_threads).
#define OB_PLC_PER 16

// Products that feed a decision (pitch lag, LPC) are kept un-fused so that they round like the reference's C code.
#ifdef __CUDA_ARCH__
#define OB_MACS(s, a, b) __fadd_rn((s), __fmul_rn((a), (b)))
#define OB_FMUL(a, b) __fmul_rn((a), (b))
#define OB_FSUB(a, b) __fsub_rn((a), (b))
#else
#define OB_MACS(s, a, b) ((s) + (a) * (b))
#define OB_FMUL(a, b) ((a) * (b))
#define OB_FSUB(a, b) ((a) - (b))
#endif

#define OB_PLC_LAG_MAX 720                   // PLC_PITCH_LAG_MAX (celt_decoder.c:58)
#define OB_PLC_LAG_MIN 100                   // PLC_PITCH_LAG_MIN (:61)
#define OB_LPC_ORDER 24                      // CELT_LPC_ORDER (celt_lpc.h:39)
#define OB_MAX_PERIOD 1024                   // MAX_PERIOD (celt_decoder.c:56 via modes.h)

template <class G> struct ObIsSolo { static constexpr bool value = false; };
template <> struct ObIsSolo<ObSolo> { static constexpr bool value = true; };
// analysis window of _celt_autocorr over n samples: the MDCT window's rising half at both ends (celt_lpc.c:299-307)
OB_DEV float ob_ac_window(int i, int n) { return i < OB_OVERLAP ? OB_WINDOW[i] : (i >= n - OB_OVERLAP ? OB_WINDOW[n - 1 - i] : 1.f); }

// ---- the loss state machine: integer only, shared by the plan pass (one thread per stream) and the synthesis block ----------
struct ObPlanState {
    uint32_t rng;              // st->rng
    int32_t loss_duration;     // st->loss_duration, 2.5 ms units, saturates at 10000
    int32_t skip_plc;          // st->skip_plc: noise PLC until two consecutive packets have arrived
    int32_t plc_end;           // st->end as left by the last decoded packet; 0 = nothing decoded yet (OpusDecoder.prev_mode == 0)
    int32_t last_fs;           // OpusDecoder.frame_size: the frame size of the last packet that arrived (120 after a reset)
};

// Size of the next concealment frame when `remaining` samples are still to be produced: never more than the last packet's frame
// size (opus_decoder.c:288-289), then the nearest CELT frame size below (opus_decoder.c:313-335).
OB_DEV int ob_plc_chunk(int remaining, int last_fs)
{
    remaining = ob_imin(remaining, last_fs);
    if (remaining >= 960) return 960;
    if (remaining > 480) return 480;
    if (remaining > 240 && remaining < 480) return 240;
    return remaining;
}
OB_DEV int ob_plc_noise_based(int loss_duration, int skip_plc) { return loss_duration >= 40 || skip_plc; }        // celt_decoder.c:646, start == 0
// LCG steps one noise-PLC frame takes (celt_decoder.c:676-692): every coded bin of every output channel.
OB_DEV uint32_t ob_plc_noise_steps(int end, int LM, int CC) { return (uint32_t)(CC * (OB_EBANDS[ob_imin(end, OB_NB)] << LM)); }
OB_DEV void ob_plc_advance(ObPlanState &p, int n, int CC)       // bookkeeping of ONE concealment frame of n samples
{
    const int LM = n == 960 ? 3 : n == 480 ? 2 : n == 240 ? 1 : 0;
    if (ob_plc_noise_based(p.loss_duration, p.skip_plc)) {
        const ObLcg j = ob_lcg_pow(ob_plc_noise_steps(p.plc_end, LM, CC));
        p.rng = j.a * p.rng + j.c;
        p.skip_plc = 1;
    }
    p.loss_duration = ob_imin(10000, p.loss_duration + (1 << LM));
}
// Advances the state past one frame slot (status / flags / final_range / end as the symbol kernel left them in the header).
OB_DEV void ob_plan_step(ObPlanState &p, int status, int flags, uint32_t final_range, int end, int CC, int LM)
{
    if (status <= 0) return;                                     // a failed frame leaves the stream untouched
    if (!(flags & OB_F_LOST) || (flags & OB_F_DTX)) p.last_fs = OB_SHORT << LM;      // st->frame_size = packet_frame_size (opus_decoder.c:778)
    if (flags & OB_F_LOST) {
        if (p.plc_end == 0) return;                              // nothing decoded yet: zeros, no state change (opus_decoder.c:302-309)
        for (int rem = status; rem > 0;) { const int n = ob_plc_chunk(rem, p.last_fs); ob_plc_advance(p, n, CC); rem -= n; }
    } else {
        if (p.loss_duration == 0) p.skip_plc = 0;                // celt_decoder.c:1103
        p.rng = final_range; p.loss_duration = 0; p.plc_end = end;
    }
}
// Stamps frame header h with the state it starts from and advances the state past it.
OB_DEV void ob_plan_frame(ObPlanState &p, ObFrameHdr &h, int CC)
{
    if (h.status > 0) { h.seed_in = p.rng; h.loss_in = p.loss_duration; h.skip_in = (uint8_t)p.skip_plc; h.end_in = (uint8_t)p.plc_end; h.lastfs_in = (uint16_t)p.last_fs; }
    ob_plan_step(p, h.status, h.flags, h.final_range, h.end, CC, h.LM);
}

// ---- prefilter_and_fold (celt_decoder.c:515-550): undo the post-filter on the concealed overlap and fold it like the TDAC would ----
template <class G, class SH>
OB_DEV void ob_prefilter_and_fold(const G &g, SH &sh, int CC)
{
    const float gains[3][3] = {{0.3066406250f, 0.2170410156f, 0.1296386719f}, {0.4638671875f, 0.2680664062f, 0.f}, {0.7998046875f, 0.1000976562f, 0.f}};
    const int T1 = ob_imax(sh.pf_period, 15), ts = sh.pf_tapset;
    const float g1 = -sh.pf_gain, g0 = -sh.pf_gain_old;
    float *etmp = sh.scanA;                                      // 120 floats
    for (int c = 0; c < CC; c++) {
        float *x = sh.buf[c] + OB_HISTK;
        // comb_filter(etmp, x, T0, T1, overlap, g0, g1, tapset0, tapset1, window = NULL, overlap = 0): no cross-fade, constant filter
        for (int i = g.lane; i < OB_OVERLAP; i += g.n) {
            float y = x[i];
            if (!(g0 == 0 && g1 == 0) && g1 != 0)
                y = x[i] + (g1 * gains[ts][0]) * x[i - T1] + (g1 * gains[ts][1]) * (x[i - T1 + 1] + x[i - T1 - 1]) + (g1 * gains[ts][2]) * (x[i - T1 + 2] + x[i - T1 - 2]);
            etmp[i] = y;
        }
        g.sync();
        for (int i = g.lane; i < OB_OVERLAP / 2; i += g.n)
            x[i] = OB_WINDOW[i] * etmp[OB_OVERLAP - 1 - i] + OB_WINDOW[OB_OVERLAP - i - 1] * etmp[i];
        g.sync();
    }
}

// ---- noise-based PLC / comfort noise (celt_decoder.c:648-699).  Leaves the denormalised spectrum source in sh.freq (normalised X)
// and the decayed energies in sh.oldBandE; the caller runs the common denormalise + IMDCT + de-emphasis. ----
template <class G, class SH>
OB_DEV void ob_plc_noise_fill(const G &g, SH &sh, int N, int LM, int loss_duration, uint32_t seed0, int end, int CC)
{
    const int effEnd = ob_imin(end, OB_NB);
    const float decay = loss_duration == 0 ? 1.5f : .5f;
    for (int t = g.lane; t < CC * OB_NB; t += g.n) {
        const int i = t % OB_NB;
        if (i < end) sh.oldBandE[t] = fmaxf(sh.backgroundLogE[t], sh.oldBandE[t] - decay);
    }
    const int W = OB_EBANDS[effEnd] << LM, T = CC * W;
    const int per = (T + g.n - 1) / g.n;
    {
        const int e0 = ob_imin(T, g.lane * per), e1 = ob_imin(T, e0 + per);
        const ObLcg j = ob_lcg_pow((uint32_t)e0);
        uint32_t seed = j.a * seed0 + j.c;
        for (int e = e0; e < e1; e++) {
            seed = 1664525u * seed + 1013904223u;
            const int c = e >= W, k = e - c * W;
            sh.freq[c][k] = (float)((int32_t)seed >> 20);
        }
    }
    g.sync();
    for (int t = g.lane; t < CC * effEnd; t += g.n) {           // renormalise_vector per band, Q15ONE gain (vq.c:383-407)
        const int c = t / effEnd, i = t - c * effEnd;
        float *X = sh.freq[c] + (OB_EBANDS[i] << LM);
        const int blen = (OB_EBANDS[i + 1] - OB_EBANDS[i]) << LM;
        float E = 1e-15f;
        for (int j = 0; j < blen; j++) E += X[j] * X[j];
        const float gg = 1.f / OB_SQRTF(E);
        for (int j = 0; j < blen; j++) X[j] = gg * X[j];
    }
    for (int c = 0; c < CC; c++) for (int k = W + g.lane; k < N; k += g.n) sh.freq[c][k] = 0.f;
    g.sync();
}

// ---- pitch search on the decoder history (celt_plc_pitch_search: pitch_downsample + pitch_search, pitch.c:140-217, :302-411) ----
// work: >= 1024 + 332 + 487 floats; xc: >= 310 floats.  Every lag's correlation is one lane's in-order sum.
template <class G, class SH>
OB_DEV int ob_plc_pitch_search(const G &g, SH &sh, int CC, float *work, float *xc)
{
    const int n = OB_RING >> 1;                                   // 1024
    float *lp = work;
    // the only reader of the full 2048-sample history: straight from the global ring (oldest sample at ring_pos)
    const float *r0 = sh.ring, *r1 = sh.ring + OB_RING;
    const int pos = sh.ring_pos;
#define OB_RG(r, i) (r)[(pos + (i)) & (OB_RING - 1)]
    for (int i = g.lane; i < n; i += g.n) {
        float v = i ? .25f * OB_RG(r0, 2 * i - 1) + .25f * OB_RG(r0, 2 * i + 1) + .5f * OB_RG(r0, 2 * i) : .25f * OB_RG(r0, 1) + .5f * OB_RG(r0, 0);
        if (CC == 2) v += i ? .25f * OB_RG(r1, 2 * i - 1) + .25f * OB_RG(r1, 2 * i + 1) + .5f * OB_RG(r1, 2 * i) : .25f * OB_RG(r1, 1) + .5f * OB_RG(r1, 0);
        lp[i] = v;
    }
#undef OB_RG
    g.sync();
    float *ac = sh.red;                                           // 5 values; g.sum() is not used below
    for (int k = g.lane; k <= 4; k += g.n) {                      // _celt_autocorr, lag 4, no window (celt_lpc.c:277-351)
        const int fastN = n - 4;
        float s = 0, d = 0;
        for (int i = 0; i < fastN; i++) s = OB_MACS(s, lp[i], lp[i + k]);
        for (int i = k + fastN; i < n; i++) d = OB_MACS(d, lp[i], lp[i - k]);
        ac[k] = s + d;
    }
    g.sync();
    float num[5];
    {                                                             // every lane: 4th-order LPC (uniform, tiny), as pitch.c:176-203
        float a[5], lpc[4] = {0, 0, 0, 0};
        for (int k = 0; k <= 4; k++) a[k] = ac[k];
        a[0] *= 1.0001f;
        for (int i = 1; i <= 4; i++) a[i] = OB_FSUB(a[i], OB_FMUL(OB_FMUL(a[i], .008f * i), .008f * i));
        float error = a[0];
        if (a[0] > 1e-10f) {
            for (int i = 0; i < 4; i++) {
                float rr = 0;
                for (int j = 0; j < i; j++) rr = OB_MACS(rr, lpc[j], a[i - j]);
                rr += a[i + 1];
                const float r = -(rr / error);
                lpc[i] = r;
                for (int j = 0; j < (i + 1) >> 1; j++) {
                    const float t1 = lpc[j], t2 = lpc[i - 1 - j];
                    lpc[j] = OB_MACS(t1, r, t2);
                    lpc[i - 1 - j] = OB_MACS(t2, r, t1);
                }
                error = OB_FSUB(error, OB_FMUL(OB_FMUL(r, r), error));
                if (error <= .001f * a[0]) break;
            }
        }
        float tmp = 1.0f;
        for (int i = 0; i < 4; i++) { tmp = OB_FMUL(.9f, tmp); lpc[i] = OB_FMUL(lpc[i], tmp); }
        const float c1 = .8f;
        num[0] = lpc[0] + .8f; num[1] = OB_MACS(lpc[1], c1, lpc[0]); num[2] = OB_MACS(lpc[2], c1, lpc[1]); num[3] = OB_MACS(lpc[3], c1, lpc[2]); num[4] = c1 * lpc[3];
    }
    g.sync();
    {                                                             // celt_fir5 (pitch.c:105-137): FIR on the ORIGINAL samples, so every output is independent
        const int per = (n + g.n - 1) / g.n;                      // <= OB_PLC_PER (blocks of >= 64 lanes); the host lane does it in order instead
        if constexpr (ObIsSolo<G>::value) {
            float m0 = 0, m1 = 0, m2 = 0, m3 = 0, m4 = 0;
            for (int i = 0; i < n; i++) {
                float sum = lp[i];
                sum = OB_MACS(sum, num[0], m0); sum = OB_MACS(sum, num[1], m1); sum = OB_MACS(sum, num[2], m2); sum = OB_MACS(sum, num[3], m3); sum = OB_MACS(sum, num[4], m4);
                m4 = m3; m3 = m2; m2 = m1; m1 = m0; m0 = lp[i];
                lp[i] = sum;
            }
        } else {
            float out[OB_PLC_PER];
            for (int k = 0; k < per && k < OB_PLC_PER; k++) {
                const int i = g.lane * per + k;
                if (i < n) {
                    float sum = lp[i];
                    for (int t = 0; t < 5; t++) sum = OB_MACS(sum, num[t], i - 1 - t >= 0 ? lp[i - 1 - t] : 0.f);
                    out[k] = sum;
                }
            }
            g.sync();
            for (int k = 0; k < per && k < OB_PLC_PER; k++) { const int i = g.lane * per + k; if (i < n) lp[i] = out[k]; }
        }
    }
    g.sync();
    // pitch_search(x_lp = lp + 360, y = lp, len = 1328, max_pitch = 620)
    const int len = OB_RING - OB_PLC_LAG_MAX, max_pitch = OB_PLC_LAG_MAX - OB_PLC_LAG_MIN, lag = len + max_pitch;
    const float *x_lp = lp + (OB_PLC_LAG_MAX >> 1), *y = lp;
    float *x4 = work + n, *y4 = x4 + (len >> 2);
    for (int j = g.lane; j < len >> 2; j += g.n) x4[j] = x_lp[2 * j];
    for (int j = g.lane; j < lag >> 2; j += g.n) y4[j] = y[2 * j];
    g.sync();
    for (int i = g.lane; i < max_pitch >> 2; i += g.n) {
        float s = 0;
        for (int j = 0; j < len >> 2; j++) s = OB_MACS(s, x4[j], y4[i + j]);
        xc[i] = s;
    }
    g.sync();
    int32_t *best = (int32_t *)(sh.red + 8);                      // best[0], best[1], result
    // find_best_pitch (pitch.c:45-103) keeps a running energy: in order, by one lane
    for (int pass = 0; pass < 2; pass++) {
        if (g.lane == 0) {
            const float *yy = pass ? y : y4;
            const int L = pass ? len >> 1 : len >> 2, P = pass ? max_pitch >> 1 : max_pitch >> 2;
            float Syy = 1, bn0 = -1, bn1 = -1, bd0 = 0, bd1 = 0;
            int b0 = 0, b1 = 1;
            for (int j = 0; j < L; j++) Syy = OB_MACS(Syy, yy[j], yy[j]);
            for (int i = 0; i < P; i++) {
                if (xc[i] > 0) {
                    float x16 = xc[i];
                    x16 *= 1e-12f;
                    const float nm = OB_FMUL(x16, x16);
                    if (OB_FMUL(nm, bd1) > OB_FMUL(bn1, Syy)) {
                        if (OB_FMUL(nm, bd0) > OB_FMUL(bn0, Syy)) { bn1 = bn0; bd1 = bd0; b1 = b0; bn0 = nm; bd0 = Syy; b0 = i; }
                        else { bn1 = nm; bd1 = Syy; b1 = i; }
                    }
                }
                Syy += OB_FSUB(OB_FMUL(yy[i + L], yy[i + L]), OB_FMUL(yy[i], yy[i]));
                Syy = fmaxf(1, Syy);
            }
            best[0] = b0; best[1] = b1;
        }
        g.sync();
        if (pass == 0) {
            const int b0 = best[0], b1 = best[1];
            g.sync();
            for (int i = g.lane; i < max_pitch >> 1; i += g.n) {
                float v = 0;
                int d0 = i - 2 * b0, d1 = i - 2 * b1;
                if (d0 < 0) d0 = -d0;
                if (d1 < 0) d1 = -d1;
                if (!(d0 > 2 && d1 > 2)) {
                    float s = 0;
                    for (int j = 0; j < len >> 1; j++) s = OB_MACS(s, x_lp[j], y[i + j]);
                    v = fmaxf(-1, s);
                }
                xc[i] = v;
            }
            g.sync();
        }
    }
    if (g.lane == 0) {
        const int b0 = best[0];
        int offset = 0;
        if (b0 > 0 && b0 < (max_pitch >> 1) - 1) {
            const float a = xc[b0 - 1], b = xc[b0], c = xc[b0 + 1];
            if ((c - a) > OB_FMUL(.7f, b - a)) offset = 1;
            else if ((a - c) > OB_FMUL(.7f, b - c)) offset = -1;
        }
        best[2] = OB_PLC_LAG_MAX - (2 * b0 - offset);
    }
    g.sync();
    const int r = best[2];
    g.sync();
    return r;
}

// ---- pitch-based PLC (celt_decoder.c:700-905): writes N + overlap concealed samples at buf[c] + HISTK ----
template <class G, class SH>
OB_DEV void ob_plc_pitch(const G &g, SH &sh, int N, int loss_duration, int CC)
{
    float *work = &sh.freq[0][0];                                  // 1920 floats, free while no spectrum is in flight
    float fade = 1.f;
    if (loss_duration == 0) {
        const int p = ob_plc_pitch_search(g, sh, CC, work, sh.scanA);
        if (g.lane == 0) sh.last_pitch_index = p;
        g.sync();
    } else fade = .8f;
    const int pitch_index = sh.last_pitch_index;
    const int exc_length = ob_imin(2 * pitch_index, OB_MAX_PERIOD);
    const int ext_len = N + OB_OVERLAP;
    float *E = work, *exc = work + OB_LPC_ORDER;                   // exc[-24 .. 1024)
    int32_t *flag = (int32_t *)(sh.red + 24);
    for (int c = 0; c < CC; c++) {
        float *buf = sh.buf[c];
        float *lpc = sh.lpc[c];
        for (int i = g.lane; i < OB_MAX_PERIOD + OB_LPC_ORDER; i += g.n) E[i] = buf[OB_HISTK - OB_MAX_PERIOD - OB_LPC_ORDER + i];
        g.sync();
        if (loss_duration == 0) {
            // _celt_autocorr(exc, ac, window, overlap, 24, 1024) with the window applied on the fly, then lag windowing and _celt_lpc
            float *ac = sh.scanB;                                 // 25 values
            const int n = OB_MAX_PERIOD, fastN = n - OB_LPC_ORDER;
            for (int k = g.lane; k <= OB_LPC_ORDER; k += g.n) {
                float s = 0, d = 0;
                for (int i = 0; i < fastN; i++) s = OB_MACS(s, exc[i] * ob_ac_window(i, n), exc[i + k] * ob_ac_window(i + k, n));
                for (int i = k + fastN; i < n; i++) d = OB_MACS(d, exc[i] * ob_ac_window(i, n), exc[i - k] * ob_ac_window(i - k, n));
                ac[k] = s + d;
            }
            g.sync();
            if (g.lane == 0) {
                ac[0] *= 1.0001f;
                for (int i = 1; i <= OB_LPC_ORDER; i++) ac[i] = OB_FSUB(ac[i], OB_FMUL(OB_FMUL(OB_FMUL(ac[i], 0.008f * 0.008f), (float)i), (float)i));
                // _celt_lpc (celt_lpc.c:37-91)
                float error = ac[0];
                for (int i = 0; i < OB_LPC_ORDER; i++) lpc[i] = 0;
                if (ac[0] > 1e-10f) {
                    for (int i = 0; i < OB_LPC_ORDER; i++) {
                        float rr = 0;
                        for (int j = 0; j < i; j++) rr = OB_MACS(rr, lpc[j], ac[i - j]);
                        rr += ac[i + 1];
                        const float r = -(rr / error);
                        lpc[i] = r;
                        for (int j = 0; j < (i + 1) >> 1; j++) {
                            const float t1 = lpc[j], t2 = lpc[i - 1 - j];
                            lpc[j] = OB_MACS(t1, r, t2);
                            lpc[i - 1 - j] = OB_MACS(t2, r, t1);
                        }
                        error = OB_FSUB(error, OB_FMUL(OB_FMUL(r, r), error));
                        if (error <= .001f * ac[0]) break;
                    }
                }
            }
            g.sync();
        }
        // celt_fir(exc + 1024 - exc_length, lpc, ., exc_length, 24): excitation of the last exc_length samples, in place via registers
        {
            float *x = exc + OB_MAX_PERIOD - exc_length;
            const int per = (exc_length + g.n - 1) / g.n;
            if constexpr (ObIsSolo<G>::value) {                    // host emulation: one lane, through a local copy
                float loc[OB_MAX_PERIOD];
                for (int i = 0; i < exc_length; i++) {
                    float sum = x[i];
                    for (int j = 0; j < OB_LPC_ORDER; j++) sum = OB_MACS(sum, lpc[OB_LPC_ORDER - 1 - j], x[i - OB_LPC_ORDER + j]);
                    loc[i] = sum;
                }
                for (int i = 0; i < exc_length; i++) x[i] = loc[i];
            } else {
                float out[OB_PLC_PER];
                for (int k = 0; k < per && k < OB_PLC_PER; k++) {
                    const int i = g.lane * per + k;
                    if (i < exc_length) {
                        float sum = x[i];
                        for (int j = 0; j < OB_LPC_ORDER; j++) sum = OB_MACS(sum, lpc[OB_LPC_ORDER - 1 - j], x[i - OB_LPC_ORDER + j]);
                        out[k] = sum;
                    }
                }
                g.sync();
                for (int k = 0; k < per && k < OB_PLC_PER; k++) { const int i = g.lane * per + k; if (i < exc_length) x[i] = out[k]; }
            }
        }
        g.sync();
        // decay of the excitation energy over the last two half-windows (celt_decoder.c:789-808)
        float decay;
        {
            const int dl = exc_length >> 1;
            float *e12 = sh.scanB + 32;
            for (int h = g.lane; h < 2; h += g.n) {
                float Es = 1;
                const float *p = exc + OB_MAX_PERIOD - (h + 1) * dl;
                for (int i = 0; i < dl; i++) Es = OB_MACS(Es, p[i], p[i]);
                e12[h] = Es;
            }
            g.sync();
            const float E1 = fminf(e12[0], e12[1]), E2 = e12[1];
            decay = OB_SQRTF(E1 / E2);
        }
        // extrapolation with period pitch_index, each period scaled by a further `decay` (celt_decoder.c:815-838)
        float S1p = 0;
        for (int i = g.lane; i < ext_len; i += g.n) {
            const int q = i / pitch_index, j = i - q * pitch_index;
            float att = fade * decay;
            for (int t = 0; t < q; t++) att = att * decay;
            buf[OB_HISTK + i] = att * exc[OB_MAX_PERIOD - pitch_index + j];
            const float tmp = buf[OB_HISTK - pitch_index + j];
            S1p += tmp * tmp;
        }
        const float S1 = g.sum(S1p);
        g.sync();
        // celt_iir (celt_lpc.c:143-275) over the concealed samples, continuing from the last 24 decoded ones; in place, one lane
        if (g.lane == 0) {
            float mem[OB_LPC_ORDER];
            for (int i = 0; i < OB_LPC_ORDER; i++) mem[i] = buf[OB_HISTK - 1 - i];
            float *y = buf + OB_HISTK;
            // Term order of the reference's 4-sample blocks (celt_lpc.c:224-262): for the k-th sample of a block the lags
            // ord .. k+1 are summed oldest first by xcorr_kernel, then the lags 1 .. k are patched in newest first.
            for (int i = 0; i < ext_len; i++) {
                const int k = i & 3;
                float sum = y[i];
                for (int L = OB_LPC_ORDER; L > k; L--) sum = OB_MACS(sum, lpc[L - 1], -mem[L - 1]);
                for (int L = 1; L <= k; L++) sum = OB_MACS(sum, -mem[L - 1], lpc[L - 1]);
                for (int j = OB_LPC_ORDER - 1; j >= 1; j--) mem[j] = mem[j - 1];
                mem[0] = sum;
                y[i] = sum;
            }
        }
        g.sync();
        // energy check of the synthesis (celt_decoder.c:858-895)
        float S2p = 0;
        for (int i = g.lane; i < ext_len; i += g.n) { const float t = buf[OB_HISTK + i]; S2p += t * t; }
        const float S2 = g.sum(S2p);
        if (g.lane == 0) flag[0] = !(S1 > 0.2f * S2) ? 1 : (S1 < S2 ? 2 : 0);
        g.sync();
        const int what = flag[0];
        if (what == 1) {
            for (int i = g.lane; i < ext_len; i += g.n) buf[OB_HISTK + i] = 0;
        } else if (what == 2) {
            const float ratio = OB_SQRTF((S1 + 1) / (S2 + 1));
            for (int i = g.lane; i < ext_len; i += g.n) {
                const float tg = i < OB_OVERLAP ? 1.f - OB_WINDOW[i] * (1.f - ratio) : ratio;
                buf[OB_HISTK + i] = tg * buf[OB_HISTK + i];
            }
        }
        g.sync();
    }
}

// ---- energy safety of the first good frame after a loss (celt_decoder.c:1171-1198); acts on both channels' energies ----
template <class G, class SH>
OB_DEV void ob_post_loss_energy(const G &g, SH &sh, int LM, int end, int loss_duration)
{
    const int missing = ob_imin(10, loss_duration >> LM);
    const float safety = LM == 0 ? 1.5f : LM == 1 ? .5f : 0.f;
    for (int t = g.lane; t < 2 * OB_NB; t += g.n) {
        const int i = t % OB_NB;
        if (i >= end) continue;
        float E0 = sh.oldBandE[t];
        const float E1 = sh.oldLogE[t], E2 = sh.oldLogE2[t];
        if (E0 < fmaxf(E1, E2)) {
            const float slope = fmaxf(E1 - E0, .5f * (E2 - E0));
            E0 -= fmaxf(0.f, (float)(1 + missing) * slope);
            E0 = fmaxf(-20.f, E0);
        } else E0 = fminf(fminf(E0, E1), E2);
        sh.oldBandE[t] = E0 - safety;
    }
    g.sync();
}
