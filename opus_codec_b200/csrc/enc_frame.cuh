// enc_frame.cuh -- per-stream CELT encoder state, the warp's working set, and the frame driver: celt_encode_with_ec
// (opus/celt/celt_encoder.c:1431-2368) for start = 0, no LFE, no surround mask, plus the thin Opus layer around it for CELT-only
// packets (opus/src/opus_encoder.c: dc_reject :430-468, byte budget :1188-1197, TOC gen_toc :299-329).
//
// ONE WARP PER STREAM.  Memory plan per warp:
//   shared (ObEncShared, ~18 KB): everything touched element by element or out of order -- the FFT / pitch / transient scratch, the band
//     being quantised with its folding source, the coder's output bytes, and the per-band tables of the frame;
//   global, per resident warp (ObEncWork, ~48 KB, stays in L2): the long time / frequency vectors that are only streamed with
//     lane-strided (coalesced) loops -- [history | frame] of the pre-filter, the MDCT input, the spectrum, the normalised spectrum;
//   global, per stream (ObEncStream + ObEncHist): state between calls, read at the start and written at the end of a launch.
#pragma once
#include "enc_quant.cuh"
#include "enc_tonal.cuh"
#include "repacketizer.cuh"

#define OB_BITRATE_MAX (-1)

// struct OpusCustomEncoder (celt_encoder.c:58-128), the scalar part, + the Opus layer's dc_reject memory.  Every lane holds a copy.
struct ObEncState {
    // configuration (CTLs)
    int32_t channels, stream_channels, complexity, bitrate, vbr, constrained_vbr, lsb_depth, end, force_intra, loss_rate, disable_inv, clip, disable_pf;
    // everything below is cleared by OPUS_RESET_STATE (celt_encoder.c:2552-2571)
    uint32_t rng;
    int32_t spread_decision;
    float delayedIntra;
    int32_t tonal_average, lastCodedBands, hf_average, tapset_decision, prefilter_period;
    float prefilter_gain;
    int32_t prefilter_tapset, consec_transient;
    float preemph_memE[2];
    int32_t vbr_reservoir, vbr_drift, vbr_offset, vbr_count;
    float overlap_max, stereo_saving;
    int32_t intensity;
    float spec_avg;
    float hp_mem[4];                                  // Opus layer: dc_reject (opus_encoder.c:430-468)
    uint32_t final_range;
    ObAnalysisInfo an;                                // st->analysis: set by the Opus layer before every frame (CELT_SET_ANALYSIS)
};
// ... and its vectors (in_mem, prefilter_mem, oldBandE, oldLogE, oldLogE2, energyError): per stream in HBM between launches
struct ObEncHist {
    float in_mem[2 * OB_OVERLAP];
    float prefilter_mem[2 * OB_MAXPERIOD];
    float oldBandE[2 * OB_NB], oldLogE[2 * OB_NB], oldLogE2[2 * OB_NB], energyError[2 * OB_NB];
};

#define OB_ENC_DELAY 192                                  // delay_compensation = Fs/250 (opus_encoder.c:282)
#define OB_ENC_BUFFER 480                                 // encoder_buffer = Fs/100 (:276)
#define OB_ENC_SCR 1280

struct ObEncShared {
    float A[OB_ENC_SCR];                              // pitch-search scratch / transient high-pass output / FFT work / small per-band scratch
    union {
        struct { float pitch_buf[(OB_MAXPERIOD + OB_MAX_N) >> 1]; } t;                                                  // time-domain phase
        ObEncBandsShared bands;                                                                                          // quantisation phase
    } u;
    float oldBandE[2 * OB_NB], oldLogE[2 * OB_NB], oldLogE2[2 * OB_NB], energyError[2 * OB_NB];
    float bandE[2 * OB_NB], bandLogE[2 * OB_NB], bandLogE2[2 * OB_NB], error[2 * OB_NB];
    int fine_quant[OB_NB], pulses[OB_NB], cap[OB_NB], offsets[OB_NB], importance[OB_NB], spread_weight[OB_NB], fine_priority[OB_NB], tf_res[OB_NB];
    ObAnalysisInfo an_tmp;                            // lane 0 -> warp hand-over of an inline analysis result
    uint8_t collapse_masks[2 * OB_NB + 2];
    uint8_t bytes[1280];                              // the range coder's buffer (payload without the TOC)
};
struct ObEncWork {
    // Buffers whose lifetimes do not overlap share storage (48 KB instead of 73 KB per resident warp: 148 x 16 slots then fit the 126 MB L2):
    union {
        float pcm_hp[2 * (OB_MAX_N + OB_ENC_DELAY)];  // [delay compensation | dc_reject / hp_cutoff output]: what CELT encodes; dead after pre-emphasis
        float pre[2 * (OB_MAX_N + OB_MAXPERIOD)];     // [pre-filter memory | frame] per channel; built after pre-emphasis, dead after the comb filter
        float freq[2 * OB_MAX_N];                     // MDCT output, written after the pre-filter and the transient analysis
    };
    union {
        float in[2 * (OB_MAX_N + OB_OVERLAP)];        // [overlap | pre-filtered frame] per channel: the MDCT's input
        float X[2 * OB_MAX_N];                        // normalised spectrum: written from freq after the last MDCT of the frame
    };
    float env[(OB_MAX_N + OB_OVERLAP + 7) >> 1];      // transient_analysis: the masking envelope (chunk-per-lane scans)
    ObEncHist hist;                                   // in_mem / prefilter_mem of the stream being coded (the four energy arrays live in shared memory)
    ObEncBandsWork bw;
    ObRepack rp;                                      // opus_packet_pad of a 'PLC frame'
    uint8_t multi_tmp[1284];                          // the 20 ms frames of a 40-120 ms packet before they are repacketized
};

OB_DEV void ob_enc_reset(ObEncState &st)
{
    st.rng = 0; st.spread_decision = 2; st.delayedIntra = 1; st.tonal_average = 256; st.lastCodedBands = 0; st.hf_average = 0;
    st.tapset_decision = 0; st.prefilter_period = 0; st.prefilter_gain = 0; st.prefilter_tapset = 0; st.consec_transient = 0;
    st.preemph_memE[0] = st.preemph_memE[1] = 0; st.vbr_reservoir = st.vbr_drift = st.vbr_offset = st.vbr_count = 0;
    st.overlap_max = 0; st.stereo_saving = 0; st.intensity = 0; st.spec_avg = 0; st.final_range = 0; st.an.valid = 0;
    for (int i = 0; i < 4; i++) st.hp_mem[i] = 0;
}
OB_DEV void ob_enc_hist_reset(ObEncHist &h)
{
    for (int i = 0; i < 2 * OB_OVERLAP; i++) h.in_mem[i] = 0;
    for (int i = 0; i < 2 * OB_MAXPERIOD; i++) h.prefilter_mem[i] = 0;
    for (int i = 0; i < 2 * OB_NB; i++) { h.oldBandE[i] = 0; h.oldLogE[i] = h.oldLogE2[i] = -28.f; h.energyError[i] = 0; }
}
// stream state <-> the warp's working set, at the two ends of a launch
template <class G>
OB_DEV void ob_enc_load_hist(const G &g, ObEncShared &sh, ObEncWork &wk, const ObEncHist &h)
{
    for (int i = g.lane; i < 2 * OB_OVERLAP; i += g.n) wk.hist.in_mem[i] = h.in_mem[i];
    for (int i = g.lane; i < 2 * OB_MAXPERIOD; i += g.n) wk.hist.prefilter_mem[i] = h.prefilter_mem[i];
    for (int i = g.lane; i < 2 * OB_NB; i += g.n) { sh.oldBandE[i] = h.oldBandE[i]; sh.oldLogE[i] = h.oldLogE[i]; sh.oldLogE2[i] = h.oldLogE2[i]; sh.energyError[i] = h.energyError[i]; }
    g.sync();
}
template <class G>
OB_DEV void ob_enc_store_hist(const G &g, const ObEncShared &sh, const ObEncWork &wk, ObEncHist &h)
{
    g.sync();
    for (int i = g.lane; i < 2 * OB_OVERLAP; i += g.n) h.in_mem[i] = wk.hist.in_mem[i];
    for (int i = g.lane; i < 2 * OB_MAXPERIOD; i += g.n) h.prefilter_mem[i] = wk.hist.prefilter_mem[i];
    for (int i = g.lane; i < 2 * OB_NB; i += g.n) { h.oldBandE[i] = sh.oldBandE[i]; h.oldLogE[i] = sh.oldLogE[i]; h.oldLogE2[i] = sh.oldLogE2[i]; h.energyError[i] = sh.energyError[i]; }
    g.sync();
}

// run_prefilter (celt_encoder.c:1188-1318)
template <class G>
OB_STAGE int ob_run_prefilter(const G &g, ObEncState &st, ObEncShared &sh, ObEncWork &wk, float *in, int CC, int N, int prefilter_tapset, int *pitch, float *gain,
        int *qgain, int enabled, int nbAvailableBytes)
{
    float *pre[2] = {wk.pre, wk.pre + (N + OB_MAXPERIOD)};
    int pitch_index, pf_on, qg;
    float gain1, pf_threshold;
    for (int c = 0; c < CC; c++) {
        float *__restrict__ dst = pre[c];
        const float *__restrict__ mem = wk.hist.prefilter_mem + c * OB_MAXPERIOD, *__restrict__ src = in + c * (N + OB_OVERLAP) + OB_OVERLAP;
        for (int i = g.lane; i < OB_MAXPERIOD; i += g.n) dst[i] = mem[i];
        for (int i = g.lane; i < N; i += g.n) dst[OB_MAXPERIOD + i] = src[i];
    }
    g.sync();
    if (enabled) {
        float *pitch_buf = sh.u.t.pitch_buf;
        g.pace(3);
        ob_pitch_downsample(g, pre[0], pre[1], pitch_buf, sh.A, OB_MAXPERIOD + N, CC);
        g.pace(4);
        ob_pitch_search(g, pitch_buf + (OB_MAXPERIOD >> 1), pitch_buf, N, OB_MAXPERIOD - 3 * OB_MINPERIOD, &pitch_index, sh.A);
        pitch_index = OB_MAXPERIOD - pitch_index;
        g.pace(5);
        gain1 = ob_remove_doubling(g, pitch_buf, OB_MAXPERIOD, OB_MINPERIOD, N, &pitch_index, st.prefilter_period, st.prefilter_gain, sh.A);
        if (pitch_index > OB_MAXPERIOD - 2) pitch_index = OB_MAXPERIOD - 2;
        gain1 = .7f * gain1;
        if (st.loss_rate > 2) gain1 = .5f * gain1;
        if (st.loss_rate > 4) gain1 = .5f * gain1;
        if (st.loss_rate > 8) gain1 = 0;
    } else { gain1 = 0; pitch_index = OB_MINPERIOD; }
    if (st.an.valid) gain1 = gain1 * st.an.max_pitch_ratio;          // celt_encoder.c:1246-1247
    pf_threshold = .2f;
    int dp = pitch_index - st.prefilter_period; if (dp < 0) dp = -dp;
    if (dp * 10 > pitch_index) pf_threshold += .2f;
    if (nbAvailableBytes < 25) pf_threshold += .1f;
    if (nbAvailableBytes < 35) pf_threshold += .1f;
    if (st.prefilter_gain > .4f) pf_threshold -= .1f;
    if (st.prefilter_gain > .55f) pf_threshold -= .1f;
    pf_threshold = ob_fmax(pf_threshold, .2f);
    if (gain1 < pf_threshold) { gain1 = 0; pf_on = 0; qg = 0; }
    else {
        if (fabsf(gain1 - st.prefilter_gain) < .1f) gain1 = st.prefilter_gain;
        qg = (int)floor((double)(.5f + gain1 * 32 / 3)) - 1;
        qg = ob_imax(0, ob_imin(7, qg));
        gain1 = 0.09375f * (qg + 1);
        pf_on = 1;
    }
    g.pace(6);
    for (int c = 0; c < CC; c++) {
        const int offset = OB_SHORT - OB_OVERLAP;      // 0 for this mode
        st.prefilter_period = ob_imax(st.prefilter_period, OB_MINPERIOD);
        float *inc = in + c * (N + OB_OVERLAP), *mem_in = wk.hist.in_mem + c * OB_OVERLAP, *mem_pf = wk.hist.prefilter_mem + c * OB_MAXPERIOD;
        for (int i = g.lane; i < OB_OVERLAP; i += g.n) inc[i] = mem_in[i];
        if (offset)
            ob_comb_filter_xy(g, inc + OB_OVERLAP, pre[c] + OB_MAXPERIOD, st.prefilter_period, st.prefilter_period, offset,
                    -st.prefilter_gain, -st.prefilter_gain, st.prefilter_tapset, st.prefilter_tapset, 0);
        ob_comb_filter_xy(g, inc + OB_OVERLAP + offset, pre[c] + OB_MAXPERIOD + offset, st.prefilter_period, pitch_index, N - offset,
                -st.prefilter_gain, -gain1, st.prefilter_tapset, prefilter_tapset, OB_OVERLAP);
        for (int i = g.lane; i < OB_OVERLAP; i += g.n) mem_in[i] = inc[N + i];
        // the new pre-filter memory is the last MAXPERIOD samples of [old memory | frame] (N <= 960 < MAXPERIOD), read from the assembled copy
        for (int i = g.lane; i < OB_MAXPERIOD; i += g.n) mem_pf[i] = pre[c][N + i];
    }
    g.sync();
    *gain = gain1; *pitch = pitch_index; *qgain = qg;
    return pf_on;
}

// celt_encode_with_ec (celt_encoder.c:1431-2368).  pcm: interleaved floats in [-1,1].  enc: coder created by the caller over the
// payload buffer (as opus_encode_frame_native does, opus_encoder.c:1791) -- tell == 1 on entry.  Returns bytes used or < 0.
template <class G>
OB_STAGE int ob_celt_encode(const G &g, ObEncState &st, ObEncShared &sh, ObEncWork &wk, const float *pcm, int frame_size, int nbCompressedBytes, ObRangeEnc &enc_io)
{
    ObRangeEnc enc = enc_io;                                      // the coder works in registers inside a stage
    const int CC = st.channels, C = st.stream_channels, end = st.end, effEnd = st.end;
    int LM, shortBlocks = 0, isTransient = 0, tf_select, codedBands, alloc_trim, pitch_index = OB_MINPERIOD, dual_stereo = 0, effectiveBytes;
    int prefilter_tapset = 0, pf_on, anti_collapse_rsv, anti_collapse_on = 0, silence = 0, tf_chan = 0, pitch_change = 0, secondMdct;
    int signalBandwidth, transient_got_disabled = 0, enable_tf_analysis, nbFilledBytes, nbAvailableBytes, dynalloc_logp;
    int32_t vbr_rate, total_bits, total_boost, balance, tell, tot_boost = 0, equiv_rate, bits;
    float gain1 = 0, tf_estimate = 0, sample_max, maxDepth, temporal_vbr = 0;
    float *bandE = sh.bandE, *bandLogE = sh.bandLogE, *bandLogE2 = sh.bandLogE2, *error = sh.error;
    int *fine_quant = sh.fine_quant, *pulses = sh.pulses, *cap = sh.cap, *offsets = sh.offsets, *importance = sh.importance, *spread_weight = sh.spread_weight,
        *fine_priority = sh.fine_priority, *tf_res = sh.tf_res;
    uint8_t *collapse_masks = sh.collapse_masks;
    float *oldBandE = sh.oldBandE, *oldLogE = sh.oldLogE, *oldLogE2 = sh.oldLogE2, *energyError = sh.energyError;
    float *in = wk.in, *freq = wk.freq, *X = wk.X;

    if (nbCompressedBytes < 2 || pcm == nullptr) return OB_BAD_ARG;
    for (LM = 0; LM <= 3; LM++) if (OB_SHORT << LM == frame_size) break;
    if (LM > 3) return OB_BAD_ARG;
    const int M = 1 << LM, N = M * OB_SHORT;
    tell = enc.tell();
    nbFilledBytes = (tell + 4) >> 3;
    nbCompressedBytes = ob_imin(nbCompressedBytes, 1275);
    nbAvailableBytes = nbCompressedBytes - nbFilledBytes;
    if (st.vbr && st.bitrate != OB_BITRATE_MAX) {
        const int32_t den = 48000 >> OB_BITRES;
        vbr_rate = (st.bitrate * frame_size + (den >> 1)) / den;
        effectiveBytes = vbr_rate >> (3 + OB_BITRES);
    } else {
        int32_t tmp;
        vbr_rate = 0;
        tmp = st.bitrate * frame_size;
        if (tell > 1) tmp += tell * 48000;
        if (st.bitrate != OB_BITRATE_MAX) {
            nbCompressedBytes = ob_imax(2, ob_imin(nbCompressedBytes, (tmp + 4 * 48000) / (8 * 48000)));
            enc.shrink((uint32_t)nbCompressedBytes);
        }
        effectiveBytes = nbCompressedBytes - nbFilledBytes;
    }
    equiv_rate = ((int32_t)nbCompressedBytes * 8 * 50 << (3 - LM)) - (40 * C + 20) * ((400 >> LM) - 50);
    if (st.bitrate != OB_BITRATE_MAX) equiv_rate = ob_imin(equiv_rate, st.bitrate - (40 * C + 20) * ((400 >> LM) - 50));
    if (vbr_rate > 0 && st.constrained_vbr) {
        const int32_t vbr_bound = vbr_rate;
        const int32_t max_allowed = ob_imin(ob_imax(tell == 1 ? 2 : 0, (vbr_rate + vbr_bound - st.vbr_reservoir) >> (OB_BITRES + 3)), nbAvailableBytes);
        if (max_allowed < nbAvailableBytes) {
            nbCompressedBytes = nbFilledBytes + max_allowed;
            nbAvailableBytes = max_allowed;
            enc.shrink((uint32_t)nbCompressedBytes);
        }
    }
    total_bits = nbCompressedBytes * 8;

    sample_max = ob_fmax(st.overlap_max, ob_maxabs(g, pcm, C * (N - OB_OVERLAP)));
    st.overlap_max = ob_maxabs(g, pcm + C * (N - OB_OVERLAP), C * OB_OVERLAP);
    sample_max = ob_fmax(sample_max, st.overlap_max);
    silence = (sample_max <= (float)1 / (1 << st.lsb_depth));
    if (tell == 1) enc.bit_logp(silence, 15);
    else silence = 0;
    if (silence) {
        if (vbr_rate > 0) {
            effectiveBytes = nbCompressedBytes = ob_imin(nbCompressedBytes, nbFilledBytes + 2);
            total_bits = nbCompressedBytes * 8;
            nbAvailableBytes = 2;
            enc.shrink((uint32_t)nbCompressedBytes);
        }
        tell = nbCompressedBytes * 8;
        enc.nbits_total += tell - enc.tell();
    }
    g.pace(2);
    for (int c = 0; c < CC; c++) {
        const int need_clip = st.clip && sample_max > 65536.f;
        ob_preemphasis(g, pcm + c, in + c * (N + OB_OVERLAP) + OB_OVERLAP, N, CC, &st.preemph_memE[c], need_clip);
    }
    {   // pitch pre-filter
        int qg;
        const int enabled = nbAvailableBytes > 12 * C && !silence && !st.disable_pf && st.complexity >= 5;
        prefilter_tapset = st.tapset_decision;
        pf_on = ob_run_prefilter(g, st, sh, wk, in, CC, N, prefilter_tapset, &pitch_index, &gain1, &qg, enabled, nbAvailableBytes);
        if ((gain1 > .4f || st.prefilter_gain > .4f) && (!st.an.valid || (double)st.an.tonality > .3)
                && (pitch_index > 1.26 * st.prefilter_period || pitch_index < .79 * st.prefilter_period))
            pitch_change = 1;
        if (pf_on == 0) {
            if (tell + 16 <= total_bits) enc.bit_logp(0, 1);
        } else {
            enc.bit_logp(1, 1);
            pitch_index += 1;
            const int octave = ob_ilog((uint32_t)pitch_index) - 5;
            enc.uint((uint32_t)octave, 6);
            enc.bits((uint32_t)(pitch_index - (16 << octave)), (uint32_t)(4 + octave));
            pitch_index -= 1;
            enc.bits((uint32_t)qg, 3);
            enc.icdf(prefilter_tapset, OB_TAPSET_ICDF, 2);
        }
    }
    g.pace(7);
    if (st.complexity >= 1) isTransient = ob_transient_analysis(g, in, N + OB_OVERLAP, CC, &tf_estimate, &tf_chan, sh.A, wk.env);
    if (LM > 0 && enc.tell() + 3 <= total_bits) { if (isTransient) shortBlocks = M; }
    else { isTransient = 0; transient_got_disabled = 1; }

    g.pace(8);
    secondMdct = shortBlocks && st.complexity >= 8;
    if (secondMdct) {
        ob_compute_mdcts(g, 0, in, freq, C, CC, LM, sh.A);
        ob_band_energies(g, freq, bandE, effEnd, C, LM);
        ob_amp2log2(g, effEnd, end, bandE, bandLogE2, C);
        for (int c = 0; c < C; c++) for (int i = 0; i < end; i++) bandLogE2[OB_NB * c + i] += .5f * (float)LM;
    }
    ob_compute_mdcts(g, shortBlocks, in, freq, C, CC, LM, sh.A);
    if (CC == 2 && C == 1) tf_chan = 0;
    ob_band_energies(g, freq, bandE, effEnd, C, LM);
    ob_amp2log2(g, effEnd, end, bandE, bandLogE, C);
    {   // temporal VBR (celt_encoder.c:1849-1865)
        float follow = -10.0f, frame_avg = 0;
        const float offset = shortBlocks ? .5f * (float)LM : 0;
        for (int i = 0; i < end; i++) {
            follow = ob_fmax(follow - 1.f, bandLogE[i] - offset);
            if (C == 2) follow = ob_fmax(follow, bandLogE[i + OB_NB] - offset);
            frame_avg += follow;
        }
        frame_avg /= (float)end;
        temporal_vbr = frame_avg - st.spec_avg;
        temporal_vbr = ob_fmin(3.f, ob_fmax(-1.5f, temporal_vbr));
        st.spec_avg += .02f * temporal_vbr;
    }
    if (!secondMdct) for (int i = 0; i < C * OB_NB; i++) bandLogE2[i] = bandLogE[i];
    if (LM > 0 && enc.tell() + 3 <= total_bits && !isTransient && st.complexity >= 5) {
        if (ob_patch_transient(bandLogE, oldBandE, end, C, sh.A)) {
            isTransient = 1;
            shortBlocks = M;
            g.sync();
            ob_compute_mdcts(g, shortBlocks, in, freq, C, CC, LM, sh.A);
            ob_band_energies(g, freq, bandE, effEnd, C, LM);
            ob_amp2log2(g, effEnd, end, bandE, bandLogE, C);
            for (int c = 0; c < C; c++) for (int i = 0; i < end; i++) bandLogE2[OB_NB * c + i] += .5f * (float)LM;
            tf_estimate = .2f;
        }
    }
    if (LM > 0 && enc.tell() + 3 <= total_bits) enc.bit_logp(isTransient, 3);

    g.pace(10);
    ob_normalise_bands(g, freq, X, bandE, effEnd, C, M);
    enable_tf_analysis = effectiveBytes >= 15 * C && st.complexity >= 2;
    maxDepth = ob_dynalloc_analysis(bandLogE, bandLogE2, oldBandE, end, C, offsets, st.lsb_depth, isTransient, st.vbr, st.constrained_vbr, LM,
            effectiveBytes, &tot_boost, importance, spread_weight, st.an.valid ? st.an.leak_boost : nullptr, sh.A);
    g.pace(11);
    if (enable_tf_analysis) {
        const int lambda = ob_imax(80, 20480 / effectiveBytes + 2);
        tf_select = ob_tf_analysis(g, effEnd, isTransient, tf_res, lambda, X, N, LM, tf_estimate, tf_chan, importance, sh.u.bands.xb, sh.u.bands.xb + OB_MAX_BAND,
                                   (int *)sh.A);
        for (int i = effEnd; i < end; i++) tf_res[i] = tf_res[effEnd - 1];
    } else {
        for (int i = 0; i < end; i++) tf_res[i] = isTransient;
        tf_select = 0;
    }
    for (int c = 0; c < C; c++) for (int i = 0; i < end; i++) {
        if (fabsf(bandLogE[i + c * OB_NB] - oldBandE[i + c * OB_NB]) < 2.f) bandLogE[i + c * OB_NB] -= energyError[i + c * OB_NB] * 0.25f;
    }
    g.pace(12);
    ob_quant_coarse_energy(g, end, effEnd, bandLogE, oldBandE, (uint32_t)total_bits, error, enc, C, LM, nbAvailableBytes, st.force_intra,
            &st.delayedIntra, st.complexity >= 4, st.loss_rate, wk.bw.bytes_save, sh.A);
    g.pace(13);
    ob_tf_encode(end, isTransient, tf_res, LM, tf_select, enc);
    if (enc.tell() + 4 <= total_bits) {
        if (shortBlocks || st.complexity < 3 || nbAvailableBytes < 10 * C) {
            if (st.complexity == 0) st.spread_decision = 0; else st.spread_decision = 2;
        } else {
            st.spread_decision = ob_spreading_decision(g, X, &st.tonal_average, st.spread_decision, &st.hf_average, &st.tapset_decision,
                    pf_on && !shortBlocks, effEnd, C, M, spread_weight);
        }
        enc.icdf(st.spread_decision, OB_SPREAD_ICDF, 5);
    }
    for (int i = 0; i < OB_NB; i++) {                                  // init_caps (celt.c:272-281)
        const int Nb = (OB_EBANDS[i + 1] - OB_EBANDS[i]) << LM;
        cap[i] = (OB_CACHE_CAPS[OB_NB * (2 * LM + C - 1) + i] + 64) * C * Nb >> 2;
    }
    dynalloc_logp = 6;
    total_bits <<= OB_BITRES;
    total_boost = 0;
    tell = (int32_t)enc.tell_frac();
    for (int i = 0; i < end; i++) {                                    // celt_encoder.c:2017-2050
        const int width = C * (OB_EBANDS[i + 1] - OB_EBANDS[i]) << LM;
        const int quanta = ob_imin(width << OB_BITRES, ob_imax(6 << OB_BITRES, width));
        int dynalloc_loop_logp = dynalloc_logp, boost = 0, j;
        for (j = 0; tell + (dynalloc_loop_logp << OB_BITRES) < total_bits - total_boost && boost < cap[i]; j++) {
            const int flag = j < offsets[i];
            enc.bit_logp(flag, (uint32_t)dynalloc_loop_logp);
            tell = (int32_t)enc.tell_frac();
            if (!flag) break;
            boost += quanta;
            total_boost += quanta;
            dynalloc_loop_logp = 1;
        }
        if (j) dynalloc_logp = ob_imax(2, dynalloc_logp - 1);
        offsets[i] = boost;
    }
    if (C == 2) {
        const float intensity_thresholds[21] = {1, 2, 3, 4, 5, 6, 7, 8, 16, 24, 36, 44, 50, 56, 62, 67, 72, 79, 88, 106, 134};
        const float intensity_histeresis[21] = {1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 3, 3, 4, 5, 6, 8, 8};
        if (LM != 0) dual_stereo = ob_stereo_analysis(g, X, LM, N);
        st.intensity = ob_hysteresis_decision((float)(equiv_rate / 1000), intensity_thresholds, intensity_histeresis, 21, st.intensity);
        st.intensity = ob_imin(end, ob_imax(0, st.intensity));
    }
    alloc_trim = 5;
    if (tell + (6 << OB_BITRES) <= total_bits - total_boost) {
        alloc_trim = ob_alloc_trim_analysis(g, X, bandLogE, end, LM, C, N, &st.stereo_saving, tf_estimate, st.intensity, equiv_rate, st.an.valid, st.an.tonality_slope);
        enc.icdf(alloc_trim, OB_TRIM_ICDF, 7);
        tell = (int32_t)enc.tell_frac();
    }
    if (vbr_rate > 0) {                                                // celt_encoder.c:2086-2195
        float alpha;
        int32_t delta, target, base_target, min_allowed;
        const int lm_diff = 3 - LM;
        nbCompressedBytes = ob_imin(nbCompressedBytes, 1275 >> (3 - LM));
        base_target = vbr_rate - ((40 * C + 20) << OB_BITRES);
        if (st.constrained_vbr) base_target += (st.vbr_offset >> lm_diff);
        target = ob_compute_vbr(base_target, LM, equiv_rate, st.lastCodedBands, C, st.intensity, st.constrained_vbr, st.stereo_saving, tot_boost,
                tf_estimate, maxDepth, temporal_vbr, st.an.valid, st.an.activity, st.an.tonality, pitch_change);
        target = target + tell;
        min_allowed = ((tell + total_boost + (1 << (OB_BITRES + 3)) - 1) >> (OB_BITRES + 3)) + 2;
        nbAvailableBytes = (target + (1 << (OB_BITRES + 2))) >> (OB_BITRES + 3);
        nbAvailableBytes = ob_imax(min_allowed, nbAvailableBytes);
        nbAvailableBytes = ob_imin(nbCompressedBytes, nbAvailableBytes);
        delta = target - vbr_rate;
        target = nbAvailableBytes << (OB_BITRES + 3);
        if (silence) { nbAvailableBytes = 2; target = 2 * 8 << OB_BITRES; delta = 0; }
        if (st.vbr_count < 970) { st.vbr_count++; alpha = 1.f / (float)(st.vbr_count + 20); }
        else alpha = .001f;
        if (st.constrained_vbr) st.vbr_reservoir += target - vbr_rate;
        if (st.constrained_vbr) {
            st.vbr_drift += (int32_t)(alpha * (float)((delta * (1 << lm_diff)) - st.vbr_offset - st.vbr_drift));
            st.vbr_offset = -st.vbr_drift;
        }
        if (st.constrained_vbr && st.vbr_reservoir < 0) {
            const int adjust = (-st.vbr_reservoir) / (8 << OB_BITRES);
            nbAvailableBytes += silence ? 0 : adjust;
            st.vbr_reservoir = 0;
        }
        nbCompressedBytes = ob_imin(nbCompressedBytes, nbAvailableBytes);
        enc.shrink((uint32_t)nbCompressedBytes);
    }
    bits = (((int32_t)nbCompressedBytes * 8) << OB_BITRES) - (int32_t)enc.tell_frac() - 1;
    anti_collapse_rsv = isTransient && LM >= 2 && bits >= ((LM + 2) << OB_BITRES) ? (1 << OB_BITRES) : 0;
    bits -= anti_collapse_rsv;
    signalBandwidth = end - 1;
    if (st.an.valid) {                                                 // celt_encoder.c:2208-2222
        int min_bandwidth;
        if (equiv_rate < (int32_t)32000 * C) min_bandwidth = 13;
        else if (equiv_rate < (int32_t)48000 * C) min_bandwidth = 16;
        else if (equiv_rate < (int32_t)60000 * C) min_bandwidth = 18;
        else if (equiv_rate < (int32_t)80000 * C) min_bandwidth = 19;
        else min_bandwidth = 20;
        signalBandwidth = ob_imax(st.an.bandwidth, min_bandwidth);
    }
    g.pace(14);
    codedBands = ob_enc_allocation(enc, end, offsets, cap, alloc_trim, &st.intensity, &dual_stereo, bits, &balance, pulses, fine_quant, fine_priority,
            C, LM, st.lastCodedBands, signalBandwidth, (int *)sh.A);
    if (st.lastCodedBands) st.lastCodedBands = ob_imin(st.lastCodedBands + 1, ob_imax(st.lastCodedBands - 1, codedBands));
    else st.lastCodedBands = codedBands;
    ob_quant_fine_energy(end, oldBandE, error, fine_quant, enc, C);
    for (int i = 0; i < 2 * OB_NB; i++) collapse_masks[i] = 0;
    g.sync();
    ob_enc_all_bands(g, end, X, C, N, collapse_masks, bandE, pulses, shortBlocks, st.spread_decision, dual_stereo, st.intensity, tf_res,
            nbCompressedBytes * (8 << OB_BITRES) - anti_collapse_rsv, balance, enc, LM, codedBands, &st.rng, st.complexity, st.disable_inv, sh.u.bands, wk.bw);
    g.pace(64 + 64 * 21);
    if (anti_collapse_rsv > 0) {
        anti_collapse_on = st.consec_transient < 2;
        enc.bits((uint32_t)anti_collapse_on, 1);
    }
    ob_quant_energy_finalise(end, oldBandE, error, fine_quant, fine_priority, nbCompressedBytes * 8 - enc.tell(), enc, C);
    for (int i = 0; i < OB_NB * CC; i++) energyError[i] = 0;
    for (int c = 0; c < C; c++) for (int i = 0; i < end; i++) energyError[i + c * OB_NB] = ob_fmax(-0.5f, ob_fmin(0.5f, error[i + c * OB_NB]));
    if (silence) for (int i = 0; i < C * OB_NB; i++) oldBandE[i] = -28.f;
    st.prefilter_period = pitch_index;
    st.prefilter_gain = gain1;
    st.prefilter_tapset = prefilter_tapset;
    if (CC == 2 && C == 1) for (int i = 0; i < OB_NB; i++) oldBandE[OB_NB + i] = oldBandE[i];
    if (!isTransient) {
        for (int i = 0; i < CC * OB_NB; i++) { oldLogE2[i] = oldLogE[i]; oldLogE[i] = oldBandE[i]; }
    } else for (int i = 0; i < CC * OB_NB; i++) oldLogE[i] = ob_fmin(oldLogE[i], oldBandE[i]);
    for (int c = 0; c < CC; c++) for (int i = end; i < OB_NB; i++) { oldBandE[c * OB_NB + i] = 0; oldLogE[c * OB_NB + i] = oldLogE2[c * OB_NB + i] = -28.f; }
    if (isTransient || transient_got_disabled) st.consec_transient++; else st.consec_transient = 0;
    st.rng = enc.rng;
    enc.done();
    enc_io = enc;
    g.sync();
    if (enc.error) return OB_INTERNAL_ERROR;
    return nbCompressedBytes;
}

// dc_reject (opus_encoder.c:430-468), float build, cutoff 3 Hz @ 48 kHz: out = x - m, m <- coef*x + 1e-30 + coef2*m: a first-order scan per channel
template <class G>
OB_STAGE void ob_dc_reject(const G &g, const float *in, float *out, float *hp_mem, int len, int channels)
{
    const float coef = 6.3f * 3 / 48000, coef2 = 1 - coef;
    for (int c = 0; c < channels; c++) {
        const float m0 = hp_mem[2 * c];
        const float *x = in + c;
        float *y = out + c;
        // y_i (stored) needs the memory BEFORE sample i: put(i, m_i) writes the output of sample i+1; sample 0 uses the carried memory
        if (g.lane == 0) y[0] = x[0] - m0;
        hp_mem[2 * c] = ob_scan1(g, len, coef2, m0, false, [&](int i) { return coef * x[channels * i] + 1e-30f; },
                                 [&](int i, float m) { if (i + 1 < len) y[channels * (i + 1)] = x[channels * (i + 1)] - m; });
    }
    g.sync();
}

// hp_cutoff (opus_encoder.c:369-404) + silk_biquad_float (:331-366), Fs = 48000.  cutoff_Hz is 60 for as long as a stream stays
// CELT-only: variable_HP_smth2_Q15 starts at lin2log(60) << 8 and in MODE_CELT_ONLY is smoothed towards the same value (:1795-1805).
// The biquad's two memories form a second-order linear recurrence: chunk-per-lane scan.
template <class G>
OB_STAGE void ob_hp_cutoff(const G &g, const float *in, float *out, float *hp_mem, int len, int channels)
{
    const int32_t cutoff_Hz = 60;
    const int32_t Fc_Q19 = (int32_t)((int16_t)2471 * (int32_t)(int16_t)cutoff_Hz) / 48;        // SILK_FIX_CONST(1.5 * 3.14159 / 1000, 19) = 2471
    const int32_t r_Q28 = (1 << 28) - 471 * Fc_Q19;                                          // SILK_FIX_CONST(0.92, 9) = 471
    const int32_t B_Q28[3] = {r_Q28, (int32_t)((uint32_t)(-r_Q28) << 1), r_Q28};
    const int32_t r_Q22 = r_Q28 >> 6;
    const int32_t fc2 = (int32_t)(((int64_t)Fc_Q19 * Fc_Q19) >> 16);                          // silk_SMULWW (64-bit form, OPUS_FAST_INT64)
    const int32_t A_Q28[2] = {(int32_t)(((int64_t)r_Q22 * (fc2 - (2 << 22))) >> 16), (int32_t)(((int64_t)r_Q22 * r_Q22) >> 16)};
    const float A0 = (float)(A_Q28[0] * (1.f / ((int32_t)1 << 28))), A1 = (float)(A_Q28[1] * (1.f / ((int32_t)1 << 28)));
    const float B0 = (float)(B_Q28[0] * (1.f / ((int32_t)1 << 28))), B1 = (float)(B_Q28[1] * (1.f / ((int32_t)1 << 28))), B2 = (float)(B_Q28[2] * (1.f / ((int32_t)1 << 28)));
    for (int c = 0; c < channels; c++) {
        ObState2 s; s.s0 = hp_mem[2 * c]; s.s1 = hp_mem[2 * c + 1];
        s = ob_scan2(g, len, -A0, 1.f, -A1, 0.f, s, [&](int k, float &S0, float &S1, bool emit) {
            const float inval = in[k * channels + c];
            const float vout = S0 + B0 * inval;
            S0 = S1 - vout * A0 + B1 * inval;
            S1 = -vout * A1 + B2 * inval + 1e-30f;
            if (emit) out[k * channels + c] = vout;
        });
        hp_mem[2 * c] = s.s0; hp_mem[2 * c + 1] = s.s1;
    }
    g.sync();
}

// compute_stereo_width (opus_encoder.c:729-809), float build: feeds the SILK / CELT mode thresholds of the non-low-delay applications
struct ObStereoWidth { float XX, XY, YY, smoothed_width, max_follower; };
template <class G>
OB_STAGE float ob_compute_stereo_width(const G &g, const float *pcm, int frame_size, ObStereoWidth &mem)
{
    const int frame_rate = 48000 / frame_size;
    const float short_alpha = 1.0f - 25 * 1.0f / ob_imax(50, frame_rate);
    float xx = 0, xy = 0, yy = 0;
    ob_psum2(g, (frame_size - 3 + 3) / 4, xx, xy, [&](int q, float &a, float &b) {
        const int i = 4 * q;
        float pxx, pxy, x, y;
        x = pcm[2 * i]; y = pcm[2 * i + 1]; pxx = x * x; pxy = x * y;
        x = pcm[2 * i + 2]; y = pcm[2 * i + 3]; pxx += x * x; pxy += x * y;
        x = pcm[2 * i + 4]; y = pcm[2 * i + 5]; pxx += x * x; pxy += x * y;
        x = pcm[2 * i + 6]; y = pcm[2 * i + 7]; pxx += x * x; pxy += x * y;
        a += pxx; b += pxy;
    });
    yy = ob_psum(g, (frame_size - 3 + 3) / 4, 0.f, [&](int q) {
        const int i = 4 * q;
        float pyy, y;
        y = pcm[2 * i + 1]; pyy = y * y;
        y = pcm[2 * i + 3]; pyy += y * y;
        y = pcm[2 * i + 5]; pyy += y * y;
        y = pcm[2 * i + 7]; pyy += y * y;
        return pyy;
    });
    if (!(xx < 1e9f) || xx != xx || !(yy < 1e9f) || yy != yy) xy = xx = yy = 0;
    mem.XX += short_alpha * (xx - mem.XX);
    mem.XY += short_alpha * (xy - mem.XY);
    mem.YY += short_alpha * (yy - mem.YY);
    mem.XX = ob_fmax(0, mem.XX); mem.XY = ob_fmax(0, mem.XY); mem.YY = ob_fmax(0, mem.YY);
    if (ob_fmax(mem.XX, mem.YY) > 8e-4f) {
        const float sqrt_xx = (float)sqrt((double)mem.XX), sqrt_yy = (float)sqrt((double)mem.YY);
        const float qrrt_xx = (float)sqrt((double)sqrt_xx), qrrt_yy = (float)sqrt((double)sqrt_yy);
        mem.XY = ob_fmin(mem.XY, sqrt_xx * sqrt_yy);
        const float corr = mem.XY / (1e-15f + sqrt_xx * sqrt_yy);
        const float ldiff = 1.0f * (float)fabs(qrrt_xx - qrrt_yy) / (1e-15f + qrrt_xx + qrrt_yy);
        const float width = (float)sqrt((double)(1.f - corr * corr)) * ldiff;
        mem.smoothed_width += (width - mem.smoothed_width) / frame_rate;
        mem.max_follower = ob_fmax(mem.max_follower - .02f / frame_rate, mem.smoothed_width);
    }
    return ob_fmin(1.0f, 20 * mem.max_follower);
}

// ---- Opus layer for CELT-only packets: OPUS_APPLICATION_RESTRICTED_LOWDELAY => MODE_CELT_ONLY (opus_encoder.c:1330-1332); AUDIO / VOIP
// for as long as the rate-dependent mode decision (:1333-1385) says MODE_CELT_ONLY -- a frame it would give to SILK is OB_UNIMPLEMENTED ----
// User-visible encoder settings (the CTLs of src/encoder.rs) + the Opus-layer state this path keeps between frames.
struct ObOpusEncCfg {
    int32_t bitrate;           // OPUS_SET_BITRATE: bits/s, or -1000 (OPUS_AUTO) / -1 (OPUS_BITRATE_MAX)
    int32_t complexity, vbr, vbr_constraint, max_bandwidth, user_bandwidth, force_channels, packet_loss, lsb_depth;
    int32_t application;       // 2048 VOIP, 2049 AUDIO, 2051 RESTRICTED_LOWDELAY (0 is read as 2051)
    int32_t signal_type;       // OPUS_SET_SIGNAL: 0 auto, 3001 voice, 3002 music
    int32_t prediction_disabled, phase_inversion_disabled, use_dtx, inband_fec;
    int32_t variable_duration; // OPUS_SET_EXPERT_FRAME_DURATION: 0 / 5000 = FRAMESIZE_ARG, 5001 (2.5 ms) ... 5009 (120 ms)
};
struct ObOpusEncState {
    int32_t stream_channels, first, auto_bandwidth, bandwidth, hybrid_stereo_width_Q14;
    int32_t voice_ratio, detected_bandwidth;     // from the signal analysis (opus_encoder.c:1146-1176); voice_ratio = -1: unknown
    ObTonalState *tonal;                         // st->analysis when the analysis runs inline (host emulation); the GPU runs it in its own kernel
    int32_t prev_mode;                           // 0 before the first packet, then 1002 (MODE_CELT_ONLY)
    int32_t nb_no_activity_ms_Q1;                // generalised DTX (opus_encoder.c:988-1013)
    float peak_signal_energy;
    ObStereoWidth width_mem;
    float *delay;                                // st->delay_buffer [OB_ENC_BUFFER * channels] of the AUDIO / VOIP applications (global memory); null: low delay
};

OB_DEV int32_t ob_compute_equiv_rate(int32_t bitrate, int channels, int frame_rate, int vbr, int celt_only, int complexity, int loss)
{                                                                    // opus_encoder.c:898-931
    int32_t equiv = bitrate;
    if (frame_rate > 50) equiv -= (40 * channels + 20) * (frame_rate - 50);
    if (!vbr) equiv -= equiv / 12;
    equiv = equiv * (90 + complexity) / 100;
    if (celt_only) { if (complexity < 5) equiv = equiv * 9 / 10; }
    else equiv -= equiv * loss / (12 * loss + 20);                   // "mode not known yet"
    return equiv;
}

template <class G>
OB_DEV void ob_stereo_fade(const G &g, float *buf, float g1, float g2, int frame_size)     // opus_encoder.c:471-501, in == out, channels == 2, Fs = 48000
{
    g1 = 1.0f - g1; g2 = 1.0f - g2;
    for (int i = g.lane; i < frame_size; i += g.n) {
        float gg = g2;
        if (i < OB_OVERLAP) { const float w = OB_WINDOW[i] * OB_WINDOW[i]; gg = w * g2 + (1.0f - w) * g1; }
        float diff = .5f * (buf[i * 2] - buf[i * 2 + 1]);
        diff = gg * diff;
        buf[i * 2] = buf[i * 2] - diff;
        buf[i * 2 + 1] = buf[i * 2 + 1] + diff;
    }
    g.sync();
}

// mean square of a frame: celt_inner_prod(pcm, pcm, n) / n (opus_encoder.c:1127-1130, :1757)
template <class G>
OB_DEV float ob_mean_square(const G &g, const float *x, int n) { return ob_psum(g, n, 0.f, [&](int i) { return x[i] * x[i]; }) / n; }

// opus_encode_frame_native (opus_encoder.c:1698-2459) for one 2.5/5/10/20 ms frame, CELT-only: DC reject, stereo fade, CELT, TOC.
// data: max_data_bytes of room for TOC + payload.  bitrate_bps / equiv_rate: the packet-level values (a frame of a multi-frame packet
// inherits them).  Returns the frame's packet length (TOC included) or a negative OPUS_* code.
template <class G>
OB_STAGE int ob_opus_encode_frame(const G &g, const ObOpusEncCfg &cfg, ObOpusEncState &os, ObEncState &st, ObEncShared &sh, ObEncWork &wk, const float *pcm, int frame_size,
        uint8_t *data, int max_data_bytes, int32_t bitrate_bps, int32_t equiv_rate, const ObAnalysisInfo &analysis_info, int is_silence)
{
    const int channels = st.channels, frame_rate = 48000 / frame_size;
    const int curr_bandwidth = os.bandwidth;
    int activity = -1;                                                                                        // VAD_NO_DECISION (:1748-1762)
    if (is_silence) activity = 0;
    else if (analysis_info.valid) {
        activity = analysis_info.activity_probability >= .1f;
        if (!activity) {
            const float noise_energy = ob_mean_square(g, pcm, frame_size * channels);
            activity = os.peak_signal_energy < (316.23f * noise_energy);
        }
    }

    g.pace(1);
    ObRangeEnc enc;
    enc.init(sh.bytes, (uint32_t)(max_data_bytes - 1));
    float *pcm_buf = wk.pcm_hp;
    const int total_buffer = os.delay ? OB_ENC_DELAY : 0;                                                     // :1740-1744
    for (int i = g.lane; i < total_buffer * channels; i += g.n) pcm_buf[i] = os.delay[(OB_ENC_BUFFER - total_buffer) * channels + i];
    float *fresh = pcm_buf + total_buffer * channels;
    if (cfg.application == 2048) ob_hp_cutoff(g, pcm, fresh, st.hp_mem, frame_size, channels);
    else ob_dc_reject(g, pcm, fresh, st.hp_mem, frame_size, channels);
    {
        const float sum = ob_psum(g, frame_size * channels, 0.f, [&](int i) { return fresh[i] * fresh[i]; });
        if (!(sum < 1e9f) || sum != sum) {
            for (int i = g.lane; i < frame_size * channels; i += g.n) fresh[i] = 0;
            st.hp_mem[0] = st.hp_mem[1] = st.hp_mem[2] = st.hp_mem[3] = 0;
            g.sync();
        }
    }
    if (os.delay) {                                                                                           // :2125-2134, before the fades
        const int keep = OB_ENC_BUFFER - (frame_size + total_buffer);
        if (keep > 0) {
            float *t = sh.A;                                                                                  // keep * channels <= 336 floats, moved through scratch
            for (int i = g.lane; i < channels * keep; i += g.n) t[i] = os.delay[channels * frame_size + i];
            g.sync();
            for (int i = g.lane; i < channels * keep; i += g.n) os.delay[i] = t[i];
            for (int i = g.lane; i < (frame_size + total_buffer) * channels; i += g.n) os.delay[channels * keep + i] = pcm_buf[i];
        } else for (int i = g.lane; i < OB_ENC_BUFFER * channels; i += g.n) os.delay[i] = pcm_buf[(frame_size + total_buffer - OB_ENC_BUFFER) * channels + i];
        g.sync();
    }
    int stereoWidth_Q14;
    if (equiv_rate > 32000) stereoWidth_Q14 = 16384;
    else if (equiv_rate < 16000) stereoWidth_Q14 = 0;
    else stereoWidth_Q14 = 16384 - 2048 * (int32_t)(32000 - equiv_rate) / (equiv_rate - 14000);
    if (channels == 2 && (os.hybrid_stereo_width_Q14 < (1 << 14) || stereoWidth_Q14 < (1 << 14))) {
        float g1 = (float)os.hybrid_stereo_width_Q14, g2 = (float)stereoWidth_Q14;
        g1 *= (1.f / 16384); g2 *= (1.f / 16384);
        ob_stereo_fade(g, pcm_buf, g1, g2, frame_size);
        os.hybrid_stereo_width_Q14 = stereoWidth_Q14;
    }
    const int nb_compr_bytes = max_data_bytes - 1;
    st.end = curr_bandwidth == 1101 ? 13 : curr_bandwidth <= 1103 ? 17 : curr_bandwidth == 1104 ? 19 : 21;
    st.stream_channels = os.stream_channels;
    st.complexity = cfg.complexity; st.lsb_depth = cfg.lsb_depth; st.loss_rate = cfg.packet_loss;
    st.vbr = cfg.vbr; st.constrained_vbr = cfg.vbr_constraint;
    st.disable_pf = st.force_intra = cfg.prediction_disabled != 0;                                            // CELT_SET_PREDICTION(reducedDependency ? 0 : 2) (:2109-2116)
    st.disable_inv = cfg.phase_inversion_disabled != 0;
    st.bitrate = cfg.vbr ? bitrate_bps : OB_BITRATE_MAX;
    st.an = analysis_info;                                                                                    // CELT_SET_ANALYSIS (:2229)
    int ret = ob_celt_encode(g, st, sh, wk, pcm_buf, frame_size, nb_compr_bytes, enc);
    if (ret < 0) return OB_INTERNAL_ERROR;
    {   // gen_toc (:299-329)
        int period = 0, fr = frame_rate;
        while (fr < 400) { fr <<= 1; period++; }
        int tmp = curr_bandwidth - 1102;
        if (tmp < 0) tmp = 0;
        if (g.lane == 0) data[0] = (uint8_t)(0x80 | tmp << 5 | period << 3 | (os.stream_channels == 2) << 2);
    }
    for (int i = g.lane; i < ret; i += g.n) data[1 + i] = sh.bytes[i];                                       // the payload leaves shared memory in one coalesced copy
    g.sync();
    st.final_range = enc.rng;
    os.first = 0;
    os.prev_mode = 1002;
    if (cfg.use_dtx && (analysis_info.valid || is_silence)) {                                                 // DTX decision (:2364-2378)
        int dtx = 0;
        if (!activity) {
            os.nb_no_activity_ms_Q1 += 2 * 1000 * frame_size / 48000;
            if (os.nb_no_activity_ms_Q1 > 10 * 20 * 2) {
                if (os.nb_no_activity_ms_Q1 <= (10 + 20) * 20 * 2) dtx = 1;
                else os.nb_no_activity_ms_Q1 = 10 * 20 * 2;
            }
        } else os.nb_no_activity_ms_Q1 = 0;
        if (dtx) { st.final_range = 0; return 1; }                                                            // data[0] already holds the TOC
    } else os.nb_no_activity_ms_Q1 = 0;
    ret += 1;
    if (!cfg.vbr && ret != max_data_bytes) return OB_UNIMPLEMENTED;     // opus_packet_pad: CELT CBR always fills its budget on this path
    return ret;
}

// opus_encode_float -> opus_encode_native -> opus_encode_frame_native for one packet of 2.5 ... 120 ms, CELT-only
// (opus_encoder.c:1057-1696, :1698-2459; the lines this path executes are listed in SURVEY 8a).  data: out_bytes capacity.
// Returns the packet length in bytes (TOC included) or a negative OPUS_* code.
// pre_info: the frame's AnalysisInfo when the analysis ran ahead of the encoder in its own kernel (ob_k_analysis); nullptr: run it
// inline on os.tonal (lane 0; the result reaches the other lanes through shared memory).
template <class G>
OB_STAGE int ob_opus_encode(const G &g, const ObOpusEncCfg &cfg, ObOpusEncState &os, ObEncState &st, ObEncShared &sh, ObEncWork &wk, const float *pcm, int frame_size,
        uint8_t *data, int out_bytes, const ObAnalysisInfo *pre_info = nullptr, int pace_base = 0)
{
    const int channels = st.channels, Fs = 48000;
    g.set_base(pace_base);
    st.final_range = 0;                                                                                       // st->rangeFinal = 0 (:1092)
    if (frame_size != 120 && frame_size != 240 && frame_size != 480 && (frame_size % 960 != 0 || frame_size <= 0 || frame_size > 5760)) return OB_BAD_ARG;
    int max_data_bytes = ob_imin(1276, out_bytes);
    if (max_data_bytes <= 0) return OB_BAD_ARG;
    if (max_data_bytes == 1 && Fs == frame_size * 10) return OB_BUFFER_TOO_SMALL;                             // "cannot encode 100 ms in 1 byte"
    const int frame_rate = Fs / frame_size;
    int32_t bitrate_bps;
    int cbr_bytes = -1;
    if (cfg.bitrate == -1000) bitrate_bps = 60 * Fs / frame_size + Fs * channels;                              // user_bitrate_to_bitrate (:639-649)
    else if (cfg.bitrate == -1) bitrate_bps = max_data_bytes * 8 * Fs / frame_size;
    else bitrate_bps = cfg.bitrate;
    if (!cfg.vbr) {                                                                                           // :1188-1197
        const int frame_rate12 = 12 * Fs / frame_size;
        cbr_bytes = ob_imin((12 * bitrate_bps / 8 + frame_rate12 / 2) / frame_rate12, max_data_bytes);
        bitrate_bps = cbr_bytes * (int32_t)frame_rate12 * 8 / 12;
        max_data_bytes = ob_imax(1, cbr_bytes);
    }
    // ---- signal analysis at complexity >= 7 (:1108-1176), on the caller's PCM, before anything else looks at the frame ----
    ObAnalysisInfo analysis_info;
    analysis_info.valid = 0;
    int read_pos_bak = -1, read_subframe_bak = -1;
    int is_silence = 0;
    {
        const int lsb_depth = cfg.lsb_depth;
        if (cfg.complexity >= 7) {
            is_silence = ob_maxabs(g, pcm, frame_size * channels) <= (float)1 / (1 << lsb_depth);
            if (pre_info) analysis_info = *pre_info;
            else {
                read_pos_bak = os.tonal->read_pos; read_subframe_bak = os.tonal->read_subframe;
                g.sync();
                {   // cooperative front (resampler scans, FFT, per-bin and per-band sums), scalar statistics + network on lane 0: the result reaches the warp through shared memory
                    ObAnalysisInfo a; a.valid = 0;
                    ob_run_analysis(g, *os.tonal, pcm, frame_size, channels, lsb_depth, a, wk.pre, sh.A);
                    if (g.lane == 0) sh.an_tmp = a;
                }
                g.sync();
                analysis_info = sh.an_tmp;
                g.sync();
            }
            if (!is_silence && analysis_info.activity_probability > .1f)                                      // peak signal energy (:1127-1130)
                os.peak_signal_energy = ob_fmax(.999f * os.peak_signal_energy, ob_mean_square(g, pcm, frame_size * channels));
        } else if (os.tonal && os.tonal->initialized) { g.sync(); if (g.lane == 0) ob_tonal_reset(*os.tonal); g.sync(); }
        if (!is_silence) os.voice_ratio = -1;
        os.detected_bandwidth = 0;
        if (analysis_info.valid) {
            const float prob = os.first ? analysis_info.music_prob : analysis_info.music_prob_max;           // prev_mode == 0 : == MODE_CELT_ONLY
            if (cfg.signal_type == 0) os.voice_ratio = (int)floor(.5 + (double)(100 * (1 - prob)));
            const int ab = analysis_info.bandwidth;
            os.detected_bandwidth = ab <= 12 ? 1101 : ab <= 14 ? 1102 : ab <= 16 ? 1103 : ab <= 18 ? 1104 : 1105;
        }
    }
    const float stereo_width = (channels == 2 && cfg.force_channels != 1 && (cfg.application == 2048 || cfg.application == 2049))
                                   ? ob_compute_stereo_width(g, pcm, frame_size, os.width_mem) : 0;             // :1181-1184 (only the mode decision reads it)
    if (max_data_bytes < 3 || bitrate_bps < 3 * frame_rate * 8 || (frame_rate < 50 && (max_data_bytes * frame_rate < 300 || bitrate_bps < 2400))) {
        // too little room to code anything: a 'PLC frame', i.e. a TOC and nothing else (:1202-1266); st->mode is still the previous packet's
        // (MODE_HYBRID = 1001 from opus_encoder_init until a packet has been coded)
        int tocmode = os.prev_mode ? os.prev_mode : 1001, bw = os.bandwidth == 0 ? 1101 : os.bandwidth, packet_code = 0, num_multiframes = 0, fr = frame_rate;
        if (fr > 100) tocmode = 1002;
        if (fr == 25 && tocmode != 1000) { fr = 50; packet_code = 1; }
        if (fr <= 16) {
            if (out_bytes == 1 || (tocmode == 1000 && fr != 10)) { tocmode = 1000; packet_code = fr <= 12; fr = fr == 12 ? 25 : 16; }
            else { num_multiframes = 50 / fr; fr = 50; packet_code = 3; }
        }
        if (tocmode == 1000 && bw > 1103) bw = 1103;
        else if (tocmode == 1002 && bw == 1102) bw = 1101;
        else if (tocmode == 1001 && bw <= 1104) bw = 1104;
        int period = 0;
        while (fr < 400) { fr <<= 1; period++; }
        int toc;                                                                                              // gen_toc (:299-329)
        if (tocmode == 1000) toc = (bw - 1101) << 5 | (period - 2) << 3;
        else if (tocmode == 1001) toc = 0x60 | (bw - 1104) << 4 | (period - 2) << 3;
        else { int tmp = bw - 1102; if (tmp < 0) tmp = 0; toc = 0x80 | tmp << 5 | period << 3; }
        toc |= (os.stream_channels == 2) << 2;
        int ret = packet_code <= 1 ? 1 : 2;
        max_data_bytes = ob_imax(max_data_bytes, ret);
        st.final_range = 0;
        int rc = ret;
        g.sync();
        if (g.lane == 0) {                                                                                    // byte work of a rare path: one lane
            data[0] = (uint8_t)(toc | packet_code);
            if (packet_code == 3) data[1] = (uint8_t)num_multiframes;
            if (!cfg.vbr && ret != max_data_bytes) {                                                          // opus_packet_pad(data, ret, max_data_bytes)
                uint8_t copy[2] = {data[0], data[1]};
                ObRepack &rp = wk.rp;
                ob_repack_init(&rp);
                ObExt none;
                if (ob_repack_cat(&rp, copy, ret, 0) != OB_OK) rc = OB_INTERNAL_ERROR;
                else if (ob_repack_out_range(ObRpLanes1(), &rp, 0, rp.nb_frames, data, max_data_bytes, 0, 1, &none, 0) <= 0) rc = OB_INTERNAL_ERROR;
                else rc = max_data_bytes;
            }
        }
        rc = ob_bcast(g, rc, 0);
        g.sync();
        return rc;
    }
    int32_t equiv_rate = ob_compute_equiv_rate(bitrate_bps, channels, frame_rate, cfg.vbr, 0, cfg.complexity, cfg.packet_loss);
    int voice_est;                                                                                            // :1276-1289
    if (cfg.signal_type == 3001) voice_est = 127;
    else if (cfg.signal_type == 3002) voice_est = 0;
    else if (os.voice_ratio >= 0) { voice_est = os.voice_ratio * 327 >> 8; if (cfg.application == 2049) voice_est = ob_imin(voice_est, 115); }
    else voice_est = cfg.application == 2048 ? 115 : 48;
    if (cfg.force_channels > 0 && channels == 2) os.stream_channels = cfg.force_channels;
    else if (channels == 2) {
        int32_t stereo_threshold = 17000 + ((voice_est * voice_est * (19000 - 17000)) >> 14);
        if (os.stream_channels == 2) stereo_threshold -= 1000; else stereo_threshold += 1000;
        os.stream_channels = (equiv_rate > stereo_threshold) ? 2 : 1;
    } else os.stream_channels = channels;
    if (cfg.application == 2048 || cfg.application == 2049) {                                                 // mode selection (:1333-1392)
        equiv_rate = ob_compute_equiv_rate(bitrate_bps, os.stream_channels, frame_rate, cfg.vbr, 0, cfg.complexity, cfg.packet_loss);
        const int32_t mode_voice = (int32_t)((1.0f - stereo_width) * 64000 + stereo_width * 44000);
        const int32_t mode_music = (int32_t)((1.0f - stereo_width) * 10000 + stereo_width * 10000);
        int32_t threshold = mode_music + ((voice_est * voice_est * (mode_voice - mode_music)) >> 14);
        if (cfg.application == 2048) threshold += 8000;
        if (os.prev_mode == 1002) threshold -= 4000;
        int celt_only = equiv_rate >= threshold;
        if (cfg.inband_fec && cfg.packet_loss > (128 - voice_est) >> 4 && (cfg.inband_fec != 2 || voice_est > 25)) celt_only = 0;
        if (cfg.use_dtx && !(analysis_info.valid || is_silence) && voice_est > 100) celt_only = 0;           // SILK's own DTX
        if (max_data_bytes < (frame_rate > 50 ? 9000 : 6000) * frame_size / (Fs * 8)) celt_only = 1;
        if (frame_size < Fs / 100) celt_only = 1;
        if (!celt_only) return OB_UNIMPLEMENTED;                                                              // SILK / hybrid: not on this path
    }
    equiv_rate = ob_compute_equiv_rate(bitrate_bps, os.stream_channels, frame_rate, cfg.vbr, 1, cfg.complexity, cfg.packet_loss);
    {   // automatic bandwidth (:1440-1490); voice and music tables differ only for WB<->SWB and SWB<->FB
        const int32_t voice_thr[8] = {9000, 700, 9000, 700, 13500, 1000, 14000, 2000}, music_thr[8] = {9000, 700, 9000, 700, 11000, 1000, 12000, 2000};
        int bandwidth = 1105;
        do {
            const int k = 2 * (bandwidth - 1102);
            int threshold = music_thr[k] + ((voice_est * voice_est * (voice_thr[k] - music_thr[k])) >> 14);
            const int hysteresis = music_thr[k + 1] + ((voice_est * voice_est * (voice_thr[k + 1] - music_thr[k + 1])) >> 14);
            if (!os.first) { if (os.auto_bandwidth >= bandwidth) threshold -= hysteresis; else threshold += hysteresis; }
            if (equiv_rate >= threshold) break;
        } while (--bandwidth > 1101);
        if (bandwidth == 1102) bandwidth = 1103;
        os.bandwidth = os.auto_bandwidth = bandwidth;
    }
    if (os.bandwidth > cfg.max_bandwidth) os.bandwidth = cfg.max_bandwidth;
    if (cfg.user_bandwidth > 0) os.bandwidth = cfg.user_bandwidth;
    if (os.detected_bandwidth && cfg.user_bandwidth <= 0) {                                                   // :1510-1530
        int min_detected;
        if (equiv_rate <= 18000 * os.stream_channels) min_detected = 1101;
        else if (equiv_rate <= 24000 * os.stream_channels) min_detected = 1102;
        else if (equiv_rate <= 30000 * os.stream_channels) min_detected = 1103;
        else if (equiv_rate <= 44000 * os.stream_channels) min_detected = 1104;
        else min_detected = 1105;
        os.detected_bandwidth = ob_imax(os.detected_bandwidth, min_detected);
        os.bandwidth = ob_imin(os.bandwidth, os.detected_bandwidth);
    }
    if (os.bandwidth == 1102) os.bandwidth = 1103;
    if (frame_size <= 960) return ob_opus_encode_frame(g, cfg, os, st, sh, wk, pcm, frame_size, data, max_data_bytes, bitrate_bps, equiv_rate, analysis_info, is_silence);

    // ---- 40-120 ms: 20 ms frames encoded one by one and repacketized into one code-1/2/3 packet (opus_encoder.c:1551-1679,
    // opus_repacketizer_out_range_impl, repacketizer.c:112-320) ----
    const int nb_frames = frame_size / 960;
    g.sync();
    if (read_pos_bak != -1 && g.lane == 0) { os.tonal->read_pos = read_pos_bak; os.tonal->read_subframe = read_subframe_bak; }     // the analysis is read one frame at a time
    g.sync();
    const int max_header_bytes = nb_frames == 2 ? 3 : (2 + (nb_frames - 1) * 2);
    const int repacketize_len = (cfg.vbr || cfg.bitrate == -1) ? out_bytes : ob_imin(cbr_bytes, out_bytes);
    const int max_len_sum = nb_frames + repacketize_len - max_header_bytes;
    if (max_len_sum > (int)sizeof(wk.multi_tmp) || max_len_sum < nb_frames) return OB_BUFFER_TOO_SMALL;
    uint8_t *curr_data = wk.multi_tmp;
    int16_t flen[6];
    int tot = 0, dtx_count = 0;
    for (int i = 0; i < nb_frames; i++) {
        int curr_max = ob_imin(3 * bitrate_bps / (3 * 8 * 48000 / 960), max_len_sum / nb_frames);
        curr_max = ob_imin(max_len_sum - tot, curr_max);
        if (read_pos_bak != -1) {
            g.sync();
            if (g.lane == 0) { ObAnalysisInfo a; ob_tonality_get_info(*os.tonal, a, 960); sh.an_tmp = a; }
            g.sync();
            analysis_info = sh.an_tmp;
            g.sync();
        }
        g.set_base(pace_base + 2048 * i);
        const int tmp_len = ob_opus_encode_frame(g, cfg, os, st, sh, wk, pcm + (size_t)i * channels * 960, 960, curr_data, curr_max, bitrate_bps, equiv_rate, analysis_info, is_silence);
        if (tmp_len < 0) return OB_INTERNAL_ERROR;
        if (tmp_len == 1) dtx_count++;
        flen[i] = (int16_t)(tmp_len - 1);                           // opus_repacketizer_cat: the frame without its TOC
        tot += tmp_len;
        curr_data += tmp_len;
    }
    g.sync();
    const int toc = wk.multi_tmp[0] & 0xFC, pad = !cfg.vbr && dtx_count != nb_frames, maxlen = repacketize_len;
    // header bytes: warp-uniform (every lane stores the same byte); frame payloads: strided copies
    uint8_t *ptr = data;
    int tot_size = 0;
    if (nb_frames == 2) {
        if (flen[1] == flen[0]) { tot_size = 2 * flen[0] + 1; if (tot_size > maxlen) return OB_INTERNAL_ERROR; *ptr++ = (uint8_t)(toc | 1); }
        else {
            tot_size = flen[0] + flen[1] + 2 + (flen[0] >= 252);
            if (tot_size > maxlen) return OB_INTERNAL_ERROR;
            *ptr++ = (uint8_t)(toc | 2);
            if (flen[0] < 252) *ptr++ = (uint8_t)flen[0];
            else { *ptr++ = (uint8_t)(252 + (flen[0] & 3)); *ptr++ = (uint8_t)((flen[0] - (252 + (flen[0] & 3))) >> 2); }
        }
    }
    if (nb_frames > 2 || (pad && tot_size < maxlen)) {                // code 3
        ptr = data;
        int vbr = 0;
        for (int i = 1; i < nb_frames; i++) if (flen[i] != flen[0]) { vbr = 1; break; }
        int hdr1;
        if (vbr) {
            tot_size = 2;
            for (int i = 0; i < nb_frames - 1; i++) tot_size += 1 + (flen[i] >= 252) + flen[i];
            tot_size += flen[nb_frames - 1];
            if (tot_size > maxlen) return OB_INTERNAL_ERROR;
            *ptr++ = (uint8_t)(toc | 3);
            hdr1 = nb_frames | 0x80;
        } else {
            tot_size = nb_frames * flen[0] + 2;
            if (tot_size > maxlen) return OB_INTERNAL_ERROR;
            *ptr++ = (uint8_t)(toc | 3);
            hdr1 = nb_frames;
        }
        const int pad_amount = pad ? (maxlen - tot_size) : 0;
        if (pad_amount != 0) hdr1 |= 0x40;
        *ptr++ = (uint8_t)hdr1;
        if (pad_amount != 0) {
            const int nb_255s = (pad_amount - 1) / 255;
            if (tot_size + nb_255s + 1 > maxlen) return OB_INTERNAL_ERROR;
            for (int i = 0; i < nb_255s; i++) *ptr++ = 255;
            *ptr++ = (uint8_t)(pad_amount - 255 * nb_255s - 1);
            tot_size += pad_amount;
        }
        if (vbr) for (int i = 0; i < nb_frames - 1; i++) {
            if (flen[i] < 252) *ptr++ = (uint8_t)flen[i];
            else { *ptr++ = (uint8_t)(252 + (flen[i] & 3)); *ptr++ = (uint8_t)((flen[i] - (252 + (flen[i] & 3))) >> 2); }
        }
    }
    {
        const uint8_t *src = wk.multi_tmp;
        for (int i = 0; i < nb_frames; i++) {
            src += 1;                                                // skip the frame's own TOC
            for (int k = g.lane; k < flen[i]; k += g.n) ptr[k] = src[k];
            ptr += flen[i];
            src += flen[i];
        }
    }
    if (pad) for (int k = g.lane; k < (int)(data + maxlen - ptr); k += g.n) ptr[k] = 0;
    g.sync();
    return tot_size;
}
