// enc_quant.cuh -- the quantisation half of the CELT encoder, ONE WARP PER STREAM.
//
// The range coder is warp-uniform (every lane carries the same coder registers; the bytes land in shared memory), so the integer
// side -- two-pass coarse energy (quant_bands.c:156-359), fine energy / finalise (:361-426), tf_encode (celt_encoder.c:756-794), the
// encoder side of the bit allocator (rate.c:248-645), the theta / split control flow of quant_all_bands (bands.c:647-1672) -- costs one
// instruction stream per stream, while everything that is a vector runs on all 32 lanes:
//   * op_pvq_search (vq.c:165-328): every lane owns N/32 positions; each greedy pulse is a per-lane best + a 5-step shuffle arg-max
//     on the reference's cross-multiplied comparison (ties to the lower index);
//   * icwrs (cwrs.c:440-456): the pulse-count suffix sums are a warp scan, the index is an integer warp reduction of table terms;
//   * exp_rotation (vq.c:47-117): the decoder's warp-scan rotation (dec_bands.cuh), run forwards;
//   * stereo_itheta / intensity_stereo / stereo_split / stereo_merge / renormalise / Haar / Hadamard: strided loops + reductions;
//   * theta RDO (bands.c:1583-1645): the two trial encodings share the warp; snapshots of X, Y, norm and the coder bytes are strided copies.
#pragma once
#include "enc_analysis.cuh"

// ---- energy quantisation ------------------------------------------------------------------------------------------------
OB_DEV float ob_loss_distortion(const float *eBands, const float *oldEBands, int end, int C)     // quant_bands.c:142-154
{
    float dist = 0;
    for (int c = 0; c < C; c++) for (int i = 0; i < end; i++) { const float d = eBands[i + c * OB_NB] - oldEBands[i + c * OB_NB]; dist = dist + d * d; }
    return ob_fmin(200, dist);
}

OB_STAGE int ob_quant_coarse_impl(int end, const float *eBands, float *oldEBands, int32_t budget, int32_t tell, const uint8_t *prob_model,
        float *error, ObRangeEnc &enc_io, int C, int LM, int intra, float max_decay)            // quant_bands.c:156-259
{
    ObRangeEnc enc = enc_io;                                      // the coder works in registers inside a stage
    int badness = 0;
    float prev[2] = {0, 0}, coef, beta;
    if (tell + 3 <= budget) enc.bit_logp(intra, 3);
    if (intra) { coef = 0; beta = OB_BETA_INTRA[0]; } else { beta = OB_BETA_COEF[LM]; coef = OB_PRED_COEF[LM]; }
    for (int i = 0; i < end; i++) for (int c = 0; c < C; c++) {
        const float x = eBands[i + c * OB_NB];
        const float oldE = ob_fmax(-9.f, oldEBands[i + c * OB_NB]);
        const float f = x - coef * oldE - prev[c];
        int qi = (int)floor((double)(.5f + f));
        const float decay_bound = ob_fmax(-28.f, oldEBands[i + c * OB_NB]) - max_decay;
        if (qi < 0 && x < decay_bound) {
            qi += (int)(decay_bound - x);
            if (qi > 0) qi = 0;
        }
        const int qi0 = qi;
        tell = enc.tell();
        const int bits_left = budget - tell - 3 * C * (end - i);
        if (i != 0 && bits_left < 30) {
            if (bits_left < 24) qi = ob_imin(1, qi);
            if (bits_left < 16) qi = ob_imax(-1, qi);
        }
        if (budget - tell >= 15) {
            const int pi = 2 * ob_imin(i, 20);
            enc.laplace(&qi, (uint32_t)prob_model[pi] << 7, (int)prob_model[pi + 1] << 6);
        } else if (budget - tell >= 2) {
            qi = ob_imax(-1, ob_imin(qi, 1));
            enc.icdf(2 * qi ^ -(qi < 0), OB_TAPSET_ICDF, 2);              // small_energy_icdf == {2,1,0}
        } else if (budget - tell >= 1) {
            qi = ob_imin(0, qi);
            enc.bit_logp(-qi, 1);
        } else qi = -1;
        error[i + c * OB_NB] = f - (float)qi;
        badness += qi0 > qi ? qi0 - qi : qi - qi0;
        const float q = (float)qi;
        const float tmp = coef * oldE + prev[c] + q;
        oldEBands[i + c * OB_NB] = tmp;
        prev[c] = prev[c] + q - beta * q;
    }
    enc_io = enc;
    return badness;
}

// quant_coarse_energy (quant_bands.c:261-359). bytes: >= 1275 bytes for the intra pass' coder bytes; fscr: >= 4 * 21 floats.
template <class G>
OB_STAGE void ob_quant_coarse_energy(const G &g, int end, int effEnd, const float *eBands, float *oldEBands, uint32_t budget, float *error, ObRangeEnc &enc,
        int C, int LM, int nbAvailableBytes, int force_intra, float *delayedIntra, int two_pass, int loss_rate, uint8_t *bytes, float *fscr)
{
    float *oldEBands_intra = fscr, *error_intra = fscr + 2 * OB_NB;
    int badness1 = 0;
    int intra = force_intra || (!two_pass && *delayedIntra > 2 * C * end && nbAvailableBytes > end * C);
    const int32_t intra_bias = (int32_t)((budget * *delayedIntra * loss_rate) / (C * 512));
    const float new_distortion = ob_loss_distortion(eBands, oldEBands, effEnd, C);
    const uint32_t tell = (uint32_t)enc.tell();
    if (tell + 3 > budget) two_pass = intra = 0;
    float max_decay = 16.f;
    if (end > 10) max_decay = ob_fmin(max_decay, .125f * nbAvailableBytes);
    const ObRangeEnc enc_start_state = enc;
    for (int i = 0; i < C * OB_NB; i++) oldEBands_intra[i] = oldEBands[i];
    if (two_pass || intra)
        badness1 = ob_quant_coarse_impl(end, eBands, oldEBands_intra, (int32_t)budget, (int32_t)tell, OB_E_PROB_MODEL + (LM * 2 + 1) * 42, error_intra, enc, C, LM, 1, max_decay);
    if (!intra) {
        const int32_t tell_intra = (int32_t)enc.tell_frac();
        const ObRangeEnc enc_intra_state = enc;
        const uint32_t nstart_bytes = enc_start_state.offs, nintra_bytes = enc_intra_state.offs;
        uint8_t *intra_buf = enc.buf + nstart_bytes;
        const uint32_t save_bytes = nintra_bytes - nstart_bytes;
        g.sync();
        for (uint32_t k = g.lane; k < save_bytes; k += g.n) bytes[k] = intra_buf[k];
        g.sync();
        enc = enc_start_state;
        const int badness2 = ob_quant_coarse_impl(end, eBands, oldEBands, (int32_t)budget, (int32_t)tell, OB_E_PROB_MODEL + (LM * 2 + intra) * 42, error, enc, C, LM, 0, max_decay);
        if (two_pass && (badness1 < badness2 || (badness1 == badness2 && ((int32_t)enc.tell_frac()) + intra_bias > tell_intra))) {
            enc = enc_intra_state;
            g.sync();
            for (uint32_t k = g.lane; k < save_bytes; k += g.n) intra_buf[k] = bytes[k];
            for (int i = 0; i < C * OB_NB; i++) { oldEBands[i] = oldEBands_intra[i]; error[i] = error_intra[i]; }
            intra = 1;
        }
    } else {
        for (int i = 0; i < C * OB_NB; i++) { oldEBands[i] = oldEBands_intra[i]; error[i] = error_intra[i]; }
    }
    if (intra) *delayedIntra = new_distortion;
    else *delayedIntra = (OB_PRED_COEF[LM] * OB_PRED_COEF[LM]) * *delayedIntra + new_distortion;
    g.sync();
}

OB_DEV void ob_quant_fine_energy(int end, float *oldEBands, float *error, const int *fine_quant, ObRangeEnc &enc, int C)   // quant_bands.c:361-395
{
    for (int i = 0; i < end; i++) {
        const int frac = 1 << fine_quant[i];
        if (fine_quant[i] <= 0) continue;
        for (int c = 0; c < C; c++) {
            int q2 = (int)floor((double)((error[i + c * OB_NB] + .5f) * frac));
            if (q2 > frac - 1) q2 = frac - 1;
            if (q2 < 0) q2 = 0;
            enc.bits((uint32_t)q2, (uint32_t)fine_quant[i]);
            const float offset = (q2 + .5f) * (1 << (14 - fine_quant[i])) * (1.f / 16384) - .5f;
            oldEBands[i + c * OB_NB] += offset;
            error[i + c * OB_NB] -= offset;
        }
    }
}

OB_DEV void ob_quant_energy_finalise(int end, float *oldEBands, float *error, const int *fine_quant, const int *fine_priority, int bits_left,
        ObRangeEnc &enc, int C)                                                                                          // quant_bands.c:397-426
{
    for (int prio = 0; prio < 2; prio++) for (int i = 0; i < end && bits_left >= C; i++) {
        if (fine_quant[i] >= 8 || fine_priority[i] != prio) continue;
        for (int c = 0; c < C; c++) {
            const int q2 = error[i + c * OB_NB] < 0 ? 0 : 1;
            enc.bits((uint32_t)q2, 1);
            const float offset = (q2 - .5f) * (1 << (14 - fine_quant[i] - 1)) * (1.f / 16384);
            oldEBands[i + c * OB_NB] += offset;
            error[i + c * OB_NB] -= offset;
            bits_left--;
        }
    }
}

OB_DEV void ob_tf_encode(int end, int isTransient, int *tf_res, int LM, int tf_select, ObRangeEnc &enc)                  // celt_encoder.c:756-794
{
    int curr = 0, tf_changed = 0, logp = isTransient ? 2 : 4;
    uint32_t budget = enc.storage * 8, tell = (uint32_t)enc.tell();
    const int tf_select_rsv = LM > 0 && tell + logp + 1 <= budget;
    budget -= tf_select_rsv;
    for (int i = 0; i < end; i++) {
        if (tell + logp <= budget) {
            enc.bit_logp(tf_res[i] ^ curr, (uint32_t)logp);
            tell = (uint32_t)enc.tell();
            curr = tf_res[i];
            tf_changed |= curr;
        } else tf_res[i] = curr;
        logp = isTransient ? 4 : 5;
    }
    if (tf_select_rsv && OB_TF_SELECT[LM * 8 + 4 * isTransient + 0 + tf_changed] != OB_TF_SELECT[LM * 8 + 4 * isTransient + 2 + tf_changed])
        enc.bit_logp(tf_select, 1);
    else tf_select = 0;
    for (int i = 0; i < end; i++) tf_res[i] = OB_TF_SELECT[LM * 8 + 4 * isTransient + 2 * tf_select + tf_res[i]];
}

// ---- bit allocation, encoder side (rate.c:248-645): integer, warp-uniform.  iscr: >= 5 * 21 ints ------------------------------------
OB_STAGE int ob_enc_allocation(ObRangeEnc &ec_io, int end, const int *offsets, const int *cap, int alloc_trim, int *intensity, int *dual_stereo,
        int32_t total, int32_t *balance_out, int *pulses, int *ebits, int *fine_priority, int C, int LM, int prev, int signalBandwidth, int *iscr)
{
    ObRangeEnc ec = ec_io;
    const int start = 0;
    int *bits1 = iscr, *bits2 = iscr + OB_NB, *thresh = iscr + 2 * OB_NB, *trim_offset = iscr + 3 * OB_NB, *bits = iscr + 4 * OB_NB;
    int lo, hi, j, skip_start = start, skip_rsv, intensity_rsv = 0, dual_stereo_rsv = 0;
    total = ob_imax(total, 0);
    skip_rsv = total >= 1 << OB_BITRES ? 1 << OB_BITRES : 0;
    total -= skip_rsv;
    if (C == 2) {
        intensity_rsv = OB_LOG2_FRAC[end - start];
        if (intensity_rsv > total) intensity_rsv = 0;
        else { total -= intensity_rsv; dual_stereo_rsv = total >= 1 << OB_BITRES ? 1 << OB_BITRES : 0; total -= dual_stereo_rsv; }
    }
    for (j = start; j < end; j++) {
        const int w = OB_EBANDS[j + 1] - OB_EBANDS[j];
        thresh[j] = ob_imax(C << OB_BITRES, (3 * w << LM << OB_BITRES) >> 4);
        trim_offset[j] = C * w * (alloc_trim - 5 - LM) * (end - j - 1) * (1 << (LM + OB_BITRES)) >> 6;
        if (w << LM == 1) trim_offset[j] -= C << OB_BITRES;
    }
    lo = 1; hi = 11 - 1;
    do {
        int done = 0, psum = 0, mid = (lo + hi) >> 1;
        for (j = end; j-- > start;) {
            const int N = OB_EBANDS[j + 1] - OB_EBANDS[j];
            int bitsj = C * N * OB_ALLOC_VECTORS[mid * OB_NB + j] << LM >> 2;
            if (bitsj > 0) bitsj = ob_imax(0, bitsj + trim_offset[j]);
            bitsj += offsets[j];
            if (bitsj >= thresh[j] || done) { done = 1; psum += ob_imin(bitsj, cap[j]); }
            else if (bitsj >= C << OB_BITRES) psum += C << OB_BITRES;
        }
        if (psum > total) hi = mid - 1; else lo = mid + 1;
    } while (lo <= hi);
    hi = lo--;
    for (j = start; j < end; j++) {
        const int N = OB_EBANDS[j + 1] - OB_EBANDS[j];
        int b1 = C * N * OB_ALLOC_VECTORS[lo * OB_NB + j] << LM >> 2;
        int b2 = hi >= 11 ? cap[j] : C * N * OB_ALLOC_VECTORS[hi * OB_NB + j] << LM >> 2;
        if (b1 > 0) b1 = ob_imax(0, b1 + trim_offset[j]);
        if (b2 > 0) b2 = ob_imax(0, b2 + trim_offset[j]);
        if (lo > 0) b1 += offsets[j];
        b2 += offsets[j];
        if (offsets[j] > 0) skip_start = j;
        b2 = ob_imax(0, b2 - b1);
        bits1[j] = b1; bits2[j] = b2;
    }
    int32_t psum, left, percoeff, balance;
    int i, coded, done;
    const int alloc_floor = C << OB_BITRES, stereo = C > 1, logM = LM << OB_BITRES;
    lo = 0; hi = 1 << 6;
    for (i = 0; i < 6; i++) {
        const int mid = (lo + hi) >> 1;
        psum = 0; done = 0;
        for (j = end; j-- > start;) {
            const int tmp = bits1[j] + (mid * (int32_t)bits2[j] >> 6);
            if (tmp >= thresh[j] || done) { done = 1; psum += ob_imin(tmp, cap[j]); }
            else if (tmp >= alloc_floor) psum += alloc_floor;
        }
        if (psum > total) hi = mid; else lo = mid;
    }
    psum = 0; done = 0;
    for (j = end; j-- > start;) {
        int tmp = bits1[j] + ((int32_t)lo * bits2[j] >> 6);
        if (tmp < thresh[j] && !done) tmp = tmp >= alloc_floor ? alloc_floor : 0;
        else done = 1;
        tmp = ob_imin(tmp, cap[j]);
        bits[j] = tmp;
        psum += tmp;
    }
    for (coded = end;; coded--) {
        int band_width, band_bits, rem;
        j = coded - 1;
        if (j <= skip_start) { total += skip_rsv; break; }
        left = total - psum;
        percoeff = (int32_t)((uint32_t)left / (uint32_t)(OB_EBANDS[coded] - OB_EBANDS[start]));
        left -= (OB_EBANDS[coded] - OB_EBANDS[start]) * percoeff;
        rem = ob_imax(left - (OB_EBANDS[j] - OB_EBANDS[start]), 0);
        band_width = OB_EBANDS[coded] - OB_EBANDS[j];
        band_bits = (int)(bits[j] + percoeff * band_width + rem);
        if (band_bits >= ob_imax(thresh[j], alloc_floor + (1 << OB_BITRES))) {
            int depth_threshold;
            if (coded > 17) depth_threshold = j < prev ? 7 : 9; else depth_threshold = 0;
            if (coded <= start + 2 || (band_bits > (depth_threshold * band_width << LM << OB_BITRES) >> 4 && j <= signalBandwidth)) {
                ec.bit_logp(1, 1);
                break;
            }
            ec.bit_logp(0, 1);
            psum += 1 << OB_BITRES;
            band_bits -= 1 << OB_BITRES;
        }
        psum -= bits[j] + intensity_rsv;
        if (intensity_rsv > 0) intensity_rsv = OB_LOG2_FRAC[j - start];
        psum += intensity_rsv;
        if (band_bits >= alloc_floor) { psum += alloc_floor; bits[j] = alloc_floor; }
        else bits[j] = 0;
    }
    if (intensity_rsv > 0) {
        *intensity = ob_imin(*intensity, coded);
        ec.uint((uint32_t)(*intensity - start), (uint32_t)(coded + 1 - start));
    } else *intensity = 0;
    if (*intensity <= start) { total += dual_stereo_rsv; dual_stereo_rsv = 0; }
    if (dual_stereo_rsv > 0) ec.bit_logp(*dual_stereo, 1);
    else *dual_stereo = 0;
    left = total - psum;
    percoeff = (int32_t)((uint32_t)left / (uint32_t)(OB_EBANDS[coded] - OB_EBANDS[start]));
    left -= (OB_EBANDS[coded] - OB_EBANDS[start]) * percoeff;
    for (j = start; j < coded; j++) bits[j] += (int)percoeff * (OB_EBANDS[j + 1] - OB_EBANDS[j]);
    for (j = start; j < coded; j++) { const int tmp = ob_imin(left, OB_EBANDS[j + 1] - OB_EBANDS[j]); bits[j] += tmp; left -= tmp; }
    balance = 0;
    for (j = start; j < coded; j++) {
        const int N0 = OB_EBANDS[j + 1] - OB_EBANDS[j], N = N0 << LM;
        int den, offset, NClogN;
        int32_t excess, bit = bits[j] + balance;
        if (N > 1) {
            excess = ob_imax(bit - cap[j], 0);
            bits[j] = bit - excess;
            den = C * N + ((C == 2 && N > 2 && !*dual_stereo && j < *intensity) ? 1 : 0);
            NClogN = den * (OB_LOGN[j] + logM);
            offset = (NClogN >> 1) - den * 21;
            if (N == 2) offset += den << OB_BITRES >> 2;
            if (bits[j] + offset < den * 2 << OB_BITRES) offset += NClogN >> 2;
            else if (bits[j] + offset < den * 3 << OB_BITRES) offset += NClogN >> 3;
            ebits[j] = ob_imax(0, bits[j] + offset + (den << (OB_BITRES - 1)));
            ebits[j] = (int)((uint32_t)ebits[j] / (uint32_t)den) >> OB_BITRES;
            if (C * ebits[j] > (bits[j] >> OB_BITRES)) ebits[j] = bits[j] >> stereo >> OB_BITRES;
            ebits[j] = ob_imin(ebits[j], 8);
            fine_priority[j] = ebits[j] * (den << OB_BITRES) >= bits[j] + offset;
            bits[j] -= C * ebits[j] << OB_BITRES;
        } else {
            excess = ob_imax(0, bit - (C << OB_BITRES));
            bits[j] = bit - excess;
            ebits[j] = 0;
            fine_priority[j] = 1;
        }
        if (excess > 0) {
            const int extra_fine = ob_imin(excess >> (stereo + OB_BITRES), 8 - ebits[j]);
            ebits[j] += extra_fine;
            const int extra_bits = extra_fine * C << OB_BITRES;
            fine_priority[j] = extra_bits >= excess - balance;
            excess -= extra_bits;
        }
        balance = excess;
    }
    *balance_out = balance;
    for (; j < end; j++) {
        ebits[j] = bits[j] >> stereo >> OB_BITRES;
        bits[j] = 0;
        fine_priority[j] = ebits[j] < 1;
    }
    for (j = 0; j < end; j++) pulses[j] = bits[j];
    ec_io = ec;
    return coded;
}

// ---- PVQ ---------------------------------------------------------------------------------------------------------------------
// exp_rotation (vq.c:74-117), both directions, on the decoder's rotation passes (dec_bands.cuh: one chain per lane, or -- one long chain --
// the warp scan ob_rot_scan).  ObSoloW reproduces the scan's arithmetic on the host.
#ifndef __CUDACC__
static inline void ob_rot_scan_emul(float *x, int L, float c, float s, int rev)
{
    const float s2 = s * s, s4 = s2 * s2, s8 = s4 * s4, s16 = s8 * s8;
    float slane[32];
    for (int lane = 0; lane < 32; lane++) {
        float v = s;
        if (lane & 1) v *= s;
        if (lane & 2) v *= s2;
        if (lane & 4) v *= s4;
        if (lane & 8) v *= s8;
        if (lane & 16) v *= s16;
        slane[lane] = v;
    }
    const float mul[5] = {s, s2, s4, s8, s16};
    float carry = 0.f;
    for (int base = 0; base < L; base += 32) {
        float a[32], an[32], y[32];
        for (int lane = 0; lane < 32; lane++) {
            const int t = base + lane;
            a[lane] = t < L ? x[rev ? L - 1 - t : t] : 0.f;
            an[lane] = t + 1 < L ? x[rev ? L - 2 - t : t + 1] : 0.f;
            y[lane] = t == 0 ? a[lane] : c * a[lane];
        }
        for (int k = 0; k < 5; k++) {
            const int o = 1 << k;
            float q[32];
            for (int lane = 0; lane < 32; lane++) q[lane] = lane >= o ? y[lane] + mul[k] * y[lane - o] : y[lane];
            for (int lane = 0; lane < 32; lane++) y[lane] = q[lane];
        }
        for (int lane = 0; lane < 32; lane++) y[lane] = y[lane] + slane[lane] * carry;
        carry = y[31];
        for (int lane = 0; lane < 32; lane++) {
            const int t = base + lane;
            if (t < L) x[rev ? L - 1 - t : t] = (t + 1 < L) ? c * y[lane] - s * an[lane] : y[lane];
        }
    }
}
static inline void ob_rot_pass(const ObSoloW &, float *X, int nblocks, int len, int stride, float c, float s)
{
    if (nblocks * stride == 1 && len >= 12) { ob_rot_scan_emul(X, len, c, s, 0); ob_rot_scan_emul(X, len - 1, c, -s, 1); return; }
    ob_rot_pass(ObSolo(), X, nblocks, len, stride, c, s);
}
#endif
template <class G>
OB_STAGE void ob_exp_rotation(const G &g, float *X, int len, int dir, int stride, int K, int spread)
{
    if (2 * K >= len || spread == 0) return;
    const int factor = spread == 1 ? 15 : spread == 2 ? 10 : 5;
    const float gain = (float)(1.0f * len) / (float)(len + factor * K);
    const float theta = .5f * (gain * gain);
    float c, s;
    int stride2 = 0;
#ifdef __CUDA_ARCH__
    if (G::n == 32) {
        // as in the decoder (dec_bands.cuh ob_exp_rotation_inv): the two cosines on two lanes at once, the stride search as one vote
        const float v = OB_COSF((.5f * 3.141592653f) * ((g.lane & 1) ? 1.0f - theta : theta));
        c = __shfl_sync(0xffffffffu, v, 0); s = __shfl_sync(0xffffffffu, v, 1);
        if (len >= 8 * stride) {
            const int k = g.lane + 1;
            stride2 = 1 + __popc(__ballot_sync(0xffffffffu, (k * k + k) * stride + (stride >> 2) < len));
        }
    } else
#endif
    {
        c = OB_COSF((.5f * 3.141592653f) * theta);
        s = OB_COSF((.5f * 3.141592653f) * (1.0f - theta));
        if (len >= 8 * stride) { stride2 = 1; while ((stride2 * stride2 + stride2) * stride + (stride >> 2) < len) stride2++; }
    }
    len >>= ob_log2i(stride);                                         // stride = short blocks spanned: a power of two
    // dir < 0: (stride2: s, c) then (1: c, s);  dir > 0: (1: c, -s) then (stride2: s, -c) -- one inlined copy of the pass for all four
#ifdef __CUDACC__
#pragma unroll 1
#endif
    for (int p = 0; p < 2; p++) {
        const int second = dir < 0 ? p : 1 - p;                       // is this the stride-1 pass?
        if (!second && !stride2) continue;
        const float sg = dir < 0 ? 1.f : -1.f;
        ob_rot_pass(g, X, stride, len, second ? 1 : stride2, second ? c : s, sg * (second ? s : c));
    }
}

// One greedy pulse of op_pvq_search (vq.c:287-320): arg max over j of (xy + |X[j]|)^2 / (yy + y[j]).
// ObSolo: the reference's scan, compared cross-multiplied, first maximum wins.  Warp: every lane finds the best of its own positions the same
// way, then the 32 lane winners are ranked by their quotient (one IEEE division per lane) with two integer warp reductions -- max of the
// quotient's bit pattern, then min position among the lanes that hold it -- instead of a 5-step shuffle tournament on (num, den, j).
struct ObPvqBest { float num, den; int j; };
OB_DEV int ob_pvq_argmax(const ObSolo &, const float *X, const float *y, int N, float xy, float yy)
{
    int best_id = 0;
    float Rxy = xy + fabsf(X[0]), Ryy = yy + y[0];
    Rxy = Rxy * Rxy;
    float best_den = Ryy, best_num = Rxy;
    for (int j = 1; j < N; j++) {
        Rxy = xy + fabsf(X[j]);
        Ryy = yy + y[j];
        Rxy = Rxy * Rxy;
        if (best_den * Rxy > Ryy * best_num) { best_den = Ryy; best_num = Rxy; best_id = j; }
    }
    return best_id;
}
OB_DEV ObPvqBest ob_pvq_lane_best(const float *X, const float *y, int N, float xy, float yy, int lane, int step)
{
    ObPvqBest b; b.num = -1.f; b.den = 1.f; b.j = 0x7fffffff;
    for (int j = lane; j < N; j += step) {
        float Rxy = xy + fabsf(X[j]);
        const float Ryy = yy + y[j];
        Rxy = Rxy * Rxy;
        if (b.den * Rxy > Ryy * b.num) { b.den = Ryy; b.num = Rxy; b.j = j; }
    }
    return b;
}
// rank key of a lane winner: 0 for a lane without positions, else 1 + the bit pattern of the (non-negative) quotient
OB_DEV uint32_t ob_pvq_key(const ObPvqBest &b)
{
    if (b.j == 0x7fffffff) return 0u;
    const float q = b.num / b.den;
    uint32_t u;
#ifdef __CUDACC__
    u = __float_as_uint(q);
#else
    memcpy(&u, &q, 4);
#endif
    return (u & 0x80000000u) ? 1u : u + 1u;                       // NaN / negative cannot happen for finite input; keep the key ordered anyway
}
#ifdef __CUDACC__
OB_DEV int ob_pvq_argmax(const ObWarp &g, const float *X, const float *y, int N, float xy, float yy)
{
    const ObPvqBest b = ob_pvq_lane_best(X, y, N, xy, yy, g.lane, 32);
    const uint32_t key = ob_pvq_key(b), top = __reduce_max_sync(0xffffffffu, key);
    return (int)__reduce_min_sync(0xffffffffu, key == top ? (uint32_t)b.j : 0x7fffffffu);
}
#else
static inline int ob_pvq_argmax(const ObSoloW &, const float *X, const float *y, int N, float xy, float yy)
{
    uint32_t top = 0; int best = 0x7fffffff;
    for (int l = 0; l < 32; l++) {
        const ObPvqBest b = ob_pvq_lane_best(X, y, N, xy, yy, l, 32);
        const uint32_t key = ob_pvq_key(b);
        if (key > top || (key == top && b.j < best)) { top = key; best = b.j; }
    }
    return best;
}
#endif

// op_pvq_search_c (vq.c:165-328).  X keeps its signs (|X| is taken on the fly).  y: N floats of scratch.  Returns yy.
template <class G>
OB_STAGE float ob_pvq_search(const G &g, float *X, int *iy, float *y, int K, int N)
{
    float xy = 0, yy = 0;
    int pulsesLeft = K;
    OB_ROLLED_G
    for (int j = g.lane; j < N; j += g.n) { iy[j] = 0; y[j] = 0; }
    g.sync();
    if (K > (N >> 1)) {
        float sum = ob_psum(g, N, 0.f, [&](int j) { return fabsf(X[j]); });
        if (!(sum > 1e-15f && sum < 64)) {
            g.sync();
            OB_ROLLED_G
            for (int j = g.lane; j < N; j += g.n) X[j] = j == 0 ? (X[0] < 0 ? -1.f : 1.f) : 0.f;
            g.sync();
            sum = 1.f;
        }
        const float rcp = (K + 0.8f) * (1.f / sum);
        OB_ROLLED_G
        for (int j = g.lane; j < N; j += g.n) { const int v = (int)floor((double)(rcp * fabsf(X[j]))); iy[j] = v; y[j] = (float)v; }
        g.sync();
        ob_psum2(g, N, yy, xy, [&](int j, float &a, float &b) { a = a + y[j] * y[j]; b = b + fabsf(X[j]) * y[j]; });
        pulsesLeft -= (int)ob_psum_u32(g, N, [&](int j) { return (uint32_t)iy[j]; });
        g.sync();
        OB_ROLLED_G
        for (int j = g.lane; j < N; j += g.n) y[j] *= 2;
        g.sync();
    }
    if (pulsesLeft > N + 3) {
        const float tmp = (float)pulsesLeft;
        yy = yy + tmp * tmp;
        yy = yy + tmp * y[0];
        g.sync();
        if (g.lane == 0) iy[0] += pulsesLeft;
        pulsesLeft = 0;
        g.sync();
    }
    for (int i = 0; i < pulsesLeft; i++) {
        yy = yy + 1;
        const int best_id = ob_pvq_argmax(g, X, y, N, xy, yy);
        xy = xy + fabsf(X[best_id]);
        yy = yy + y[best_id];
        g.sync();
        if (g.lane == 0) { y[best_id] += 2; iy[best_id]++; }
        g.sync();
    }
    OB_ROLLED_G
    for (int j = g.lane; j < N; j += g.n) if (X[j] < 0) iy[j] = -iy[j];
    g.sync();
    return yy;
}

// icwrs (cwrs.c:440-456): index = [y_{n-1} < 0] + sum over j < n-1 of U(n-j, k_j) + [y_j < 0] U(n-j, k_j + |y_j| + 1), k_j = sum_{t > j} |y_t|
OB_DEV uint32_t ob_icwrs_serial(int n, const int *y)
{
    int j = n - 1, k = y[j] < 0 ? -y[j] : y[j];
    uint32_t i = y[j] < 0;
    do {
        j--;
        i += ob_pvq_u(n - j, k);
        k += y[j] < 0 ? -y[j] : y[j];
        if (y[j] < 0) i += ob_pvq_u(n - j, k + 1);
    } while (j > 0);
    return i;
}
OB_DEV uint32_t ob_icwrs(const ObSolo &, int n, const int *y) { return ob_icwrs_serial(n, y); }
#ifdef __CUDACC__
OB_DEV uint32_t ob_icwrs(const ObWarp &g, int n, const int *y)
{
    // lane l owns positions [lo, hi) (ascending chunks); suffix sums run from the high end
    const int L = (n + 31) >> 5, lo = ob_imin(n, g.lane * L), hi = ob_imin(n, lo + L);
    int own = 0;
    for (int j = lo; j < hi; j++) own += y[j] < 0 ? -y[j] : y[j];
    int suf = own;                                                   // inclusive suffix over lanes >= this one
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int p = __shfl_down_sync(0xffffffffu, suf, o); if (g.lane + o < 32) suf += p; }
    int k = suf - own;                                               // pulses strictly above this lane's chunk
    uint32_t i = 0;
    for (int j = hi - 1; j >= lo; j--) {
        const int a = y[j] < 0 ? -y[j] : y[j];
        if (j == n - 1) i += y[j] < 0;
        else { i += ob_pvq_u(n - j, k); if (y[j] < 0) i += ob_pvq_u(n - j, k + a + 1); }
        k += a;
    }
    return __reduce_add_sync(0xffffffffu, i);
}
#else
static inline uint32_t ob_icwrs(const ObSoloW &, int n, const int *y) { return ob_icwrs_serial(n, y); }
#endif

template <class G>
OB_DEV uint32_t ob_collapse_mask(const G &g, const int *iy, int N, int B)                      // vq.c:143-163
{
    if (B <= 1) return 1;
    const ObDiv dv = ob_div_make(N / B);
    return ob_por_u32(g, N, [&](int j) { return iy[j] != 0 ? 1u << ob_div(j, dv) : 0u; });
}

// Per-warp working set of quant_all_bands(encode = 1) in SHARED memory: the band being coded, the folding source and the search scratch.
struct ObEncPartFrame {
    float *X, *lowband;
    int N, b, B, B0, LM, fill, mbits, sbits, itheta, stage, mid_first;
    int32_t rebalance;
    float gain, mid, side;
    uint32_t cm;
};
struct ObEncBandsShared {
    float xb[2 * OB_MAX_BAND];                // the band: X at [0,176), Y at [176,352)
    float lowband_scratch[OB_MAX_BAND];
    float tmp[OB_MAX_BAND];                   // Hadamard permutation buffer
    float y[OB_MAX_BAND];                     // op_pvq_search: 2 * pulses per position
    int iy[OB_MAX_BAND];
    ObEncPartFrame part[6];
};
// ... and in GLOBAL memory (per warp slot): the theta-RDO snapshots, touched by strided copies only
struct ObEncBandsWork {
    float norm[2 * OB_NORM_LEN];              // folding source per channel (bands.c:1438): written and read with strided loops (theta RDO only)
    float X_save[OB_MAX_BAND], Y_save[OB_MAX_BAND], X_save2[OB_MAX_BAND], Y_save2[OB_MAX_BAND], norm_save2[OB_MAX_BAND];
    uint8_t bytes_save[1280];
};

// alg_quant (vq.c:330-359)
template <class G>
OB_DEV uint32_t ob_alg_quant(const G &g, ObEncBandsShared &S, float *X, int N, int K, int spread, int B, ObRangeEnc &enc, float gain, int resynth)
{
    int *iy = S.iy;
    ob_exp_rotation(g, X, N, 1, B, K, spread);
    const float yy = ob_pvq_search(g, X, iy, S.y, K, N);
    enc.uint(ob_icwrs(g, N, iy), ob_pvq_v(N, K));
    if (resynth) {
        const float gg = (1.f / sqrtf(yy)) * gain;                      // normalise_residual (vq.c:121-141)
        OB_ROLLED_G
        for (int i = g.lane; i < N; i += g.n) X[i] = gg * (float)iy[i];
        g.sync();
        ob_exp_rotation(g, X, N, -1, B, K, spread);
    }
    return ob_collapse_mask(g, iy, N, B);
}

template <class G>
OB_DEV void ob_renormalise(const G &g, float *X, int N, float gain)                              // vq.c:383-407
{
    const float E = 1e-15f + ob_psum(g, N, 0.f, [&](int i) { return X[i] * X[i]; });
    const float gg = (1.f / sqrtf(E)) * gain;
    OB_ROLLED_G
    for (int i = g.lane; i < N; i += g.n) X[i] = gg * X[i];
    g.sync();
}

// stereo_itheta (vq.c:410-441) with fast_atan2f (mathops.h:54-73)
OB_DEV float ob_fast_atan2f(float y, float x)
{
    const float cA = 0.43157974f, cB = 0.67848403f, cC = 0.08595542f, cE = 3.141592653f / 2;
    const float x2 = x * x, y2 = y * y;
    if (x2 + y2 < 1e-18f) return 0;
    if (x2 < y2) { const float den = (y2 + cB * x2) * (y2 + cC * x2); return -x * y * (y2 + cA * x2) / den + (y < 0 ? -cE : cE); }
    else { const float den = (x2 + cB * y2) * (x2 + cC * y2); return x * y * (x2 + cA * y2) / den + (y < 0 ? -cE : cE) - (x * y < 0 ? -cE : cE); }
}
template <class G>
OB_DEV int ob_stereo_itheta(const G &g, const float *X, const float *Y, int stereo, int N)
{
    float Emid = 1e-15f, Eside = 1e-15f;
    if (stereo) {
        ob_psum2(g, N, Emid, Eside, [&](int i, float &a, float &b) { const float m = X[i] + Y[i], s = X[i] - Y[i]; a = a + m * m; b = b + s * s; });
    } else {
        Emid += ob_psum(g, N, 0.f, [&](int i) { return X[i] * X[i]; });
        Eside += ob_psum(g, N, 0.f, [&](int i) { return Y[i] * Y[i]; });
    }
    const float mid = sqrtf(Emid), side = sqrtf(Eside);
    return (int)floor((double)(.5f + 16384 * 0.63662f * ob_fast_atan2f(side, mid)));
}

// ---- band encoding (bands.c) ----------------------------------------------------------------------------------------------------
struct ObEncBandCtx {
    ObRangeEnc *ec;
    const float *bandE;
    int band, intensity, spread, tf_change, disable_inv, resynth, theta_round, avoid_split_noise;
    int32_t remaining_bits;
    uint32_t seed;
    int pace;                                 // ObWarpPaced: stage id of the next pace point inside the band being coded
};

template <class G>
OB_DEV void ob_intensity_stereo(const G &g, float *X, const float *Y, const float *bandE, int i, int N)      // bands.c:388-410
{
    const float left = bandE[i], right = bandE[i + OB_NB];
    const float norm = 1e-15f + sqrtf(1e-15f + left * left + right * right);
    const float a1 = left / norm, a2 = right / norm;
    OB_ROLLED_G
    for (int j = g.lane; j < N; j += g.n) X[j] = a1 * X[j] + a2 * Y[j];
    g.sync();
}
template <class G>
OB_DEV void ob_stereo_split(const G &g, float *X, float *Y, int N)                                           // bands.c:412-424
{
    OB_ROLLED_G
    for (int j = g.lane; j < N; j += g.n) { const float l = .70710678f * X[j], r = .70710678f * Y[j]; X[j] = l + r; Y[j] = r - l; }
    g.sync();
}
template <class G>
OB_DEV void ob_enc_stereo_merge(const G &g, float *X, float *Y, float mid, int N)                            // bands.c:426-476
{
    float xp = 0, side = 0;
    ob_psum2(g, N, xp, side, [&](int j, float &a, float &b) { a = a + Y[j] * X[j]; b = b + Y[j] * Y[j]; });
    xp = mid * xp;
    const float El = mid * mid + side - 2 * xp, Er = mid * mid + side + 2 * xp;
    if (Er < 6e-4f || El < 6e-4f) { for (int j = g.lane; j < N; j += g.n) Y[j] = X[j]; g.sync(); return; }
    const float lgain = 1.f / sqrtf(El), rgain = 1.f / sqrtf(Er);
    OB_ROLLED_G
    for (int j = g.lane; j < N; j += g.n) { const float l = mid * X[j], r = Y[j]; X[j] = lgain * (l - r); Y[j] = rgain * (l + r); }
    g.sync();
}

// compute_theta (bands.c:700-903), encoder + decoder semantics of the encoder build (encode = 1)
template <class G>
OB_DEV void ob_enc_theta(const G &g, ObEncBandCtx &ctx, ObSplit &sp, float *X, float *Y, int N, int *b, int B, int B0, int LM, int stereo, int *fill)
{
    ObRangeEnc &ec = *ctx.ec;
    int itheta, inv = 0, imid, iside, delta, qn;
    const int i = ctx.band;
    const int pulse_cap = OB_LOGN[i] + LM * (1 << OB_BITRES);
    const int offset = (pulse_cap >> 1) - (stereo && N == 2 ? 16 : 4);
    qn = ob_compute_qn(N, *b, offset, pulse_cap, stereo);
    if (stereo && i >= ctx.intensity) qn = 1;
    itheta = ob_stereo_itheta(g, X, Y, stereo, N);
    const int32_t tell = (int32_t)ec.tell_frac();
    if (qn != 1) {
        if (!stereo || ctx.theta_round == 0) {
            itheta = (itheta * (int32_t)qn + 8192) >> 14;
            if (!stereo && ctx.avoid_split_noise && itheta > 0 && itheta < qn) {
                const int unquantized = (int)((uint32_t)(itheta * 16384) / (uint32_t)qn);
                imid = ob_bitexact_cos((int16_t)unquantized);
                iside = ob_bitexact_cos((int16_t)(16384 - unquantized));
                delta = ob_frac_mul16((N - 1) << 7, ob_bitexact_log2tan(iside, imid));
                if (delta > *b) itheta = qn;
                else if (delta < -*b) itheta = 0;
            }
        } else {
            const int bias = itheta > 8192 ? 32767 / qn : -32767 / qn;
            const int down = ob_imin(qn - 1, ob_imax(0, (itheta * (int32_t)qn + bias) >> 14));
            itheta = ctx.theta_round < 0 ? down : down + 1;
        }
        if (stereo && N > 2) {
            const int p0 = 3, x = itheta, x0 = qn / 2, ft = p0 * (x0 + 1) + x0;
            ec.encode((uint32_t)(x <= x0 ? p0 * x : (x - 1 - x0) + (x0 + 1) * p0), (uint32_t)(x <= x0 ? p0 * (x + 1) : (x - x0) + (x0 + 1) * p0), (uint32_t)ft);
        } else if (B0 > 1 || stereo) {
            ec.uint((uint32_t)itheta, (uint32_t)qn + 1);
        } else {
            const int ft = ((qn >> 1) + 1) * ((qn >> 1) + 1);
            const int fs = itheta <= (qn >> 1) ? itheta + 1 : qn + 1 - itheta;
            const int fl = itheta <= (qn >> 1) ? itheta * (itheta + 1) >> 1 : ft - ((qn + 1 - itheta) * (qn + 2 - itheta) >> 1);
            ec.encode((uint32_t)fl, (uint32_t)(fl + fs), (uint32_t)ft);
        }
        itheta = (int)((uint32_t)(itheta * 16384) / (uint32_t)qn);
        if (stereo) {
            if (itheta == 0) ob_intensity_stereo(g, X, Y, ctx.bandE, i, N);
            else ob_stereo_split(g, X, Y, N);
        }
    } else if (stereo) {
        inv = itheta > 8192 && !ctx.disable_inv;
        if (inv) { for (int j = g.lane; j < N; j += g.n) Y[j] = -Y[j]; g.sync(); }
        ob_intensity_stereo(g, X, Y, ctx.bandE, i, N);
        if (*b > 2 << OB_BITRES && ctx.remaining_bits > 2 << OB_BITRES) ec.bit_logp(inv, 2);
        else inv = 0;
        if (ctx.disable_inv) inv = 0;
        itheta = 0;
    }
    sp.qalloc = (int)((int32_t)ec.tell_frac() - tell);
    *b -= sp.qalloc;
    if (itheta == 0) { imid = 32767; iside = 0; *fill &= (1 << B) - 1; delta = -16384; }
    else if (itheta == 16384) { imid = 0; iside = 32767; *fill &= ((1 << B) - 1) << B; delta = 16384; }
    else {
        imid = ob_bitexact_cos((int16_t)itheta);
        iside = ob_bitexact_cos((int16_t)(16384 - itheta));
        delta = ob_frac_mul16((N - 1) << 7, ob_bitexact_log2tan(iside, imid));
    }
    sp.inv = inv; sp.imid = imid; sp.iside = iside; sp.delta = delta; sp.itheta = itheta;
}

// quant_partition (bands.c:943-1105), encode = 1.  The reference recurses (depth <= LM+1 <= 4); here the recursion is an explicit stack
// (in shared memory, warp-uniform) whose leaves are the cooperative stages: PVQ search + index, or noise / folding fill + renormalise.
template <class G>
OB_DEV uint32_t ob_enc_partition(const G &g, ObEncBandsShared &S, ObEncBandCtx &ctx, float *X0, int N0, int b0, int Bin, float *lowband0, int LM0, float gain0, int fill0)
{
    ObEncPartFrame *st = S.part;
    int sp = 0;
    uint32_t ret = 0;
    st[0].X = X0; st[0].lowband = lowband0; st[0].N = N0; st[0].b = b0; st[0].B = Bin; st[0].LM = LM0; st[0].gain = gain0; st[0].fill = fill0; st[0].stage = 0;
    while (sp >= 0) {
        ObEncPartFrame &f = st[sp];
        if (f.stage == 0) {
            const uint8_t *cache = ob_pcache(ctx.band, f.LM);
            if (f.LM != -1 && f.b > cache[cache[0]] + 12 && f.N > 2) {
                ObSplit s;
                const int n = f.N >> 1, lm = f.LM - 1;
                int bb = f.b, fl = f.fill;
                f.B0 = f.B;
                if (f.B == 1) fl = (fl & 1) | (fl << 1);
                const int B1 = (f.B + 1) >> 1;
                ob_enc_theta(g, ctx, s, f.X, f.X + n, n, &bb, B1, f.B0, lm, 0, &fl);
                int delta = s.delta;
                f.mid = (1.f / 32768) * s.imid;
                f.side = (1.f / 32768) * s.iside;
                if (f.B0 > 1 && (s.itheta & 0x3fff)) {
                    if (s.itheta > 8192) delta -= delta >> (4 - lm);
                    else delta = ob_imin(0, delta + (n << OB_BITRES >> (5 - lm)));
                }
                f.mbits = ob_imax(0, ob_imin(bb, (bb - delta) / 2));
                f.sbits = bb - f.mbits;
                ctx.remaining_bits -= s.qalloc;
                f.rebalance = ctx.remaining_bits;
                f.itheta = s.itheta;
                f.fill = fl; f.N = n; f.LM = lm; f.B = B1;              // from here on the frame describes its two children
                f.mid_first = f.mbits >= f.sbits;
                f.stage = 1;
                ObEncPartFrame &c = st[++sp];
                c.N = n; c.B = B1; c.LM = lm; c.stage = 0;
                if (f.mid_first) { c.X = f.X; c.lowband = f.lowband; c.b = f.mbits; c.gain = f.gain * f.mid; c.fill = fl; }
                else { c.X = f.X + n; c.lowband = f.lowband ? f.lowband + n : nullptr; c.b = f.sbits; c.gain = f.gain * f.side; c.fill = fl >> B1; }
            } else {
                uint32_t cm = 0;
                int q = ob_bits2pulses(cache, f.b);
                int curr_bits = ob_pulses2bits(cache, q);
                ctx.remaining_bits -= curr_bits;
                while (ctx.remaining_bits < 0 && q > 0) {
                    ctx.remaining_bits += curr_bits;
                    q--;
                    curr_bits = ob_pulses2bits(cache, q);
                    ctx.remaining_bits -= curr_bits;
                }
                float *X = f.X;
                const int N = f.N;
                if ((ctx.pace & 63) < 62) g.pace(ctx.pace++);       // the block's warps start their k-th leaf of this band together
                if (q != 0) {
                    cm = ob_alg_quant(g, S, X, N, ob_get_pulses(q), ctx.spread, f.B, *ctx.ec, f.gain, ctx.resynth);
                } else if (ctx.resynth) {
                    const uint32_t cm_mask = (1u << f.B) - 1;
                    const int fl = f.fill & (int)cm_mask;
                    if (!fl) { for (int j = g.lane; j < N; j += g.n) X[j] = 0; g.sync(); }
                    else {
                        // noise / folded spectrum (bands.c:1070-1098): sample j uses LCG state seed * a^(j+1) + ..., reached by a closed-form jump per lane
                        const float *lowband = f.lowband;
                        const ObLcg first = ob_lcg_pow((uint32_t)g.lane + 1u), step = ob_lcg_pow((uint32_t)g.n);
                        uint32_t seed = first.a * ctx.seed + first.c;
                        OB_ROLLED_G
                        for (int j = g.lane; j < N; j += g.n) {
                            if (lowband == nullptr) X[j] = (float)((int32_t)seed >> 20);
                            else X[j] = lowband[j] + ((seed & 0x8000u) ? (1.0f / 256) : -(1.0f / 256));
                            seed = step.a * seed + step.c;
                        }
                        const ObLcg all = ob_lcg_pow((uint32_t)N);
                        ctx.seed = all.a * ctx.seed + all.c;
                        cm = lowband == nullptr ? cm_mask : (uint32_t)fl;
                        g.sync();
                        ob_renormalise(g, X, N, f.gain);
                    }
                }
                ret = cm;
                sp--;
            }
        } else if (f.stage == 1) {
            ObEncPartFrame &c = st[sp + 1];
            c.N = f.N; c.B = f.B; c.LM = f.LM; c.stage = 0;
            if (f.mid_first) {
                f.cm = ret;
                const int32_t rb = f.mbits - (f.rebalance - ctx.remaining_bits);
                if (rb > 3 << OB_BITRES && f.itheta != 0) f.sbits += rb - (3 << OB_BITRES);
                c.X = f.X + f.N; c.lowband = f.lowband ? f.lowband + f.N : nullptr; c.b = f.sbits; c.gain = f.gain * f.side; c.fill = f.fill >> f.B;
            } else {
                f.cm = ret << (f.B0 >> 1);
                const int32_t rb = f.sbits - (f.rebalance - ctx.remaining_bits);
                if (rb > 3 << OB_BITRES && f.itheta != 16384) f.mbits += rb - (3 << OB_BITRES);
                c.X = f.X; c.lowband = f.lowband; c.b = f.mbits; c.gain = f.gain * f.mid; c.fill = f.fill;
            }
            f.stage = 2;
            sp++;
        } else {
            ret = f.mid_first ? (f.cm | ret << (f.B0 >> 1)) : (f.cm | ret);
            sp--;
        }
    }
    return ret;
}

OB_DEV uint32_t ob_enc_band_n1(ObEncBandCtx &ctx, float *X, float *Y, float *lowband_out)       // bands.c:904-937 (warp-uniform: every lane writes the same value)
{
    float *x = X;
    for (int c = 0; c < 1 + (Y != nullptr); c++) {
        int sign = 0;
        if (ctx.remaining_bits >= 1 << OB_BITRES) { sign = x[0] < 0; ctx.ec->bits((uint32_t)sign, 1); ctx.remaining_bits -= 1 << OB_BITRES; }
        if (ctx.resynth) x[0] = sign ? -1.f : 1.f;
        x = Y;
    }
    if (lowband_out) lowband_out[0] = X[0];
    return 1;
}

// quant_band (bands.c:1109-1231), encode = 1
template <class G>
OB_DEV uint32_t ob_enc_band_impl(const G &g, ObEncBandsShared &S, ObEncBandCtx &ctx, float *X, int N, int b, int B, float *lowband, int LM, float *lowband_out, float gain,
        float *lowband_scratch, int fill)
{
    const int N0 = N, longBlocks = B == 1;
    int N_B = N >> ob_log2i(B), N_B0, B0 = B, time_divide = 0, recombine = 0, tf_change = ctx.tf_change;      // B (short blocks) is a power of two
    uint32_t cm;
    if (N == 1) { g.sync(); cm = ob_enc_band_n1(ctx, X, nullptr, lowband_out); g.sync(); return cm; }
    if (tf_change > 0) recombine = tf_change;
    if (lowband_scratch && lowband && (recombine || ((N_B & 1) == 0 && tf_change < 0) || B0 > 1)) {
        OB_ROLLED_G
        for (int j = g.lane; j < N; j += g.n) lowband_scratch[j] = lowband[j];
        g.sync();
        lowband = lowband_scratch;
    }
    for (int k = 0; k < recombine; k++) {
        ob_haar1(g, X, N >> k, 1 << k);
        if (lowband) ob_haar1(g, lowband, N >> k, 1 << k);
        const int lo = fill & 0xF, hi = fill >> 4;
        const int a = (lo & 3 ? 1 : 0) | (lo & 12 ? 2 : 0), c = (hi & 3 ? 1 : 0) | (hi & 12 ? 2 : 0);
        fill = a | c << 2;
    }
    B >>= recombine;
    N_B <<= recombine;
    while ((N_B & 1) == 0 && tf_change < 0) {
        ob_haar1(g, X, N_B, B);
        if (lowband) ob_haar1(g, lowband, N_B, B);
        fill |= fill << B;
        B <<= 1; N_B >>= 1;
        time_divide++; tf_change++;
    }
    B0 = B; N_B0 = N_B;
    if (B0 > 1) {
        ob_hadamard(g, X, S.tmp, N_B >> recombine, B0 << recombine, longBlocks, 0);
        if (lowband) ob_hadamard(g, lowband, S.tmp, N_B >> recombine, B0 << recombine, longBlocks, 0);
    }
    cm = ob_enc_partition(g, S, ctx, X, N, b, B, lowband, LM, gain, fill);
    if (ctx.resynth) {
        if (B0 > 1) ob_hadamard(g, X, S.tmp, N_B >> recombine, B0 << recombine, longBlocks, 1);
        N_B = N_B0; B = B0;
        for (int k = 0; k < time_divide; k++) { B >>= 1; N_B <<= 1; cm |= cm >> B; ob_haar1(g, X, N_B, B); }
        for (int k = 0; k < recombine; k++) {
            uint32_t r = 0;
            for (int j = 0; j < 4; j++) if (cm & (1u << j)) r |= 3u << (2 * j);
            cm = r;
            ob_haar1(g, X, N0 >> k, 1 << k);
        }
        B <<= recombine;
        if (lowband_out) {
            const float n = sqrtf((float)N0);
            OB_ROLLED_G
            for (int j = g.lane; j < N0; j += g.n) lowband_out[j] = n * X[j];
            g.sync();
        }
        cm &= (1u << B) - 1;
    }
    return cm;
}

// quant_band is used from two places (mono / dual-stereo bands, and the mid and side of a joint-stereo band): ONE out-of-line copy -- the
// split tree, the PVQ stage and the transforms around them are ~4 k instructions, and the warp-per-stream kernel is bound by its instruction
// footprint.  As in every OB_STAGE the range coder works on a register copy of its state inside.
template <class G>
OB_STAGE uint32_t ob_enc_band(const G &g, ObEncBandsShared &S, ObEncBandCtx &ctx_io, float *X, int N, int b, int B, float *lowband, int LM, float *lowband_out, float gain,
        float *lowband_scratch, int fill)
{
    ObEncBandCtx ctx = ctx_io;
    ObRangeEnc ec = *ctx_io.ec;
    ctx.ec = &ec;
    const uint32_t cm = ob_enc_band_impl(g, S, ctx, X, N, b, B, lowband, LM, lowband_out, gain, lowband_scratch, fill);
    *ctx_io.ec = ec;
    ctx_io.remaining_bits = ctx.remaining_bits; ctx_io.seed = ctx.seed; ctx_io.pace = ctx.pace;
    return cm;
}

// quant_band_stereo (bands.c:1235-1381), encode = 1
template <class G>
OB_DEV uint32_t ob_enc_band_stereo(const G &g, ObEncBandsShared &S, ObEncBandCtx &ctx, float *X, float *Y, int N, int b, int B, float *lowband, int LM, float *lowband_out,
        float *lowband_scratch, int fill)
{
    ObSplit sp;
    uint32_t cm;
    const int orig_fill = fill;
    if (N == 1) { g.sync(); cm = ob_enc_band_n1(ctx, X, Y, lowband_out); g.sync(); return cm; }
    ob_enc_theta(g, ctx, sp, X, Y, N, &b, B, B, LM, 1, &fill);
    const int inv = sp.inv, delta = sp.delta, itheta = sp.itheta;
    const float mid = (1.f / 32768) * sp.imid, side = (1.f / 32768) * sp.iside;
    // N == 2 (bands.c:1273-1323): one quant_band on x2 (the mid, or the side when it is the larger), the other channel follows from a sign.
    // N > 2 (:1324-1368): mid and side, the one with the larger budget first.  One ob_enc_band call site for all of them.
    int mbits, sbits, sign = 0, c = 0, mid_first = 1;
    int32_t rebalance = 0;
    float *x2 = X, *y2 = Y;
    if (N == 2) {
        mbits = b; sbits = 0;
        if (itheta != 0 && itheta != 16384) sbits = 1 << OB_BITRES;
        mbits -= sbits;
        c = itheta > 8192;
        ctx.remaining_bits -= sp.qalloc + sbits;
        x2 = c ? Y : X; y2 = c ? X : Y;
        if (sbits) { sign = x2[0] * y2[1] - x2[1] * y2[0] < 0; ctx.ec->bits((uint32_t)sign, 1); }
        sign = 1 - 2 * sign;
    } else {
        mbits = ob_imax(0, ob_imin(b, (b - delta) / 2)); sbits = b - mbits;
        ctx.remaining_bits -= sp.qalloc;
        rebalance = ctx.remaining_bits;
        mid_first = mbits >= sbits;
    }
    cm = 0;
#ifdef __CUDACC__
#pragma unroll 1
#endif
    for (int q = 0; q < (N == 2 ? 1 : 2); q++) {
        const int do_mid = q == 0 ? mid_first : !mid_first;
        if (q == 1) {
            if (mid_first) { rebalance = mbits - (rebalance - ctx.remaining_bits); if (rebalance > 3 << OB_BITRES && itheta != 0) sbits += rebalance - (3 << OB_BITRES); }
            else { rebalance = sbits - (rebalance - ctx.remaining_bits); if (rebalance > 3 << OB_BITRES && itheta != 16384) mbits += rebalance - (3 << OB_BITRES); }
        }
        cm |= ob_enc_band(g, S, ctx, N == 2 ? x2 : (do_mid ? X : Y), N, do_mid ? mbits : sbits, B, do_mid ? lowband : nullptr, LM, do_mid ? lowband_out : nullptr,
                          (do_mid || N == 2) ? 1.0f : side, do_mid ? lowband_scratch : nullptr, N == 2 ? orig_fill : (do_mid ? fill : fill >> B));
    }
    if (N == 2) {
        g.sync();
        if (g.lane == 0) {
            y2[0] = -sign * x2[1];
            y2[1] = sign * x2[0];
            if (ctx.resynth) {
                X[0] = mid * X[0]; X[1] = mid * X[1];
                Y[0] = side * Y[0]; Y[1] = side * Y[1];
                float t = X[0]; X[0] = t - Y[0]; Y[0] = t + Y[0];
                t = X[1]; X[1] = t - Y[1]; Y[1] = t + Y[1];
            }
        }
        g.sync();
    }
    if (ctx.resynth) {
        if (N != 2) ob_enc_stereo_merge(g, X, Y, mid, N);
        if (inv) { for (int j = g.lane; j < N; j += g.n) Y[j] = -Y[j]; g.sync(); }
    }
    return cm;
}

// quant_all_bands (bands.c:1398-1672), encode = 1, start = 0.  Xg: the normalised spectrum in global memory (channel c at c*N); every band is
// staged into shared memory, coded there, and -- the encoder never reads X again -- not written back.
template <class G>
OB_STAGE void ob_enc_all_bands(const G &g, int end, const float *Xg, int C, int N_, uint8_t *collapse_masks, const float *bandE, const int *pulses, int shortBlocks,
        int spread, int dual_stereo, int intensity, const int *tf_res, int32_t total_bits, int32_t balance, ObRangeEnc &ec_io, int LM, int codedBands,
        uint32_t *seed, int complexity, int disable_inv, ObEncBandsShared &S, ObEncBandsWork &W)
{
    ObRangeEnc ec = ec_io;                                        // the coder works in registers inside a stage
    const int M = 1 << LM, B = shortBlocks ? M : 1, norm_offset = 0;
    float *norm = W.norm, *norm2 = norm + M * OB_EBANDS[OB_NB - 1] - norm_offset;
    int lowband_offset = 0, update_lowband = 1;
    const int theta_rdo = C == 2 && !dual_stereo && complexity >= 8;
    const int resynth = theta_rdo;
    float *lowband_scratch = S.lowband_scratch;
    ObEncBandCtx ctx;
    ctx.bandE = bandE; ctx.ec = &ec; ctx.intensity = intensity; ctx.seed = *seed; ctx.spread = spread; ctx.disable_inv = disable_inv;
    ctx.resynth = resynth; ctx.theta_round = 0; ctx.avoid_split_noise = B > 1;
    for (int i = 0; i < end; i++) {
        int32_t tell, remaining_bits, curr_balance;
        int b, N, effective_lowband = -1, tf_change;
        uint32_t x_cm, y_cm;
        const int last = (i == end - 1);
        float *X = S.xb, *Y = C == 2 ? S.xb + OB_MAX_BAND : nullptr;
        ctx.band = i;
        N = M * OB_EBANDS[i + 1] - M * OB_EBANDS[i];
        ctx.pace = 64 + 64 * i;
        g.pace(ctx.pace++);
        g.sync();
        OB_ROLLED_G
        for (int j = g.lane; j < N; j += g.n) { X[j] = Xg[M * OB_EBANDS[i] + j]; if (C == 2) Y[j] = Xg[N_ + M * OB_EBANDS[i] + j]; }
        g.sync();
        tell = (int32_t)ec.tell_frac();
        if (i != 0) balance -= tell;
        remaining_bits = total_bits - tell - 1;
        ctx.remaining_bits = remaining_bits;
        if (i <= codedBands - 1) {
            curr_balance = balance / ob_imin(3, codedBands - i);
            b = ob_imax(0, ob_imin(16383, ob_imin(remaining_bits + 1, pulses[i] + curr_balance)));
        } else b = 0;
        if (resynth && (M * OB_EBANDS[i] - N >= M * OB_EBANDS[0] || i == 1) && (update_lowband || lowband_offset == 0)) lowband_offset = i;
        tf_change = tf_res[i];
        ctx.tf_change = tf_change;
        float *lscratch = lowband_scratch;
        if (last && !theta_rdo) lscratch = nullptr;
        if (!resynth) lscratch = nullptr;                              // without resynthesis there is no folding source to transform
        if (lowband_offset != 0 && (spread != 3 || B > 1 || tf_change < 0)) {
            int fold_start, fold_end, fold_i;
            effective_lowband = ob_imax(0, M * OB_EBANDS[lowband_offset] - norm_offset - N);
            fold_start = lowband_offset;
            while (M * OB_EBANDS[--fold_start] > effective_lowband + norm_offset) ;
            fold_end = lowband_offset - 1;
            while (++fold_end < i && M * OB_EBANDS[fold_end] < effective_lowband + norm_offset + N) ;
            x_cm = y_cm = 0;
            fold_i = fold_start;
            do { x_cm |= collapse_masks[fold_i * C + 0]; y_cm |= collapse_masks[fold_i * C + C - 1]; } while (++fold_i < fold_end);
        } else x_cm = y_cm = (1u << B) - 1;
        if (dual_stereo && i == intensity) {
            dual_stereo = 0;
            if (resynth) { for (int j = g.lane; j < M * OB_EBANDS[i] - norm_offset; j += g.n) norm[j] = .5f * (norm[j] + norm2[j]); g.sync(); }
        }
        float *lb1 = effective_lowband != -1 ? norm + effective_lowband : nullptr;
        float *lb2 = effective_lowband != -1 ? norm2 + effective_lowband : nullptr;
        float *lo1 = last ? nullptr : norm + M * OB_EBANDS[i] - norm_offset, *lo2 = last ? nullptr : norm2 + M * OB_EBANDS[i] - norm_offset;
        if (dual_stereo || Y == nullptr) {
            // mono: one quant_band on X; dual stereo: X then Y, each with its own folding source and half the bits -- one call site
            const int ncall = dual_stereo ? 2 : 1;
            const uint32_t cm_in = x_cm | y_cm;
#ifdef __CUDACC__
#pragma unroll 1
#endif
            for (int q = 0; q < ncall; q++) {
                const uint32_t r = ob_enc_band(g, S, ctx, q ? Y : X, N, dual_stereo ? b / 2 : b, B, q ? lb2 : lb1, LM, q ? lo2 : lo1, 1.0f, lscratch,
                                               (int)(dual_stereo ? (q ? y_cm : x_cm) : cm_in));
                if (q) y_cm = r; else x_cm = r;
            }
            if (!dual_stereo) y_cm = x_cm;
        } else {
            // joint stereo; with theta RDO (bands.c:1583-1645) the band is coded twice from a snapshot -- theta rounded down, then up -- and the
            // trial with the larger weighted correlation to the input is kept.  One ob_enc_band_stereo call site for all three uses.
            const int rdo = theta_rdo && i < intensity;
            const uint32_t cm = x_cm | y_cm;
            float w0 = 0, w1 = 0, dist0 = 0, dist1 = 0;
            uint32_t cm2 = 0;
            ObRangeEnc ec_save, ec_save2;
            ObEncBandCtx ctx_save, ctx_save2;
            uint8_t *bytes_buf = nullptr;
            uint32_t save_bytes = 0;
            if (rdo) {
                {   // compute_channel_weights (bands.c:371-386)
                    float Ex = bandE[i], Ey = bandE[i + OB_NB];
                    const float minE = ob_fmin(Ex, Ey);
                    Ex = Ex + minE / 3; Ey = Ey + minE / 3;
                    w0 = Ex; w1 = Ey;
                }
                ec_save = ec; ctx_save = ctx;
                bytes_buf = ec_save.buf + ec_save.offs;
                save_bytes = ec_save.storage - ec_save.offs;
                OB_ROLLED_G
                for (int j = g.lane; j < N; j += g.n) { W.X_save[j] = X[j]; W.Y_save[j] = Y[j]; }
                g.sync();
            }
#ifdef __CUDACC__
#pragma unroll 1
#endif
            for (int t = 0; t < 1 + rdo; t++) {
                if (t == 1) {                                          // keep trial 0's result, rewind to the snapshot
                    cm2 = x_cm; ec_save2 = ec; ctx_save2 = ctx;
                    g.sync();
                    OB_ROLLED_G
                    for (int j = g.lane; j < N; j += g.n) { W.X_save2[j] = X[j]; W.Y_save2[j] = Y[j]; if (!last) W.norm_save2[j] = lo1[j]; }
                    for (uint32_t k = g.lane; k < save_bytes; k += g.n) W.bytes_save[k] = bytes_buf[k];
                    g.sync();
                    { const int pc = ctx.pace; ec = ec_save; ctx = ctx_save; ctx.pace = pc; }
                    OB_ROLLED_G
                    for (int j = g.lane; j < N; j += g.n) { X[j] = W.X_save[j]; Y[j] = W.Y_save[j]; }
                    g.sync();
                }
                ctx.theta_round = rdo ? (t ? 1 : -1) : 0;
                x_cm = ob_enc_band_stereo(g, S, ctx, X, Y, N, b, B, lb1, LM, lo1, lscratch, (int)cm);
                if (rdo) {
                    const float d = w0 * ob_psum(g, N, 0.f, [&](int j) { return W.X_save[j] * X[j]; }) + w1 * ob_psum(g, N, 0.f, [&](int j) { return W.Y_save[j] * Y[j]; });
                    if (t) dist1 = d; else dist0 = d;
                }
            }
            if (rdo && dist0 >= dist1) {
                { const int pc = ctx.pace; x_cm = cm2; ec = ec_save2; ctx = ctx_save2; ctx.pace = pc; }
                g.sync();
                OB_ROLLED_G
                for (int j = g.lane; j < N; j += g.n) { X[j] = W.X_save2[j]; Y[j] = W.Y_save2[j]; if (!last) lo1[j] = W.norm_save2[j]; }
                for (uint32_t k = g.lane; k < save_bytes; k += g.n) bytes_buf[k] = W.bytes_save[k];
                g.sync();
            }
            y_cm = x_cm;
        }
        collapse_masks[i * C + 0] = (uint8_t)x_cm;
        collapse_masks[i * C + C - 1] = (uint8_t)y_cm;
        balance += pulses[i] + tell;
        update_lowband = b > (N << OB_BITRES);
        ctx.avoid_split_noise = 0;
    }
    *seed = ctx.seed;
    ec_io = ec;
    g.sync();
}
