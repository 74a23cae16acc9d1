// enc_quant.cuh -- the quantisation half of the CELT encoder, per stream:
// coarse / fine / final band-energy quantisation with the two-pass intra/inter search (quant_bands.c:156-426),
// tf_encode (celt_encoder.c:756-794), the encoder side of the bit allocator (rate.c:248-645), and quant_all_bands
// with encode=1 (bands.c:647-1672): theta quantisation (stereo_itheta, vq.c:410-441), PVQ search (vq.c:165-328),
// pulse-vector indexing (cwrs.c:440-461), and -- for stereo at complexity >= 8 -- the theta RDO that encodes every
// band twice from a snapshot of the coder and keeps the better one (bands.c:1583-1645).
#pragma once
#include "enc_analysis.cuh"

// ---- energy quantisation ------------------------------------------------------------------------------------------------
OB_DEV float ob_loss_distortion(const float *eBands, const float *oldEBands, int end, int C)     // quant_bands.c:142-154
{
    float dist = 0;
    for (int c = 0; c < C; c++) for (int i = 0; i < end; i++) { const float d = eBands[i + c * OB_NB] - oldEBands[i + c * OB_NB]; dist = dist + d * d; }
    return ob_fmin(200, dist);
}

OB_DEV int ob_quant_coarse_impl(int end, const float *eBands, float *oldEBands, int32_t budget, int32_t tell, const uint8_t *prob_model,
        float *error, ObRangeEnc &enc, int C, int LM, int intra, float max_decay)               // quant_bands.c:156-259
{
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
    return badness;
}

// quant_coarse_energy (quant_bands.c:261-359). scratch: >= 1275 bytes for the intra pass' coder bytes.
OB_DEV void ob_quant_coarse_energy(int end, int effEnd, const float *eBands, float *oldEBands, uint32_t budget, float *error, ObRangeEnc &enc,
        int C, int LM, int nbAvailableBytes, int force_intra, float *delayedIntra, int two_pass, int loss_rate, uint8_t *scratch)
{
    float oldEBands_intra[2 * OB_NB], error_intra[2 * OB_NB];
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
        for (uint32_t k = 0; k < save_bytes; k++) scratch[k] = intra_buf[k];
        enc = enc_start_state;
        const int badness2 = ob_quant_coarse_impl(end, eBands, oldEBands, (int32_t)budget, (int32_t)tell, OB_E_PROB_MODEL + (LM * 2 + intra) * 42, error, enc, C, LM, 0, max_decay);
        if (two_pass && (badness1 < badness2 || (badness1 == badness2 && ((int32_t)enc.tell_frac()) + intra_bias > tell_intra))) {
            enc = enc_intra_state;
            for (uint32_t k = 0; k < save_bytes; k++) intra_buf[k] = scratch[k];
            for (int i = 0; i < C * OB_NB; i++) { oldEBands[i] = oldEBands_intra[i]; error[i] = error_intra[i]; }
            intra = 1;
        }
    } else {
        for (int i = 0; i < C * OB_NB; i++) { oldEBands[i] = oldEBands_intra[i]; error[i] = error_intra[i]; }
    }
    if (intra) *delayedIntra = new_distortion;
    else *delayedIntra = (OB_PRED_COEF[LM] * OB_PRED_COEF[LM]) * *delayedIntra + new_distortion;
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

// ---- bit allocation, encoder side (rate.c:248-645) ---------------------------------------------------------------------------
OB_DEV int ob_enc_allocation(ObRangeEnc &ec, int end, const int *offsets, const int *cap, int alloc_trim, int *intensity, int *dual_stereo,
        int32_t total, int32_t *balance_out, int *pulses, int *ebits, int *fine_priority, int C, int LM, int prev, int signalBandwidth)
{
    const int start = 0;
    int bits1[OB_NB], bits2[OB_NB], thresh[OB_NB], trim_offset[OB_NB], bits[OB_NB];
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
    return coded;
}

// ---- PVQ ---------------------------------------------------------------------------------------------------------------------
OB_DEV void ob_exp_rotation1_s(float *X, int len, int stride, float c, float s)                 // vq.c:47-71
{
    const float ms = -s;
    if (stride == 1) {                                     // the value written to p[1] is the next step's p[0]: carried in a register (same arithmetic)
        if (len < 2) return;
        float x1 = X[0];
        for (int i = 0; i < len - 1; i++) { const float x2 = X[i + 1]; X[i] = c * x1 + ms * x2; x1 = c * x2 + s * x1; }
        X[len - 1] = x1;
        if (len < 3) return;
        float x2 = X[len - 2];
        for (int i = len - 3; i >= 0; i--) { const float a = X[i]; X[i + 1] = c * x2 + s * a; x2 = c * a + ms * x2; }
        X[0] = x2;
        return;
    }
    float *p = X;
    for (int i = 0; i < len - stride; i++) { const float x1 = p[0], x2 = p[stride]; p[stride] = c * x2 + s * x1; *p++ = c * x1 + ms * x2; }
    p = &X[len - 2 * stride - 1];
    for (int i = len - 2 * stride - 1; i >= 0; i--) { const float x1 = p[0], x2 = p[stride]; p[stride] = c * x2 + s * x1; *p-- = c * x1 + ms * x2; }
}
OB_DEV void ob_exp_rotation_s(float *X, int len, int dir, int stride, int K, int spread)        // vq.c:74-117
{
    if (2 * K >= len || spread == 0) return;
    const int factor = spread == 1 ? 15 : spread == 2 ? 10 : 5;
    const float gain = (float)(1.0f * len) / (float)(len + factor * K);
    const float theta = .5f * (gain * gain);
    const float c = (float)cos((double)((.5f * 3.141592653f) * theta));
    const float s = (float)cos((double)((.5f * 3.141592653f) * (1.0f - theta)));
    int stride2 = 0;
    if (len >= 8 * stride) { stride2 = 1; while ((stride2 * stride2 + stride2) * stride + (stride >> 2) < len) stride2++; }
    len = len / stride;
    for (int i = 0; i < stride; i++) {
        if (dir < 0) {
            if (stride2) ob_exp_rotation1_s(X + i * len, len, stride2, s, c);
            ob_exp_rotation1_s(X + i * len, len, 1, c, s);
        } else {
            ob_exp_rotation1_s(X + i * len, len, 1, c, -s);
            if (stride2) ob_exp_rotation1_s(X + i * len, len, stride2, s, -c);
        }
    }
}

// op_pvq_search_c (vq.c:165-328).  X is overwritten with |X|.  Returns yy.
OB_DEV float ob_pvq_search(float *X, int *iy, int K, int N)
{
    float y[OB_MAX_BAND];
    uint8_t signx[OB_MAX_BAND];
    float sum = 0, xy = 0, yy = 0;
    int pulsesLeft = K;
    for (int j = 0; j < N; j++) { signx[j] = X[j] < 0; X[j] = fabsf(X[j]); iy[j] = 0; y[j] = 0; }
    if (K > (N >> 1)) {
        for (int j = 0; j < N; j++) sum += X[j];
        if (!(sum > 1e-15f && sum < 64)) {
            X[0] = 1.f;
            for (int j = 1; j < N; j++) X[j] = 0;
            sum = 1.f;
        }
        const float rcp = (K + 0.8f) * (1.f / sum);
        for (int j = 0; j < N; j++) {
            iy[j] = (int)floor((double)(rcp * X[j]));
            y[j] = (float)iy[j];
            yy = yy + y[j] * y[j];
            xy = xy + X[j] * y[j];
            y[j] *= 2;
            pulsesLeft -= iy[j];
        }
    }
    if (pulsesLeft > N + 3) {
        const float tmp = (float)pulsesLeft;
        yy = yy + tmp * tmp;
        yy = yy + tmp * y[0];
        iy[0] += pulsesLeft;
        pulsesLeft = 0;
    }
    for (int i = 0; i < pulsesLeft; i++) {
        int best_id = 0;
        yy = yy + 1;
        float Rxy = xy + X[0], Ryy = yy + y[0];
        Rxy = Rxy * Rxy;
        float best_den = Ryy, best_num = Rxy;
        for (int j = 1; j < N; j++) {
            Rxy = xy + X[j];
            Ryy = yy + y[j];
            Rxy = Rxy * Rxy;
            if (best_den * Rxy > Ryy * best_num) { best_den = Ryy; best_num = Rxy; best_id = j; }
        }
        xy = xy + X[best_id];
        yy = yy + y[best_id];
        y[best_id] += 2;
        iy[best_id]++;
    }
    for (int j = 0; j < N; j++) iy[j] = (iy[j] ^ -(int)signx[j]) + signx[j];
    return yy;
}

OB_DEV uint32_t ob_icwrs(int n, const int *y)                                                    // cwrs.c:440-456
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

OB_DEV uint32_t ob_collapse_mask(const int *iy, int N, int B)                                    // vq.c:143-163
{
    if (B <= 1) return 1;
    const int N0 = N / B;
    uint32_t mask = 0;
    for (int i = 0; i < B; i++) { uint32_t t = 0; for (int j = 0; j < N0; j++) t |= (uint32_t)iy[i * N0 + j]; mask |= (uint32_t)(t != 0) << i; }
    return mask;
}

// alg_quant (vq.c:330-359)
OB_DEV uint32_t ob_alg_quant(float *X, int N, int K, int spread, int B, ObRangeEnc &enc, float gain, int resynth)
{
    int iy[OB_MAX_BAND + 3];
    ob_exp_rotation_s(X, N, 1, B, K, spread);
    const float yy = ob_pvq_search(X, iy, K, N);
    enc.uint(ob_icwrs(N, iy), ob_pvq_v(N, K));
    if (resynth) {
        const float g = (1.f / sqrtf(yy)) * gain;                       // normalise_residual (vq.c:121-141)
        for (int i = 0; i < N; i++) X[i] = g * (float)iy[i];
        ob_exp_rotation_s(X, N, -1, B, K, spread);
    }
    return ob_collapse_mask(iy, N, B);
}

OB_DEV void ob_renormalise_s(float *X, int N, float gain)                                        // vq.c:383-407
{
    float E = 1e-15f + ob_inner_prod(X, X, N);
    const float g = (1.f / sqrtf(E)) * gain;
    for (int i = 0; i < N; i++) X[i] = g * X[i];
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
OB_DEV int ob_stereo_itheta(const float *X, const float *Y, int stereo, int N)
{
    float Emid = 1e-15f, Eside = 1e-15f;
    if (stereo) {
        for (int i = 0; i < N; i++) { const float m = X[i] + Y[i], s = X[i] - Y[i]; Emid = Emid + m * m; Eside = Eside + s * s; }
    } else { Emid += ob_inner_prod(X, X, N); Eside += ob_inner_prod(Y, Y, N); }
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
};

OB_DEV void ob_intensity_stereo(float *X, const float *Y, const float *bandE, int i, int N)      // bands.c:388-410
{
    const float left = bandE[i], right = bandE[i + OB_NB];
    const float norm = 1e-15f + sqrtf(1e-15f + left * left + right * right);
    const float a1 = left / norm, a2 = right / norm;
    for (int j = 0; j < N; j++) X[j] = a1 * X[j] + a2 * Y[j];
}
OB_DEV void ob_stereo_split(float *X, float *Y, int N)                                           // bands.c:412-424
{
    for (int j = 0; j < N; j++) { const float l = .70710678f * X[j], r = .70710678f * Y[j]; X[j] = l + r; Y[j] = r - l; }
}
OB_DEV void ob_stereo_merge_s(float *X, float *Y, float mid, int N)                              // bands.c:426-476
{
    float xp = 0, side = 0;
    for (int j = 0; j < N; j++) { xp = xp + Y[j] * X[j]; side = side + Y[j] * Y[j]; }
    xp = mid * xp;
    const float El = mid * mid + side - 2 * xp, Er = mid * mid + side + 2 * xp;
    if (Er < 6e-4f || El < 6e-4f) { for (int j = 0; j < N; j++) Y[j] = X[j]; return; }
    const float lgain = 1.f / sqrtf(El), rgain = 1.f / sqrtf(Er);
    for (int j = 0; j < N; j++) { const float l = mid * X[j], r = Y[j]; X[j] = lgain * (l - r); Y[j] = rgain * (l + r); }
}
OB_DEV void ob_hadamard_s(float *X, int N0, int stride, int hadamard, int interleave)            // bands.c:583-630
{
    float tmp[OB_MAX_BAND];
    const int N = N0 * stride;
    for (int i = 0; i < stride; i++) {
        const int row = hadamard ? ob_ordery(stride, i) : i;
        for (int j = 0; j < N0; j++) {
            if (interleave) tmp[j * stride + i] = X[row * N0 + j];
            else tmp[row * N0 + j] = X[j * stride + i];
        }
    }
    for (int j = 0; j < N; j++) X[j] = tmp[j];
}

// compute_theta (bands.c:700-903), encoder + decoder semantics of the encoder build (encode = 1)
OB_DEV void ob_enc_theta(ObEncBandCtx &ctx, ObSplit &sp, float *X, float *Y, int N, int *b, int B, int B0, int LM, int stereo, int *fill)
{
    ObRangeEnc &ec = *ctx.ec;
    int itheta, inv = 0, imid, iside, delta, qn;
    const int i = ctx.band;
    const int pulse_cap = OB_LOGN[i] + LM * (1 << OB_BITRES);
    const int offset = (pulse_cap >> 1) - (stereo && N == 2 ? 16 : 4);
    qn = ob_compute_qn(N, *b, offset, pulse_cap, stereo);
    if (stereo && i >= ctx.intensity) qn = 1;
    itheta = ob_stereo_itheta(X, Y, stereo, N);
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
            if (itheta == 0) ob_intensity_stereo(X, Y, ctx.bandE, i, N);
            else ob_stereo_split(X, Y, N);
        }
    } else if (stereo) {
        inv = itheta > 8192 && !ctx.disable_inv;
        if (inv) for (int j = 0; j < N; j++) Y[j] = -Y[j];
        ob_intensity_stereo(X, Y, ctx.bandE, i, N);
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

// quant_partition (bands.c:943-1105), encode = 1.  The reference recurses (depth <= LM+1 <= 4); here the recursion is an explicit
// stack, as in the decoder's symbol kernel: recursion inside a huge divergent thread-per-stream kernel makes nvcc's
// convergence-barrier bookkeeping (and with it its use of uniform registers) unreliable on sm_100a.
struct ObEncPartFrame {
    float *X, *lowband;
    int N, b, B, B0, LM, fill, mbits, sbits, itheta, stage, mid_first;
    int32_t rebalance;
    float gain, mid, side;
    uint32_t cm;
};

OB_DEV_NOINLINE uint32_t ob_enc_partition(ObEncBandCtx &ctx, float *X0, int N0, int b0, int Bin, float *lowband0, int LM0, float gain0, int fill0)
{
    ObEncPartFrame st[6];
    int sp = 0;
    uint32_t ret = 0;
    st[0].X = X0; st[0].lowband = lowband0; st[0].N = N0; st[0].b = b0; st[0].B = Bin; st[0].LM = LM0; st[0].gain = gain0; st[0].fill = fill0; st[0].stage = 0;
    // Two alternating phases, as in the decoder's symbol kernel: every thread first walks its cheap split / merge states until it
    // stands on a leaf, then the threads of the warp run the expensive leaf (PVQ search, rotation, index coding) together.
    while (sp >= 0) {
      bool at_leaf = false;
      while (sp >= 0 && !at_leaf) {
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
                ob_enc_theta(ctx, s, f.X, f.X + n, n, &bb, B1, f.B0, lm, 0, &fl);
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
            } else at_leaf = true;
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
      if (at_leaf) {
            ObEncPartFrame &f = st[sp];
            const uint8_t *cache = ob_pcache(ctx.band, f.LM);
            {
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
                if (q != 0) {
                    cm = ob_alg_quant(f.X, f.N, ob_get_pulses(q), ctx.spread, f.B, *ctx.ec, f.gain, ctx.resynth);
                } else if (ctx.resynth) {
                    const uint32_t cm_mask = (1u << f.B) - 1;
                    const int fl = f.fill & (int)cm_mask;
                    float *X = f.X;
                    if (!fl) { for (int j = 0; j < f.N; j++) X[j] = 0; }
                    else {
                        if (f.lowband == nullptr) {
                            for (int j = 0; j < f.N; j++) { ctx.seed = 1664525u * ctx.seed + 1013904223u; X[j] = (float)((int32_t)ctx.seed >> 20); }
                            cm = cm_mask;
                        } else {
                            for (int j = 0; j < f.N; j++) {
                                ctx.seed = 1664525u * ctx.seed + 1013904223u;
                                const float tmp = (ctx.seed & 0x8000u) ? (1.0f / 256) : -(1.0f / 256);
                                X[j] = f.lowband[j] + tmp;
                            }
                            cm = (uint32_t)fl;
                        }
                        ob_renormalise_s(X, f.N, f.gain);
                    }
                }
                ret = cm;
                sp--;
            }
      }
    }
    return ret;
}

OB_DEV uint32_t ob_enc_band_n1(ObEncBandCtx &ctx, float *X, float *Y, float *lowband_out)       // bands.c:904-937
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
OB_DEV_NOINLINE uint32_t ob_enc_band(ObEncBandCtx &ctx, float *X, int N, int b, int B, float *lowband, int LM, float *lowband_out, float gain,
        float *lowband_scratch, int fill)
{
    const int N0 = N, longBlocks = B == 1;
    int N_B = N / B, N_B0, B0 = B, time_divide = 0, recombine = 0, tf_change = ctx.tf_change;
    uint32_t cm;
    if (N == 1) return ob_enc_band_n1(ctx, X, nullptr, lowband_out);
    if (tf_change > 0) recombine = tf_change;
    if (lowband_scratch && lowband && (recombine || ((N_B & 1) == 0 && tf_change < 0) || B0 > 1)) {
        for (int j = 0; j < N; j++) lowband_scratch[j] = lowband[j];
        lowband = lowband_scratch;
    }
    for (int k = 0; k < recombine; k++) {
        ob_haar1_s(X, N >> k, 1 << k);
        if (lowband) ob_haar1_s(lowband, N >> k, 1 << k);
        const int lo = fill & 0xF, hi = fill >> 4;
        const int a = (lo & 3 ? 1 : 0) | (lo & 12 ? 2 : 0), c = (hi & 3 ? 1 : 0) | (hi & 12 ? 2 : 0);
        fill = a | c << 2;
    }
    B >>= recombine;
    N_B <<= recombine;
    while ((N_B & 1) == 0 && tf_change < 0) {
        ob_haar1_s(X, N_B, B);
        if (lowband) ob_haar1_s(lowband, N_B, B);
        fill |= fill << B;
        B <<= 1; N_B >>= 1;
        time_divide++; tf_change++;
    }
    B0 = B; N_B0 = N_B;
    if (B0 > 1) {
        ob_hadamard_s(X, N_B >> recombine, B0 << recombine, longBlocks, 0);
        if (lowband) ob_hadamard_s(lowband, N_B >> recombine, B0 << recombine, longBlocks, 0);
    }
    cm = ob_enc_partition(ctx, X, N, b, B, lowband, LM, gain, fill);
    if (ctx.resynth) {
        if (B0 > 1) ob_hadamard_s(X, N_B >> recombine, B0 << recombine, longBlocks, 1);
        N_B = N_B0; B = B0;
        for (int k = 0; k < time_divide; k++) { B >>= 1; N_B <<= 1; cm |= cm >> B; ob_haar1_s(X, N_B, B); }
        for (int k = 0; k < recombine; k++) {
            uint32_t r = 0;
            for (int j = 0; j < 4; j++) if (cm & (1u << j)) r |= 3u << (2 * j);
            cm = r;
            ob_haar1_s(X, N0 >> k, 1 << k);
        }
        B <<= recombine;
        if (lowband_out) {
            const float n = sqrtf((float)N0);
            for (int j = 0; j < N0; j++) lowband_out[j] = n * X[j];
        }
        cm &= (1u << B) - 1;
    }
    return cm;
}

// quant_band_stereo (bands.c:1235-1381), encode = 1
OB_DEV_NOINLINE uint32_t ob_enc_band_stereo(ObEncBandCtx &ctx, float *X, float *Y, int N, int b, int B, float *lowband, int LM, float *lowband_out,
        float *lowband_scratch, int fill)
{
    ObSplit sp;
    uint32_t cm;
    const int orig_fill = fill;
    if (N == 1) return ob_enc_band_n1(ctx, X, Y, lowband_out);
    ob_enc_theta(ctx, sp, X, Y, N, &b, B, B, LM, 1, &fill);
    const int inv = sp.inv, delta = sp.delta, itheta = sp.itheta;
    const float mid = (1.f / 32768) * sp.imid, side = (1.f / 32768) * sp.iside;
    if (N == 2) {
        int mbits = b, sbits = 0, sign = 0;
        if (itheta != 0 && itheta != 16384) sbits = 1 << OB_BITRES;
        mbits -= sbits;
        const int c = itheta > 8192;
        ctx.remaining_bits -= sp.qalloc + sbits;
        float *x2 = c ? Y : X, *y2 = c ? X : Y;
        if (sbits) { sign = x2[0] * y2[1] - x2[1] * y2[0] < 0; ctx.ec->bits((uint32_t)sign, 1); }
        sign = 1 - 2 * sign;
        cm = ob_enc_band(ctx, x2, N, mbits, B, lowband, LM, lowband_out, 1.0f, lowband_scratch, orig_fill);
        y2[0] = -sign * x2[1];
        y2[1] = sign * x2[0];
        if (ctx.resynth) {
            X[0] = mid * X[0]; X[1] = mid * X[1];
            Y[0] = side * Y[0]; Y[1] = side * Y[1];
            float t = X[0]; X[0] = t - Y[0]; Y[0] = t + Y[0];
            t = X[1]; X[1] = t - Y[1]; Y[1] = t + Y[1];
        }
    } else {
        int mbits = ob_imax(0, ob_imin(b, (b - delta) / 2)), sbits = b - mbits;
        ctx.remaining_bits -= sp.qalloc;
        int32_t rebalance = ctx.remaining_bits;
        if (mbits >= sbits) {
            cm = ob_enc_band(ctx, X, N, mbits, B, lowband, LM, lowband_out, 1.0f, lowband_scratch, fill);
            rebalance = mbits - (rebalance - ctx.remaining_bits);
            if (rebalance > 3 << OB_BITRES && itheta != 0) sbits += rebalance - (3 << OB_BITRES);
            cm |= ob_enc_band(ctx, Y, N, sbits, B, nullptr, LM, nullptr, side, nullptr, fill >> B);
        } else {
            cm = ob_enc_band(ctx, Y, N, sbits, B, nullptr, LM, nullptr, side, nullptr, fill >> B);
            rebalance = sbits - (rebalance - ctx.remaining_bits);
            if (rebalance > 3 << OB_BITRES && itheta != 16384) mbits += rebalance - (3 << OB_BITRES);
            cm |= ob_enc_band(ctx, X, N, mbits, B, lowband, LM, lowband_out, 1.0f, lowband_scratch, fill);
        }
    }
    if (ctx.resynth) {
        if (N != 2) ob_stereo_merge_s(X, Y, mid, N);
        if (inv) for (int j = 0; j < N; j++) Y[j] = -Y[j];
    }
    return cm;
}

// Scratch memory of quant_all_bands(encode=1): the folding source and, for theta RDO, the snapshots.
struct ObEncBandsScratch {
    float norm[2 * OB_NORM_LEN];
    float lowband_scratch[OB_MAX_BAND];
    float X_save[OB_MAX_BAND], Y_save[OB_MAX_BAND], X_save2[OB_MAX_BAND], Y_save2[OB_MAX_BAND], norm_save2[OB_MAX_BAND];
    uint8_t bytes_save[1275];
};

// quant_all_bands (bands.c:1398-1672), encode = 1, start = 0
OB_DEV_NOINLINE void ob_enc_all_bands(int end, float *X_, float *Y_, uint8_t *collapse_masks, const float *bandE, const int *pulses, int shortBlocks,
        int spread, int dual_stereo, int intensity, const int *tf_res, int32_t total_bits, int32_t balance, ObRangeEnc &ec, int LM, int codedBands,
        uint32_t *seed, int complexity, int disable_inv, ObEncBandsScratch &S)
{
    const int M = 1 << LM, B = shortBlocks ? M : 1, C = Y_ != nullptr ? 2 : 1, norm_offset = 0;
    float *norm = S.norm, *norm2 = norm + M * OB_EBANDS[OB_NB - 1] - norm_offset;
    int lowband_offset = 0, update_lowband = 1;
    const int theta_rdo = Y_ != nullptr && !dual_stereo && complexity >= 8;
    const int resynth = theta_rdo;
    float *lowband_scratch = resynth ? S.lowband_scratch : X_ + M * OB_EBANDS[OB_NB - 1];
    ObEncBandCtx ctx;
    ctx.bandE = bandE; ctx.ec = &ec; ctx.intensity = intensity; ctx.seed = *seed; ctx.spread = spread; ctx.disable_inv = disable_inv;
    ctx.resynth = resynth; ctx.theta_round = 0; ctx.avoid_split_noise = B > 1;
    for (int i = 0; i < end; i++) {
        int32_t tell, remaining_bits, curr_balance;
        int b, N, effective_lowband = -1, tf_change;
        uint32_t x_cm, y_cm;
        const int last = (i == end - 1);
        float *X = X_ + M * OB_EBANDS[i], *Y = Y_ ? Y_ + M * OB_EBANDS[i] : nullptr;
        ctx.band = i;
        N = M * OB_EBANDS[i + 1] - M * OB_EBANDS[i];
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
            if (resynth) for (int j = 0; j < M * OB_EBANDS[i] - norm_offset; j++) norm[j] = .5f * (norm[j] + norm2[j]);
        }
        float *lb1 = effective_lowband != -1 ? norm + effective_lowband : nullptr;
        float *lb2 = effective_lowband != -1 ? norm2 + effective_lowband : nullptr;
        float *lo1 = last ? nullptr : norm + M * OB_EBANDS[i] - norm_offset, *lo2 = last ? nullptr : norm2 + M * OB_EBANDS[i] - norm_offset;
        if (dual_stereo) {
            x_cm = ob_enc_band(ctx, X, N, b / 2, B, lb1, LM, lo1, 1.0f, lscratch, (int)x_cm);
            y_cm = ob_enc_band(ctx, Y, N, b / 2, B, lb2, LM, lo2, 1.0f, lscratch, (int)y_cm);
        } else {
            if (Y != nullptr) {
                if (theta_rdo && i < intensity) {
                    float w[2];
                    {   // compute_channel_weights (bands.c:371-386)
                        float Ex = bandE[i], Ey = bandE[i + OB_NB];
                        const float minE = ob_fmin(Ex, Ey);
                        Ex = Ex + minE / 3; Ey = Ey + minE / 3;
                        w[0] = Ex; w[1] = Ey;
                    }
                    const uint32_t cm = x_cm | y_cm;
                    const ObRangeEnc ec_save = ec;
                    const ObEncBandCtx ctx_save = ctx;
                    for (int j = 0; j < N; j++) { S.X_save[j] = X[j]; S.Y_save[j] = Y[j]; }
                    ctx.theta_round = -1;
                    x_cm = ob_enc_band_stereo(ctx, X, Y, N, b, B, lb1, LM, lo1, lscratch, (int)cm);
                    const float dist0 = w[0] * ob_inner_prod(S.X_save, X, N) + w[1] * ob_inner_prod(S.Y_save, Y, N);
                    const uint32_t cm2 = x_cm;
                    const ObRangeEnc ec_save2 = ec;
                    const ObEncBandCtx ctx_save2 = ctx;
                    for (int j = 0; j < N; j++) { S.X_save2[j] = X[j]; S.Y_save2[j] = Y[j]; }
                    if (!last) for (int j = 0; j < N; j++) S.norm_save2[j] = lo1[j];
                    const uint32_t nstart_bytes = ec_save.offs, nend_bytes = ec_save.storage;
                    uint8_t *bytes_buf = ec_save.buf + nstart_bytes;
                    const uint32_t save_bytes = nend_bytes - nstart_bytes;
                    for (uint32_t k = 0; k < save_bytes; k++) S.bytes_save[k] = bytes_buf[k];
                    ec = ec_save;
                    ctx = ctx_save;
                    for (int j = 0; j < N; j++) { X[j] = S.X_save[j]; Y[j] = S.Y_save[j]; }
                    ctx.theta_round = 1;
                    x_cm = ob_enc_band_stereo(ctx, X, Y, N, b, B, lb1, LM, lo1, lscratch, (int)cm);
                    const float dist1 = w[0] * ob_inner_prod(S.X_save, X, N) + w[1] * ob_inner_prod(S.Y_save, Y, N);
                    if (dist0 >= dist1) {
                        x_cm = cm2;
                        ec = ec_save2;
                        ctx = ctx_save2;
                        for (int j = 0; j < N; j++) { X[j] = S.X_save2[j]; Y[j] = S.Y_save2[j]; }
                        if (!last) for (int j = 0; j < N; j++) lo1[j] = S.norm_save2[j];
                        for (uint32_t k = 0; k < save_bytes; k++) bytes_buf[k] = S.bytes_save[k];
                    }
                } else {
                    ctx.theta_round = 0;
                    x_cm = ob_enc_band_stereo(ctx, X, Y, N, b, B, lb1, LM, lo1, lscratch, (int)(x_cm | y_cm));
                }
            } else {
                x_cm = ob_enc_band(ctx, X, N, b, B, lb1, LM, lo1, 1.0f, lscratch, (int)(x_cm | y_cm));
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
}
