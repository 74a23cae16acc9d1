// dec_bands.cuh -- float reconstruction of the normalised spectrum X from the symbol IR, cooperatively
// by one warp per (stream, frame).
//
// Replays, band by band, what the decoder side of quant_all_bands / quant_band / quant_band_stereo /
// quant_partition does to X (opus/celt/bands.c:1109-1672) once every integer decision is known:
//   PVQ pulses -> unit-norm vector (vq.c:121-141) + spreading rotation (vq.c:47-117), noise fill / spectral
//   folding + renormalise (bands.c:1063-1099, vq.c:383-407), Hadamard (de)interleave and Haar recombination
//   (bands.c:583-645, :1147-1220), stereo merge / inversion (bands.c:426-476, :1371-1380), and the folding
//   source `norm` (bands.c:1222-1229).
// Vector steps are strided over the lanes; the LCG noise (bands.c:61-64) is evaluated with a closed-form jump
// so every lane gets its own sample without a serial chain.
#pragma once
#include "ob_ir.h"
#include "ob_group.cuh"
#include "dec_symbols.cuh"   // OB_DEV, tables

// lane-strided loops over a band run 1-6 iterations on a warp: kept rolled (code size; see OB_ROLLED in ob_coop.cuh)
#if defined(__CUDACC__) && !defined(OB_BANDS_UNROLL)
#define OB_ROLLED_BANDS _Pragma("unroll 1")
#else
#define OB_ROLLED_BANDS
#endif

#ifdef __CUDACC__
#define OB_SQRTF(x) sqrtf(x)
#else
#include <math.h>
#define OB_SQRTF(x) sqrtf(x)
#endif

// celt_lcg_rand^k as an affine map s -> a*s + c (mod 2^32): composed by binary powering.
struct ObLcg { uint32_t a, c; };
OB_DEV ObLcg ob_lcg_pow(uint32_t k)
{
    ObLcg r = {1u, 0u};
    uint32_t pa = 1664525u, pc = 1013904223u;
    while (k) {
        if (k & 1) { r.c = pa * r.c + pc; r.a = pa * r.a; }
        pc = pa * pc + pc;
        pa = pa * pa;
        k >>= 1;
    }
    return r;
}

// Exact t / d for 0 <= t, d < 2^16 with one multiply: magic = floor(2^32 / d) + 1 (d >= 2).
struct ObDiv { uint32_t d, magic; };
OB_DEV ObDiv ob_div_make(int d) { ObDiv r; r.d = (uint32_t)d; r.magic = d > 1 ? 0xFFFFFFFFu / (uint32_t)d + 1u : 0u; return r; }
// divisor for splitting a work-item index into (block, item) over nblk blocks of d items: with one block the quotient is always 0,
// so the magic (a real division to compute) is not needed -- magic 0 makes ob_div_q return 0
OB_DEV ObDiv ob_div_make_blocks(int d, int nblk) { ObDiv r; r.d = (uint32_t)d; r.magic = nblk > 1 ? 0xFFFFFFFFu / (uint32_t)d + 1u : 0u; return r; }
// the quotient alone, for d >= 2 (one multiply-high on the device)
OB_DEV int ob_div_q(int t, const ObDiv &dv)
{
#ifdef __CUDA_ARCH__
    return (int)__umulhi((uint32_t)t, dv.magic);
#else
    return (int)(((uint64_t)(uint32_t)t * dv.magic) >> 32);
#endif
}
OB_DEV int ob_div(int t, const ObDiv &dv)
{
    if (dv.d <= 1) return t;
#ifdef __CUDACC__
    return (int)__umulhi((uint32_t)t, dv.magic);
#else
    return (int)(((uint64_t)(uint32_t)t * dv.magic) >> 32);
#endif
}
OB_DEV int ob_log2i(int v) { return 31 - OB_CLZ((uint32_t)v); }
#ifdef __CUDACC__
#define OB_RSQRTF(x) rsqrtf(x)
#define OB_COSF(x) cosf(x)
#else
#define OB_RSQRTF(x) (1.f / sqrtf(x))
#define OB_COSF(x) ((float)cos((double)(x)))
#endif

// haar1 (bands.c:632-645): N0*stride/2 independent butterflies.  stride is always a power of two here.
template <class G>
OB_DEV void ob_haar1(const G &g, float *X, int N0, int stride)
{
    N0 >>= 1;
    const int total = N0 * stride, ls = ob_log2i(stride);
    OB_ROLLED_BANDS
    for (int t = g.lane; t < total; t += g.n) {
        const int i = t & (stride - 1), j = t >> ls;
        const float t1 = .70710678f * X[stride * 2 * j + i], t2 = .70710678f * X[stride * (2 * j + 1) + i];
        X[stride * 2 * j + i] = t1 + t2;
        X[stride * (2 * j + 1) + i] = t1 - t2;
    }
    g.sync();
}

OB_DEV int ob_ordery(int stride, int i)
{
    // ordery_table (bands.c:576-581) rows for stride 2, 4, 8, 16
    switch (stride) {
    case 2: return i == 0 ? 1 : 0;
    case 4: { const int t[4] = {3, 0, 2, 1}; return t[i]; }
    case 8: { const int t[8] = {7, 0, 4, 3, 6, 1, 5, 2}; return t[i]; }
    default: { const int t[16] = {15, 0, 8, 7, 12, 3, 11, 4, 14, 1, 9, 6, 13, 2, 10, 5}; return t[i & 15]; }
    }
}

// deinterleave_hadamard / interleave_hadamard (bands.c:583-630) through a scratch buffer.
template <class G>
OB_DEV void ob_hadamard(const G &g, float *X, float *tmp, int N0, int stride, int hadamard, int interleave)
{
    const int N = N0 * stride, ls = ob_log2i(stride);                 // stride is a power of two
    OB_ROLLED_BANDS
    for (int t = g.lane; t < N; t += g.n) {
        const int i = t & (stride - 1), j = t >> ls;      // i: block, j: position inside block (any enumeration of the pairs will do)
        const int row = hadamard ? ob_ordery(stride, i) : i;
        if (interleave) tmp[j * stride + i] = X[row * N0 + j];
        else tmp[row * N0 + j] = X[j * stride + i];
    }
    g.sync();
    OB_ROLLED_BANDS
    for (int t = g.lane; t < N; t += g.n) X[t] = tmp[t];
    g.sync();
}

// exp_rotation(dir=-1) (vq.c:74-117).  Each (block, residue-class) chain of exp_rotation1 is an independent
// serial recurrence; chains are spread over the lanes.
#ifdef __CUDACC__
// One exp_rotation1 sweep over a single chain as a first-order linear recurrence, evaluated with a warp scan:
//   u_0 = a_0, u_t = s*u_{t-1} + c*a_t;  out_t = c*u_t - s*a_{t+1} (t <= L-2), out_{L-1} = u_{L-1}
// (the forward loop of vq.c:56-63; the backward loop :64-71 is the same recurrence on the reversed array with -s).
// elem(t) = x[rev ? L-1-t : t].
__device__ __forceinline__ void ob_rot_scan(float *x, int L, float c, float s, int rev)
{
    const int lane = (int)(threadIdx.x & 31);
    const float s2 = s * s, s4 = s2 * s2, s8 = s4 * s4, s16 = s8 * s8;
    float slane = s;                                   // s^(lane+1)
    if (lane & 1) slane *= s;
    if (lane & 2) slane *= s2;
    if (lane & 4) slane *= s4;
    if (lane & 8) slane *= s8;
    if (lane & 16) slane *= s16;
    float carry = 0.f;
    for (int base = 0; base < L; base += 32) {
        const int t = base + lane;
        const bool valid = t < L;
        const float a = valid ? x[rev ? L - 1 - t : t] : 0.f;
        const float an = (t + 1 < L) ? x[rev ? L - 2 - t : t + 1] : 0.f;
        float y = (t == 0) ? a : c * a;
        float p;
        p = __shfl_up_sync(0xffffffffu, y, 1); if (lane >= 1) y += s * p;
        p = __shfl_up_sync(0xffffffffu, y, 2); if (lane >= 2) y += s2 * p;
        p = __shfl_up_sync(0xffffffffu, y, 4); if (lane >= 4) y += s4 * p;
        p = __shfl_up_sync(0xffffffffu, y, 8); if (lane >= 8) y += s8 * p;
        p = __shfl_up_sync(0xffffffffu, y, 16); if (lane >= 16) y += s16 * p;
        y += slane * carry;
        carry = __shfl_sync(0xffffffffu, y, 31);
        __syncwarp();
        if (valid) x[rev ? L - 1 - t : t] = (t + 1 < L) ? c * y - s * an : y;
        __syncwarp();
    }
}
#endif

template <class G>
OB_DEV void ob_rot_pass(const G &g, float *X, int nblocks, int len, int stride, float c, float s)
{
    const int chains = nblocks * stride;
#ifdef __CUDACC__
    if (G::n == 32 && chains == 1 && len >= 12) {
        ob_rot_scan(X, len, c, s, 0);                  // forward sweep i = 0 .. len-2
        ob_rot_scan(X, len - 1, c, -s, 1);             // backward sweep i = len-3 .. 0 (x[len-1] untouched)
        return;
    }
#endif
    const int lgb = ob_log2i(nblocks);                                // nblocks (short blocks) is a power of two: chain t -> (block, residue) by shifts
    OB_ROLLED_BANDS
    for (int t = g.lane; t < chains; t += g.n) {
        float *x = X + (t & (nblocks - 1)) * len;
        const int r = t >> lgb;
        // forward sweep over i = r, r+stride, ... < len-stride: the rotated x[i+stride] is carried in a register
        int i = r;
        if (i < len - stride) {
            float x1 = x[i];
            for (; i < len - stride; i += stride) {
                const float x2 = x[i + stride];
                x[i] = c * x1 - s * x2;
                x1 = c * x2 + s * x1;
            }
            x[i] = x1;
        }
        // backward sweep covers i = len-2*stride-1 .. 0, restricted to this residue class; x[i] is carried downwards.  The forward sweep
        // stopped at the first i >= len-stride of the class, so the largest one <= len-2*stride-1 is two strides below it.
        i -= 2 * stride;
        if (i >= 0) {
            float x2 = x[i + stride];
            for (; i >= 0; i -= stride) {
                const float x1 = x[i];
                x[i + stride] = c * x2 + s * x1;
                x2 = c * x1 - s * x2;
            }
            x[i + stride] = x2;
        }
    }
    g.sync();
}

template <class G>
OB_DEV void ob_exp_rotation_inv(const G &g, float *X, int len, int stride, int K, int spread)
{
    if (2 * K >= len || spread == 0) return;
    const int factor = spread == 1 ? 15 : spread == 2 ? 10 : 5;
    const float gain = (float)(1.0f * len) / (float)(len + factor * K);
    const float theta = .5f * (gain * gain);
    float c, s;
    int stride2 = 0;
#ifdef __CUDA_ARCH__
    if (G::n == 32) {
        // the two cosines on two lanes at once, the stride search as one vote: lane k tests candidate k+1 (the predicate is monotone, stride2 <= 13)
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
        if (len >= 8 * stride) {
            stride2 = 1;
            while ((stride2 * stride2 + stride2) * stride + (stride >> 2) < len) stride2++;
        }
    }
    len >>= ob_log2i(stride);                                         // stride = short blocks spanned: a power of two
    // one inlined copy of the pass for both strides (code size: see ob_reconstruct_bands)
#ifdef __CUDACC__
#pragma unroll 1
#endif
    for (int p = stride2 ? 0 : 1; p < 2; p++) ob_rot_pass(g, X, stride, len, p ? 1 : stride2, p ? c : s, p ? s : c);
}

#define OB_LEAF_WIN 64
// Per-warp shared-memory working set of the band-reconstruction stage.  CH = 1: room for mono frames only (5.3 KB instead of 8.5 KB per
// warp: 30 instead of 24 resident warps per SM, the kernel is occupancy bound); the second-channel halves are then never touched.
template <int CH>
struct ObBandsSharedT {
    static constexpr int channels = CH;
    float xb[CH * OB_MAX_BAND];         // current band: channel 0 at [0,176), channel 1 at [176,352)
    float norm[CH * OB_NORM_LEN];       // folding source per channel (bands.c:1438)
    float scratch[OB_MAX_BAND];         // transformed copy of the folding source (lowband_scratch)
    float tmp[OB_MAX_BAND];             // Hadamard permutation buffer
    ObLeaf leaves[OB_LEAF_WIN];         // window of the frame's leaf records: filled when a band's leaves (<= 16 per quant_band call, two calls) are not all inside
    ObBand bands[OB_NB];
};
typedef ObBandsSharedT<2> ObBandsShared;

// Fills one terminal partition (bands.c:1038-1103).  X already holds (float)iy for PULSES leaves.
// lowband: folding source aligned with the band start.
template <class G>
OB_DEV void ob_fill_leaf(const G &g, const ObLeaf &lf, float *Xb, int off_in_band, const float *lowband,
        uint32_t seed_in, const ObLcg &step, int spread)
{
    float *X = Xb + off_in_band;
    const int n = lf.n;
    if (lf.kind == OB_LEAF_PULSES) {
        const float ryy = (float)lf.lcg_before;                            // the vector's squared norm, summed (exactly: small integers) by the symbol stage
        const float gg = OB_RSQRTF(ryy) * lf.gain;                         // normalise_residual (vq.c:121-141)
        OB_ROLLED_BANDS
        for (int j = g.lane; j < n; j += g.n) X[j] = gg * X[j];
        g.sync();
        ob_exp_rotation_inv(g, X, n, lf.B, lf.K, spread);
    } else if (lf.kind == OB_LEAF_ZERO) {
        OB_ROLLED_BANDS
        for (int j = g.lane; j < n; j += g.n) X[j] = 0.f;
        g.sync();
    } else if (lf.kind == OB_LEAF_ONE) {
        if (g.lane == 0) X[0] = lf.K ? -1.f : 1.f;
        g.sync();
    } else {
        // noise / folded spectrum + renormalise (bands.c:1070-1098)
        const ObLcg first = ob_lcg_pow((uint32_t)lf.lcg_before + (uint32_t)g.lane + 1u);
        uint32_t seed = first.a * seed_in + first.c;
        float e = 0.f;
        OB_ROLLED_BANDS
        for (int j = g.lane; j < n; j += g.n) {
            float v;
            if (lf.kind == OB_LEAF_NOISE) v = (float)((int32_t)seed >> 20);
            else v = lowband[off_in_band + j] + ((seed & 0x8000u) ? (1.0f / 256) : -(1.0f / 256));
            X[j] = v;
            e += v * v;
            seed = step.a * seed + step.c;
        }
        e = 1e-15f + g.sum(e);
        const float gg = OB_RSQRTF(e) * lf.gain;                           // renormalise_vector (vq.c:383-407)
        OB_ROLLED_BANDS
        for (int j = g.lane; j < n; j += g.n) X[j] = gg * X[j];
        g.sync();
    }
}

// quant_band, resynthesis side (bands.c:1109-1231).  Xb: band buffer of this channel (already holding (float)iy);
// leaves: staged leaf records [0, leaf_cnt); lowband: source in norm[] or nullptr; lowband_out: destination in norm[] or nullptr.
template <class G, class SH>
OB_DEV void ob_band_call(const G &g, SH &sh, const ObLeaf *leaves, int leaf_cnt, int band_off_abs, float *Xb, int N, int B,
        int tf_change, const float *lowband, float *lowband_out, uint32_t seed_in, const ObLcg &step, int spread)
{
    const int N0 = N, longBlocks = B == 1;
    int N_B = N >> ob_log2i(B), time_divide = 0, recombine = 0;          // B (short blocks) is a power of two
    float *scratch = sh.scratch, *tmp = sh.tmp;
    if (N == 1) {
        ob_fill_leaf(g, leaves[0], Xb, 0, nullptr, seed_in, step, spread);
        if (lowband_out && g.lane == 0) lowband_out[0] = Xb[0];
        g.sync();
        return;
    }
    if (tf_change > 0) recombine = tf_change;
    float *lb = nullptr;
    if (lowband) {
        if (recombine || ((N_B & 1) == 0 && tf_change < 0) || B > 1) {
            OB_ROLLED_BANDS
            for (int j = g.lane; j < N; j += g.n) scratch[j] = lowband[j];
            g.sync();
            lb = scratch;
        } else lb = const_cast<float *>(lowband);          // used read-only in this case
    }
    for (int k = 0; k < recombine; k++) if (lb) ob_haar1(g, lb, N >> k, 1 << k);
    B >>= recombine;
    N_B <<= recombine;
    while ((N_B & 1) == 0 && tf_change < 0) {
        if (lb) ob_haar1(g, lb, N_B, B);
        B <<= 1; N_B >>= 1;
        time_divide++; tf_change++;
    }
    const int B0 = B, N_B0 = N_B;
    if (B0 > 1 && lb) ob_hadamard(g, lb, tmp, N_B >> recombine, B0 << recombine, longBlocks, 0);

    for (int l = 0; l < leaf_cnt; l++) {
        const ObLeaf lf = leaves[l];
        ob_fill_leaf(g, lf, Xb, (int)lf.off - band_off_abs, lb, seed_in, step, spread);
    }

    if (B0 > 1) ob_hadamard(g, Xb, tmp, N_B >> recombine, B0 << recombine, longBlocks, 1);
    N_B = N_B0; B = B0;
#ifdef __CUDACC__
#pragma unroll 1
#endif
    for (int k = 0; k < time_divide + recombine; k++) {              // undo the time divisions, then the recombinations: one ob_haar1 site
        int hn, hs;
        if (k < time_divide) { B >>= 1; N_B <<= 1; hn = N_B; hs = B; }
        else { hn = N0 >> (k - time_divide); hs = 1 << (k - time_divide); }
        ob_haar1(g, Xb, hn, hs);
    }
    if (lowband_out) {
        const float n = OB_SQRTF((float)N0);
        OB_ROLLED_BANDS
        for (int j = g.lane; j < N0; j += g.n) lowband_out[j] = n * Xb[j];
        g.sync();
    }
}

// stereo_merge (bands.c:426-476)
template <class G>
OB_DEV void ob_stereo_merge(const G &g, float *X, float *Y, float mid, int N)
{
    float xp = 0.f, side = 0.f;
    OB_ROLLED_BANDS
    for (int j = g.lane; j < N; j += g.n) { xp += Y[j] * X[j]; side += Y[j] * Y[j]; }
    xp = g.sum(xp); side = g.sum(side);
    xp = mid * xp;
    const float El = mid * mid + side - 2 * xp, Er = mid * mid + side + 2 * xp;
    if (Er < 6e-4f || El < 6e-4f) {
        OB_ROLLED_BANDS
        for (int j = g.lane; j < N; j += g.n) Y[j] = X[j];
        g.sync();
        return;
    }
    const float lgain = OB_RSQRTF(El), rgain = OB_RSQRTF(Er);
    OB_ROLLED_BANDS
    for (int j = g.lane; j < N; j += g.n) {
        const float l = mid * X[j], r = Y[j];
        X[j] = lgain * (l - r);
        Y[j] = rgain * (l + r);
    }
    g.sync();
}

// Reconstructs the normalised spectrum of one frame band by band and writes it to Xout (C*N floats, channel c at
// c*N; coefficients of bands >= end are NOT written).  seed_in: the stream's range-coder state left by the previous
// frame (celt_decoder.c:1275 passes &st->rng).
template <class G, class SH>
OB_DEV void ob_reconstruct_bands(const G &g, const ObFrameIR *ir, uint32_t seed_in, SH &sh, float *Xout)
{
    const ObFrameHdr &h = ir->hdr;
    const int LM = h.LM, M = 1 << LM, C = SH::channels == 1 ? 1 : h.C, N = OB_SHORT << LM, end = h.end;     // mono-only instantiation: the stereo paths compile away
    const int Bfr = (h.flags & OB_F_TRANSIENT) ? M : 1, spread = h.spread;
    const ObLcg step = ob_lcg_pow((uint32_t)g.n);
    float *norm = sh.norm, *norm2 = sh.norm + OB_NORM_LEN;
    float *Xb = sh.xb, *Yb = sh.xb + OB_MAX_BAND;
    {   // band records -> shared (16-byte records as 4 words each)
        const uint32_t *src = (const uint32_t *)ir->bands;
        uint32_t *dst = (uint32_t *)sh.bands;
        OB_ROLLED_BANDS
        for (int j = g.lane; j < (int)(sizeof(ObBand) / 4) * OB_NB; j += g.n) dst[j] = src[j];
    }
    g.sync();
    int win0 = -2 * OB_LEAF_WIN;                                      // first leaf of the staged window: none yet
    for (int i = 0; i < end; i++) {
        const ObBand br = sh.bands[i];
        const int boff = M * OB_EBANDS[i], Nb = M * (OB_EBANDS[i + 1] - OB_EBANDS[i]);
        const int last = i == end - 1, tf_change = h.tf_change[i];
        const int na = br.leaf_cnt_a, nb = br.leaf_cnt_b;
        const ObLeaf *la, *lbv;
        {   // this band's leaves must lie in the staged window (12-byte records as 3 words each: ~one refill per frame); its pulse vector (as float)
            int l0 = OB_MAX_LEAVES, l1 = 0;
            if (na) { l0 = br.leaf_begin_a; l1 = l0 + na; }
            if (nb) { l0 = ob_imin(l0, (int)br.leaf_begin_b); l1 = ob_imax(l1, (int)br.leaf_begin_b + nb); }
            la = sh.leaves + ((int)br.leaf_begin_a - win0); lbv = sh.leaves + ((int)br.leaf_begin_b - win0);
            if (l1 - l0 > OB_LEAF_WIN) {                             // the two calls' leaves are not neighbours (never seen): stage them one by one
                const uint32_t *sa = (const uint32_t *)(ir->leaves + br.leaf_begin_a), *sb = (const uint32_t *)(ir->leaves + br.leaf_begin_b);
                uint32_t *dst = (uint32_t *)sh.leaves;
                OB_ROLLED_BANDS
                for (int j = g.lane; j < 3 * na; j += g.n) dst[j] = sa[j];
                OB_ROLLED_BANDS
                for (int j = g.lane; j < 3 * nb; j += g.n) dst[3 * (OB_LEAF_WIN / 2) + j] = sb[j];
                la = sh.leaves; lbv = sh.leaves + OB_LEAF_WIN / 2;
                win0 = -2 * OB_LEAF_WIN;
            } else if (l1 > l0 && (l0 < win0 || l1 > win0 + OB_LEAF_WIN)) {
                win0 = l0;
                const uint32_t *src = (const uint32_t *)(ir->leaves + win0);
                uint32_t *dst = (uint32_t *)sh.leaves;
                const int cnt = 3 * ob_imin(OB_LEAF_WIN, OB_MAX_LEAVES - win0);
                OB_ROLLED_BANDS
                for (int j = g.lane; j < cnt; j += g.n) dst[j] = src[j];
                la = sh.leaves + ((int)br.leaf_begin_a - win0); lbv = sh.leaves + ((int)br.leaf_begin_b - win0);
            }
            const int16_t *iy = ir->iy + boff;
            OB_ROLLED_BANDS
            for (int j = g.lane; j < Nb; j += g.n) Xb[j] = (float)iy[j];
            if (C == 2) for (int j = g.lane; j < Nb; j += g.n) Yb[j] = (float)iy[N + j];
        }
        if (br.flags & 8) {                                          // leaving dual stereo (bands.c:1551-1558)
            OB_ROLLED_BANDS
            for (int j = g.lane; j < boff; j += g.n) norm[j] = .5f * (norm[j] + norm2[j]);
        }
        g.sync();
        const float *lb1 = br.eff_lowband >= 0 ? norm + br.eff_lowband : nullptr;
        const float *lb2 = br.eff_lowband >= 0 ? norm2 + br.eff_lowband : nullptr;
        float *lo1 = last ? nullptr : norm + boff, *lo2 = last ? nullptr : norm2 + boff;
        // One ob_band_call site for every mode (the loop is kept rolled): the call inlines ~5 k instructions, and with one copy per mode
        // `no_instruction` was the band kernel's largest stall on stereo frames (8.9 cycles per issued instruction, profiles/r01n_*).
        const int mode = SH::channels == 1 ? (int)OB_BAND_MONO : (int)br.mode;       // mono-sized instantiation: mono frames only
        if (mode != OB_BAND_MONO && mode != OB_BAND_DUAL && Nb == 1) {                // quant_band_n1 with Y (bands.c:904-937)
            ob_fill_leaf(g, la[0], Xb, 0, nullptr, seed_in, step, spread);
            ob_fill_leaf(g, la[1], Yb, 0, nullptr, seed_in, step, spread);
            if (lo1 && g.lane == 0) lo1[0] = Xb[0];
            g.sync();
        } else {
            // MONO: X.  DUAL: X then Y, each with its own folding source.  JOINT_N2 (bands.c:1273-1323): x2 only, which is Y when c is set.
            // JOINT, N > 2 (bands.c:1324-1381): mid on X with folding, side on Y without.
            const int c = mode == OB_BAND_JOINT_N2 ? (br.flags >> 1) & 1 : 0;
            const int ncall = (mode == OB_BAND_DUAL || mode == OB_BAND_JOINT) ? 2 : 1;
#ifdef __CUDACC__
#pragma unroll 1
#endif
            for (int q = 0; q < ncall; q++) {
                const int onY = q | c;
                const float *lbq = q ? (mode == OB_BAND_DUAL ? lb2 : nullptr) : lb1;
                float *loq = q ? (mode == OB_BAND_DUAL ? lo2 : nullptr) : lo1;
                ob_band_call(g, sh, q ? lbv : la, q ? nb : na, onY ? N + boff : boff, onY ? Yb : Xb, Nb, Bfr, tf_change, lbq, loq, seed_in, step, spread);
            }
            if (mode == OB_BAND_JOINT_N2) {
                const int sign = 1 - 2 * ((br.flags >> 2) & 1);
                float *x2 = c ? Yb : Xb, *y2 = c ? Xb : Yb;
                if (g.lane == 0) {
                    const float mid = (1.f / 32768) * br.imid, side = (1.f / 32768) * br.iside;
                    y2[0] = -sign * x2[1];
                    y2[1] = sign * x2[0];
                    Xb[0] = mid * Xb[0]; Xb[1] = mid * Xb[1];
                    Yb[0] = side * Yb[0]; Yb[1] = side * Yb[1];
                    float t = Xb[0]; Xb[0] = t - Yb[0]; Yb[0] = t + Yb[0];
                    t = Xb[1]; Xb[1] = t - Yb[1]; Yb[1] = t + Yb[1];
                    if (br.flags & 1) { Yb[0] = -Yb[0]; Yb[1] = -Yb[1]; }
                }
                g.sync();
            } else if (mode == OB_BAND_JOINT) {
                const float mid = (1.f / 32768) * br.imid;
                ob_stereo_merge(g, Xb, Yb, mid, Nb);
                if (br.flags & 1) {
                    OB_ROLLED_BANDS
                    for (int j = g.lane; j < Nb; j += g.n) Yb[j] = -Yb[j];
                    g.sync();
                }
            }
        }
        OB_ROLLED_BANDS
        for (int j = g.lane; j < Nb; j += g.n) Xout[boff + j] = Xb[j];
        if (C == 2) for (int j = g.lane; j < Nb; j += g.n) Xout[N + boff + j] = Yb[j];
        g.sync();
    }
}
