/* oracle/celt_oracle.c -- TEST INFRASTRUCTURE ONLY (see celt_oracle.h).
 *
 * Scalar CPU restatement of the reference's CELT-only 48 kHz decoder.  Every function names the
 * reference lines (relative to /root/reference) whose behaviour it restates.  It is written to be
 * arithmetically identical to the reference's float build (same operation order, libm in double where
 * the reference uses double), so that integer results are bit-exact and PCM agrees to float rounding
 * (the reference's SSE inner products / comb filter reorder a few sums).  Constant tables are data,
 * dumped from the reference by oracle/gen_tables.c.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "celt_oracle.h"

#define OB_TABLE(type, name, n) static const type name[n]
#include "../opus_codec_b200/csrc/celt_tables.inc"

#define NB 21          /* bands in the 48 kHz mode (static_modes_float.h:866-888) */
#define OVL 120        /* overlap */
#define SHORT 120      /* shortMdctSize */
#define HIST 2048      /* DECODE_BUFFER_SIZE (celt_decoder.c:72) */
#define BITRES 3
#define MAX_FINE_BITS 8
#define FINE_OFFSET 21
#define QTHETA_OFFSET 4
#define QTHETA_OFFSET_TWOPHASE 16
#define MINPERIOD 15
#define SPREAD_NONE 0
#define SPREAD_NORMAL 2
#define SPREAD_AGGRESSIVE 3

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int ilog(uint32_t v) { return v ? 32 - __builtin_clz(v) : 0; }   /* EC_ILOG, ecintrin.h:86 */

/* ================================================================================================
 * Range decoder  (opus/celt/entdec.c, entcode.c, entcode.h)
 * ============================================================================================== */
typedef struct {
    const uint8_t *buf;
    uint32_t storage, end_offs, end_window, offs, rng, val, ext;
    int nend_bits, nbits_total, rem, error;
} rdec;

static int rd_byte(rdec *d) { return d->offs < d->storage ? d->buf[d->offs++] : 0; }            /* entdec.c:91-93 */
static int rd_byte_end(rdec *d) { return d->end_offs < d->storage ? d->buf[d->storage - ++d->end_offs] : 0; } /* :95-98 */

static void rd_norm(rdec *d)                                                                   /* entdec.c:102-117 */
{
    while (d->rng <= 0x800000u) {
        int sym;
        d->nbits_total += 8;
        d->rng <<= 8;
        sym = d->rem;
        d->rem = rd_byte(d);
        sym = (sym << 8 | d->rem) >> 1;
        d->val = ((d->val << 8) + (255 & ~sym)) & 0x7FFFFFFFu;
    }
}

static void rd_init(rdec *d, const uint8_t *buf, uint32_t len)                                 /* entdec.c:119-137 */
{
    d->buf = buf; d->storage = len; d->end_offs = 0; d->end_window = 0; d->nend_bits = 0;
    d->nbits_total = 32 + 1 - ((32 - 7) / 8) * 8;   /* = 9 */
    d->offs = 0; d->rng = 1u << 7;
    d->rem = rd_byte(d);
    d->val = d->rng - 1 - (d->rem >> 1);
    d->error = 0; d->ext = 0;
    rd_norm(d);
}

static int rd_tell(const rdec *d) { return d->nbits_total - ilog(d->rng); }                    /* entcode.h:111-113 */

static uint32_t rd_tell_frac(const rdec *d)                                                    /* entcode.c:69-84 */
{
    static const unsigned corr[8] = {35733, 38967, 42495, 46340, 50535, 55109, 60097, 65535};
    uint32_t nbits = (uint32_t)d->nbits_total << BITRES;
    int l = ilog(d->rng);
    uint32_t r = d->rng >> (l - 16);
    unsigned b = (r >> 12) - 8;
    b += r > corr[b];
    return nbits - (uint32_t)((l << 3) + (int)b);
}

static unsigned rd_decode(rdec *d, unsigned ft)                                                /* entdec.c:139-144 */
{
    unsigned s;
    d->ext = d->rng / ft;
    s = d->val / d->ext;
    return ft - (s + 1 < ft ? s + 1 : ft);
}
static unsigned rd_decode_bin(rdec *d, unsigned bits)                                          /* entdec.c:146-151 */
{
    unsigned s;
    d->ext = d->rng >> bits;
    s = d->val / d->ext;
    return (1u << bits) - (s + 1u < (1u << bits) ? s + 1u : (1u << bits));
}
static void rd_update(rdec *d, unsigned fl, unsigned fh, unsigned ft)                          /* entdec.c:153-159 */
{
    uint32_t s = d->ext * (ft - fh);
    d->val -= s;
    d->rng = fl > 0 ? d->ext * (fh - fl) : d->rng - s;
    rd_norm(d);
}
static int rd_bit_logp(rdec *d, unsigned logp)                                                 /* entdec.c:162-176 */
{
    uint32_t r = d->rng, v = d->val, s = r >> logp;
    int ret = v < s;
    if (!ret) d->val = v - s;
    d->rng = ret ? s : r - s;
    rd_norm(d);
    return ret;
}
static int rd_icdf(rdec *d, const uint8_t *icdf, unsigned ftb)                                 /* entdec.c:178-196 */
{
    uint32_t s = d->rng, v = d->val, r = s >> ftb, t;
    int ret = -1;
    do { t = s; s = r * icdf[++ret]; } while (v < s);
    d->val = v - s;
    d->rng = t - s;
    rd_norm(d);
    return ret;
}
static uint32_t rd_bits(rdec *d, unsigned bits)                                                /* entdec.c:246-266 */
{
    uint32_t window = d->end_window, ret;
    int avail = d->nend_bits;
    if ((unsigned)avail < bits) {
        do { window |= (uint32_t)rd_byte_end(d) << avail; avail += 8; } while (avail <= 32 - 8);
    }
    ret = window & ((1u << bits) - 1u);
    window >>= bits; avail -= bits;
    d->end_window = window; d->nend_bits = avail; d->nbits_total += bits;
    return ret;
}
static uint32_t rd_uint(rdec *d, uint32_t ft)                                                  /* entdec.c:219-244 */
{
    unsigned s; int ftb;
    ft--;
    ftb = ilog(ft);
    if (ftb > 8) {
        uint32_t t; unsigned f;
        ftb -= 8;
        f = (unsigned)(ft >> ftb) + 1;
        s = rd_decode(d, f);
        rd_update(d, s, s + 1, f);
        t = (uint32_t)s << ftb | rd_bits(d, ftb);
        if (t <= ft) return t;
        d->error = 1;
        return ft;
    }
    ft++;
    s = rd_decode(d, (unsigned)ft);
    rd_update(d, s, s + 1, (unsigned)ft);
    return s;
}

/* Laplace-distributed integer (opus/celt/laplace.c:94-134, helper :44-49) */
static int rd_laplace(rdec *d, unsigned fs, int decay)
{
    int val = 0;
    unsigned fl = 0, fm = rd_decode_bin(d, 15);
    if (fm >= fs) {
        val++;
        fl = fs;
        fs = ((32768 - 32 - fs) * (int32_t)(16384 - decay) >> 15) + 1;
        while (fs > 1 && fm >= fl + 2 * fs) {
            fs *= 2; fl += fs;
            fs = ((fs - 2) * (int32_t)decay) >> 15;
            fs += 1;
            val++;
        }
        if (fs <= 1) {
            int di = (fm - fl) >> 1;
            val += di;
            fl += 2 * di;
        }
        if (fm < fl + fs) val = -val; else fl += fs;
    }
    rd_update(d, fl, fl + fs < 32768 ? fl + fs : 32768, 32768);
    return val;
}

/* ================================================================================================
 * Bit-exact integer helpers (opus/celt/bands.c:61-91, mathops.c:43-66, mathops.h:44)
 * ============================================================================================== */
static inline int frac_mul16(int a, int b) { return (16384 + ((int32_t)(int16_t)a * (int16_t)b)) >> 15; }

int co_bitexact_cos(int x)
{
    int32_t tmp = (4096 + ((int32_t)(int16_t)x * (int16_t)x)) >> 13;
    int16_t x2 = (int16_t)tmp;
    x2 = (int16_t)((32767 - x2) + frac_mul16(x2, (-7651 + frac_mul16(x2, (8277 + frac_mul16(-626, x2))))));
    return 1 + x2;
}
int co_bitexact_log2tan(int isin, int icos)
{
    int lc = ilog((uint32_t)icos), ls = ilog((uint32_t)isin);
    icos <<= 15 - lc;
    isin <<= 15 - ls;
    return (ls - lc) * (1 << 11) + frac_mul16(isin, frac_mul16(isin, -2597) + 7932)
                                 - frac_mul16(icos, frac_mul16(icos, -2597) + 7932);
}
unsigned co_isqrt32(uint32_t v)
{
    unsigned g = 0;
    int bshift = (ilog(v) - 1) >> 1;
    unsigned b = 1u << bshift;
    do {
        uint32_t t = (((uint32_t)g << 1) + b) << bshift;
        if (t <= v) { g += b; v -= t; }
        b >>= 1; bshift--;
    } while (bshift >= 0);
    return g;
}
static inline uint32_t lcg(uint32_t s) { return 1664525u * s + 1013904223u; }                 /* bands.c:61-64 */

/* ================================================================================================
 * PVQ codebook enumeration (opus/celt/cwrs.c:430-541)
 * ============================================================================================== */
static inline uint32_t pvq_u(int n, int k) { int a = imin(n, k), b = imax(n, k); return OB_PVQ_U_DATA[OB_PVQ_U_ROW[a] + b]; }
uint32_t co_pvq_v(int n, int k) { return pvq_u(n, k) + pvq_u(n, k + 1); }

uint32_t co_cwrsi(int n, int k, uint32_t i, int *y)                                            /* cwrs.c:463-537 */
{
    uint32_t p, yy = 0; int s, k0; int val;
    while (n > 2) {
        uint32_t q;
        if (k >= n) {                        /* many pulses: walk along row n */
            const uint32_t *row = OB_PVQ_U_DATA + OB_PVQ_U_ROW[n];
            p = row[k + 1];
            s = -(i >= p);
            i -= p & s;
            k0 = k;
            q = row[n];
            if (q > i) {
                k = n;
                do p = OB_PVQ_U_DATA[OB_PVQ_U_ROW[--k] + n]; while (p > i);
            } else for (p = row[k]; p > i; p = row[k]) k--;
            i -= p;
            val = (k0 - k + s) ^ s;
            *y++ = val; yy += (uint32_t)(val * val);
        } else {                             /* many dimensions: walk along column n */
            p = OB_PVQ_U_DATA[OB_PVQ_U_ROW[k] + n];
            q = OB_PVQ_U_DATA[OB_PVQ_U_ROW[k + 1] + n];
            if (p <= i && i < q) { i -= p; *y++ = 0; }
            else {
                s = -(i >= q);
                i -= q & s;
                k0 = k;
                do p = OB_PVQ_U_DATA[OB_PVQ_U_ROW[--k] + n]; while (p > i);
                i -= p;
                val = (k0 - k + s) ^ s;
                *y++ = val; yy += (uint32_t)(val * val);
            }
        }
        n--;
    }
    p = 2 * k + 1;                            /* n == 2 */
    s = -(i >= p);
    i -= p & s;
    k0 = k;
    k = (i + 1) >> 1;
    if (k) i -= 2 * k - 1;
    val = (k0 - k + s) ^ s;
    *y++ = val; yy += (uint32_t)(val * val);
    s = -(int)i;                              /* n == 1 */
    val = (k + s) ^ s;
    *y = val; yy += (uint32_t)(val * val);
    return yy;
}

uint32_t co_icwrs(int n, const int *y)                                                         /* cwrs.c:440-456 */
{
    int j = n - 1, k = abs(y[j]);
    uint32_t i = y[j] < 0;
    do {
        j--;
        i += pvq_u(n - j, k);
        k += abs(y[j]);
        if (y[j] < 0) i += pvq_u(n - j, k + 1);
    } while (j > 0);
    return i;
}

/* ================================================================================================
 * Pulse cache lookups (opus/celt/rate.h:48-87)
 * ============================================================================================== */
static const uint8_t *pcache(int band, int LM) { return OB_CACHE_BITS + OB_CACHE_INDEX[(LM + 1) * NB + band]; }
static int get_pulses(int i) { return i < 8 ? i : (8 + (i & 7)) << ((i >> 3) - 1); }
static int bits2pulses(int band, int LM, int bits)
{
    const uint8_t *cache = pcache(band, LM);
    int lo = 0, hi = cache[0], i;
    bits--;
    for (i = 0; i < 6; i++) {
        int mid = (lo + hi + 1) >> 1;
        if ((int)cache[mid] >= bits) hi = mid; else lo = mid;
    }
    return (bits - (lo == 0 ? -1 : (int)cache[lo]) <= (int)cache[hi] - bits) ? lo : hi;
}
static int pulses2bits(int band, int LM, int pulses) { return pulses == 0 ? 0 : pcache(band, LM)[pulses] + 1; }

/* ================================================================================================
 * Bit allocation (opus/celt/rate.c:248-645, opus/celt/celt.c:272-281) -- int32 only, bit-exact
 * ============================================================================================== */
static void init_caps(int *cap, int LM, int C)
{
    int i;
    for (i = 0; i < NB; i++) {
        int N = (OB_EBANDS[i + 1] - OB_EBANDS[i]) << LM;
        cap[i] = (OB_CACHE_CAPS[NB * (2 * LM + C - 1) + i] + 64) * C * N >> 2;
    }
}

static int interp_bits2pulses(int start, int end, int skip_start, const int *bits1, const int *bits2,
        const int *thresh, const int *cap, int32_t total, int32_t *balance_out, int skip_rsv, int *intensity,
        int intensity_rsv, int *dual_stereo, int dual_stereo_rsv, int *bits, int *ebits, int *fine_priority,
        int C, int LM, rdec *ec)
{
    int32_t psum, left, percoeff, balance;
    int lo = 0, hi = 1 << 6, i, j, coded, done;
    const int alloc_floor = C << BITRES, stereo = C > 1, logM = LM << BITRES;

    for (i = 0; i < 6; i++) {                                    /* rate.c:269-293 */
        int mid = (lo + hi) >> 1;
        psum = 0; done = 0;
        for (j = end; j-- > start;) {
            int tmp = bits1[j] + (mid * (int32_t)bits2[j] >> 6);
            if (tmp >= thresh[j] || done) { done = 1; psum += imin(tmp, cap[j]); }
            else if (tmp >= alloc_floor) psum += alloc_floor;
        }
        if (psum > total) hi = mid; else lo = mid;
    }
    psum = 0; done = 0;
    for (j = end; j-- > start;) {                                /* rate.c:297-312 */
        int tmp = bits1[j] + ((int32_t)lo * bits2[j] >> 6);
        if (tmp < thresh[j] && !done) tmp = tmp >= alloc_floor ? alloc_floor : 0;
        else done = 1;
        tmp = imin(tmp, cap[j]);
        bits[j] = tmp;
        psum += tmp;
    }
    for (coded = end;; coded--) {                                /* band skipping, rate.c:315-391 */
        int band_width, band_bits, rem;
        j = coded - 1;
        if (j <= skip_start) { total += skip_rsv; break; }
        left = total - psum;
        percoeff = (uint32_t)left / (uint32_t)(OB_EBANDS[coded] - OB_EBANDS[start]);
        left -= (OB_EBANDS[coded] - OB_EBANDS[start]) * percoeff;
        rem = imax(left - (OB_EBANDS[j] - OB_EBANDS[start]), 0);
        band_width = OB_EBANDS[coded] - OB_EBANDS[j];
        band_bits = (int)(bits[j] + percoeff * band_width + rem);
        if (band_bits >= imax(thresh[j], alloc_floor + (1 << BITRES))) {
            if (rd_bit_logp(ec, 1)) break;
            psum += 1 << BITRES;
            band_bits -= 1 << BITRES;
        }
        psum -= bits[j] + intensity_rsv;
        if (intensity_rsv > 0) intensity_rsv = OB_LOG2_FRAC[j - start];
        psum += intensity_rsv;
        if (band_bits >= alloc_floor) { psum += alloc_floor; bits[j] = alloc_floor; }
        else bits[j] = 0;
    }
    if (intensity_rsv > 0) *intensity = start + (int)rd_uint(ec, coded + 1 - start);   /* rate.c:395-420 */
    else *intensity = 0;
    if (*intensity <= start) { total += dual_stereo_rsv; dual_stereo_rsv = 0; }
    if (dual_stereo_rsv > 0) *dual_stereo = rd_bit_logp(ec, 1);
    else *dual_stereo = 0;

    left = total - psum;                                         /* rate.c:423-434 */
    percoeff = (uint32_t)left / (uint32_t)(OB_EBANDS[coded] - OB_EBANDS[start]);
    left -= (OB_EBANDS[coded] - OB_EBANDS[start]) * percoeff;
    for (j = start; j < coded; j++) bits[j] += (int)percoeff * (OB_EBANDS[j + 1] - OB_EBANDS[j]);
    for (j = start; j < coded; j++) {
        int tmp = imin(left, OB_EBANDS[j + 1] - OB_EBANDS[j]);
        bits[j] += tmp;
        left -= tmp;
    }
    balance = 0;
    for (j = start; j < coded; j++) {                            /* fine/PVQ split, rate.c:438-516 */
        int N0 = OB_EBANDS[j + 1] - OB_EBANDS[j], N = N0 << LM, den, offset, NClogN;
        int32_t excess, bit = (int32_t)bits[j] + balance;
        if (N > 1) {
            excess = imax(bit - cap[j], 0);
            bits[j] = bit - excess;
            den = C * N + ((C == 2 && N > 2 && !*dual_stereo && j < *intensity) ? 1 : 0);
            NClogN = den * (OB_LOGN[j] + logM);
            offset = (NClogN >> 1) - den * FINE_OFFSET;
            if (N == 2) offset += den << BITRES >> 2;
            if (bits[j] + offset < den * 2 << BITRES) offset += NClogN >> 2;
            else if (bits[j] + offset < den * 3 << BITRES) offset += NClogN >> 3;
            ebits[j] = imax(0, bits[j] + offset + (den << (BITRES - 1)));
            ebits[j] = (int)((uint32_t)ebits[j] / (uint32_t)den) >> BITRES;
            if (C * ebits[j] > (bits[j] >> BITRES)) ebits[j] = bits[j] >> stereo >> BITRES;
            ebits[j] = imin(ebits[j], MAX_FINE_BITS);
            fine_priority[j] = ebits[j] * (den << BITRES) >= bits[j] + offset;
            bits[j] -= C * ebits[j] << BITRES;
        } else {
            excess = imax(0, bit - (C << BITRES));
            bits[j] = bit - excess;
            ebits[j] = 0;
            fine_priority[j] = 1;
        }
        if (excess > 0) {
            int extra_fine = imin(excess >> (stereo + BITRES), MAX_FINE_BITS - ebits[j]);
            int extra_bits;
            ebits[j] += extra_fine;
            extra_bits = extra_fine * C << BITRES;
            fine_priority[j] = extra_bits >= excess - balance;
            excess -= extra_bits;
        }
        balance = excess;
    }
    *balance_out = balance;
    for (; j < end; j++) {                                       /* skipped bands, rate.c:522-528 */
        ebits[j] = bits[j] >> stereo >> BITRES;
        bits[j] = 0;
        fine_priority[j] = ebits[j] < 1;
    }
    return coded;
}

static int compute_allocation(int start, int end, const int *offsets, const int *cap, int alloc_trim,
        int *intensity, int *dual_stereo, int32_t total, int32_t *balance, int *pulses, int *ebits,
        int *fine_priority, int C, int LM, rdec *ec)                                           /* rate.c:534-645 */
{
    int lo, hi, j, skip_start = start, skip_rsv, intensity_rsv = 0, dual_stereo_rsv = 0;
    int bits1[NB], bits2[NB], thresh[NB], trim_offset[NB];
    total = imax(total, 0);
    skip_rsv = total >= 1 << BITRES ? 1 << BITRES : 0;
    total -= skip_rsv;
    if (C == 2) {
        intensity_rsv = OB_LOG2_FRAC[end - start];
        if (intensity_rsv > total) intensity_rsv = 0;
        else {
            total -= intensity_rsv;
            dual_stereo_rsv = total >= 1 << BITRES ? 1 << BITRES : 0;
            total -= dual_stereo_rsv;
        }
    }
    for (j = start; j < end; j++) {
        int w = OB_EBANDS[j + 1] - OB_EBANDS[j];
        thresh[j] = imax(C << BITRES, (3 * w << LM << BITRES) >> 4);
        trim_offset[j] = C * w * (alloc_trim - 5 - LM) * (end - j - 1) * (1 << (LM + BITRES)) >> 6;
        if (w << LM == 1) trim_offset[j] -= C << BITRES;
    }
    lo = 1; hi = 11 - 1;
    do {
        int done = 0, psum = 0, mid = (lo + hi) >> 1;
        for (j = end; j-- > start;) {
            int N = OB_EBANDS[j + 1] - OB_EBANDS[j];
            int bitsj = C * N * OB_ALLOC_VECTORS[mid * NB + j] << LM >> 2;
            if (bitsj > 0) bitsj = imax(0, bitsj + trim_offset[j]);
            bitsj += offsets[j];
            if (bitsj >= thresh[j] || done) { done = 1; psum += imin(bitsj, cap[j]); }
            else if (bitsj >= C << BITRES) psum += C << BITRES;
        }
        if (psum > total) hi = mid - 1; else lo = mid + 1;
    } while (lo <= hi);
    hi = lo--;
    for (j = start; j < end; j++) {
        int N = OB_EBANDS[j + 1] - OB_EBANDS[j];
        int b1 = C * N * OB_ALLOC_VECTORS[lo * NB + j] << LM >> 2;
        int b2 = hi >= 11 ? cap[j] : C * N * OB_ALLOC_VECTORS[hi * NB + j] << LM >> 2;
        if (b1 > 0) b1 = imax(0, b1 + trim_offset[j]);
        if (b2 > 0) b2 = imax(0, b2 + trim_offset[j]);
        if (lo > 0) b1 += offsets[j];
        b2 += offsets[j];
        if (offsets[j] > 0) skip_start = j;
        b2 = imax(0, b2 - b1);
        bits1[j] = b1; bits2[j] = b2;
    }
    return interp_bits2pulses(start, end, skip_start, bits1, bits2, thresh, cap, total, balance, skip_rsv,
            intensity, intensity_rsv, dual_stereo, dual_stereo_rsv, pulses, ebits, fine_priority, C, LM, ec);
}

/* ================================================================================================
 * Float vector helpers (opus/celt/vq.c, bands.c)
 * ============================================================================================== */
static void exp_rotation1(float *X, int len, int stride, float c, float s)                     /* vq.c:47-71 */
{
    int i; float ms = -s; float *p = X;
    for (i = 0; i < len - stride; i++) {
        float x1 = p[0], x2 = p[stride];
        p[stride] = c * x2 + s * x1;
        *p++ = c * x1 + ms * x2;
    }
    p = &X[len - 2 * stride - 1];
    for (i = len - 2 * stride - 1; i >= 0; i--) {
        float x1 = p[0], x2 = p[stride];
        p[stride] = c * x2 + s * x1;
        *p-- = c * x1 + ms * x2;
    }
}
static void exp_rotation_inv(float *X, int len, int stride, int K, int spread)                 /* vq.c:74-117, dir=-1 */
{
    static const int SPREAD_FACTOR[3] = {15, 10, 5};
    int i, stride2 = 0; float c, s, gain, theta;
    if (2 * K >= len || spread == SPREAD_NONE) return;
    gain = (float)(1.0f * len) / (float)(len + SPREAD_FACTOR[spread - 1] * K);
    theta = .5f * (gain * gain);
    c = (float)cos((.5f * 3.141592653f) * theta);
    s = (float)cos((.5f * 3.141592653f) * (1.0f - theta));
    if (len >= 8 * stride) {
        stride2 = 1;
        while ((stride2 * stride2 + stride2) * stride + (stride >> 2) < len) stride2++;
    }
    len = len / stride;
    for (i = 0; i < stride; i++) {
        if (stride2) exp_rotation1(X + i * len, len, stride2, s, c);
        exp_rotation1(X + i * len, len, 1, c, s);
    }
}
static void renormalise(float *X, int N, float gain)                                           /* vq.c:383-407 */
{
    int i; float E = 1e-15f, g;
    for (i = 0; i < N; i++) E += X[i] * X[i];
    g = (1.f / (float)sqrt(E)) * gain;
    for (i = 0; i < N; i++) X[i] = g * X[i];
}
static void haar1(float *X, int N0, int stride)                                                /* bands.c:632-645 */
{
    int i, j;
    N0 >>= 1;
    for (i = 0; i < stride; i++) for (j = 0; j < N0; j++) {
        float t1 = .70710678f * X[stride * 2 * j + i], t2 = .70710678f * X[stride * (2 * j + 1) + i];
        X[stride * 2 * j + i] = t1 + t2;
        X[stride * (2 * j + 1) + i] = t1 - t2;
    }
}
static const int ordery_table[] = {1, 0, 3, 0, 2, 1, 7, 0, 4, 3, 6, 1, 5, 2,
                                   15, 0, 8, 7, 12, 3, 11, 4, 14, 1, 9, 6, 13, 2, 10, 5};   /* bands.c:576-581 */
static void deinterleave_hadamard(float *X, int N0, int stride, int hadamard)                  /* bands.c:583-608 */
{
    float tmp[176]; int i, j, N = N0 * stride;
    if (hadamard) { const int *o = ordery_table + stride - 2;
        for (i = 0; i < stride; i++) for (j = 0; j < N0; j++) tmp[o[i] * N0 + j] = X[j * stride + i]; }
    else for (i = 0; i < stride; i++) for (j = 0; j < N0; j++) tmp[i * N0 + j] = X[j * stride + i];
    memcpy(X, tmp, sizeof(float) * N);
}
static void interleave_hadamard(float *X, int N0, int stride, int hadamard)                    /* bands.c:610-630 */
{
    float tmp[176]; int i, j, N = N0 * stride;
    if (hadamard) { const int *o = ordery_table + stride - 2;
        for (i = 0; i < stride; i++) for (j = 0; j < N0; j++) tmp[j * stride + i] = X[o[i] * N0 + j]; }
    else for (i = 0; i < stride; i++) for (j = 0; j < N0; j++) tmp[j * stride + i] = X[i * N0 + j];
    memcpy(X, tmp, sizeof(float) * N);
}
static void stereo_merge(float *X, float *Y, float mid, int N)                                 /* bands.c:426-476 */
{
    int j; float xp = 0, side = 0, El, Er, lgain, rgain;
    for (j = 0; j < N; j++) { xp += Y[j] * X[j]; side += Y[j] * Y[j]; }
    xp = mid * xp;
    El = mid * mid + side - 2 * xp;
    Er = mid * mid + side + 2 * xp;
    if (Er < 6e-4f || El < 6e-4f) { memcpy(Y, X, sizeof(float) * N); return; }
    lgain = 1.f / (float)sqrt(El);
    rgain = 1.f / (float)sqrt(Er);
    for (j = 0; j < N; j++) {
        float l = mid * X[j], r = Y[j];
        X[j] = lgain * (l - r);
        Y[j] = rgain * (l + r);
    }
}

/* ================================================================================================
 * Band decoding (opus/celt/bands.c:647-1672, decode side only)
 * ============================================================================================== */
typedef struct {
    rdec *ec; int band, intensity, spread, tf_change, disable_inv;
    int32_t remaining_bits; uint32_t seed;
} bctx;
typedef struct { int inv, imid, iside, delta, itheta, qalloc; } split_t;

static int compute_qn(int N, int b, int offset, int pulse_cap, int stereo)                     /* bands.c:647-671 */
{
    static const int16_t exp2_table8[8] = {16384, 17866, 19483, 21247, 23170, 25267, 27554, 30048};
    int qn, qb, N2 = 2 * N - 1;
    if (stereo && N == 2) N2--;
    qb = (b + N2 * offset) / N2;               /* celt_sudiv: C signed division */
    qb = imin(b - pulse_cap - (4 << BITRES), qb);
    qb = imin(8 << BITRES, qb);
    if (qb < (1 << BITRES >> 1)) qn = 1;
    else { qn = exp2_table8[qb & 7] >> (14 - (qb >> BITRES)); qn = (qn + 1) >> 1 << 1; }
    return qn;
}

static void decode_theta(bctx *ctx, split_t *sp, int N, int *b, int B, int B0, int LM, int stereo, int *fill)
{                                                                                              /* bands.c:700-903 */
    rdec *ec = ctx->ec;
    int itheta = 0, inv = 0, imid, iside, delta, qn, i = ctx->band;
    int pulse_cap = OB_LOGN[i] + LM * (1 << BITRES);
    int offset = (pulse_cap >> 1) - (stereo && N == 2 ? QTHETA_OFFSET_TWOPHASE : QTHETA_OFFSET);
    int32_t tell;
    qn = compute_qn(N, *b, offset, pulse_cap, stereo);
    if (stereo && i >= ctx->intensity) qn = 1;
    tell = (int32_t)rd_tell_frac(ec);
    if (qn != 1) {
        if (stereo && N > 2) {                 /* step pdf */
            int p0 = 3, x0 = qn / 2, ft = p0 * (x0 + 1) + x0, x;
            int fs = (int)rd_decode(ec, ft);
            if (fs < (x0 + 1) * p0) x = fs / p0; else x = x0 + 1 + (fs - (x0 + 1) * p0);
            rd_update(ec, x <= x0 ? p0 * x : (x - 1 - x0) + (x0 + 1) * p0,
                          x <= x0 ? p0 * (x + 1) : (x - x0) + (x0 + 1) * p0, ft);
            itheta = x;
        } else if (B0 > 1 || stereo) {         /* uniform pdf */
            itheta = (int)rd_uint(ec, qn + 1);
        } else {                               /* triangular pdf */
            int fs, fl, ft = ((qn >> 1) + 1) * ((qn >> 1) + 1);
            int fm = (int)rd_decode(ec, ft);
            if (fm < ((qn >> 1) * ((qn >> 1) + 1) >> 1)) {
                itheta = (int)(co_isqrt32(8 * (uint32_t)fm + 1) - 1) >> 1;
                fs = itheta + 1;
                fl = itheta * (itheta + 1) >> 1;
            } else {
                itheta = (int)(2 * (qn + 1) - co_isqrt32(8 * (uint32_t)(ft - fm - 1) + 1)) >> 1;
                fs = qn + 1 - itheta;
                fl = ft - ((qn + 1 - itheta) * (qn + 2 - itheta) >> 1);
            }
            rd_update(ec, fl, fl + fs, ft);
        }
        itheta = (int)((uint32_t)((int32_t)itheta * 16384) / (uint32_t)qn);
    } else if (stereo) {
        if (*b > 2 << BITRES && ctx->remaining_bits > 2 << BITRES) inv = rd_bit_logp(ec, 2);
        else inv = 0;
        if (ctx->disable_inv) inv = 0;
        itheta = 0;
    }
    sp->qalloc = (int)((int32_t)rd_tell_frac(ec) - tell);
    *b -= sp->qalloc;
    if (itheta == 0) { imid = 32767; iside = 0; *fill &= (1 << B) - 1; delta = -16384; }
    else if (itheta == 16384) { imid = 0; iside = 32767; *fill &= ((1 << B) - 1) << B; delta = 16384; }
    else {
        imid = co_bitexact_cos((int16_t)itheta);
        iside = co_bitexact_cos((int16_t)(16384 - itheta));
        delta = frac_mul16((N - 1) << 7, co_bitexact_log2tan(iside, imid));
    }
    sp->inv = inv; sp->imid = imid; sp->iside = iside; sp->delta = delta; sp->itheta = itheta;
}

static unsigned collapse_mask_of(const int *iy, int N, int B)                                  /* vq.c:143-163 */
{
    unsigned mask = 0; int N0, i, j;
    if (B <= 1) return 1;
    N0 = N / B;
    for (i = 0; i < B; i++) { unsigned t = 0; for (j = 0; j < N0; j++) t |= (unsigned)iy[i * N0 + j]; mask |= (unsigned)(t != 0) << i; }
    return mask;
}

/* pulse-vector tap: where decode_partition records every decoded iy[] (index = position in the frame's X, channel c at c*N) */
static __thread int16_t *g_iy_tap = NULL;
static __thread uint8_t *g_iy_set = NULL;
static __thread const float *g_iy_base = NULL;

static unsigned decode_partition(bctx *ctx, float *X, int N, int b, int B, float *lowband, int LM, float gain, int fill)
{                                                                                              /* bands.c:943-1105 */
    const uint8_t *cache = pcache(ctx->band, LM);
    unsigned cm = 0;
    int B0 = B;
    if (LM != -1 && b > cache[cache[0]] + 12 && N > 2) {
        split_t sp; float mid, side, *Y, *next_lowband2 = NULL;
        int mbits, sbits, delta, itheta; int32_t rebalance;
        N >>= 1; Y = X + N; LM -= 1;
        if (B == 1) fill = (fill & 1) | (fill << 1);
        B = (B + 1) >> 1;
        decode_theta(ctx, &sp, N, &b, B, B0, LM, 0, &fill);
        delta = sp.delta; itheta = sp.itheta;
        mid = (1.f / 32768) * sp.imid;
        side = (1.f / 32768) * sp.iside;
        if (B0 > 1 && (itheta & 0x3fff)) {
            if (itheta > 8192) delta -= delta >> (4 - LM);
            else delta = imin(0, delta + (N << BITRES >> (5 - LM)));
        }
        mbits = imax(0, imin(b, (b - delta) / 2));
        sbits = b - mbits;
        ctx->remaining_bits -= sp.qalloc;
        if (lowband) next_lowband2 = lowband + N;
        rebalance = ctx->remaining_bits;
        if (mbits >= sbits) {
            cm = decode_partition(ctx, X, N, mbits, B, lowband, LM, gain * mid, fill);
            rebalance = mbits - (rebalance - ctx->remaining_bits);
            if (rebalance > 3 << BITRES && itheta != 0) sbits += rebalance - (3 << BITRES);
            cm |= decode_partition(ctx, Y, N, sbits, B, next_lowband2, LM, gain * side, fill >> B) << (B0 >> 1);
        } else {
            cm = decode_partition(ctx, Y, N, sbits, B, next_lowband2, LM, gain * side, fill >> B) << (B0 >> 1);
            rebalance = sbits - (rebalance - ctx->remaining_bits);
            if (rebalance > 3 << BITRES && itheta != 16384) mbits += rebalance - (3 << BITRES);
            cm |= decode_partition(ctx, X, N, mbits, B, lowband, LM, gain * mid, fill);
        }
    } else {
        int q = bits2pulses(ctx->band, LM, b);
        int curr_bits = pulses2bits(ctx->band, LM, q);
        ctx->remaining_bits -= curr_bits;
        while (ctx->remaining_bits < 0 && q > 0) {
            ctx->remaining_bits += curr_bits;
            q--;
            curr_bits = pulses2bits(ctx->band, LM, q);
            ctx->remaining_bits -= curr_bits;
        }
        if (q != 0) {                           /* alg_unquant, vq.c:363-380 */
            int K = get_pulses(q), iy[176], j; uint32_t Ryy; float g;
            Ryy = co_cwrsi(N, K, rd_uint(ctx->ec, co_pvq_v(N, K)), iy);
            if (g_iy_tap) for (j = 0; j < N; j++) { g_iy_tap[(X - g_iy_base) + j] = (int16_t)iy[j]; g_iy_set[(X - g_iy_base) + j] = 1; }
            g = (1.f / (float)sqrt((float)Ryy)) * gain;      /* normalise_residual, vq.c:121-141 */
            for (j = 0; j < N; j++) X[j] = g * iy[j];
            exp_rotation_inv(X, N, B, K, ctx->spread);
            cm = collapse_mask_of(iy, N, B);
        } else {
            unsigned cm_mask = (unsigned)(1UL << B) - 1; int j;
            fill &= (int)cm_mask;
            if (!fill) memset(X, 0, sizeof(float) * N);
            else {
                if (lowband == NULL) {
                    for (j = 0; j < N; j++) { ctx->seed = lcg(ctx->seed); X[j] = (float)((int32_t)ctx->seed >> 20); }
                    cm = cm_mask;
                } else {
                    for (j = 0; j < N; j++) {
                        float tmp = 1.0f / 256;
                        ctx->seed = lcg(ctx->seed);
                        tmp = (ctx->seed & 0x8000) ? tmp : -tmp;
                        X[j] = lowband[j] + tmp;
                    }
                    cm = (unsigned)fill;
                }
                renormalise(X, N, gain);
            }
        }
    }
    return cm;
}

static unsigned decode_band_n1(bctx *ctx, float *X, float *Y, float *lowband_out)              /* bands.c:904-937 */
{
    float *x = X; int c;
    for (c = 0; c < 1 + (Y != NULL); c++) {
        int sign = 0;
        if (ctx->remaining_bits >= 1 << BITRES) { sign = (int)rd_bits(ctx->ec, 1); ctx->remaining_bits -= 1 << BITRES; }
        x[0] = sign ? -1.f : 1.f;
        x = Y;
    }
    if (lowband_out) lowband_out[0] = X[0];
    return 1;
}

static unsigned decode_band(bctx *ctx, float *X, int N, int b, int B, float *lowband, int LM, float *lowband_out,
        float gain, float *lowband_scratch, int fill)                                          /* bands.c:1109-1231 */
{
    static const uint8_t bit_interleave[16] = {0, 1, 1, 1, 2, 3, 3, 3, 2, 3, 3, 3, 2, 3, 3, 3};
    static const uint8_t bit_deinterleave[16] = {0x00, 0x03, 0x0C, 0x0F, 0x30, 0x33, 0x3C, 0x3F,
                                                 0xC0, 0xC3, 0xCC, 0xCF, 0xF0, 0xF3, 0xFC, 0xFF};
    int N0 = N, N_B = N / B, N_B0, B0 = B, time_divide = 0, recombine = 0, longBlocks = B0 == 1, k;
    int tf_change = ctx->tf_change;
    unsigned cm;
    if (N == 1) return decode_band_n1(ctx, X, NULL, lowband_out);
    if (tf_change > 0) recombine = tf_change;
    if (lowband_scratch && lowband && (recombine || ((N_B & 1) == 0 && tf_change < 0) || B0 > 1)) {
        memcpy(lowband_scratch, lowband, sizeof(float) * N);
        lowband = lowband_scratch;
    }
    for (k = 0; k < recombine; k++) {
        if (lowband) haar1(lowband, N >> k, 1 << k);
        fill = bit_interleave[fill & 0xF] | bit_interleave[fill >> 4] << 2;
    }
    B >>= recombine;
    N_B <<= recombine;
    while ((N_B & 1) == 0 && tf_change < 0) {
        if (lowband) haar1(lowband, N_B, B);
        fill |= fill << B;
        B <<= 1; N_B >>= 1;
        time_divide++; tf_change++;
    }
    B0 = B; N_B0 = N_B;
    if (B0 > 1 && lowband) deinterleave_hadamard(lowband, N_B >> recombine, B0 << recombine, longBlocks);
    cm = decode_partition(ctx, X, N, b, B, lowband, LM, gain, fill);
    if (B0 > 1) interleave_hadamard(X, N_B >> recombine, B0 << recombine, longBlocks);
    N_B = N_B0; B = B0;
    for (k = 0; k < time_divide; k++) { B >>= 1; N_B <<= 1; cm |= cm >> B; haar1(X, N_B, B); }
    for (k = 0; k < recombine; k++) { cm = bit_deinterleave[cm]; haar1(X, N0 >> k, 1 << k); }
    B <<= recombine;
    if (lowband_out) {
        float n = (float)sqrt((float)N0); int j;
        for (j = 0; j < N0; j++) lowband_out[j] = n * X[j];
    }
    cm &= (1u << B) - 1;
    return cm;
}

static unsigned decode_band_stereo(bctx *ctx, float *X, float *Y, int N, int b, int B, float *lowband, int LM,
        float *lowband_out, float *lowband_scratch, int fill)                                  /* bands.c:1235-1381 */
{
    split_t sp; float mid, side; unsigned cm; int mbits, sbits, delta, itheta, inv, orig_fill = fill, j;
    if (N == 1) return decode_band_n1(ctx, X, Y, lowband_out);
    decode_theta(ctx, &sp, N, &b, B, B, LM, 1, &fill);
    inv = sp.inv; delta = sp.delta; itheta = sp.itheta;
    mid = (1.f / 32768) * sp.imid;
    side = (1.f / 32768) * sp.iside;
    if (N == 2) {
        int c, sign = 0; float *x2, *y2;
        mbits = b; sbits = 0;
        if (itheta != 0 && itheta != 16384) sbits = 1 << BITRES;
        mbits -= sbits;
        c = itheta > 8192;
        ctx->remaining_bits -= sp.qalloc + sbits;
        x2 = c ? Y : X; y2 = c ? X : Y;
        if (sbits) sign = (int)rd_bits(ctx->ec, 1);
        sign = 1 - 2 * sign;
        cm = decode_band(ctx, x2, N, mbits, B, lowband, LM, lowband_out, 1.0f, lowband_scratch, orig_fill);
        y2[0] = -sign * x2[1];
        y2[1] = sign * x2[0];
        X[0] = mid * X[0]; X[1] = mid * X[1];
        Y[0] = side * Y[0]; Y[1] = side * Y[1];
        { float t = X[0]; X[0] = t - Y[0]; Y[0] = t + Y[0]; t = X[1]; X[1] = t - Y[1]; Y[1] = t + Y[1]; }
    } else {
        int32_t rebalance;
        mbits = imax(0, imin(b, (b - delta) / 2));
        sbits = b - mbits;
        ctx->remaining_bits -= sp.qalloc;
        rebalance = ctx->remaining_bits;
        if (mbits >= sbits) {
            cm = decode_band(ctx, X, N, mbits, B, lowband, LM, lowband_out, 1.0f, lowband_scratch, fill);
            rebalance = mbits - (rebalance - ctx->remaining_bits);
            if (rebalance > 3 << BITRES && itheta != 0) sbits += rebalance - (3 << BITRES);
            cm |= decode_band(ctx, Y, N, sbits, B, NULL, LM, NULL, side, NULL, fill >> B);
        } else {
            cm = decode_band(ctx, Y, N, sbits, B, NULL, LM, NULL, side, NULL, fill >> B);
            rebalance = sbits - (rebalance - ctx->remaining_bits);
            if (rebalance > 3 << BITRES && itheta != 16384) mbits += rebalance - (3 << BITRES);
            cm |= decode_band(ctx, X, N, mbits, B, lowband, LM, lowband_out, 1.0f, lowband_scratch, fill);
        }
    }
    if (N != 2) stereo_merge(X, Y, mid, N);
    if (inv) for (j = 0; j < N; j++) Y[j] = -Y[j];
    return cm;
}

static void decode_all_bands(int start, int end, float *X_, float *Y_, uint8_t *collapse_masks, const int *pulses,
        int shortBlocks, int spread, int dual_stereo, int intensity, const int *tf_res, int32_t total_bits,
        int32_t balance, rdec *ec, int LM, int codedBands, uint32_t *seed, int disable_inv)   /* bands.c:1398-1672 */
{
    const int M = 1 << LM, B = shortBlocks ? M : 1, C = Y_ ? 2 : 1, norm_offset = M * OB_EBANDS[start];
    float normbuf[2 * 8 * 100];
    float *norm = normbuf, *norm2 = norm + M * OB_EBANDS[NB - 1] - norm_offset;
    float *lowband_scratch = X_ + M * OB_EBANDS[NB - 1];
    int i, lowband_offset = 0, update_lowband = 1;
    bctx ctx;
    ctx.ec = ec; ctx.intensity = intensity; ctx.seed = *seed; ctx.spread = spread; ctx.disable_inv = disable_inv;
    for (i = start; i < end; i++) {
        int32_t tell, remaining_bits, curr_balance;
        int b, N, effective_lowband = -1, tf_change, last = (i == end - 1);
        unsigned x_cm, y_cm;
        float *X = X_ + M * OB_EBANDS[i], *Y = Y_ ? Y_ + M * OB_EBANDS[i] : NULL;
        ctx.band = i;
        N = M * OB_EBANDS[i + 1] - M * OB_EBANDS[i];
        tell = (int32_t)rd_tell_frac(ec);
        if (i != start) balance -= tell;
        remaining_bits = total_bits - tell - 1;
        ctx.remaining_bits = remaining_bits;
        if (i <= codedBands - 1) {
            curr_balance = balance / imin(3, codedBands - i);          /* celt_sudiv */
            b = imax(0, imin(16383, imin(remaining_bits + 1, pulses[i] + curr_balance)));
        } else b = 0;
        if ((M * OB_EBANDS[i] - N >= M * OB_EBANDS[start] || i == start + 1) && (update_lowband || lowband_offset == 0))
            lowband_offset = i;
        /* special_hybrid_folding (bands.c:1384-1395) copies nothing for start==0 */
        tf_change = tf_res[i];
        ctx.tf_change = tf_change;
        if (last) lowband_scratch = NULL;
        if (lowband_offset != 0 && (spread != SPREAD_AGGRESSIVE || B > 1 || tf_change < 0)) {
            int fold_start, fold_end, fold_i;
            effective_lowband = imax(0, M * OB_EBANDS[lowband_offset] - norm_offset - N);
            fold_start = lowband_offset;
            while (M * OB_EBANDS[--fold_start] > effective_lowband + norm_offset) ;
            fold_end = lowband_offset - 1;
            while (++fold_end < i && M * OB_EBANDS[fold_end] < effective_lowband + norm_offset + N) ;
            x_cm = y_cm = 0;
            fold_i = fold_start;
            do { x_cm |= collapse_masks[fold_i * C + 0]; y_cm |= collapse_masks[fold_i * C + C - 1]; } while (++fold_i < fold_end);
        } else x_cm = y_cm = (1u << B) - 1;
        if (dual_stereo && i == intensity) {
            int j;
            dual_stereo = 0;
            for (j = 0; j < M * OB_EBANDS[i] - norm_offset; j++) norm[j] = .5f * (norm[j] + norm2[j]);
        }
        if (dual_stereo) {
            x_cm = decode_band(&ctx, X, N, b / 2, B, effective_lowband != -1 ? norm + effective_lowband : NULL, LM,
                    last ? NULL : norm + M * OB_EBANDS[i] - norm_offset, 1.0f, lowband_scratch, (int)x_cm);
            y_cm = decode_band(&ctx, Y, N, b / 2, B, effective_lowband != -1 ? norm2 + effective_lowband : NULL, LM,
                    last ? NULL : norm2 + M * OB_EBANDS[i] - norm_offset, 1.0f, lowband_scratch, (int)y_cm);
        } else {
            if (Y != NULL)
                x_cm = decode_band_stereo(&ctx, X, Y, N, b, B, effective_lowband != -1 ? norm + effective_lowband : NULL, LM,
                        last ? NULL : norm + M * OB_EBANDS[i] - norm_offset, lowband_scratch, (int)(x_cm | y_cm));
            else
                x_cm = decode_band(&ctx, X, N, b, B, effective_lowband != -1 ? norm + effective_lowband : NULL, LM,
                        last ? NULL : norm + M * OB_EBANDS[i] - norm_offset, 1.0f, lowband_scratch, (int)(x_cm | y_cm));
            y_cm = x_cm;
        }
        collapse_masks[i * C + 0] = (uint8_t)x_cm;
        collapse_masks[i * C + C - 1] = (uint8_t)y_cm;
        balance += pulses[i] + tell;
        update_lowband = b > (N << BITRES);
    }
    *seed = ctx.seed;
}

static void anti_collapse(float *X_, const uint8_t *collapse_masks, int LM, int C, int size, int start, int end,
        const float *logE, const float *prev1logE, const float *prev2logE, const int *pulses, uint32_t seed)
{                                                                                              /* bands.c:268-362 */
    int c, i, j, k;
    for (i = start; i < end; i++) {
        int N0 = OB_EBANDS[i + 1] - OB_EBANDS[i];
        int depth = (int)((uint32_t)(1 + pulses[i]) / (uint32_t)N0) >> LM;
        float thresh = .5f * (float)exp(0.6931471805599453094 * (-.125f * depth));
        float sqrt_1 = 1.f / (float)sqrt((float)(N0 << LM));
        for (c = 0; c < C; c++) {
            float prev1 = prev1logE[c * NB + i], prev2 = prev2logE[c * NB + i], Ediff, r, *X;
            int renorm = 0;
            if (C == 1) { prev1 = fmaxf(prev1, prev1logE[NB + i]); prev2 = fmaxf(prev2, prev2logE[NB + i]); }
            Ediff = logE[c * NB + i] - fminf(prev1, prev2);
            Ediff = fmaxf(0, Ediff);
            r = 2.f * (float)exp(0.6931471805599453094 * (-Ediff));
            if (LM == 3) r *= 1.41421356f;
            r = fminf(thresh, r);
            r = r * sqrt_1;
            X = X_ + c * size + (OB_EBANDS[i] << LM);
            for (k = 0; k < 1 << LM; k++) {
                if (!(collapse_masks[i * C + c] & 1 << k)) {
                    for (j = 0; j < N0; j++) { seed = lcg(seed); X[(j << LM) + k] = (seed & 0x8000) ? r : -r; }
                    renorm = 1;
                }
            }
            if (renorm) renormalise(X, N0 << LM, 1.0f);
        }
    }
}

/* ================================================================================================
 * Synthesis: FFT, IMDCT, denormalise, comb filter, de-emphasis
 * ============================================================================================== */
typedef struct { float r, i; } cpx;
static const int16_t *fft_factors(int shift) { return shift == 0 ? OB_FFT_FACTORS480 : shift == 1 ? OB_FFT_FACTORS240 : shift == 2 ? OB_FFT_FACTORS120 : OB_FFT_FACTORS60; }
static const int16_t *fft_bitrev(int shift) { return shift == 0 ? OB_FFT_BITREV480 : shift == 1 ? OB_FFT_BITREV240 : shift == 2 ? OB_FFT_BITREV120 : OB_FFT_BITREV60; }
#define TW(k) (((const cpx *)OB_FFT_TWIDDLES)[k])
#define CMUL(m, a, b) do { (m).r = (a).r * (b).r - (a).i * (b).i; (m).i = (a).r * (b).i + (a).i * (b).r; } while (0)

static void bfly2(cpx *F, int m, int N)                                                        /* kiss_fft.c:48-100 (m==4) */
{
    const float tw = 0.7071067812f; int i; (void)m;
    for (i = 0; i < N; i++) {
        cpx *F2 = F + 4, t;
        t = F2[0]; F2[0].r = F[0].r - t.r; F2[0].i = F[0].i - t.i; F[0].r += t.r; F[0].i += t.i;
        t.r = (F2[1].r + F2[1].i) * tw; t.i = (F2[1].i - F2[1].r) * tw;
        F2[1].r = F[1].r - t.r; F2[1].i = F[1].i - t.i; F[1].r += t.r; F[1].i += t.i;
        t.r = F2[2].i; t.i = -F2[2].r;
        F2[2].r = F[2].r - t.r; F2[2].i = F[2].i - t.i; F[2].r += t.r; F[2].i += t.i;
        t.r = (F2[3].i - F2[3].r) * tw; t.i = (-(F2[3].i + F2[3].r)) * tw;
        F2[3].r = F[3].r - t.r; F2[3].i = F[3].i - t.i; F[3].r += t.r; F[3].i += t.i;
        F += 8;
    }
}
static void bfly4(cpx *F, int fstride, int m, int N, int mm)                                   /* kiss_fft.c:102-171 */
{
    int i, j;
    if (m == 1) {
        for (i = 0; i < N; i++) {
            cpx s0, s1;
            s0.r = F[0].r - F[2].r; s0.i = F[0].i - F[2].i;
            F[0].r += F[2].r; F[0].i += F[2].i;
            s1.r = F[1].r + F[3].r; s1.i = F[1].i + F[3].i;
            F[2].r = F[0].r - s1.r; F[2].i = F[0].i - s1.i;
            F[0].r += s1.r; F[0].i += s1.i;
            s1.r = F[1].r - F[3].r; s1.i = F[1].i - F[3].i;
            F[1].r = s0.r + s1.i; F[1].i = s0.i - s1.r;
            F[3].r = s0.r - s1.i; F[3].i = s0.i + s1.r;
            F += 4;
        }
    } else {
        cpx *beg = F; const int m2 = 2 * m, m3 = 3 * m;
        for (i = 0; i < N; i++) {
            int t1 = 0, t2 = 0, t3 = 0;
            F = beg + i * mm;
            for (j = 0; j < m; j++) {
                cpx s0, s1, s2, s3, s4, s5;
                CMUL(s0, F[m], TW(t1)); CMUL(s1, F[m2], TW(t2)); CMUL(s2, F[m3], TW(t3));
                s5.r = F->r - s1.r; s5.i = F->i - s1.i;
                F->r += s1.r; F->i += s1.i;
                s3.r = s0.r + s2.r; s3.i = s0.i + s2.i;
                s4.r = s0.r - s2.r; s4.i = s0.i - s2.i;
                F[m2].r = F->r - s3.r; F[m2].i = F->i - s3.i;
                t1 += fstride; t2 += fstride * 2; t3 += fstride * 3;
                F->r += s3.r; F->i += s3.i;
                F[m].r = s5.r + s4.i; F[m].i = s5.i - s4.r;
                F[m3].r = s5.r - s4.i; F[m3].i = s5.i + s4.r;
                ++F;
            }
        }
    }
}
static void bfly3(cpx *F, int fstride, int m, int N, int mm)                                   /* kiss_fft.c:176-236 */
{
    int i, k; const int m2 = 2 * m; cpx *beg = F; const float epi3i = TW(fstride * m).i;
    for (i = 0; i < N; i++) {
        int t1 = 0, t2 = 0;
        F = beg + i * mm;
        for (k = m; k; k--) {
            cpx s0, s1, s2, s3;
            CMUL(s1, F[m], TW(t1)); CMUL(s2, F[m2], TW(t2));
            s3.r = s1.r + s2.r; s3.i = s1.i + s2.i;
            s0.r = s1.r - s2.r; s0.i = s1.i - s2.i;
            t1 += fstride; t2 += fstride * 2;
            F[m].r = F->r - s3.r * .5f; F[m].i = F->i - s3.i * .5f;
            s0.r *= epi3i; s0.i *= epi3i;
            F->r += s3.r; F->i += s3.i;
            F[m2].r = F[m].r + s0.i; F[m2].i = F[m].i - s0.r;
            F[m].r = F[m].r - s0.i; F[m].i = F[m].i + s0.r;
            ++F;
        }
    }
}
static void bfly5(cpx *F, int fstride, int m, int N, int mm)                                   /* kiss_fft.c:240-308 */
{
    int i, u; cpx *beg = F; const cpx ya = TW(fstride * m), yb = TW(fstride * 2 * m);
    for (i = 0; i < N; i++) {
        cpx *F0 = beg + i * mm, *F1 = F0 + m, *F2 = F0 + 2 * m, *F3 = F0 + 3 * m, *F4 = F0 + 4 * m;
        for (u = 0; u < m; ++u) {
            cpx s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12;
            s0 = *F0;
            CMUL(s1, *F1, TW(u * fstride)); CMUL(s2, *F2, TW(2 * u * fstride));
            CMUL(s3, *F3, TW(3 * u * fstride)); CMUL(s4, *F4, TW(4 * u * fstride));
            s7.r = s1.r + s4.r; s7.i = s1.i + s4.i; s10.r = s1.r - s4.r; s10.i = s1.i - s4.i;
            s8.r = s2.r + s3.r; s8.i = s2.i + s3.i; s9.r = s2.r - s3.r; s9.i = s2.i - s3.i;
            F0->r = F0->r + (s7.r + s8.r); F0->i = F0->i + (s7.i + s8.i);
            s5.r = s0.r + (s7.r * ya.r + s8.r * yb.r); s5.i = s0.i + (s7.i * ya.r + s8.i * yb.r);
            s6.r = s10.i * ya.i + s9.i * yb.i; s6.i = -(s10.r * ya.i + s9.r * yb.i);
            F1->r = s5.r - s6.r; F1->i = s5.i - s6.i; F4->r = s5.r + s6.r; F4->i = s5.i + s6.i;
            s11.r = s0.r + (s7.r * yb.r + s8.r * ya.r); s11.i = s0.i + (s7.i * yb.r + s8.i * ya.r);
            s12.r = s9.i * ya.i - s10.i * yb.i; s12.i = s10.r * yb.i - s9.r * ya.i;
            F2->r = s11.r + s12.r; F2->i = s11.i + s12.i; F3->r = s11.r - s12.r; F3->i = s11.i - s12.i;
            ++F0; ++F1; ++F2; ++F3; ++F4;
        }
    }
}
/* In-place decimation-in-time FFT on bit-reversed input (opus_fft_impl, kiss_fft.c:521-567). */
static void fft_impl(cpx *fout, int shift)
{
    const int16_t *fac = fft_factors(shift);
    int fstride[9], L = 0, m, m2, p, i;
    fstride[0] = 1;
    do { p = fac[2 * L]; m = fac[2 * L + 1]; fstride[L + 1] = fstride[L] * p; L++; } while (m != 1);
    m = fac[2 * L - 1];
    for (i = L - 1; i >= 0; i--) {
        m2 = i != 0 ? fac[2 * i - 1] : 1;
        switch (fac[2 * i]) {
        case 2: bfly2(fout, m, fstride[i]); break;
        case 4: bfly4(fout, fstride[i] << shift, m, fstride[i], m2); break;
        case 3: bfly3(fout, fstride[i] << shift, m, fstride[i], m2); break;
        case 5: bfly5(fout, fstride[i] << shift, m, fstride[i], m2); break;
        }
        m = m2;
    }
}
void co_fft(float *data, int shift)   /* natural-order in/out, unscaled (opus_fft_c without the 1/nfft scale) */
{
    int n = 480 >> shift, i; cpx tmp[480]; const int16_t *br = fft_bitrev(shift);
    for (i = 0; i < n; i++) tmp[br[i]] = ((cpx *)data)[i];
    fft_impl(tmp, shift);
    memcpy(data, tmp, sizeof(cpx) * n);
}

void co_mdct_backward(const float *in, float *out, int shift, int stride)                      /* mdct.c:242-342 */
{
    int i, N = 1920, N2, N4; const float *trig = OB_MDCT_TRIG;
    for (i = 0; i < shift; i++) { N >>= 1; trig += N; }
    N2 = N >> 1; N4 = N >> 2;
    {
        const float *xp1 = in, *xp2 = in + stride * (N2 - 1);
        float *yp = out + (OVL >> 1);
        const int16_t *br = fft_bitrev(shift);
        for (i = 0; i < N4; i++) {
            int rev = *br++;
            float yr = *xp2 * trig[i] + *xp1 * trig[N4 + i];
            float yi = *xp1 * trig[i] - *xp2 * trig[N4 + i];
            yp[2 * rev + 1] = yr;
            yp[2 * rev] = yi;
            xp1 += 2 * stride; xp2 -= 2 * stride;
        }
    }
    fft_impl((cpx *)(out + (OVL >> 1)), shift);
    {
        float *yp0 = out + (OVL >> 1), *yp1 = out + (OVL >> 1) + N2 - 2;
        for (i = 0; i < (N4 + 1) >> 1; i++) {
            float re = yp0[1], im = yp0[0], t0 = trig[i], t1 = trig[N4 + i], yr, yi;
            yr = re * t0 + im * t1;
            yi = re * t1 - im * t0;
            re = yp1[1]; im = yp1[0];
            yp0[0] = yr; yp1[1] = yi;
            t0 = trig[N4 - i - 1]; t1 = trig[N2 - i - 1];
            yr = re * t0 + im * t1;
            yi = re * t1 - im * t0;
            yp1[0] = yr; yp0[1] = yi;
            yp0 += 2; yp1 -= 2;
        }
    }
    {
        float *xp1 = out + OVL - 1, *yp1 = out; const float *wp1 = OB_WINDOW, *wp2 = OB_WINDOW + OVL - 1;
        for (i = 0; i < OVL / 2; i++) {
            float x1 = *xp1, x2 = *yp1;
            *yp1++ = *wp2 * x2 - *wp1 * x1;
            *xp1-- = *wp1 * x2 + *wp2 * x1;
            wp1++; wp2--;
        }
    }
}

static void denormalise(const float *X, float *freq, const float *bandLogE, int start, int end, int M, int silence)
{                                                                                              /* bands.c:196-265 */
    int i, N = M * SHORT, bound = M * OB_EBANDS[end]; float *f = freq; const float *x;
    if (silence) { bound = 0; start = end = 0; }
    x = X + M * OB_EBANDS[start];
    for (i = 0; i < M * OB_EBANDS[start]; i++) *f++ = 0;
    for (i = start; i < end; i++) {
        int j = M * OB_EBANDS[i], band_end = M * OB_EBANDS[i + 1];
        float lg = bandLogE[i] + OB_EMEANS[i];
        float g = (float)exp(0.6931471805599453094 * (lg < 32.f ? lg : 32.f));
        do { *f++ = *x++ * g; } while (++j < band_end);
    }
    memset(&freq[bound], 0, sizeof(float) * (N - bound));
}

static void comb_filter_inplace(float *x, int T0, int T1, int N, float g0, float g1, int tapset0, int tapset1, int overlap)
{                                                                                              /* celt.c:190-256, y==x */
    static const float gains[3][3] = {{0.3066406250f, 0.2170410156f, 0.1296386719f},
                                      {0.4638671875f, 0.2680664062f, 0.f}, {0.7998046875f, 0.1000976562f, 0.f}};
    float g00, g01, g02, g10, g11, g12, x0, x1, x2, x3, x4; int i;
    if (g0 == 0 && g1 == 0) return;
    T0 = imax(T0, MINPERIOD); T1 = imax(T1, MINPERIOD);
    g00 = g0 * gains[tapset0][0]; g01 = g0 * gains[tapset0][1]; g02 = g0 * gains[tapset0][2];
    g10 = g1 * gains[tapset1][0]; g11 = g1 * gains[tapset1][1]; g12 = g1 * gains[tapset1][2];
    x1 = x[-T1 + 1]; x2 = x[-T1]; x3 = x[-T1 - 1]; x4 = x[-T1 - 2];
    if (g0 == g1 && T0 == T1 && tapset0 == tapset1) overlap = 0;
    for (i = 0; i < overlap; i++) {
        float f = OB_WINDOW[i] * OB_WINDOW[i];
        x0 = x[i - T1 + 2];
        x[i] = x[i] + ((1.0f - f) * g00) * x[i - T0] + ((1.0f - f) * g01) * (x[i - T0 + 1] + x[i - T0 - 1])
                    + ((1.0f - f) * g02) * (x[i - T0 + 2] + x[i - T0 - 2])
                    + (f * g10) * x2 + (f * g11) * (x1 + x3) + (f * g12) * (x0 + x4);
        x4 = x3; x3 = x2; x2 = x1; x1 = x0;
    }
    if (g1 == 0) return;
    /* comb_filter_const (celt.c:162-185) */
    x4 = x[i - T1 - 2]; x3 = x[i - T1 - 1]; x2 = x[i - T1]; x1 = x[i - T1 + 1];
    for (; i < N; i++) {
        x0 = x[i - T1 + 2];
        x[i] = x[i] + g10 * x2 + g11 * (x1 + x3) + g12 * (x0 + x4);
        x4 = x3; x3 = x2; x2 = x1; x1 = x0;
    }
}

/* ================================================================================================
 * Decoder object and frame driver
 * ============================================================================================== */
struct co_decoder {
    int channels;                         /* CC: channels of the decoder object                         */
    /* everything below is cleared by reset (celt_decoder.c:1514-1529) */
    uint32_t rng, final_range;
    int pf_period, pf_period_old, pf_tapset, pf_tapset_old;
    float pf_gain, pf_gain_old;
    float preemph_mem[2];
    float mem[2][HIST + OVL];
    float oldBandE[2 * NB], oldLogE[2 * NB], oldLogE2[2 * NB], backgroundLogE[2 * NB];
    co_tap_t *tap;
};

void co_decoder_reset(co_decoder *d)
{
    int ch = d->channels, i; co_tap_t *tap = d->tap;
    memset(d, 0, sizeof(*d));
    d->channels = ch; d->tap = tap;
    for (i = 0; i < 2 * NB; i++) d->oldLogE[i] = d->oldLogE2[i] = -28.f;
}
co_decoder *co_decoder_create(int channels)
{
    co_decoder *d;
    if (channels != 1 && channels != 2) return NULL;
    d = (co_decoder *)calloc(1, sizeof(*d));
    if (!d) return NULL;
    d->channels = channels;
    co_decoder_reset(d);
    return d;
}
void co_decoder_destroy(co_decoder *d) { free(d); }
uint32_t co_decoder_final_range(const co_decoder *d) { return d->final_range; }
void co_decoder_set_tap(co_decoder *d, co_tap_t *tap) { d->tap = tap; }
int co_tap_size(void) { return (int)sizeof(co_tap_t); }

static void tf_decode(int start, int end, int isTransient, int *tf_res, int LM, rdec *dec)    /* celt_decoder.c:460-497 */
{
    int i, curr = 0, tf_select = 0, tf_changed = 0, logp = isTransient ? 2 : 4, tf_select_rsv;
    uint32_t budget = dec->storage * 8, tell = (uint32_t)rd_tell(dec);
    tf_select_rsv = LM > 0 && tell + logp + 1 <= budget;
    budget -= tf_select_rsv;
    for (i = start; i < end; i++) {
        if (tell + logp <= budget) {
            curr ^= rd_bit_logp(dec, logp);
            tell = (uint32_t)rd_tell(dec);
            tf_changed |= curr;
        }
        tf_res[i] = curr;
        logp = isTransient ? 4 : 5;
    }
    if (tf_select_rsv && OB_TF_SELECT[LM * 8 + 4 * isTransient + 0 + tf_changed] != OB_TF_SELECT[LM * 8 + 4 * isTransient + 2 + tf_changed])
        tf_select = rd_bit_logp(dec, 1);
    for (i = start; i < end; i++) tf_res[i] = OB_TF_SELECT[LM * 8 + 4 * isTransient + 2 * tf_select + tf_res[i]];
}

/* celt_decode_with_ec (celt_decoder.c:970-1369), data!=NULL && len>1, start=0, downsample=1. */
static int celt_decode_frame(co_decoder *st, const uint8_t *data, int len, float *pcm, int frame_size, int C, int end)
{
    const int CC = st->channels, start = 0;
    int LM, M, N, c, i, silence, isTransient = 0, shortBlocks, intra_ener, spread, alloc_trim, codedBands;
    int pf_pitch = 0, pf_tapset = 0, pf_qg = -1, pf_on = 0, intensity = 0, dual_stereo = 0, anti_collapse_rsv, anti_collapse_on = 0;
    int tf_res[NB], cap[NB], offsets[NB], fine_quant[NB], pulses[NB], fine_priority[NB], coarse_qi[2 * NB];
    int32_t total_bits, tell, bits, balance; float pf_gain = 0;
    uint8_t collapse_masks[2 * NB];
    float X[2 * 960], freq[960];
    float *out_syn[2];
    float *oldBandE = st->oldBandE, *oldLogE = st->oldLogE, *oldLogE2 = st->oldLogE2, *bgLogE = st->backgroundLogE;
    rdec dec; uint32_t seed_in;
    co_tap_t *tap = st->tap;

    for (LM = 0; LM <= 3; LM++) if (SHORT << LM == frame_size) break;
    if (LM > 3) return CO_BAD_ARG;
    M = 1 << LM;
    if (len < 0 || len > 1275 || pcm == NULL) return CO_BAD_ARG;
    N = M * SHORT;
    for (c = 0; c < CC; c++) out_syn[c] = st->mem[c] + HIST - N;
    rd_init(&dec, data, (uint32_t)len);
    memset(coarse_qi, 0, sizeof(coarse_qi)); memset(tf_res, 0, sizeof(tf_res)); memset(pulses, 0, sizeof(pulses));
    memset(fine_quant, 0, sizeof(fine_quant)); memset(fine_priority, 0, sizeof(fine_priority)); memset(offsets, 0, sizeof(offsets));
    memset(X, 0, sizeof(X));

    if (C == 1) for (i = 0; i < NB; i++) oldBandE[i] = fmaxf(oldBandE[i], oldBandE[NB + i]);   /* :1114-1118 */
    total_bits = len * 8;
    tell = rd_tell(&dec);
    if (tell >= total_bits) silence = 1;
    else if (tell == 1) silence = rd_bit_logp(&dec, 15);
    else silence = 0;
    if (silence) { tell = len * 8; dec.nbits_total += tell - rd_tell(&dec); }                  /* :1129-1134 */
    if (start == 0 && tell + 16 <= total_bits) {                                               /* :1139-1152 */
        if (rd_bit_logp(&dec, 1)) {
            int octave = (int)rd_uint(&dec, 6);
            pf_on = 1;
            pf_pitch = (16 << octave) + (int)rd_bits(&dec, 4 + octave) - 1;
            pf_qg = (int)rd_bits(&dec, 3);
            if (rd_tell(&dec) + 2 <= total_bits) pf_tapset = rd_icdf(&dec, OB_TAPSET_ICDF, 2);
            pf_gain = .09375f * (pf_qg + 1);
        }
        tell = rd_tell(&dec);
    }
    if (LM > 0 && tell + 3 <= total_bits) { isTransient = rd_bit_logp(&dec, 3); tell = rd_tell(&dec); }
    shortBlocks = isTransient ? M : 0;
    intra_ener = tell + 3 <= total_bits ? rd_bit_logp(&dec, 3) : 0;

    {   /* unquant_coarse_energy (quant_bands.c:428-491) */
        const uint8_t *prob = OB_E_PROB_MODEL + (LM * 2 + intra_ener) * 42;
        float prev[2] = {0, 0}, coef, beta; int32_t budget = (int32_t)dec.storage * 8;
        if (intra_ener) { coef = 0; beta = OB_BETA_INTRA[0]; } else { beta = OB_BETA_COEF[LM]; coef = OB_PRED_COEF[LM]; }
        for (i = start; i < end; i++) for (c = 0; c < C; c++) {
            int qi; float q, tmp;
            tell = rd_tell(&dec);
            if (budget - tell >= 15) { int pi = 2 * imin(i, 20); qi = rd_laplace(&dec, prob[pi] << 7, prob[pi + 1] << 6); }
            else if (budget - tell >= 2) { static const uint8_t small_icdf[3] = {2, 1, 0}; qi = rd_icdf(&dec, small_icdf, 2); qi = (qi >> 1) ^ -(qi & 1); }
            else if (budget - tell >= 1) qi = -rd_bit_logp(&dec, 1);
            else qi = -1;
            coarse_qi[c * NB + i] = qi;
            q = (float)qi;
            oldBandE[i + c * NB] = fmaxf(-9.f, oldBandE[i + c * NB]);
            tmp = coef * oldBandE[i + c * NB] + prev[c] + q;
            oldBandE[i + c * NB] = tmp;
            prev[c] = prev[c] + q - beta * q;
        }
    }
    tf_decode(start, end, isTransient, tf_res, LM, &dec);
    tell = rd_tell(&dec);
    spread = SPREAD_NORMAL;
    if (tell + 4 <= total_bits) spread = rd_icdf(&dec, OB_SPREAD_ICDF, 5);
    init_caps(cap, LM, C);
    {   /* dynalloc (celt_decoder.c:1217-1246) */
        int dynalloc_logp = 6;
        total_bits <<= BITRES;
        tell = (int32_t)rd_tell_frac(&dec);
        for (i = start; i < end; i++) {
            int width = C * (OB_EBANDS[i + 1] - OB_EBANDS[i]) << LM;
            int quanta = imin(width << BITRES, imax(6 << BITRES, width));
            int loop_logp = dynalloc_logp, boost = 0;
            while (tell + (loop_logp << BITRES) < total_bits && boost < cap[i]) {
                int flag = rd_bit_logp(&dec, loop_logp);
                tell = (int32_t)rd_tell_frac(&dec);
                if (!flag) break;
                boost += quanta;
                total_bits -= quanta;
                loop_logp = 1;
            }
            offsets[i] = boost;
            if (boost > 0) dynalloc_logp = imax(2, dynalloc_logp - 1);
        }
    }
    alloc_trim = tell + (6 << BITRES) <= total_bits ? rd_icdf(&dec, OB_TRIM_ICDF, 7) : 5;
    bits = (((int32_t)len * 8) << BITRES) - (int32_t)rd_tell_frac(&dec) - 1;
    anti_collapse_rsv = isTransient && LM >= 2 && bits >= ((LM + 2) << BITRES) ? (1 << BITRES) : 0;
    bits -= anti_collapse_rsv;
    codedBands = compute_allocation(start, end, offsets, cap, alloc_trim, &intensity, &dual_stereo, bits, &balance,
            pulses, fine_quant, fine_priority, C, LM, &dec);
    for (i = start; i < end; i++) {                                                            /* unquant_fine_energy, quant_bands.c:493-514 */
        if (fine_quant[i] <= 0) continue;
        for (c = 0; c < C; c++) {
            int q2 = (int)rd_bits(&dec, fine_quant[i]);
            float offset = (q2 + .5f) * (1 << (14 - fine_quant[i])) * (1.f / 16384) - .5f;
            oldBandE[i + c * NB] += offset;
        }
    }
    for (c = 0; c < CC; c++) memmove(st->mem[c], st->mem[c] + N, sizeof(float) * (HIST - N + OVL));   /* :1265-1267 */

    seed_in = st->rng;
    memset(collapse_masks, 0, sizeof(collapse_masks));
    if (tap) { memset(tap->iy_set, 0, sizeof(tap->iy_set)); g_iy_tap = tap->iy; g_iy_set = tap->iy_set; g_iy_base = X; }
    decode_all_bands(start, end, X, C == 2 ? X + N : NULL, collapse_masks, pulses, shortBlocks, spread, dual_stereo,
            intensity, tf_res, len * (8 << BITRES) - anti_collapse_rsv, balance, &dec, LM, codedBands, &st->rng,
            CC == 1);
    g_iy_tap = NULL;
    if (tap) { memcpy(tap->X, X, sizeof(float) * N); if (C == 2) memcpy(tap->X + N, X + N, sizeof(float) * N); tap->seed_out = st->rng; }
    if (anti_collapse_rsv > 0) anti_collapse_on = (int)rd_bits(&dec, 1);
    {   /* unquant_energy_finalise (quant_bands.c:516-542) */
        int bits_left = len * 8 - rd_tell(&dec), prio;
        for (prio = 0; prio < 2; prio++) for (i = start; i < end && bits_left >= C; i++) {
            if (fine_quant[i] >= MAX_FINE_BITS || fine_priority[i] != prio) continue;
            for (c = 0; c < C; c++) {
                int q2 = (int)rd_bits(&dec, 1);
                float offset = (q2 - .5f) * (1 << (14 - fine_quant[i] - 1)) * (1.f / 16384);
                oldBandE[i + c * NB] += offset;
                bits_left--;
            }
        }
    }
    if (anti_collapse_on)
        anti_collapse(X, collapse_masks, LM, C, N, start, end, oldBandE, oldLogE, oldLogE2, pulses, st->rng);
    if (silence) for (i = 0; i < C * NB; i++) oldBandE[i] = -28.f;
    if (tap) memcpy(tap->bandLogE, oldBandE, sizeof(float) * 2 * NB);

    {   /* celt_synthesis (celt_decoder.c:382-458) */
        int B, NBk, shift, b;
        if (isTransient) { B = M; NBk = SHORT; shift = 3; } else { B = 1; NBk = SHORT << LM; shift = 3 - LM; }
        if (CC == 2 && C == 1) {
            float *freq2 = out_syn[1] + OVL / 2;
            denormalise(X, freq, oldBandE, start, end, M, silence);
            memcpy(freq2, freq, sizeof(float) * N);
            if (tap) memcpy(tap->freq, freq, sizeof(float) * N);
            for (b = 0; b < B; b++) co_mdct_backward(&freq2[b], out_syn[0] + NBk * b, shift, B);
            for (b = 0; b < B; b++) co_mdct_backward(&freq[b], out_syn[1] + NBk * b, shift, B);
        } else if (CC == 1 && C == 2) {
            float *freq2 = out_syn[0] + OVL / 2;
            denormalise(X, freq, oldBandE, start, end, M, silence);
            denormalise(X + N, freq2, oldBandE + NB, start, end, M, silence);
            for (i = 0; i < N; i++) freq[i] = .5f * freq[i] + .5f * freq2[i];
            if (tap) memcpy(tap->freq, freq, sizeof(float) * N);
            for (b = 0; b < B; b++) co_mdct_backward(&freq[b], out_syn[0] + NBk * b, shift, B);
        } else {
            for (c = 0; c < CC; c++) {
                denormalise(X + c * N, freq, oldBandE + c * NB, start, end, M, silence);
                if (tap) memcpy(tap->freq + c * N, freq, sizeof(float) * N);
                for (b = 0; b < B; b++) co_mdct_backward(&freq[b], out_syn[c] + NBk * b, shift, B);
            }
        }
    }
    for (c = 0; c < CC; c++) {                                                                 /* :1301-1314 */
        if (tap) memcpy(tap->presyn + c * (960 + OVL), out_syn[c], sizeof(float) * (N + OVL));
        st->pf_period = imax(st->pf_period, MINPERIOD);
        st->pf_period_old = imax(st->pf_period_old, MINPERIOD);
        comb_filter_inplace(out_syn[c], st->pf_period_old, st->pf_period, SHORT, st->pf_gain_old, st->pf_gain,
                st->pf_tapset_old, st->pf_tapset, OVL);
        if (LM != 0)
            comb_filter_inplace(out_syn[c] + SHORT, st->pf_period, pf_pitch, N - SHORT, st->pf_gain, pf_gain,
                    st->pf_tapset, pf_tapset, OVL);
    }
    st->pf_period_old = st->pf_period; st->pf_gain_old = st->pf_gain; st->pf_tapset_old = st->pf_tapset;
    st->pf_period = pf_pitch; st->pf_gain = pf_gain; st->pf_tapset = pf_tapset;
    if (LM != 0) { st->pf_period_old = st->pf_period; st->pf_gain_old = st->pf_gain; st->pf_tapset_old = st->pf_tapset; }

    if (C == 1) memcpy(&oldBandE[NB], oldBandE, sizeof(float) * NB);                           /* :1327-1357 */
    if (!isTransient) {
        memcpy(oldLogE2, oldLogE, sizeof(float) * 2 * NB);
        memcpy(oldLogE, oldBandE, sizeof(float) * 2 * NB);
    } else for (i = 0; i < 2 * NB; i++) oldLogE[i] = fminf(oldLogE[i], oldBandE[i]);
    {
        float max_bg_inc = imin(160, 0 + M) * 0.001f;      /* loss_duration is 0 on this path */
        for (i = 0; i < 2 * NB; i++) bgLogE[i] = fminf(bgLogE[i] + max_bg_inc, oldBandE[i]);
    }
    for (c = 0; c < 2; c++) {
        for (i = 0; i < start; i++) { oldBandE[c * NB + i] = 0; oldLogE[c * NB + i] = oldLogE2[c * NB + i] = -28.f; }
        for (i = end; i < NB; i++) { oldBandE[c * NB + i] = 0; oldLogE[c * NB + i] = oldLogE2[c * NB + i] = -28.f; }
    }
    st->rng = dec.rng;                                                                         /* :1358 */

    for (c = 0; c < CC; c++) {                                                                 /* deemphasis, :249-377 */
        float m = st->preemph_mem[c]; const float coef0 = OB_PREEMPH[0]; const float *x = out_syn[c]; int j;
        for (j = 0; j < N; j++) {
            float tmp = x[j] + 1e-30f + m;
            m = coef0 * tmp;
            pcm[j * CC + c] = tmp * (1.f / 32768.f);
        }
        st->preemph_mem[c] = m;
    }
    if (tap) {
        tap->LM = LM; tap->C = C; tap->end = end; tap->silence = silence; tap->transient = isTransient; tap->intra = intra_ener;
        tap->spread = spread; tap->trim = alloc_trim; tap->coded_bands = codedBands; tap->intensity = intensity;
        tap->dual_stereo = dual_stereo; tap->pf_on = pf_on; tap->pf_pitch = pf_pitch; tap->pf_tapset = pf_tapset; tap->pf_qg = pf_qg;
        tap->anti_collapse_on = anti_collapse_on; tap->anti_collapse_rsv = anti_collapse_rsv;
        tap->total_bits_q3 = len * (8 << BITRES) - anti_collapse_rsv; tap->balance = balance;
        memcpy(tap->tf_res, tf_res, sizeof(tf_res)); memcpy(tap->pulses, pulses, sizeof(pulses));
        memcpy(tap->fine_quant, fine_quant, sizeof(fine_quant)); memcpy(tap->fine_priority, fine_priority, sizeof(fine_priority));
        memcpy(tap->offsets, offsets, sizeof(offsets)); memcpy(tap->coarse_qi, coarse_qi, sizeof(coarse_qi));
        memcpy(tap->collapse_masks, collapse_masks, sizeof(collapse_masks));
        tap->seed_in = seed_in; tap->final_range = dec.rng;
    }
    st->final_range = dec.rng;
    if (rd_tell(&dec) > 8 * len) return CO_INTERNAL_ERROR;
    return frame_size;
}

/* opus_decode_native / opus_decode_frame for CELT-only, code-0 packets (opus_decoder.c:670-811, :237-668;
 * TOC layout opus.c:173-192, opus_decoder.c:1077-1107). */
int co_decode_float(co_decoder *d, const uint8_t *pkt, int len, float *pcm, int frame_size)
{
    int toc, C, end, fs, bw;
    if (!d || !pcm || frame_size <= 0) return CO_BAD_ARG;
    if (pkt == NULL || len <= 0) return CO_UNIMPLEMENTED;          /* packet loss concealment: not on this path */
    toc = pkt[0];
    if (!(toc & 0x80)) return CO_UNIMPLEMENTED;                      /* SILK / hybrid */
    if (toc & 0x3) return CO_UNIMPLEMENTED;                          /* multi-frame packets (codes 1-3) */
    fs = SHORT << ((toc >> 3) & 0x3);
    if (fs > frame_size) return CO_BUFFER_TOO_SMALL;
    if (len <= 2) return CO_UNIMPLEMENTED;                           /* payload <= 1 byte => DTX/PLC (opus_decoder.c:284-290) */
    C = (toc & 0x4) ? 2 : 1;
    bw = (toc >> 5) & 0x3;                                           /* 0 NB, 1 WB, 2 SWB, 3 FB */
    end = bw == 0 ? 13 : bw == 1 ? 17 : bw == 2 ? 19 : 21;
    return celt_decode_frame(d, pkt + 1, len - 1, pcm, fs, C, end);
}

int co_decode_stream(const uint8_t *pkts, const int *lens, int stride, int nframes, int frame_size, int channels,
                     float *pcm_out, uint32_t *ranges, int *samples, co_tap_t *taps)
{
    int f; co_decoder *d = co_decoder_create(channels);
    if (!d) return CO_BAD_ARG;
    for (f = 0; f < nframes; f++) {
        int n;
        co_decoder_set_tap(d, taps ? &taps[f] : NULL);
        n = co_decode_float(d, pkts + (size_t)f * stride, lens[f], pcm_out + (size_t)f * frame_size * channels, frame_size);
        if (samples) samples[f] = n;
        if (ranges) ranges[f] = d->final_range;
        if (n < 0 && !samples) { co_decoder_destroy(d); return n; }
    }
    co_decoder_destroy(d);
    return 0;
}
