// dec_symbols.cuh -- CELT symbol decoding, ONE (stream, frame) PER THREAD.
//
// The whole range-coded side of a CELT frame (flags, coarse/fine energy indices, tf, spread, dynalloc,
// trim, bit allocation, every theta / PVQ index / sign of quant_all_bands, anti-collapse and energy
// finalise bits) is a pure integer function of the packet bytes, so it is decoded here
// with no access to per-stream state and emitted as an ObFrameIR for the parallel kernels.
//
// The range decoder state lives in registers; the split recursion of quant_partition
// (opus/celt/bands.c:943-1105) is an explicit 5-deep stack; cwrsi (opus/celt/cwrs.c:463-537) writes the
// pulse vector straight to the IR as int16.
//
// This header is device code, but it is written so that g++ can compile it as plain C++ as well
// (OB_DEV -> static inline): tests/host_emul uses that to check the per-thread logic against the oracle
// on the CPU-only build box.  The product never runs it on the host.
#pragma once
#include "ob_ir.h"

#ifdef __CUDACC__
#define OB_DEV __device__ __forceinline__
#define OB_MEM __device__ __forceinline__
#define OB_DEV_NOINLINE static __device__ __noinline__
#define OB_STAGE static __device__ __noinline__      // a whole encoder stage: compiled once, called (warp-uniformly) from the frame driver
#define OB_TABLE(type, name, n) static __device__ const type name[n]
#define OB_CLZ(x) __clz((int)(x))
#else
#define OB_DEV static inline
#define OB_MEM inline
#define OB_DEV_NOINLINE static
#define OB_STAGE static
#define OB_TABLE(type, name, n) static const type name[n]
#define OB_CLZ(x) __builtin_clz(x)
#endif
#include "celt_tables.inc"

#define OB_BITRES 3

OB_DEV int ob_imin(int a, int b) { return a < b ? a : b; }
OB_DEV int ob_imax(int a, int b) { return a > b ? a : b; }
OB_DEV int ob_ilog(uint32_t v) { return v ? 32 - OB_CLZ(v) : 0; }              // EC_ILOG (ecintrin.h:86)

// ------------------------------------------------------------------------------------------------
// Range decoder (opus/celt/entdec.c; struct ec_ctx opus/celt/entcode.h:62-91)
// ------------------------------------------------------------------------------------------------
struct ObRangeDec {
    const uint8_t *buf;
    uint32_t storage, offs, end_offs, end_window, rng, val, ext;
    int nend_bits, nbits_total, rem, error;

    OB_MEM int read_byte() { return offs < storage ? buf[offs++] : 0; }                         // entdec.c:91-93
    OB_MEM int read_byte_end() { return end_offs < storage ? buf[storage - ++end_offs] : 0; }   // entdec.c:95-98
    OB_MEM void normalize()                                                                     // entdec.c:102-117
    {
        while (rng <= 0x800000u) {
            nbits_total += 8;
            rng <<= 8;
            int sym = rem;
            rem = read_byte();
            sym = (sym << 8 | rem) >> 1;
            val = ((val << 8) + (255u & ~(uint32_t)sym)) & 0x7FFFFFFFu;
        }
    }
    OB_MEM void init(const uint8_t *b, uint32_t len)                                            // entdec.c:119-137
    {
        buf = b; storage = len; end_offs = 0; end_window = 0; nend_bits = 0; nbits_total = 9; offs = 0;
        rng = 128u; ext = 0; error = 0;
        rem = read_byte();
        val = rng - 1 - (uint32_t)(rem >> 1);
        normalize();
    }
    OB_MEM int tell() const { return nbits_total - ob_ilog(rng); }                             // entcode.h:111-113
    OB_MEM uint32_t tell_frac() const                                                          // entcode.c:69-84
    {
        int l = ob_ilog(rng);
        uint32_t r = rng >> (l - 16);
        uint32_t b = (r >> 12) - 8;
        // correction[] = {35733, 38967, 42495, 46340, 50535, 55109, 60097, 65535}
        const uint32_t corr = b == 0 ? 35733u : b == 1 ? 38967u : b == 2 ? 42495u : b == 3 ? 46340u
                            : b == 4 ? 50535u : b == 5 ? 55109u : b == 6 ? 60097u : 65535u;
        b += r > corr;
        return ((uint32_t)nbits_total << OB_BITRES) - (uint32_t)((l << 3) + (int)b);
    }
    OB_MEM uint32_t decode(uint32_t ft)                                                         // entdec.c:139-144
    {
        ext = rng / ft;
        uint32_t s = val / ext;
        return ft - (s + 1 < ft ? s + 1 : ft);
    }
    OB_MEM uint32_t decode_bin(uint32_t bits)                                                   // entdec.c:146-151
    {
        ext = rng >> bits;
        uint32_t s = val / ext;
        uint32_t top = 1u << bits;
        return top - (s + 1u < top ? s + 1u : top);
    }
    OB_MEM void update(uint32_t fl, uint32_t fh, uint32_t ft)                                   // entdec.c:153-159
    {
        uint32_t s = ext * (ft - fh);
        val -= s;
        rng = fl > 0 ? ext * (fh - fl) : rng - s;
        normalize();
    }
    OB_MEM int bit_logp(uint32_t logp)                                                          // entdec.c:162-176
    {
        uint32_t s = rng >> logp;
        int ret = val < s;
        if (!ret) val -= s;
        rng = ret ? s : rng - s;
        normalize();
        return ret;
    }
    OB_MEM int icdf(const uint8_t *tab, uint32_t ftb)                                           // entdec.c:178-196
    {
        uint32_t s = rng, r = s >> ftb, t;
        int ret = -1;
        do { t = s; s = r * tab[++ret]; } while (val < s);
        val -= s;
        rng = t - s;
        normalize();
        return ret;
    }
    OB_MEM uint32_t bits(uint32_t n)                                                            // entdec.c:246-266
    {
        uint32_t window = end_window;
        int avail = nend_bits;
        if ((uint32_t)avail < n) {
            do { window |= (uint32_t)read_byte_end() << avail; avail += 8; } while (avail <= 24);
        }
        uint32_t ret = window & ((1u << n) - 1u);
        end_window = window >> n;
        nend_bits = avail - (int)n;
        nbits_total += (int)n;
        return ret;
    }
    OB_MEM uint32_t uint(uint32_t ft)                                                           // entdec.c:219-244
    {
        ft--;
        int ftb = ob_ilog(ft);
        if (ftb > 8) {
            ftb -= 8;
            uint32_t f = (ft >> ftb) + 1;
            uint32_t s = decode(f);
            update(s, s + 1, f);
            uint32_t t = s << ftb | bits((uint32_t)ftb);
            if (t <= ft) return t;
            error = 1;
            return ft;
        }
        ft++;
        uint32_t s = decode(ft);
        update(s, s + 1, ft);
        return s;
    }
    OB_MEM int laplace(uint32_t fs, int decay)                                                  // laplace.c:94-134
    {
        int v = 0;
        uint32_t fl = 0, fm = decode_bin(15);
        if (fm >= fs) {
            v++;
            fl = fs;
            fs = (uint32_t)(((int32_t)(32768 - 32 - fs) * (int32_t)(16384 - decay)) >> 15) + 1;
            while (fs > 1 && fm >= fl + 2 * fs) {
                fs *= 2; fl += fs;
                fs = (uint32_t)(((int32_t)(fs - 2) * (int32_t)decay) >> 15);
                fs += 1;
                v++;
            }
            if (fs <= 1) {
                int di = (int)((fm - fl) >> 1);
                v += di;
                fl += 2 * (uint32_t)di;
            }
            if (fm < fl + fs) v = -v; else fl += fs;
        }
        update(fl, fl + fs < 32768u ? fl + fs : 32768u, 32768u);
        return v;
    }
};

// ------------------------------------------------------------------------------------------------
// Bit-exact integer helpers (opus/celt/bands.c:66-91, mathops.c:43-66, mathops.h:44)
// ------------------------------------------------------------------------------------------------
OB_DEV int ob_frac_mul16(int a, int b) { return (16384 + ((int32_t)(int16_t)a * (int32_t)(int16_t)b)) >> 15; }
OB_DEV int ob_bitexact_cos(int x)
{
    int32_t tmp = (4096 + ((int32_t)(int16_t)x * (int32_t)(int16_t)x)) >> 13;
    int16_t x2 = (int16_t)tmp;
    x2 = (int16_t)((32767 - x2) + ob_frac_mul16(x2, (-7651 + ob_frac_mul16(x2, (8277 + ob_frac_mul16(-626, x2))))));
    return 1 + x2;
}
OB_DEV int ob_bitexact_log2tan(int isin, int icos)
{
    int lc = ob_ilog((uint32_t)icos), ls = ob_ilog((uint32_t)isin);
    icos <<= 15 - lc;
    isin <<= 15 - ls;
    return (ls - lc) * (1 << 11) + ob_frac_mul16(isin, ob_frac_mul16(isin, -2597) + 7932)
                                 - ob_frac_mul16(icos, ob_frac_mul16(icos, -2597) + 7932);
}
OB_DEV uint32_t ob_isqrt32(uint32_t v)
{
    uint32_t g = 0;
    int bshift = (ob_ilog(v) - 1) >> 1;
    uint32_t b = 1u << bshift;
    do {
        uint32_t t = ((g << 1) + b) << bshift;
        if (t <= v) { g += b; v -= t; }
        b >>= 1; bshift--;
    } while (bshift >= 0);
    return g;
}

// ------------------------------------------------------------------------------------------------
// PVQ: U(N,K) table, V(N,K), index -> pulse vector (opus/celt/cwrs.c:430-541)
// ------------------------------------------------------------------------------------------------
// U(n,k) = U(k,n): the reference keeps rows a = min <= 14 of the triangle behind row pointers (two dependent loads); OB_PVQ_U_RECT holds the same
// numbers at [a * 177 + b] (oracle/gen_tables.c), one load -- the symbol kernel is bound by exactly these dependent table loads.
OB_DEV uint32_t ob_pvq_u(int n, int k) { const int a = ob_imin(n, k), b = ob_imax(n, k); return OB_PVQ_U_RECT[a * 177 + b]; }
OB_DEV uint32_t ob_pvq_u_tri(int n, int k) { const int a = ob_imin(n, k), b = ob_imax(n, k); return OB_PVQ_U_DATA[OB_PVQ_U_ROW[a] + b]; }      // the reference's layout (tests)
OB_DEV uint32_t ob_pvq_v(int n, int k) { return ob_pvq_u(n, k) + ob_pvq_u(n, k + 1); }

// Writes n pulse counts to y (int16) and returns the collapse mask of extract_collapse_mask (vq.c:143-163)
// for B blocks of n/B coefficients, computed on the fly; yy: the vector's squared norm (what decode_pulses returns, cwrs.c:535).
OB_DEV uint32_t ob_cwrsi(int n, int k, uint32_t i, int16_t *y, int B, uint32_t &yy)
{
    yy = 0;
    const int blk = n >> (31 - OB_CLZ((uint32_t)B));   // coefficients per short block (B is a power of two)
    int left = blk;                      // coefficients left in the current block
    uint32_t bit = 1, mask = 0;
    uint32_t p;
    int s, k0, val;
#define OB_EMIT(v) do { *y++ = (int16_t)(v); yy += (uint32_t)((v) * (v)); if (v) mask |= bit; if (--left == 0) { left = blk; bit <<= 1; } } while (0)
    // One coefficient per iteration: the sign from U(n,k+1), then the largest kk <= k with U(n,kk) <= i (cwrs.c:474-518 does the same search
    // as two specialised linear scans).  This kernel waits on dependent table loads, not on issue slots, so the first three candidates are
    // fetched together with U(n,k+1) -- one load round trip decides most coefficients -- and only longer runs fall back to the scan.
    while (n > 2) {
        const int km1 = k > 0 ? k - 1 : 0, km2 = k > 1 ? k - 2 : 0;
        const uint32_t pk1 = ob_pvq_u(n, k + 1), u0 = ob_pvq_u(n, k), u1 = ob_pvq_u(n, km1), u2 = ob_pvq_u(n, km2);
        s = -(int)(i >= pk1);
        i -= pk1 & (uint32_t)s;
        k0 = k;
        if (u0 <= i) p = u0;
        else if (u1 <= i) { p = u1; k = km1; }
        else if (u2 <= i) { p = u2; k = km2; }
        else { k -= 3; for (p = ob_pvq_u(n, k); p > i; p = ob_pvq_u(n, k)) k--; }
        i -= p;
        val = (k0 - k + s) ^ s;
        OB_EMIT(val);
        n--;
    }
    p = 2 * (uint32_t)k + 1;
    s = -(int)(i >= p);
    i -= p & (uint32_t)s;
    k0 = k;
    k = (int)((i + 1) >> 1);
    if (k) i -= 2 * (uint32_t)k - 1;
    val = (k0 - k + s) ^ s;
    OB_EMIT(val);
    s = -(int)i;
    val = (k + s) ^ s;
    OB_EMIT(val);
#undef OB_EMIT
    return B > 1 ? mask : 1u;
}

// ------------------------------------------------------------------------------------------------
// Pulse cache (opus/celt/rate.h:48-87)
// ------------------------------------------------------------------------------------------------
OB_DEV const uint8_t *ob_pcache(int band, int LM) { return OB_CACHE_BITS + OB_CACHE_INDEX[(LM + 1) * OB_NB + band]; }
OB_DEV int ob_get_pulses(int i) { return i < 8 ? i : (8 + (i & 7)) << ((i >> 3) - 1); }
OB_DEV int ob_bits2pulses(const uint8_t *cache, int bits)
{
    int lo = 0, hi = cache[0];
    bits--;
    for (int i = 0; i < 6; i++) {
        int mid = (lo + hi + 1) >> 1;
        if ((int)cache[mid] >= bits) hi = mid; else lo = mid;
    }
    return (bits - (lo == 0 ? -1 : (int)cache[lo]) <= (int)cache[hi] - bits) ? lo : hi;
}
OB_DEV int ob_pulses2bits(const uint8_t *cache, int pulses) { return pulses == 0 ? 0 : cache[pulses] + 1; }

// ------------------------------------------------------------------------------------------------
// Bit allocation (opus/celt/rate.c:248-645) -- int32 only, must be bit-exact
// ------------------------------------------------------------------------------------------------
struct ObAlloc {
    int16_t pulses[OB_NB];     // PVQ bits per band, 1/8 bit
    uint8_t ebits[OB_NB];      // fine energy bits
    uint8_t fine_priority[OB_NB];
    int32_t balance;
    int coded_bands, intensity, dual_stereo;
};

OB_DEV_NOINLINE void ob_compute_allocation(ObRangeDec &ec, int end, const int16_t *offsets, const int16_t *cap,
        int alloc_trim, int32_t total, int C, int LM, ObAlloc &out)
{
    const int start = 0;
    int16_t bits1[OB_NB], bits2[OB_NB], thresh[OB_NB], trim_offset[OB_NB];
    int32_t bits[OB_NB];
    int lo, hi, j, skip_start = start, skip_rsv, intensity_rsv = 0, dual_stereo_rsv = 0;
    total = ob_imax(total, 0);
    skip_rsv = total >= 1 << OB_BITRES ? 1 << OB_BITRES : 0;
    total -= skip_rsv;
    if (C == 2) {
        intensity_rsv = OB_LOG2_FRAC[end - start];
        if (intensity_rsv > total) intensity_rsv = 0;
        else {
            total -= intensity_rsv;
            dual_stereo_rsv = total >= 1 << OB_BITRES ? 1 << OB_BITRES : 0;
            total -= dual_stereo_rsv;
        }
    }
    for (j = start; j < end; j++) {
        int w = OB_EBANDS[j + 1] - OB_EBANDS[j];
        thresh[j] = (int16_t)ob_imax(C << OB_BITRES, (3 * w << LM << OB_BITRES) >> 4);
        int to = C * w * (alloc_trim - 5 - LM) * (end - j - 1) * (1 << (LM + OB_BITRES)) >> 6;
        if (w << LM == 1) to -= C << OB_BITRES;
        trim_offset[j] = (int16_t)to;
    }
    lo = 1; hi = 11 - 1;
    do {
        int done = 0, psum = 0, mid = (lo + hi) >> 1;
        for (j = end; j-- > start;) {
            int N = OB_EBANDS[j + 1] - OB_EBANDS[j];
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
        int N = OB_EBANDS[j + 1] - OB_EBANDS[j];
        int b1 = C * N * OB_ALLOC_VECTORS[lo * OB_NB + j] << LM >> 2;
        int b2 = hi >= 11 ? cap[j] : C * N * OB_ALLOC_VECTORS[hi * OB_NB + j] << LM >> 2;
        if (b1 > 0) b1 = ob_imax(0, b1 + trim_offset[j]);
        if (b2 > 0) b2 = ob_imax(0, b2 + trim_offset[j]);
        if (lo > 0) b1 += offsets[j];
        b2 += offsets[j];
        if (offsets[j] > 0) skip_start = j;
        b2 = ob_imax(0, b2 - b1);
        bits1[j] = (int16_t)b1; bits2[j] = (int16_t)b2;
    }
    // ---- interp_bits2pulses (rate.c:248-532) ----
    {
        int32_t psum, left, percoeff, balance;
        int i, coded, done;
        const int alloc_floor = C << OB_BITRES, stereo = C > 1, logM = LM << OB_BITRES;
        lo = 0; hi = 1 << 6;
        for (i = 0; i < 6; i++) {
            int mid = (lo + hi) >> 1;
            psum = 0; done = 0;
            for (j = end; j-- > start;) {
                int tmp = bits1[j] + (mid * (int32_t)bits2[j] >> 6);
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
                if (ec.bit_logp(1)) break;
                psum += 1 << OB_BITRES;
                band_bits -= 1 << OB_BITRES;
            }
            psum -= bits[j] + intensity_rsv;
            if (intensity_rsv > 0) intensity_rsv = OB_LOG2_FRAC[j - start];
            psum += intensity_rsv;
            if (band_bits >= alloc_floor) { psum += alloc_floor; bits[j] = alloc_floor; }
            else bits[j] = 0;
        }
        if (intensity_rsv > 0) out.intensity = start + (int)ec.uint((uint32_t)(coded + 1 - start));
        else out.intensity = 0;
        if (out.intensity <= start) { total += dual_stereo_rsv; dual_stereo_rsv = 0; }
        if (dual_stereo_rsv > 0) out.dual_stereo = ec.bit_logp(1);
        else out.dual_stereo = 0;

        left = total - psum;
        percoeff = (int32_t)((uint32_t)left / (uint32_t)(OB_EBANDS[coded] - OB_EBANDS[start]));
        left -= (OB_EBANDS[coded] - OB_EBANDS[start]) * percoeff;
        for (j = start; j < coded; j++) bits[j] += (int)percoeff * (OB_EBANDS[j + 1] - OB_EBANDS[j]);
        for (j = start; j < coded; j++) {
            int tmp = ob_imin(left, OB_EBANDS[j + 1] - OB_EBANDS[j]);
            bits[j] += tmp;
            left -= tmp;
        }
        balance = 0;
        for (j = start; j < coded; j++) {
            int N0 = OB_EBANDS[j + 1] - OB_EBANDS[j], N = N0 << LM, den, offset, NClogN, eb, fp;
            int32_t excess, bit = bits[j] + balance, bj;
            if (N > 1) {
                excess = ob_imax(bit - cap[j], 0);
                bj = bit - excess;
                den = C * N + ((C == 2 && N > 2 && !out.dual_stereo && j < out.intensity) ? 1 : 0);
                NClogN = den * (OB_LOGN[j] + logM);
                offset = (NClogN >> 1) - den * 21;                       // FINE_OFFSET
                if (N == 2) offset += den << OB_BITRES >> 2;
                if (bj + offset < den * 2 << OB_BITRES) offset += NClogN >> 2;
                else if (bj + offset < den * 3 << OB_BITRES) offset += NClogN >> 3;
                eb = ob_imax(0, bj + offset + (den << (OB_BITRES - 1)));
                eb = (int)((uint32_t)eb / (uint32_t)den) >> OB_BITRES;
                if (C * eb > (bj >> OB_BITRES)) eb = bj >> stereo >> OB_BITRES;
                eb = ob_imin(eb, 8);                                     // MAX_FINE_BITS
                fp = eb * (den << OB_BITRES) >= bj + offset;
                bj -= C * eb << OB_BITRES;
            } else {
                excess = ob_imax(0, bit - (C << OB_BITRES));
                bj = bit - excess;
                eb = 0;
                fp = 1;
            }
            if (excess > 0) {
                int extra_fine = ob_imin(excess >> (stereo + OB_BITRES), 8 - eb);
                eb += extra_fine;
                int extra_bits = extra_fine * C << OB_BITRES;
                fp = extra_bits >= excess - balance;
                excess -= extra_bits;
            }
            balance = excess;
            out.pulses[j] = (int16_t)bj; out.ebits[j] = (uint8_t)eb; out.fine_priority[j] = (uint8_t)fp;
        }
        out.balance = balance;
        for (; j < end; j++) {
            int eb = bits[j] >> stereo >> OB_BITRES;
            out.ebits[j] = (uint8_t)eb;
            out.pulses[j] = 0;
            out.fine_priority[j] = (uint8_t)(eb < 1);
        }
        out.coded_bands = coded;
    }
}

// ------------------------------------------------------------------------------------------------
// Band symbol decoding: compute_theta + quant_partition + quant_band(+_stereo) + quant_all_bands
// (opus/celt/bands.c:647-1672), decode side, integer results only.
// ------------------------------------------------------------------------------------------------
struct ObBandCtx {
    ObRangeDec *ec;
    ObFrameIR *ir;
    int band, intensity, disable_inv, tf_change;
    int32_t remaining_bits;
    uint32_t lcg;          // LCG steps so far
    int n_leaves;
};
struct ObSplit { int inv, imid, iside, delta, itheta, qalloc; };

OB_DEV int ob_compute_qn(int N, int b, int offset, int pulse_cap, int stereo)                  // bands.c:647-671
{
    int qn, qb, N2 = 2 * N - 1;
    if (stereo && N == 2) N2--;
    qb = (b + N2 * offset) / N2;
    qb = ob_imin(b - pulse_cap - (4 << OB_BITRES), qb);
    qb = ob_imin(8 << OB_BITRES, qb);
    if (qb < (1 << OB_BITRES >> 1)) qn = 1;
    else {
        // exp2_table8 = {16384, 17866, 19483, 21247, 23170, 25267, 27554, 30048}
        const int f = qb & 7;
        const int e = f == 0 ? 16384 : f == 1 ? 17866 : f == 2 ? 19483 : f == 3 ? 21247 : f == 4 ? 23170 : f == 5 ? 25267 : f == 6 ? 27554 : 30048;
        qn = e >> (14 - (qb >> OB_BITRES));
        qn = (qn + 1) >> 1 << 1;
    }
    return qn;
}

OB_DEV_NOINLINE void ob_decode_theta(ObBandCtx &ctx, ObSplit &sp, int N, int &b, int B, int B0, int LM, int stereo, int &fill)
{                                                                                              // bands.c:700-903
    ObRangeDec &ec = *ctx.ec;
    int itheta = 0, inv = 0, imid, iside, delta, qn, i = ctx.band;
    int pulse_cap = OB_LOGN[i] + LM * (1 << OB_BITRES);
    int offset = (pulse_cap >> 1) - (stereo && N == 2 ? 16 : 4);
    qn = ob_compute_qn(N, b, offset, pulse_cap, stereo);
    if (stereo && i >= ctx.intensity) qn = 1;
    int32_t tell = (int32_t)ec.tell_frac();
    if (qn != 1) {
        if (stereo && N > 2) {
            const int p0 = 3, x0 = qn / 2, ft = p0 * (x0 + 1) + x0;
            int x, fs = (int)ec.decode((uint32_t)ft);
            if (fs < (x0 + 1) * p0) x = fs / p0; else x = x0 + 1 + (fs - (x0 + 1) * p0);
            ec.update((uint32_t)(x <= x0 ? p0 * x : (x - 1 - x0) + (x0 + 1) * p0),
                      (uint32_t)(x <= x0 ? p0 * (x + 1) : (x - x0) + (x0 + 1) * p0), (uint32_t)ft);
            itheta = x;
        } else if (B0 > 1 || stereo) {
            itheta = (int)ec.uint((uint32_t)qn + 1);
        } else {
            int fs, fl, ft = ((qn >> 1) + 1) * ((qn >> 1) + 1);
            int fm = (int)ec.decode((uint32_t)ft);
            if (fm < ((qn >> 1) * ((qn >> 1) + 1) >> 1)) {
                itheta = (int)(ob_isqrt32(8 * (uint32_t)fm + 1) - 1) >> 1;
                fs = itheta + 1;
                fl = itheta * (itheta + 1) >> 1;
            } else {
                itheta = (int)(2 * (qn + 1) - ob_isqrt32(8 * (uint32_t)(ft - fm - 1) + 1)) >> 1;
                fs = qn + 1 - itheta;
                fl = ft - ((qn + 1 - itheta) * (qn + 2 - itheta) >> 1);
            }
            ec.update((uint32_t)fl, (uint32_t)(fl + fs), (uint32_t)ft);
        }
        itheta = (int)((uint32_t)(itheta * 16384) / (uint32_t)qn);
    } else if (stereo) {
        if (b > 2 << OB_BITRES && ctx.remaining_bits > 2 << OB_BITRES) inv = ec.bit_logp(2);
        else inv = 0;
        if (ctx.disable_inv) inv = 0;
        itheta = 0;
    }
    sp.qalloc = (int)((int32_t)ec.tell_frac() - tell);
    b -= sp.qalloc;
    if (itheta == 0) { imid = 32767; iside = 0; fill &= (1 << B) - 1; delta = -16384; }
    else if (itheta == 16384) { imid = 0; iside = 32767; fill &= ((1 << B) - 1) << B; delta = 16384; }
    else {
        imid = ob_bitexact_cos((int16_t)itheta);
        iside = ob_bitexact_cos((int16_t)(16384 - itheta));
        delta = ob_frac_mul16((N - 1) << 7, ob_bitexact_log2tan(iside, imid));
    }
    sp.inv = inv; sp.imid = imid; sp.iside = iside; sp.delta = delta; sp.itheta = itheta;
}

OB_DEV void ob_emit_leaf(ObBandCtx &ctx, int off, int n, int K, int kind, int B, float gain, uint32_t yy = 0)
{
    if (ctx.n_leaves < OB_MAX_LEAVES) {
        ObLeaf &l = ctx.ir->leaves[ctx.n_leaves];
        l.off = (uint16_t)off; l.n = (uint8_t)n; l.K = (uint8_t)K; l.kind = (uint8_t)kind; l.B = (uint8_t)B;
        l.lcg_before = (uint16_t)(kind == OB_LEAF_PULSES ? yy : ctx.lcg); l.gain = gain;
    }
    ctx.n_leaves++;
}

// quant_partition (bands.c:943-1105) with the recursion unrolled onto an explicit stack.
struct ObPartFrame {
    int16_t off, N;
    int32_t b, mbits, sbits, rebalance;
    int32_t fill;
    float gain, mid, side;
    uint32_t cm;
    int8_t B, B0, LM, stage, mid_first;
    int16_t itheta;
};

OB_DEV_NOINLINE uint32_t ob_decode_partition(ObBandCtx &ctx, int off, int N, int b, int B, int has_lowband, int LM, float gain, int fill)
{
    ObPartFrame st[6];
    int sp = 0;
    uint32_t ret = 0;
    st[0].off = (int16_t)off; st[0].N = (int16_t)N; st[0].b = b; st[0].B = (int8_t)B; st[0].LM = (int8_t)LM;
    st[0].gain = gain; st[0].fill = fill; st[0].stage = 0;
    // The walk over the split tree is a small state machine.  It is run in two alternating phases so that the threads of a warp
    // (one frame each) meet again at the expensive step: first every thread advances through its cheap split / merge states
    // until it stands on a leaf, then all threads that have a leaf decode it together (the PVQ index decode is ~40 % of this
    // kernel's instructions; interleaved with other threads' split states it ran at 4-9 active lanes).
    while (sp >= 0) {
      bool at_leaf = false;
      while (sp >= 0 && !at_leaf) {
        ObPartFrame &f = st[sp];
        if (f.stage == 0) {
            const uint8_t *cache = ob_pcache(ctx.band, f.LM);
            if (f.LM != -1 && f.b > cache[cache[0]] + 12 && f.N > 2) {
                ObSplit s;
                int n = f.N >> 1, lm = f.LM - 1, bb = f.b, fl = f.fill, B1;
                f.B0 = f.B;
                if (f.B == 1) fl = (fl & 1) | (fl << 1);
                B1 = (f.B + 1) >> 1;
                ob_decode_theta(ctx, s, n, bb, B1, f.B0, lm, 0, fl);
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
                f.itheta = (int16_t)s.itheta;
                f.fill = fl; f.N = (int16_t)n; f.LM = (int8_t)lm; f.B = (int8_t)B1;     // now describe the children
                f.mid_first = f.mbits >= f.sbits;
                f.stage = 1;
                ObPartFrame &c = st[++sp];
                c.N = (int16_t)n; c.B = (int8_t)B1; c.LM = (int8_t)lm; c.stage = 0;
                if (f.mid_first) { c.off = f.off; c.b = f.mbits; c.gain = f.gain * f.mid; c.fill = fl; }
                else { c.off = (int16_t)(f.off + n); c.b = f.sbits; c.gain = f.gain * f.side; c.fill = fl >> B1; }
            } else at_leaf = true;
        } else if (f.stage == 1) {
            // first child done: rebalance (bands.c:1017-1035) and launch the second
            ObPartFrame &c = st[sp + 1];
            c.N = f.N; c.B = f.B; c.LM = f.LM; c.stage = 0;
            if (f.mid_first) {
                f.cm = ret;
                int32_t rb = f.mbits - (f.rebalance - ctx.remaining_bits);
                if (rb > 3 << OB_BITRES && f.itheta != 0) f.sbits += rb - (3 << OB_BITRES);
                c.off = (int16_t)(f.off + f.N); c.b = f.sbits; c.gain = f.gain * f.side; c.fill = f.fill >> f.B;
            } else {
                f.cm = ret << (f.B0 >> 1);
                int32_t rb = f.sbits - (f.rebalance - ctx.remaining_bits);
                if (rb > 3 << OB_BITRES && f.itheta != 16384) f.mbits += rb - (3 << OB_BITRES);
                c.off = f.off; c.b = f.mbits; c.gain = f.gain * f.mid; c.fill = f.fill;
            }
            f.stage = 2;
            sp++;
        } else {
            ret = f.mid_first ? (f.cm | ret << (f.B0 >> 1)) : (f.cm | ret);
            sp--;
        }
      }
      if (at_leaf) {
            ObPartFrame &f = st[sp];
            const uint8_t *cache = ob_pcache(ctx.band, f.LM);
            {
                // leaf: bands.c:1038-1103
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
                    int K = ob_get_pulses(q);
                    uint32_t idx = ctx.ec->uint(ob_pvq_v(f.N, K));
                    uint32_t yy;
                    cm = ob_cwrsi(f.N, K, idx, ctx.ir->iy + f.off, f.B, yy);
                    ob_emit_leaf(ctx, f.off, f.N, K, OB_LEAF_PULSES, f.B, f.gain, yy);
                } else {
                    uint32_t cm_mask = (1u << f.B) - 1;
                    int fl = f.fill & (int)cm_mask;
                    if (!fl) ob_emit_leaf(ctx, f.off, f.N, 0, OB_LEAF_ZERO, f.B, f.gain);
                    else if (!has_lowband) {
                        ob_emit_leaf(ctx, f.off, f.N, 0, OB_LEAF_NOISE, f.B, f.gain);
                        ctx.lcg += (uint32_t)f.N;
                        cm = cm_mask;
                    } else {
                        ob_emit_leaf(ctx, f.off, f.N, 0, OB_LEAF_FOLD, f.B, f.gain);
                        ctx.lcg += (uint32_t)f.N;
                        cm = (uint32_t)fl;
                    }
                }
                ret = cm;
                sp--;
            }
      }
    }
    return ret;
}

// quant_band (bands.c:1109-1231): the integer side -- fill / collapse-mask bookkeeping around the partition.
// Returns cm; appends leaves [*leaf_begin, +*leaf_cnt).
OB_DEV uint32_t ob_decode_band(ObBandCtx &ctx, int off, int N, int b, int B, int has_lowband, int LM, float gain, int fill,
        uint16_t *leaf_begin, uint8_t *leaf_cnt)
{
    int N_B = N / B, B0, time_divide = 0, recombine = 0, k, tf_change = ctx.tf_change;
    uint32_t cm;
    *leaf_begin = (uint16_t)ctx.n_leaves;
    if (N == 1) {                                        // quant_band_n1 (bands.c:904-937), mono call
        int sign = 0;
        if (ctx.remaining_bits >= 1 << OB_BITRES) { sign = (int)ctx.ec->bits(1); ctx.remaining_bits -= 1 << OB_BITRES; }
        ob_emit_leaf(ctx, off, 1, sign, OB_LEAF_ONE, 1, 1.0f);
        *leaf_cnt = 1;
        return 1;
    }
    if (tf_change > 0) recombine = tf_change;
    for (k = 0; k < recombine; k++) {
        // bit_interleave_table = {0,1,1,1,2,3,3,3,2,3,3,3,2,3,3,3}
        int lo = fill & 0xF, hi = fill >> 4;
        int a = (lo & 3 ? 1 : 0) | (lo & 12 ? 2 : 0), c = (hi & 3 ? 1 : 0) | (hi & 12 ? 2 : 0);
        fill = a | c << 2;
    }
    B >>= recombine;
    N_B <<= recombine;
    while ((N_B & 1) == 0 && tf_change < 0) {
        fill |= fill << B;
        B <<= 1; N_B >>= 1;
        time_divide++; tf_change++;
    }
    B0 = B;
    cm = ob_decode_partition(ctx, off, N, b, B, has_lowband, LM, gain, fill);
    for (k = 0; k < time_divide; k++) { B >>= 1; cm |= cm >> B; }
    for (k = 0; k < recombine; k++) {
        // bit_deinterleave_table[cm]: bit j of cm -> bits 2j and 2j+1
        uint32_t r = 0;
        for (int j = 0; j < 4; j++) if (cm & (1u << j)) r |= 3u << (2 * j);
        cm = r;
    }
    B <<= recombine;
    cm &= (1u << B) - 1;
    (void)B0;
    *leaf_cnt = (uint8_t)(ctx.n_leaves - *leaf_begin);
    return cm;
}

// quant_all_bands (bands.c:1398-1672), decode side.
OB_DEV_NOINLINE void ob_decode_all_bands(ObRangeDec &ec, ObFrameIR *ir, int end, int C, const ObAlloc &al, int shortBlocks,
        int spread, const int8_t *tf_res, int32_t total_bits, int LM, int disable_inv, uint8_t *collapse_masks)
{
    const int start = 0, M = 1 << LM, B = shortBlocks ? M : 1, norm_offset = 0;
    int i, lowband_offset = 0, update_lowband = 1, dual_stereo = al.dual_stereo;
    int32_t balance = al.balance;
    ObBandCtx ctx;
    const int N_frame = OB_SHORT << LM;
    ctx.ec = &ec; ctx.ir = ir; ctx.intensity = al.intensity; ctx.disable_inv = disable_inv; ctx.lcg = 0; ctx.n_leaves = 0;
    for (i = start; i < end; i++) {
        ObBand &br = ir->bands[i];
        int32_t tell, remaining_bits, curr_balance;
        int b, N, effective_lowband = -1, tf_change;
        uint32_t x_cm, y_cm;
        const int xoff = M * OB_EBANDS[i], yoff = N_frame + M * OB_EBANDS[i];
        ctx.band = i;
        N = M * OB_EBANDS[i + 1] - M * OB_EBANDS[i];
        tell = (int32_t)ec.tell_frac();
        if (i != start) balance -= tell;
        remaining_bits = total_bits - tell - 1;
        ctx.remaining_bits = remaining_bits;
        if (i <= al.coded_bands - 1) {
            curr_balance = balance / ob_imin(3, al.coded_bands - i);
            b = ob_imax(0, ob_imin(16383, ob_imin(remaining_bits + 1, al.pulses[i] + curr_balance)));
        } else b = 0;
        if ((M * OB_EBANDS[i] - N >= M * OB_EBANDS[start] || i == start + 1) && (update_lowband || lowband_offset == 0))
            lowband_offset = i;
        tf_change = tf_res[i];
        ctx.tf_change = tf_change;
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
        br.flags = 0; br.imid = 0; br.iside = 0; br.leaf_cnt_b = 0; br.leaf_begin_b = 0;
        if (dual_stereo && i == al.intensity) {
            dual_stereo = 0;
            br.flags |= 8;            // norm[] <- (norm + norm2)/2 before this band (bands.c:1551-1558)
        }
        br.eff_lowband = (int16_t)effective_lowband;
        const int has_lb = effective_lowband != -1;
        if (dual_stereo) {
            br.mode = OB_BAND_DUAL;
            x_cm = ob_decode_band(ctx, xoff, N, b / 2, B, has_lb, LM, 1.0f, (int)x_cm, &br.leaf_begin_a, &br.leaf_cnt_a);
            y_cm = ob_decode_band(ctx, yoff, N, b / 2, B, has_lb, LM, 1.0f, (int)y_cm, &br.leaf_begin_b, &br.leaf_cnt_b);
        } else if (C == 2) {
            // quant_band_stereo (bands.c:1235-1381)
            int fill = (int)(x_cm | y_cm);
            if (N == 1) {
                br.mode = OB_BAND_JOINT;
                br.leaf_begin_a = (uint16_t)ctx.n_leaves;
                for (int c = 0; c < 2; c++) {
                    int sign = 0;
                    if (ctx.remaining_bits >= 1 << OB_BITRES) { sign = (int)ec.bits(1); ctx.remaining_bits -= 1 << OB_BITRES; }
                    ob_emit_leaf(ctx, c ? yoff : xoff, 1, sign, OB_LEAF_ONE, 1, 1.0f);
                }
                br.leaf_cnt_a = 2;
                br.imid = 32767;     // mid=~1, no merge for N==1: K2 treats mode JOINT with n==1 specially
                x_cm = 1;
            } else {
                ObSplit s;
                int orig_fill = fill, mbits, sbits, bs = b;      // quant_band_stereo takes b by value (update_lowband below uses the original)
                ob_decode_theta(ctx, s, N, bs, B, B, LM, 1, fill);
                const float mid = (1.f / 32768) * s.imid, side = (1.f / 32768) * s.iside;
                (void)mid;
                br.imid = (int16_t)s.imid; br.iside = (int16_t)s.iside;
                if (s.inv) br.flags |= 1;
                if (N == 2) {
                    int sign = 0;
                    br.mode = OB_BAND_JOINT_N2;
                    mbits = bs; sbits = 0;
                    if (s.itheta != 0 && s.itheta != 16384) sbits = 1 << OB_BITRES;
                    mbits -= sbits;
                    const int c = s.itheta > 8192;
                    ctx.remaining_bits -= s.qalloc + sbits;
                    if (sbits) sign = (int)ec.bits(1);
                    if (c) br.flags |= 2;
                    if (sign) br.flags |= 4;
                    x_cm = ob_decode_band(ctx, c ? yoff : xoff, N, mbits, B, has_lb, LM, 1.0f, orig_fill, &br.leaf_begin_a, &br.leaf_cnt_a);
                } else {
                    br.mode = OB_BAND_JOINT;
                    mbits = ob_imax(0, ob_imin(bs, (bs - s.delta) / 2));
                    sbits = bs - mbits;
                    ctx.remaining_bits -= s.qalloc;
                    int32_t rebalance = ctx.remaining_bits;
                    if (mbits >= sbits) {
                        x_cm = ob_decode_band(ctx, xoff, N, mbits, B, has_lb, LM, 1.0f, fill, &br.leaf_begin_a, &br.leaf_cnt_a);
                        rebalance = mbits - (rebalance - ctx.remaining_bits);
                        if (rebalance > 3 << OB_BITRES && s.itheta != 0) sbits += rebalance - (3 << OB_BITRES);
                        x_cm |= ob_decode_band(ctx, yoff, N, sbits, B, 0, LM, side, fill >> B, &br.leaf_begin_b, &br.leaf_cnt_b);
                    } else {
                        x_cm = ob_decode_band(ctx, yoff, N, sbits, B, 0, LM, side, fill >> B, &br.leaf_begin_b, &br.leaf_cnt_b);
                        rebalance = sbits - (rebalance - ctx.remaining_bits);
                        if (rebalance > 3 << OB_BITRES && s.itheta != 16384) mbits += rebalance - (3 << OB_BITRES);
                        x_cm |= ob_decode_band(ctx, xoff, N, mbits, B, has_lb, LM, 1.0f, fill, &br.leaf_begin_a, &br.leaf_cnt_a);
                    }
                }
            }
            y_cm = x_cm;
        } else {
            br.mode = OB_BAND_MONO;
            x_cm = ob_decode_band(ctx, xoff, N, b, B, has_lb, LM, 1.0f, (int)(x_cm | y_cm), &br.leaf_begin_a, &br.leaf_cnt_a);
            y_cm = x_cm;
        }
        collapse_masks[i * C + 0] = (uint8_t)x_cm;
        collapse_masks[i * C + C - 1] = (uint8_t)y_cm;
        balance += al.pulses[i] + tell;
        update_lowband = b > (N << OB_BITRES);
    }
    ir->hdr.n_leaves = (uint16_t)ob_imin(ctx.n_leaves, 65535);
    ir->hdr.lcg_total = ctx.lcg;
    if (ctx.n_leaves > OB_MAX_LEAVES) ir->hdr.status = OB_INTERNAL_ERROR;
}

// ------------------------------------------------------------------------------------------------
// Frame driver: Opus TOC (opus/src/opus_decoder.c:733-741, opus.c:173-192) + the symbol side of
// celt_decode_with_ec (opus/celt/celt_decoder.c:1100-1290).
// pkt points at the TOC byte; len includes it.  max_frame is the PCM capacity per channel.
// ------------------------------------------------------------------------------------------------
// ---- framing: opus_packet_parse_impl (opus/src/opus.c:194-353, self_delimited = 0) ----------------------------------------------
OB_DEV int ob_parse_size(const uint8_t *data, int len, int *size)                   // opus.c:146-170
{
    if (len < 1) { *size = -1; return -1; }
    if (data[0] < 252) { *size = data[0]; return 1; }
    if (len < 2) { *size = -1; return -1; }
    *size = 4 * data[1] + data[0];
    return 2;
}
// data/len: the packet with its TOC.  Returns the frame count (1..48) or an OPUS_* error; sizes[i] = payload bytes of frame i,
// *first_off = offset of frame 0's payload from data (the frames follow each other; padding, if any, is behind the last one).
OB_DEV int ob_parse_packet(const uint8_t *data, int len, int16_t *sizes, int *first_off)
{
    if (len < 0) return OB_BAD_ARG;
    if (len == 0) return OB_INVALID_PACKET;
    const uint8_t *p = data;
    const int toc = *p++;
    const int framesize = (toc & 0x80) ? (48000 << ((toc >> 3) & 3)) / 400 : ((toc & 0x60) == 0x60 ? ((toc & 8) ? 960 : 480) : (((toc >> 3) & 3) == 3 ? 2880 : (48000 << ((toc >> 3) & 3)) / 100));
    int count, cbr = 0, last_size, bytes, sz;
    len--;
    last_size = len;
    switch (toc & 3) {
    case 0: count = 1; break;
    case 1:
        count = 2; cbr = 1;
        if (len & 1) return OB_INVALID_PACKET;
        last_size = len / 2;
        sizes[0] = (int16_t)last_size;
        break;
    case 2:
        count = 2;
        bytes = ob_parse_size(p, len, &sz);
        len -= bytes;
        if (sz < 0 || sz > len) return OB_INVALID_PACKET;
        sizes[0] = (int16_t)sz;
        p += bytes;
        last_size = len - sz;
        break;
    default: {
        if (len < 1) return OB_INVALID_PACKET;
        const int ch = *p++;
        count = ch & 0x3F;
        if (count <= 0 || framesize * count > 5760) return OB_INVALID_PACKET;
        len--;
        if (ch & 0x40) {                                         // padding
            int q;
            do {
                if (len <= 0) return OB_INVALID_PACKET;
                q = *p++;
                len--;
                const int tmp = q == 255 ? 254 : q;
                len -= tmp;
            } while (q == 255);
        }
        if (len < 0) return OB_INVALID_PACKET;
        cbr = !(ch & 0x80);
        if (!cbr) {                                              // VBR: every frame but the last carries its size
            last_size = len;
            for (int i = 0; i < count - 1; i++) {
                bytes = ob_parse_size(p, len, &sz);
                len -= bytes;
                if (sz < 0 || sz > len) return OB_INVALID_PACKET;
                sizes[i] = (int16_t)sz;
                p += bytes;
                last_size -= bytes + sz;
            }
            if (last_size < 0) return OB_INVALID_PACKET;
        } else {
            last_size = len / count;
            if (last_size * count != len) return OB_INVALID_PACKET;
            for (int i = 0; i < count - 1; i++) sizes[i] = (int16_t)last_size;
        }
    } }
    if (last_size > 1275) return OB_INVALID_PACKET;
    sizes[count - 1] = (int16_t)last_size;
    *first_off = (int)(p - data);
    return count;
}

// The packets of ONE stream for one call -> frame slots (the frame loop of opus_decode_native, opus_decoder.c:715-799).  frame_size:
// the caller's per-packet PCM slot.  Returns the number of slots written (<= cap); a packet that does not fit gets one error slot.
// ds = 48000 / output rate: frame_size and the slots' sample offsets count OUTPUT samples, concealment lengths 48 kHz samples.
// fec: opus_decode(..., decode_fec = 1): a packet that parses and is CELT-only carries no FEC, so the whole slot is concealed as if it
// were lost (opus_decoder.c:744-750).
OB_DEV int ob_frame_packets(const uint8_t *packets, const int32_t *offsets, const int32_t *lens, int F, int frame_size, ObSlot *slots, int cap, int ds = 1,
                            int fec = 0)
{
    int n = 0;
    int16_t sizes[48];
    for (int f = 0; f < F; f++) {
        const int len = lens[f];
        // every packet still to come keeps one slot in reserve (cap >= F), so each packet of the call gets at least its error slot and
        // its samples / range / status are always written: a multi-frame packet that would eat into the reserve is OB_BUFFER_TOO_SMALL
        const int avail = cap - n - (F - 1 - f);
        if (avail < 1) break;                                    // unreachable for cap >= F
        ObSlot one;
        one.off = 0; one.len = 0; one.toc = 0; one.flags = OB_SLOT_FIRST | OB_SLOT_LAST; one.pkt = (uint16_t)f; one.sample_off = 0; one.status = 0;
        if (len <= 0) {                                          // lost packet: conceal the whole slot (opus_decoder.c:684-688, :715-729)
            one.status = len < 0 || frame_size <= 0 || frame_size % (OB_SHORT / ds) != 0 ? OB_BAD_ARG : frame_size * ds;
            slots[n++] = one;
            continue;
        }
        const uint8_t *data = packets + offsets[f];
        const int toc = data[0];
        int first_off = 0;
        const int count = (toc & 0x80) ? ob_parse_packet(data, len, sizes, &first_off) : OB_UNIMPLEMENTED;     // SILK / hybrid: not on this path
        const int N = OB_SHORT << ((toc >> 3) & 3);
        if (fec && (count >= 0 || frame_size <= 0 || frame_size % (OB_SHORT / ds) != 0)) {
            one.status = frame_size <= 0 || frame_size % (OB_SHORT / ds) != 0 ? OB_BAD_ARG : frame_size * ds;       // (:683-684)
            slots[n++] = one;
            continue;
        }
        if (count < 0) one.status = count;
        else if (count * N > frame_size * ds) one.status = OB_BUFFER_TOO_SMALL;                                // opus_decoder.c:764-765
        else if (count > avail) one.status = OB_BUFFER_TOO_SMALL;                                              // decoder created with too few frame slots
        if (one.status < 0) { slots[n++] = one; continue; }
        uint32_t off = (uint32_t)offsets[f] + (uint32_t)first_off;
        for (int i = 0; i < count; i++) {
            ObSlot sl;
            sl.off = off; sl.len = sizes[i]; sl.toc = (uint8_t)toc; sl.pkt = (uint16_t)f; sl.sample_off = (uint16_t)(i * N / ds);
            sl.flags = (uint8_t)((i == 0 ? OB_SLOT_FIRST : 0) | (i == count - 1 ? OB_SLOT_LAST : 0));
            sl.status = sizes[i] <= 1 ? N : 0;                   // payloads of <= 1 byte are DTX: concealed for one frame (opus_decoder.c:284-290)
            slots[n++] = sl;
            off += (uint32_t)sizes[i];
        }
    }
    return n;
}

// One frame slot -> IR.  pay/paylen: the frame's payload (no TOC); toc: the packet's TOC; conceal > 0: nothing to decode, that
// many samples are to be concealed (lost packet, DTX frame).
OB_DEV_NOINLINE void ob_decode_symbols(const uint8_t *pay, int paylen, int toc, int conceal, int dec_channels, ObFrameIR *ir, int phase_inv_disabled = 0)
{
    ObFrameHdr &h = ir->hdr;
    h.final_range = 0; h.n_leaves = 0; h.flags = 0; h.lcg_total = 0; h.LM = 0; h.C = 1; h.end = 0;
    const int LM = (toc >> 3) & 3, M = 1 << LM, N = OB_SHORT << LM;
    if (conceal > 0) {                                           // toc == 0: the packet itself is lost; else a DTX frame of a packet that arrived
        h.status = conceal; h.flags = (uint8_t)(OB_F_LOST | (toc ? OB_F_DTX : 0)); h.LM = (uint8_t)LM;
        return;
    }
    int len = paylen;
    const int C = (toc & 4) ? 2 : 1;
    const int bw = (toc >> 5) & 3;
    const int end = bw == 0 ? 13 : bw == 1 ? 17 : bw == 2 ? 19 : 21;
    const int start = 0;
    h.LM = (uint8_t)LM; h.C = (uint8_t)C; h.end = (uint8_t)end; h.status = N;

    ObRangeDec ec;
    ec.init(pay, (uint32_t)len);
    int32_t total_bits = len * 8;
    int32_t tell = ec.tell();
    int silence, isTransient = 0, flags = 0;
    if (tell >= total_bits) silence = 1;
    else if (tell == 1) silence = ec.bit_logp(15);
    else silence = 0;
    if (silence) { tell = len * 8; ec.nbits_total += tell - ec.tell(); flags |= OB_F_SILENCE; }
    h.pf_pitch = 0; h.pf_tapset = 0; h.pf_qg = 0;
    if (start == 0 && tell + 16 <= total_bits) {
        if (ec.bit_logp(1)) {
            int octave = (int)ec.uint(6);
            h.pf_pitch = (uint16_t)((16 << octave) + (int)ec.bits((uint32_t)(4 + octave)) - 1);
            h.pf_qg = (uint8_t)ec.bits(3);
            if (ec.tell() + 2 <= total_bits) h.pf_tapset = (uint8_t)ec.icdf(OB_TAPSET_ICDF, 2);
            flags |= OB_F_POSTFILTER;
        }
        tell = ec.tell();
    }
    if (LM > 0 && tell + 3 <= total_bits) { isTransient = ec.bit_logp(3); tell = ec.tell(); }
    if (isTransient) flags |= OB_F_TRANSIENT;
    const int intra = tell + 3 <= total_bits ? ec.bit_logp(3) : 0;
    if (intra) flags |= OB_F_INTRA;

    // coarse energy indices (quant_bands.c:428-491): only the integer qi, the float recurrence runs in the synthesis kernel
    {
        const uint8_t *prob = OB_E_PROB_MODEL + (LM * 2 + intra) * 42;
        const int32_t budget = len * 8;
        for (int i = 0; i < 2 * OB_NB; i++) h.coarse_qi[i] = 0;
        for (int i = start; i < end; i++) for (int c = 0; c < C; c++) {
            int qi;
            tell = ec.tell();
            if (budget - tell >= 15) { int pi = 2 * ob_imin(i, 20); qi = ec.laplace((uint32_t)prob[pi] << 7, (int)prob[pi + 1] << 6); }
            else if (budget - tell >= 2) { const int t = ec.icdf(OB_TAPSET_ICDF, 2); qi = (t >> 1) ^ -(t & 1); }  // small_energy_icdf == {2,1,0}
            else if (budget - tell >= 1) qi = -ec.bit_logp(1);
            else qi = -1;
            h.coarse_qi[c * OB_NB + i] = (int16_t)qi;
        }
    }
    // tf_decode (celt_decoder.c:460-497)
    int8_t tf_res[OB_NB];
    {
        int curr = 0, tf_select = 0, tf_changed = 0, logp = isTransient ? 2 : 4;
        uint32_t budget = (uint32_t)len * 8, t = (uint32_t)ec.tell();
        const int tf_select_rsv = LM > 0 && t + logp + 1 <= budget;
        budget -= tf_select_rsv;
        for (int i = 0; i < OB_NB; i++) tf_res[i] = 0;
        for (int i = start; i < end; i++) {
            if (t + logp <= budget) { curr ^= ec.bit_logp((uint32_t)logp); t = (uint32_t)ec.tell(); tf_changed |= curr; }
            tf_res[i] = (int8_t)curr;
            logp = isTransient ? 4 : 5;
        }
        if (tf_select_rsv && OB_TF_SELECT[LM * 8 + 4 * isTransient + 0 + tf_changed] != OB_TF_SELECT[LM * 8 + 4 * isTransient + 2 + tf_changed])
            tf_select = ec.bit_logp(1);
        for (int i = start; i < end; i++) tf_res[i] = OB_TF_SELECT[LM * 8 + 4 * isTransient + 2 * tf_select + tf_res[i]];
        for (int i = 0; i < OB_NB; i++) h.tf_change[i] = tf_res[i];
    }
    tell = ec.tell();
    int spread = 2;
    if (tell + 4 <= total_bits) spread = ec.icdf(OB_SPREAD_ICDF, 5);
    h.spread = (uint8_t)spread;

    int16_t cap[OB_NB], offsets[OB_NB];
    for (int i = 0; i < OB_NB; i++) {                                             // init_caps (celt.c:272-281)
        int Nb = (OB_EBANDS[i + 1] - OB_EBANDS[i]) << LM;
        cap[i] = (int16_t)((OB_CACHE_CAPS[OB_NB * (2 * LM + C - 1) + i] + 64) * C * Nb >> 2);
        offsets[i] = 0;
    }
    {   // dynalloc (celt_decoder.c:1217-1246)
        int dynalloc_logp = 6;
        total_bits <<= OB_BITRES;
        tell = (int32_t)ec.tell_frac();
        for (int i = start; i < end; i++) {
            int width = C * (OB_EBANDS[i + 1] - OB_EBANDS[i]) << LM;
            int quanta = ob_imin(width << OB_BITRES, ob_imax(6 << OB_BITRES, width));
            int loop_logp = dynalloc_logp, boost = 0;
            while (tell + (loop_logp << OB_BITRES) < total_bits && boost < cap[i]) {
                int flag = ec.bit_logp((uint32_t)loop_logp);
                tell = (int32_t)ec.tell_frac();
                if (!flag) break;
                boost += quanta;
                total_bits -= quanta;
                loop_logp = 1;
            }
            offsets[i] = (int16_t)boost;
            if (boost > 0) dynalloc_logp = ob_imax(2, dynalloc_logp - 1);
        }
    }
    const int alloc_trim = tell + (6 << OB_BITRES) <= total_bits ? ec.icdf(OB_TRIM_ICDF, 7) : 5;
    int32_t bits = ((len * 8) << OB_BITRES) - (int32_t)ec.tell_frac() - 1;
    const int anti_collapse_rsv = isTransient && LM >= 2 && bits >= ((LM + 2) << OB_BITRES) ? (1 << OB_BITRES) : 0;
    bits -= anti_collapse_rsv;

    ObAlloc al;
    for (int i = 0; i < OB_NB; i++) { al.pulses[i] = 0; al.ebits[i] = 0; al.fine_priority[i] = 0; }
    ob_compute_allocation(ec, end, offsets, cap, alloc_trim, bits, C, LM, al);
    h.coded_bands = (uint8_t)al.coded_bands; h.intensity = (uint8_t)al.intensity; h.dual_stereo = (uint8_t)al.dual_stereo;
    for (int i = 0; i < OB_NB; i++) { h.pulses[i] = al.pulses[i]; h.fine_quant[i] = al.ebits[i]; }

    for (int i = 0; i < 2 * OB_NB; i++) { h.fine_q2[i] = 0; h.final_bit[i] = -1; }
    for (int i = start; i < end; i++) {                                           // unquant_fine_energy (quant_bands.c:493-514)
        if (al.ebits[i] == 0) continue;
        for (int c = 0; c < C; c++) h.fine_q2[c * OB_NB + i] = (uint8_t)ec.bits(al.ebits[i]);
    }

    uint8_t collapse_masks[2 * OB_NB];
    for (int i = 0; i < 2 * OB_NB; i++) collapse_masks[i] = 0;
    (void)M;
    ob_decode_all_bands(ec, ir, end, C, al, isTransient ? M : 0, spread, tf_res, len * (8 << OB_BITRES) - anti_collapse_rsv,
            LM, dec_channels == 1 || phase_inv_disabled, collapse_masks);      // st->disable_inv (celt_decoder.c:227, :1567)
    for (int i = 0; i < 2 * OB_NB; i++) h.collapse_masks[i] = collapse_masks[i];

    if (anti_collapse_rsv > 0 && ec.bits(1)) flags |= OB_F_ANTICOLLAPSE;
    {   // unquant_energy_finalise (quant_bands.c:516-542)
        int bits_left = len * 8 - ec.tell();
        for (int prio = 0; prio < 2; prio++) for (int i = start; i < end && bits_left >= C; i++) {
            if (al.ebits[i] >= 8 || al.fine_priority[i] != prio) continue;
            for (int c = 0; c < C; c++) { h.final_bit[c * OB_NB + i] = (int8_t)ec.bits(1); bits_left--; }
        }
    }
    h.flags = (uint8_t)flags;
    h.final_range = ec.rng;
    if (h.status > 0 && ec.tell() > 8 * len) h.status = OB_INTERNAL_ERROR;       // celt_decoder.c:1364
}
