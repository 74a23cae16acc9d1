// enc_range.cuh -- range ENCODER (opus/celt/entenc.c:60-305) and Laplace encoder (opus/celt/laplace.c:51-92),
// one coder per thread.  The coder state is a plain struct so that the two-pass coarse-energy search
// (quant_bands.c:296-346) and the stereo theta RDO (bands.c:1583-1645) can snapshot and restore it by value.
#pragma once
#include "dec_symbols.cuh"

struct ObRangeEnc {
    uint8_t *buf;
    uint32_t storage, end_offs, end_window, offs, rng, val, ext;
    int nend_bits, nbits_total, rem, error;

    OB_MEM int write_byte(uint32_t v)                                                           // entenc.c:60-64
    {
        if (offs + end_offs >= storage) return -1;
        buf[offs++] = (uint8_t)v;
        return 0;
    }
    OB_MEM int write_byte_at_end(uint32_t v)                                                    // entenc.c:66-70
    {
        if (offs + end_offs >= storage) return -1;
        buf[storage - ++end_offs] = (uint8_t)v;
        return 0;
    }
    OB_MEM void carry_out(int c)                                                                // entenc.c:85-105
    {
        if (c != 255) {
            const int carry = c >> 8;
            if (rem >= 0) error |= write_byte((uint32_t)(rem + carry));
            if (ext > 0) {
                const uint32_t sym = (255u + (uint32_t)carry) & 255u;
                do error |= write_byte(sym); while (--ext > 0);
            }
            rem = c & 255;
        } else ext++;
    }
    OB_MEM void normalize()                                                                     // entenc.c:107-116
    {
        while (rng <= 0x800000u) {
            carry_out((int)(val >> 23));
            val = (val << 8) & 0x7FFFFFFFu;
            rng <<= 8;
            nbits_total += 8;
        }
    }
    OB_MEM void init(uint8_t *b, uint32_t size)                                                 // entenc.c:118-131
    {
        buf = b; end_offs = 0; end_window = 0; nend_bits = 0; nbits_total = 33; offs = 0;
        rng = 0x80000000u; rem = -1; val = 0; ext = 0; storage = size; error = 0;
    }
    OB_MEM int tell() const { return nbits_total - ob_ilog(rng); }
    OB_MEM uint32_t tell_frac() const
    {
        int l = ob_ilog(rng);
        uint32_t r = rng >> (l - 16);
        uint32_t b = (r >> 12) - 8;
        const uint32_t corr = b == 0 ? 35733u : b == 1 ? 38967u : b == 2 ? 42495u : b == 3 ? 46340u
                            : b == 4 ? 50535u : b == 5 ? 55109u : b == 6 ? 60097u : 65535u;
        b += r > corr;
        return ((uint32_t)nbits_total << OB_BITRES) - (uint32_t)((l << 3) + (int)b);
    }
    OB_MEM void encode(uint32_t fl, uint32_t fh, uint32_t ft)                                   // entenc.c:133-142
    {
        const uint32_t r = rng / ft;
        if (fl > 0) { val += rng - r * (ft - fl); rng = r * (fh - fl); }
        else rng -= r * (ft - fh);
        normalize();
    }
    OB_MEM void encode_bin(uint32_t fl, uint32_t fh, uint32_t bits)                             // entenc.c:144-153
    {
        const uint32_t r = rng >> bits;
        if (fl > 0) { val += rng - r * ((1u << bits) - fl); rng = r * (fh - fl); }
        else rng -= r * ((1u << bits) - fh);
        normalize();
    }
    OB_MEM void bit_logp(int v, uint32_t logp)                                                  // entenc.c:156-168
    {
        uint32_t r = rng;
        const uint32_t l = val, s = r >> logp;
        r -= s;
        if (v) val = l + r;
        rng = v ? s : r;
        normalize();
    }
    OB_MEM void icdf(int s, const uint8_t *tab, uint32_t ftb)                                   // entenc.c:170-179
    {
        const uint32_t r = rng >> ftb;
        if (s > 0) { val += rng - r * tab[s - 1]; rng = r * (uint32_t)(tab[s - 1] - tab[s]); }
        else rng -= r * tab[s];
        normalize();
    }
    OB_MEM void bits(uint32_t fl, uint32_t n)                                                   // entenc.c:211-231
    {
        uint32_t window = end_window;
        int used = nend_bits;
        if (used + (int)n > 32) {
            do { error |= write_byte_at_end(window & 255u); window >>= 8; used -= 8; } while (used >= 8);
        }
        window |= fl << used;
        used += (int)n;
        end_window = window; nend_bits = used; nbits_total += (int)n;
    }
    OB_MEM void uint(uint32_t fl, uint32_t ft)                                                  // entenc.c:192-209
    {
        ft--;
        int ftb = ob_ilog(ft);
        if (ftb > 8) {
            ftb -= 8;
            const uint32_t f = (ft >> ftb) + 1, l = fl >> ftb;
            encode(l, l + 1, f);
            bits(fl & ((1u << ftb) - 1u), (uint32_t)ftb);
        } else encode(fl, fl + 1, ft + 1);
    }
    OB_MEM void shrink(uint32_t size)                                                           // entenc.c:254-260
    {
        // OPUS_MOVE(buf+size-end_offs, buf+storage-end_offs, end_offs): moving towards lower addresses
        for (uint32_t k = 0; k < end_offs; k++) buf[size - end_offs + k] = buf[storage - end_offs + k];
        storage = size;
    }
    OB_MEM void done()                                                                          // entenc.c:262-305
    {
        int l = 32 - ob_ilog(rng);
        uint32_t msk = 0x7FFFFFFFu >> l;
        uint32_t end = (val + msk) & ~msk;
        if ((end | msk) >= val + rng) {
            l++;
            msk >>= 1;
            end = (val + msk) & ~msk;
        }
        while (l > 0) {
            carry_out((int)(end >> 23));
            end = (end << 8) & 0x7FFFFFFFu;
            l -= 8;
        }
        if (rem >= 0 || ext > 0) carry_out(0);
        uint32_t window = end_window;
        int used = nend_bits;
        while (used >= 8) { error |= write_byte_at_end(window & 255u); window >>= 8; used -= 8; }
        if (!error) {
            for (uint32_t k = offs; k < storage - end_offs; k++) buf[k] = 0;
            if (used > 0) {
                if (end_offs >= storage) error = -1;
                else {
                    l = -l;
                    if (offs + end_offs >= storage && l < used) { window &= (1u << l) - 1; error = -1; }
                    buf[storage - end_offs - 1] |= (uint8_t)window;
                }
            }
        }
    }
    OB_MEM void laplace(int *value, uint32_t fs, int decay)                                     // laplace.c:51-92
    {
        uint32_t fl = 0;
        int v = *value;
        if (v) {
            const int s = -(v < 0);
            int i;
            v = (v + s) ^ s;
            fl = fs;
            fs = (uint32_t)(((int32_t)(32768 - 32 - fs) * (int32_t)(16384 - decay)) >> 15);
            for (i = 1; fs > 0 && i < v; i++) {
                fs *= 2;
                fl += fs + 2;
                fs = (uint32_t)(((int32_t)fs * (int32_t)decay) >> 15);
            }
            if (!fs) {
                int ndi_max = (int)(32768 - fl + 1 - 1) >> 0;
                ndi_max = (ndi_max - s) >> 1;
                const int di = ob_imin(v - i, ndi_max - 1);
                fl += (uint32_t)((2 * di + 1 + s) * 1);
                fs = ob_imin(1, (int)(32768 - fl));
                *value = (i + di + s) ^ s;
            } else {
                fs += 1;
                fl += fs & ~(uint32_t)s;
            }
        }
        encode_bin(fl, fl + fs, 15);
    }
};
