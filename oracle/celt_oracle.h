/* oracle/celt_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, scalar, one stream at a time) of the reference's CELT-only 48 kHz
 * decode path: Decoder::decode_float -> opus_decode_float -> celt_decode_with_ec.
 * Parity status: PINNED -- checked bit-for-bit (final range, energies, pulse spectrum) and to float
 * rounding (PCM) against the unmodified reference built in oracle/_ref, see tests/test_oracle.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this.  The product
 * library (opus_codec_b200/csrc) never links or calls it.
 */
#ifndef CELT_ORACLE_H
#define CELT_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CO_OK 0
#define CO_BAD_ARG (-1)
#define CO_BUFFER_TOO_SMALL (-2)
#define CO_INTERNAL_ERROR (-3)
#define CO_INVALID_PACKET (-4)
#define CO_UNIMPLEMENTED (-5)

/* What one decoded frame looked like inside (for stage-wise parity of the CUDA kernels). */
typedef struct {
    int LM, C, end, silence, transient, intra, spread, trim, coded_bands, intensity, dual_stereo;
    int pf_on, pf_pitch, pf_tapset, pf_qg, anti_collapse_on, anti_collapse_rsv;
    int total_bits_q3, balance;
    int tf_res[21], pulses[21], fine_quant[21], fine_priority[21], offsets[21];
    int coarse_qi[42];             /* Laplace-decoded coarse energy deltas [c*21+i] */
    uint8_t collapse_masks[42];    /* [i*C+c] as in the reference */
    uint32_t seed_in, seed_out, final_range;
    float X[2 * 960];              /* normalised spectrum after all bands (+anti-collapse NOT applied) */
    float bandLogE[42];            /* energies after finalise */
    float freq[2 * 960];           /* denormalised MDCT coefficients */
    float presyn[2 * (960 + 120)]; /* IMDCT output (incl. overlap tail) before the comb filter */
    int16_t iy[2 * 960];           /* decoded PVQ pulse vectors at their position in X (transformed domain of the leaf), where iy_set != 0 */
    uint8_t iy_set[2 * 960];
} co_tap_t;

typedef struct co_decoder co_decoder;

co_decoder *co_decoder_create(int channels);
void co_decoder_destroy(co_decoder *d);
void co_decoder_reset(co_decoder *d);
/* Mirrors opus_decode_float for CELT-only code-0 packets at Fs=48000: returns samples per channel
 * or a negative CO_* (== OPUS_*) code.  pcm is interleaved, frame_size is the capacity per channel. */
int co_decode_float(co_decoder *d, const uint8_t *pkt, int len, float *pcm, int frame_size);
uint32_t co_decoder_final_range(const co_decoder *d);
void co_decoder_set_tap(co_decoder *d, co_tap_t *tap);
int co_tap_size(void);

/* Whole-stream helper: pkts [nframes][stride]. */
int co_decode_stream(const uint8_t *pkts, const int *lens, int stride, int nframes, int frame_size, int channels,
                     float *pcm_out, uint32_t *ranges, int *samples, co_tap_t *taps);

/* Known-answer helpers exposed for tests (opus/celt/tests/test_unit_mathops.c:89-140, test_unit_cwrs32.c). */
int co_bitexact_cos(int x);
int co_bitexact_log2tan(int isin, int icos);
unsigned co_isqrt32(uint32_t v);
uint32_t co_pvq_v(int n, int k);
uint32_t co_cwrsi(int n, int k, uint32_t idx, int *y); /* returns sum y^2 */
uint32_t co_icwrs(int n, const int *y);
void co_mdct_backward(const float *in, float *out, int shift, int stride); /* out must hold N2+overlap, pre-zeroed overlap */
void co_fft(float *cpx, int shift);  /* in-place forward complex FFT, natural order in, like opus_fft_c without scaling */

#ifdef __cplusplus
}
#endif
#endif
