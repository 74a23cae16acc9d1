/* oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Thin helpers around the UNMODIFIED reference (libopus 1.5.2 compiled from /root/reference/opus by
 * oracle/Makefile into oracle/_ref/libopus_ref.so).  Nothing here re-implements codec arithmetic: every
 * sample and every packet byte comes out of the reference's own public entry points, the very ones the
 * crate's safe wrappers call:
 *     Decoder::decode_float -> opus_decode_float   (src/decoder.rs:162-175)
 *     Encoder::encode_float -> opus_encode_float   (src/encoder.rs:234-242)
 *     final_range()         -> OPUS_GET_FINAL_RANGE (src/decoder.rs:302-312, src/encoder.rs:411-419)
 *
 * Besides whole-stream encode/decode helpers it offers
 *   - "taps": link-time --wrap interposers on three internal libopus functions that record what
 *     passed through them (normalised spectrum after quant_all_bands, band energies + MDCT input at
 *     denormalise_bands, pre-post-filter time signal at comb_filter), so stage-wise parity of the CUDA
 *     kernels can be checked without touching reference source;
 *   - a pthread pool that runs one stream per thread for the CPU baseline.
 */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <pthread.h>
#include <time.h>
#include "opus.h"
#include "opus_private.h" /* OPUS_SET_FORCE_MODE, MODE_CELT_ONLY */
#include "arch.h"
#include "modes.h"
#include "entcode.h"

#define REF_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------
 * taps
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    /* quant_all_bands (opus/celt/bands.c:1398) */
    int   n_qab;                 /* calls seen */
    int   LM, C, shortBlocks, spread, dual_stereo, intensity, codedBands;
    int   total_bits, balance;
    int   pulses[21], tf_res[21];
    uint8_t collapse_masks[42];
    uint32_t seed_in, seed_out;
    float X[2 * 960];            /* normalised spectrum after the call: X then Y */
    /* denormalise_bands (opus/celt/bands.c:196), last call per channel */
    int   n_denorm;
    float bandLogE[42];
    float freq[2 * 960];
    /* comb_filter (opus/celt/celt.c:190): input of the FIRST call per channel = IMDCT output */
    int   n_comb;
    float presyn[2 * (960 + 120)];
    int   comb_T0[4], comb_T1[4], comb_N[4];
    float comb_g0[4], comb_g1[4];
} ref_tap_t;

static __thread ref_tap_t *g_tap = NULL;

REF_EXPORT int ref_tap_size(void) { return (int)sizeof(ref_tap_t); }
REF_EXPORT void ref_tap_set(void *p) { g_tap = (ref_tap_t *)p; if (g_tap) { g_tap->n_qab = g_tap->n_denorm = g_tap->n_comb = 0; } }

void __real_quant_all_bands(int encode, const CELTMode *m, int start, int end, celt_norm *X_, celt_norm *Y_,
        unsigned char *collapse_masks, const celt_ener *bandE, int *pulses, int shortBlocks, int spread,
        int dual_stereo, int intensity, int *tf_res, opus_int32 total_bits, opus_int32 balance, ec_ctx *ec,
        int LM, int codedBands, opus_uint32 *seed, int complexity, int arch, int disable_inv);

void __wrap_quant_all_bands(int encode, const CELTMode *m, int start, int end, celt_norm *X_, celt_norm *Y_,
        unsigned char *collapse_masks, const celt_ener *bandE, int *pulses, int shortBlocks, int spread,
        int dual_stereo, int intensity, int *tf_res, opus_int32 total_bits, opus_int32 balance, ec_ctx *ec,
        int LM, int codedBands, opus_uint32 *seed, int complexity, int arch, int disable_inv)
{
    ref_tap_t *t = g_tap;
    uint32_t seed_in = *seed;
    __real_quant_all_bands(encode, m, start, end, X_, Y_, collapse_masks, bandE, pulses, shortBlocks, spread,
            dual_stereo, intensity, tf_res, total_bits, balance, ec, LM, codedBands, seed, complexity, arch,
            disable_inv);
    if (t && !encode) {
        int N = 120 << LM, C = Y_ ? 2 : 1, i;
        t->n_qab++;
        t->LM = LM; t->C = C; t->shortBlocks = shortBlocks; t->spread = spread;
        t->dual_stereo = dual_stereo; t->intensity = intensity; t->codedBands = codedBands;
        t->total_bits = total_bits; t->balance = balance;
        for (i = 0; i < 21; i++) { t->pulses[i] = pulses[i]; t->tf_res[i] = tf_res[i]; }
        memcpy(t->collapse_masks, collapse_masks, (size_t)C * 21);
        t->seed_in = seed_in; t->seed_out = *seed;
        memcpy(t->X, X_, sizeof(float) * N);
        if (Y_) memcpy(t->X + N, Y_, sizeof(float) * N);
    }
}

void __real_denormalise_bands(const CELTMode *m, const celt_norm *X, celt_sig *freq, const opus_val16 *bandLogE,
        int start, int end, int M, int downsample, int silence);
void __wrap_denormalise_bands(const CELTMode *m, const celt_norm *X, celt_sig *freq, const opus_val16 *bandLogE,
        int start, int end, int M, int downsample, int silence)
{
    ref_tap_t *t = g_tap;
    __real_denormalise_bands(m, X, freq, bandLogE, start, end, M, downsample, silence);
    if (t) {
        int c = t->n_denorm & 1, N = 120 * M;
        memcpy(t->bandLogE + 21 * c, bandLogE, sizeof(float) * 21);
        memcpy(t->freq + N * c, freq, sizeof(float) * N);
        t->n_denorm++;
    }
}

void __real_comb_filter(opus_val32 *y, opus_val32 *x, int T0, int T1, int N, opus_val16 g0, opus_val16 g1,
        int tapset0, int tapset1, const opus_val16 *window, int overlap, int arch);
void __wrap_comb_filter(opus_val32 *y, opus_val32 *x, int T0, int T1, int N, opus_val16 g0, opus_val16 g1,
        int tapset0, int tapset1, const opus_val16 *window, int overlap, int arch)
{
    ref_tap_t *t = g_tap;
    if (t && t->n_comb < 4) {
        int k = t->n_comb;
        t->comb_T0[k] = T0; t->comb_T1[k] = T1; t->comb_N[k] = N; t->comb_g0[k] = g0; t->comb_g1[k] = g1;
    }
    if (t) {
        int per = t->LM ? 2 : 1, k = t->n_comb;
        if (k % per == 0 && k / per < 2)   /* first call of a channel: x[0..Nframe+overlap) is the raw IMDCT output */
            memcpy(t->presyn + (k / per) * (960 + 120), x, sizeof(float) * ((120 << t->LM) + 120));
        t->n_comb++;
    }
    __real_comb_filter(y, x, T0, T1, N, g0, g1, tapset0, tapset1, window, overlap, arch);
}

/* ------------------------------------------------------------------------------------------------
 * whole-stream helpers
 * ---------------------------------------------------------------------------------------------- */

/* Creates an encoder the way the benchmarks configure it:
 * OPUS_APPLICATION_RESTRICTED_LOWDELAY forces MODE_CELT_ONLY (opus/src/opus_encoder.c:1330-1332).
 * vbr: 0 = CBR, 1 = VBR, 2 = constrained VBR. */
/* Optional extra CTLs applied by make_encoder (0 / OPUS_AUTO = leave alone): OPUS_SET_BANDWIDTH
 * (1101..1105) and OPUS_SET_FORCE_CHANNELS.  Used to produce NB/WB/SWB and mono-in-stereo test packets. */
static int g_extra_bandwidth = 0, g_extra_force_channels = 0, g_force_celt = 1;
/* 1 (default): VOIP / AUDIO encoders get OPUS_SET_FORCE_MODE(MODE_CELT_ONLY) (how the CELT goldens were made); 0: the encoder's own
 * SILK / hybrid / CELT decision, i.e. what a user of the crate gets. */
REF_EXPORT void ref_set_encoder_force_celt(int on) { g_force_celt = on; }
REF_EXPORT void ref_set_encoder_extras(int bandwidth, int force_channels)
{
    g_extra_bandwidth = bandwidth; g_extra_force_channels = force_channels;
}

/* More optional encoder CTLs (all 0 / OPUS_AUTO = leave alone): OPUS_SET_SIGNAL (3001 voice / 3002 music), OPUS_SET_PREDICTION_DISABLED,
 * OPUS_SET_PHASE_INVERSION_DISABLED, OPUS_SET_DTX, OPUS_SET_INBAND_FEC, OPUS_SET_EXPERT_FRAME_DURATION (5000..5009), OPUS_SET_PACKET_LOSS_PERC. */
static int g_x_signal = 0, g_x_pred_disabled = 0, g_x_phase_inv_disabled = 0, g_x_dtx = 0, g_x_fec = 0, g_x_duration = 0, g_x_loss = 0;
REF_EXPORT void ref_set_encoder_extras2(int signal, int pred_disabled, int phase_inv_disabled, int dtx, int fec, int duration, int loss)
{
    g_x_signal = signal; g_x_pred_disabled = pred_disabled; g_x_phase_inv_disabled = phase_inv_disabled; g_x_dtx = dtx; g_x_fec = fec;
    g_x_duration = duration; g_x_loss = loss;
}
/* OPUS_GET_IN_DTX / OPUS_GET_LOOKAHEAD of the encoder after the last ref_encode_stream call. */
static int g_last_in_dtx = 0, g_last_lookahead = 0;
REF_EXPORT int ref_last_in_dtx(void) { return g_last_in_dtx; }
REF_EXPORT int ref_last_lookahead(void) { return g_last_lookahead; }

static OpusEncoder *make_encoder(int channels, int application, int bitrate, int vbr, int complexity)
{
    int err = 0;
    OpusEncoder *e = opus_encoder_create(48000, channels, application, &err);
    if (!e || err != OPUS_OK) return NULL;
    opus_encoder_ctl(e, OPUS_SET_BITRATE(bitrate));
    opus_encoder_ctl(e, OPUS_SET_COMPLEXITY(complexity));
    opus_encoder_ctl(e, OPUS_SET_VBR(vbr != 0));
    opus_encoder_ctl(e, OPUS_SET_VBR_CONSTRAINT(vbr == 2));
    if (application != OPUS_APPLICATION_RESTRICTED_LOWDELAY && g_force_celt)
        opus_encoder_ctl(e, OPUS_SET_FORCE_MODE(MODE_CELT_ONLY));
    if (g_extra_bandwidth) opus_encoder_ctl(e, OPUS_SET_BANDWIDTH(g_extra_bandwidth));
    if (g_extra_force_channels) opus_encoder_ctl(e, OPUS_SET_FORCE_CHANNELS(g_extra_force_channels));
    if (g_x_signal) opus_encoder_ctl(e, OPUS_SET_SIGNAL(g_x_signal));
    if (g_x_pred_disabled) opus_encoder_ctl(e, OPUS_SET_PREDICTION_DISABLED(1));
    if (g_x_phase_inv_disabled) opus_encoder_ctl(e, OPUS_SET_PHASE_INVERSION_DISABLED(1));
    if (g_x_dtx) opus_encoder_ctl(e, OPUS_SET_DTX(1));
    if (g_x_fec) opus_encoder_ctl(e, OPUS_SET_INBAND_FEC(g_x_fec));
    if (g_x_duration) opus_encoder_ctl(e, OPUS_SET_EXPERT_FRAME_DURATION(g_x_duration));
    if (g_x_loss) opus_encoder_ctl(e, OPUS_SET_PACKET_LOSS_PERC(g_x_loss));
    return e;
}

/* pcm: nframes*frame_size*channels interleaved floats. out: nframes slots of max_bytes each.
 * Returns 0 or a negative OPUS_* code. */
REF_EXPORT int ref_encode_stream(const float *pcm, int nframes, int frame_size, int channels, int application,
        int bitrate, int vbr, int complexity, unsigned char *out, int max_bytes, int *lens, uint32_t *ranges)
{
    int f;
    OpusEncoder *e = make_encoder(channels, application, bitrate, vbr, complexity);
    if (!e) return OPUS_ALLOC_FAIL;
    for (f = 0; f < nframes; f++) {
        opus_uint32 rng = 0;
        int n = opus_encode_float(e, pcm + (size_t)f * frame_size * channels, frame_size,
                out + (size_t)f * max_bytes, max_bytes);
        if (n < 0) { opus_encoder_destroy(e); return n; }
        lens[f] = n;
        opus_encoder_ctl(e, OPUS_GET_FINAL_RANGE(&rng));
        if (ranges) ranges[f] = rng;
    }
    { opus_int32 v = 0; opus_encoder_ctl(e, OPUS_GET_IN_DTX(&v)); g_last_in_dtx = v; opus_encoder_ctl(e, OPUS_GET_LOOKAHEAD(&v)); g_last_lookahead = v; }
    opus_encoder_destroy(e);
    return 0;
}

/* Same through the int16 API (opus_encode). */
REF_EXPORT int ref_encode_stream_i16(const opus_int16 *pcm, int nframes, int frame_size, int channels, int application,
        int bitrate, int vbr, int complexity, unsigned char *out, int max_bytes, int *lens, uint32_t *ranges)
{
    int f;
    OpusEncoder *e = make_encoder(channels, application, bitrate, vbr, complexity);
    if (!e) return OPUS_ALLOC_FAIL;
    for (f = 0; f < nframes; f++) {
        opus_uint32 rng = 0;
        int n = opus_encode(e, pcm + (size_t)f * frame_size * channels, frame_size, out + (size_t)f * max_bytes, max_bytes);
        if (n < 0) { opus_encoder_destroy(e); return n; }
        lens[f] = n;
        opus_encoder_ctl(e, OPUS_GET_FINAL_RANGE(&rng));
        if (ranges) ranges[f] = rng;
    }
    opus_encoder_destroy(e);
    return 0;
}

/* pkts: nframes slots of `stride` bytes; lens[f] bytes valid (0 => packet loss / PLC).
 * dec_channels: channel count of the decoder object.  taps (optional): nframes ref_tap_t records. */
/* Optional decoder CTLs applied by ref_decode_stream: OPUS_SET_GAIN (Q8 dB) and OPUS_SET_PHASE_INVERSION_DISABLED. */
static int g_dec_gain = 0, g_dec_phase_inv_disabled = 0, g_dec_fs = 48000;
/* Optional per-frame decode_fec flags for ref_decode_stream (NULL = all 0). */
static const int *g_dec_fec_flags = 0;
REF_EXPORT void ref_set_decoder_fec_flags(const int *flags) { g_dec_fec_flags = flags; }
/* Optional per-frame OPUS_GET_PITCH output of ref_decode_stream (NULL = off). */
static int *g_dec_pitch_out = 0;
REF_EXPORT void ref_set_decoder_pitch_out(int *buf) { g_dec_pitch_out = buf; }
REF_EXPORT void ref_set_decoder_fs(int fs) { g_dec_fs = fs; }
REF_EXPORT void ref_set_decoder_extras(int gain_q8, int phase_inv_disabled)
{
    g_dec_gain = gain_q8; g_dec_phase_inv_disabled = phase_inv_disabled;
}

REF_EXPORT int ref_decode_stream(const unsigned char *pkts, const int *lens, int stride, int nframes, int frame_size,
        int dec_channels, float *pcm_out, uint32_t *ranges, int *samples, void *taps)
{
    int f, err = 0;
    OpusDecoder *d = opus_decoder_create(g_dec_fs, dec_channels, &err);
    if (!d || err != OPUS_OK) return OPUS_ALLOC_FAIL;
    if (g_dec_gain) opus_decoder_ctl(d, OPUS_SET_GAIN(g_dec_gain));
    if (g_dec_phase_inv_disabled) opus_decoder_ctl(d, OPUS_SET_PHASE_INVERSION_DISABLED(1));
    for (f = 0; f < nframes; f++) {
        opus_uint32 rng = 0;
        int n;
        if (taps) ref_tap_set((char *)taps + (size_t)f * sizeof(ref_tap_t));
        n = opus_decode_float(d, lens[f] > 0 ? pkts + (size_t)f * stride : NULL, lens[f],
                pcm_out + (size_t)f * frame_size * dec_channels, frame_size, g_dec_fec_flags ? g_dec_fec_flags[f] : 0);
        if (taps) ref_tap_set(NULL);
        if (samples) samples[f] = n;
        if (g_dec_pitch_out) { opus_int32 v = 0; opus_decoder_ctl(d, OPUS_GET_PITCH(&v)); g_dec_pitch_out[f] = v; }
        if (n < 0 && !samples) { opus_decoder_destroy(d); return n; }
        opus_decoder_ctl(d, OPUS_GET_FINAL_RANGE(&rng));
        if (ranges) ranges[f] = rng;
    }
    opus_decoder_destroy(d);
    return 0;
}

/* Same through the int16 API (opus_decode: float decode + opus_pcm_soft_clip + FLOAT2INT16). */
REF_EXPORT int ref_decode_stream_i16(const unsigned char *pkts, const int *lens, int stride, int nframes, int frame_size,
        int dec_channels, opus_int16 *pcm_out, uint32_t *ranges, int *samples)
{
    int f, err = 0;
    OpusDecoder *d = opus_decoder_create(48000, dec_channels, &err);
    if (!d || err != OPUS_OK) return OPUS_ALLOC_FAIL;
    if (g_dec_gain) opus_decoder_ctl(d, OPUS_SET_GAIN(g_dec_gain));
    if (g_dec_phase_inv_disabled) opus_decoder_ctl(d, OPUS_SET_PHASE_INVERSION_DISABLED(1));
    for (f = 0; f < nframes; f++) {
        opus_uint32 rng = 0;
        int n = opus_decode(d, lens[f] > 0 ? pkts + (size_t)f * stride : NULL, lens[f],
                pcm_out + (size_t)f * frame_size * dec_channels, frame_size, 0);
        if (samples) samples[f] = n;
        opus_decoder_ctl(d, OPUS_GET_FINAL_RANGE(&rng));
        if (ranges) ranges[f] = rng;
    }
    opus_decoder_destroy(d);
    return 0;
}

/* Decode ONE packet with a fresh decoder (used for stateless final-range checks). */
REF_EXPORT int ref_decode_packet_fresh(const unsigned char *pkt, int len, int frame_size, int dec_channels,
        float *pcm_out, uint32_t *range)
{
    int lens[1]; lens[0] = len;
    return ref_decode_stream(pkt, lens, len, 1, frame_size, dec_channels, pcm_out, range, NULL, NULL);
}

/* ------------------------------------------------------------------------------------------------
 * CPU baseline: a pthread pool, one stream per thread at a time.
 * Streams are handed out round-robin; object creation is outside the timed region.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int tid, nthreads, nstreams, nframes, frame_size, channels;
    const unsigned char *pkts; const int *lens; int stride;   /* decode: [stream][frame][stride] */
    const float *pcm_in;                                      /* encode: [stream][frame][fs*ch]  */
    float *pcm_out; unsigned char *pkt_out; int *len_out;     /* optional outputs                */
    uint32_t *ranges;                                         /* [stream][frame], optional       */
    int application, bitrate, vbr, complexity;
    int encode;
    pthread_barrier_t *bar;
    double t0, t1;
    int err;
} pool_arg_t;

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

static void *pool_worker(void *vp)
{
    pool_arg_t *a = (pool_arg_t *)vp;
    int mine = 0, s, f, k, err = 0;
    void **objs;
    float *scratch = (float *)malloc(sizeof(float) * 960 * 2);
    unsigned char pkt[1500];
    for (s = a->tid; s < a->nstreams; s += a->nthreads) mine++;
    objs = (void **)calloc((size_t)(mine > 0 ? mine : 1), sizeof(void *));
    for (k = 0; k < mine; k++) {
        if (a->encode) objs[k] = make_encoder(a->channels, a->application, a->bitrate, a->vbr, a->complexity);
        else objs[k] = opus_decoder_create(48000, a->channels, &err);
        if (!objs[k]) a->err = OPUS_ALLOC_FAIL;
    }
    pthread_barrier_wait(a->bar);
    a->t0 = now_s();
    if (!a->err) for (k = 0, s = a->tid; s < a->nstreams; s += a->nthreads, k++) {
        for (f = 0; f < a->nframes; f++) {
            size_t sf = (size_t)s * a->nframes + f;
            opus_uint32 rng;
            int n;
            if (a->encode) {
                unsigned char *dst = a->pkt_out ? a->pkt_out + sf * a->stride : pkt;
                n = opus_encode_float((OpusEncoder *)objs[k], a->pcm_in + sf * a->frame_size * a->channels,
                        a->frame_size, dst, a->stride);
                if (a->len_out) a->len_out[sf] = n;
                if (a->ranges) { opus_encoder_ctl((OpusEncoder *)objs[k], OPUS_GET_FINAL_RANGE(&rng)); a->ranges[sf] = rng; }
            } else {
                float *dst = a->pcm_out ? a->pcm_out + sf * a->frame_size * a->channels : scratch;
                n = opus_decode_float((OpusDecoder *)objs[k], a->pkts + sf * a->stride, a->lens[sf], dst,
                        a->frame_size, 0);
                if (a->ranges) { opus_decoder_ctl((OpusDecoder *)objs[k], OPUS_GET_FINAL_RANGE(&rng)); a->ranges[sf] = rng; }
            }
            if (n < 0) a->err = n;
        }
    }
    a->t1 = now_s();
    pthread_barrier_wait(a->bar);
    for (k = 0; k < mine; k++) {
        if (!objs[k]) continue;
        if (a->encode) opus_encoder_destroy((OpusEncoder *)objs[k]); else opus_decoder_destroy((OpusDecoder *)objs[k]);
    }
    free(objs); free(scratch);
    return NULL;
}

static double run_pool(pool_arg_t *proto, int nthreads, int *err_out)
{
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    pool_arg_t *args = (pool_arg_t *)malloc(sizeof(pool_arg_t) * nthreads);
    pthread_barrier_t bar;
    double t0 = 1e300, t1 = 0;
    int i, err = 0;
    pthread_barrier_init(&bar, NULL, nthreads);
    for (i = 0; i < nthreads; i++) {
        args[i] = *proto; args[i].tid = i; args[i].nthreads = nthreads; args[i].bar = &bar; args[i].err = 0;
        pthread_create(&th[i], NULL, pool_worker, &args[i]);
    }
    for (i = 0; i < nthreads; i++) {
        pthread_join(th[i], NULL);
        if (args[i].t0 < t0) t0 = args[i].t0;
        if (args[i].t1 > t1) t1 = args[i].t1;
        if (args[i].err) err = args[i].err;
    }
    pthread_barrier_destroy(&bar);
    free(th); free(args);
    if (err_out) *err_out = err;
    return t1 - t0;
}

/* Returns wall seconds from the first thread's first frame to the last thread's last frame (<0 on error). */
REF_EXPORT double ref_decode_pool(const unsigned char *pkts, const int *lens, int stride, int nstreams, int nframes,
        int frame_size, int channels, int nthreads, float *pcm_out, uint32_t *ranges)
{
    pool_arg_t a; int err = 0; double t;
    memset(&a, 0, sizeof(a));
    a.nstreams = nstreams; a.nframes = nframes; a.frame_size = frame_size; a.channels = channels;
    a.pkts = pkts; a.lens = lens; a.stride = stride; a.pcm_out = pcm_out; a.ranges = ranges; a.encode = 0;
    t = run_pool(&a, nthreads, &err);
    return err ? (double)err : t;
}

REF_EXPORT double ref_encode_pool(const float *pcm, int nstreams, int nframes, int frame_size, int channels,
        int application, int bitrate, int vbr, int complexity, int nthreads,
        unsigned char *pkt_out, int stride, int *len_out, uint32_t *ranges)
{
    pool_arg_t a; int err = 0; double t;
    memset(&a, 0, sizeof(a));
    a.nstreams = nstreams; a.nframes = nframes; a.frame_size = frame_size; a.channels = channels;
    a.pcm_in = pcm; a.pkt_out = pkt_out; a.stride = stride > 0 ? stride : 1500; a.len_out = len_out; a.ranges = ranges;
    a.application = application; a.bitrate = bitrate; a.vbr = vbr; a.complexity = complexity; a.encode = 1;
    t = run_pool(&a, nthreads, &err);
    return err ? (double)err : t;
}

/* ------------------------------------------------------------------------------------------------
 * CELT-level encode: drives the reference's celt_encode_with_ec() directly (the CELT encoder object embedded in
 * OpusEncoder is configured the same way by opus_encoder.c:2085-2116, :2255-2300) WITHOUT the Opus-layer signal
 * analysis (analysis.valid == 0, SURVEY 2.2 "Signal analysis ... OUT OF SCOPE").  Packets carry no TOC byte.
 * nbytes: bytes per frame handed to the CELT encoder (CBR size, or the VBR ceiling).
 * ---------------------------------------------------------------------------------------------- */
#include "celt.h"
#include "entenc.h"
REF_EXPORT int ref_celt_encode_stream(const float *pcm, int nframes, int frame_size, int channels, int bitrate, int vbr,
        int complexity, int nbytes, unsigned char *out, int max_bytes, int *lens, uint32_t *ranges)
{
    int f;
    CELTEncoder *e = (CELTEncoder *)malloc((size_t)celt_encoder_get_size(channels));
    if (!e) return OPUS_ALLOC_FAIL;
    if (celt_encoder_init(e, 48000, channels, opus_select_arch()) != OPUS_OK) { free(e); return OPUS_INTERNAL_ERROR; }
    celt_encoder_ctl(e, CELT_SET_SIGNALLING(0));
    celt_encoder_ctl(e, OPUS_SET_COMPLEXITY(complexity));
    celt_encoder_ctl(e, OPUS_SET_LSB_DEPTH(24));
    celt_encoder_ctl(e, OPUS_SET_VBR(vbr != 0));
    celt_encoder_ctl(e, OPUS_SET_VBR_CONSTRAINT(vbr == 2));
    /* exactly what opus_encode_frame_native does: CBR leaves the CELT bitrate at MAX and sizes the frame through the
       byte budget (opus_encoder.c:2108, :2262-2273); the range coder is created by the caller (:1791, :2299) */
    celt_encoder_ctl(e, OPUS_SET_BITRATE(vbr ? bitrate : OPUS_BITRATE_MAX));
    if (nbytes > max_bytes) nbytes = max_bytes;
    for (f = 0; f < nframes; f++) {
        opus_uint32 rng = 0;
        ec_enc enc;
        int n;
        ec_enc_init(&enc, out + (size_t)f * max_bytes, (opus_uint32)nbytes);
        n = celt_encode_with_ec(e, pcm + (size_t)f * frame_size * channels, frame_size, NULL, nbytes, &enc);
        if (n < 0) { free(e); return n; }
        lens[f] = n;
        celt_encoder_ctl(e, OPUS_GET_FINAL_RANGE(&rng));
        if (ranges) ranges[f] = rng;
    }
    free(e);
    return 0;
}

/* CELT-level decode of TOC-less packets (celt_decode_with_ec), for round trips of the above. */
REF_EXPORT int ref_celt_decode_stream(const unsigned char *pkts, const int *lens, int stride, int nframes, int frame_size,
        int channels, float *pcm_out, uint32_t *ranges)
{
    int f;
    CELTDecoder *d = (CELTDecoder *)malloc((size_t)celt_decoder_get_size(channels));
    if (!d) return OPUS_ALLOC_FAIL;
    if (celt_decoder_init(d, 48000, channels) != OPUS_OK) { free(d); return OPUS_INTERNAL_ERROR; }
    celt_decoder_ctl(d, CELT_SET_SIGNALLING(0));
    for (f = 0; f < nframes; f++) {
        opus_uint32 rng = 0;
        int n = celt_decode_with_ec(d, pkts + (size_t)f * stride, lens[f], pcm_out + (size_t)f * frame_size * channels, frame_size, NULL, 0);
        if (n < 0) { free(d); return n; }
        celt_decoder_ctl(d, OPUS_GET_FINAL_RANGE(&rng));
        if (ranges) ranges[f] = rng;
    }
    free(d);
    return 0;
}

REF_EXPORT const char *ref_version(void) { return opus_get_version_string(); }
