/* opus_b200.h -- C ABI of libopus_b200.so: batched CELT-only Opus decoding (48 kHz) on NVIDIA B200.
 *
 * This is the drop-in boundary for the hot path of the Deniskore/opus-codec crate.  Each entry point names
 * the reference interface it replaces; the Rust-side binding a maintainer would add beside src/bindings.rs
 * is shown in INTEGRATION.md.  Plain pointers and sizes only; no variadic CTLs (the reference's
 * opus_decoder_ctl(st, request, ...) at src/bindings.rs:404-411 is awkward to bind for a batch).
 *
 * Error convention: identical to libopus (src/bindings.rs:103-109 <-> src/error.rs:36-62):
 *   0 OK, -1 BAD_ARG, -2 BUFFER_TOO_SMALL, -3 INTERNAL_ERROR, -4 INVALID_PACKET, -5 UNIMPLEMENTED,
 *   -6 INVALID_STATE, -7 ALLOC_FAIL.  Batch calls return the call-level status; per-stream results
 *   (samples per channel, or a negative code) are written to samples_out[].
 *
 * Scope of this version: output rates 48 / 24 / 16 / 12 / 8 kHz (the decoder runs at 48 kHz inside), CELT-only TOC (config >= 16), every packet code (0-3: several frames per packet,
 * CBR / VBR sizes, padding), fec = false.  Lost packets (len 0) and DTX frames (<= 1 byte) are concealed like the
 * reference does (celt_decode_lost).  A SILK / hybrid TOC gives that (stream, packet) the status OPUS_UNIMPLEMENTED (-5),
 * a malformed packet OPUS_INVALID_PACKET (-4), both leaving the stream's state untouched; other streams are unaffected.
 * There is no CPU fallback: every entry point that decodes fails with OPUS_INTERNAL_ERROR if no CUDA device is usable.
 */
#ifndef OPUS_B200_H
#define OPUS_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OB_ABI_VERSION 7

typedef struct ObDecoder ObDecoder;

/* Replaces n x opus_decoder_create(Fs, channels, &err) (src/bindings.rs:366-373; Decoder::new src/decoder.rs:35-63).
 * n_streams independent decoders with `channels` output channels each live on CUDA device `device`; fs is the OUTPUT sample
 * rate (48000, 24000, 16000, 12000 or 8000): frame_size arguments and sample counts are in samples at that rate.
 * max_frames: the largest number of CELT frames per stream any single call will hold (>= 1): one per code-0 packet, up to 48
 * for a code-3 packet.  A packet whose frames do not fit any more gets OPUS_BUFFER_TOO_SMALL.
 * Returns NULL and sets *error on failure. */
ObDecoder *ob_decoder_create(int32_t n_streams, int32_t fs, int32_t channels, int32_t device, int32_t max_frames, int32_t *error);

/* Replaces opus_decoder_destroy (src/bindings.rs:412-415; Drop for Decoder src/decoder.rs:411-417). */
void ob_decoder_destroy(ObDecoder *dec);

/* Replaces n x opus_decode_float(st, data, len, pcm, frame_size, 0) (src/bindings.rs:393-403; Decoder::decode_float
 * src/decoder.rs:134-182): one packet per stream.
 *   packets        concatenated packet bytes (host memory), TOC byte included
 *   offsets[s]     byte offset of stream s's packet inside `packets`
 *   lens[s]        its length in bytes; 0 = lost packet: frame_size samples are concealed (must be a multiple of 120),
 *                  like decode_float(&[], out, false) (src/decoder.rs:134-182 -> opus_decoder.c:715-729)
 *   pcm_out        host buffer, n_streams * frame_size * channels floats, interleaved per stream
 *   frame_size     capacity per channel of each stream's slot (120/240/480/960...; a packet longer than this
 *                  gives OPUS_BUFFER_TOO_SMALL for that stream, like src/decoder.rs:149-160)
 *   samples_out[s] samples per channel decoded for stream s, or a negative OPUS_* code */
int32_t ob_decode_float(ObDecoder *dec, const uint8_t *packets, const int32_t *offsets, const int32_t *lens,
                        float *pcm_out, int32_t frame_size, int32_t *samples_out);

/* n_frames consecutive packets per stream in one call (what a jitter buffer or an offline transcode hands over):
 * equivalent to n_frames calls of ob_decode_float.  offsets/lens/samples_out/ranges_out are [n_streams][n_frames],
 * pcm_out is [n_streams][n_frames][frame_size*channels].  ranges_out (optional) receives OPUS_GET_FINAL_RANGE
 * after every frame (src/decoder.rs:302-312). n_frames <= max_frames. */
int32_t ob_decode_float_multi(ObDecoder *dec, int32_t n_frames, const uint8_t *packets, const int32_t *offsets,
                              const int32_t *lens, float *pcm_out, int32_t frame_size, int32_t *samples_out,
                              uint32_t *ranges_out);

/* Replaces n x opus_decode(st, data, len, pcm, frame_size, 0) (src/bindings.rs:382-392; Decoder::decode src/decoder.rs:75-127):
 * int16 PCM.  As in the reference's float build the packet is decoded in float, run through opus_pcm_soft_clip (whose gain is
 * carried from packet to packet per stream) and rounded to nearest with saturation.  pcm_out: [n_streams][n_frames][frame_size*channels]. */
int32_t ob_decode(ObDecoder *dec, const uint8_t *packets, const int32_t *offsets, const int32_t *lens,
                  int16_t *pcm_out, int32_t frame_size, int32_t *samples_out);
int32_t ob_decode_multi(ObDecoder *dec, int32_t n_frames, const uint8_t *packets, const int32_t *offsets, const int32_t *lens,
                        int16_t *pcm_out, int32_t frame_size, int32_t *samples_out, uint32_t *ranges_out);

/* Pipelined form of ob_decode_float_multi for callers that keep two calls in flight (a media server draining jitter buffers):
 * returns once the work is enqueued; the outputs are valid after ob_decoder_wait.  While call n's PCM travels to the host,
 * call n+1's kernels already run.  At most two calls may be outstanding: before submitting call n+2 (or touching call n's output
 * or input buffers) call ob_decoder_wait(dec, 1), which returns when everything except the most recent call has completed;
 * ob_decoder_wait(dec, 0) drains all.  Host buffers should be pinned.  Results are identical to the blocking call. */
int32_t ob_decode_float_multi_async(ObDecoder *dec, int32_t n_frames, const uint8_t *packets, const int32_t *offsets,
                                    const int32_t *lens, float *pcm_out, int32_t frame_size, int32_t *samples_out,
                                    uint32_t *ranges_out);
int32_t ob_decode_multi_async(ObDecoder *dec, int32_t n_frames, const uint8_t *packets, const int32_t *offsets,
                              const int32_t *lens, int16_t *pcm_out, int32_t frame_size, int32_t *samples_out,
                              uint32_t *ranges_out);                     /* the int16 form: half the bytes over PCIe */
int32_t ob_decoder_wait(ObDecoder *dec, int32_t keep_in_flight);

/* Same, with every pointer a DEVICE pointer on the decoder's device (inputs already resident in HBM, outputs
 * left there); asynchronous on the decoder's stream unless sync != 0. */
int32_t ob_decode_float_device(ObDecoder *dec, int32_t n_frames, const uint8_t *d_packets, const int32_t *d_offsets,
                               const int32_t *d_lens, float *d_pcm_out, int32_t frame_size, int32_t *d_samples_out,
                               uint32_t *d_ranges_out, int32_t sync);

/* Replaces n x opus_decoder_ctl(st, OPUS_GET_FINAL_RANGE_REQUEST, &v) (Decoder::final_range src/decoder.rs:302-312):
 * out[s] = final range of the last packet decoded for stream s (0 before any packet). */
int32_t ob_decoder_final_range(ObDecoder *dec, uint32_t *out);

/* Replaces opus_decoder_ctl(st, OPUS_RESET_STATE) (Decoder::reset src/decoder.rs:241-255) for the streams listed in
 * idx[0..n) (idx == NULL: all streams). */
int32_t ob_decoder_reset(ObDecoder *dec, const int32_t *idx, int32_t n);

/* Replaces opus_decoder_ctl(st, OPUS_GET_LAST_PACKET_DURATION_REQUEST, &v) (Decoder::get_last_packet_duration src/decoder.rs:294-296). */
int32_t ob_decoder_last_packet_duration(ObDecoder *dec, int32_t *out);
/* OPUS_GET_PITCH (Decoder::get_pitch, src/decoder.rs): the post-filter period of the last decoded frame, 0 before the first. */
int32_t ob_decoder_get_pitch(ObDecoder *dec, int32_t *out);

/* Replaces opus_decoder_ctl(st, OPUS_SET_GAIN / OPUS_GET_GAIN) (Decoder::set_gain src/decoder.rs:318-320, gain :325-327):
 * Q8 dB in [-32768, 32767], one value for the whole batch, applied to every decoded and concealed sample. */
int32_t ob_decoder_set_gain(ObDecoder *dec, int32_t gain_q8);
int32_t ob_decoder_get_gain(ObDecoder *dec, int32_t *value);
/* Replaces OPUS_SET / GET_PHASE_INVERSION_DISABLED (Decoder::set_phase_inversion_disabled src/decoder.rs:341-346,
 * phase_inversion_disabled :333-335). */
int32_t ob_decoder_set_phase_inversion_disabled(ObDecoder *dec, int32_t disabled);
int32_t ob_decoder_get_phase_inversion_disabled(ObDecoder *dec, int32_t *value);
/* The decode_fec argument of opus_decode / opus_decode_float (Decoder::decode(.., fec), src/decoder.rs:75-182) for the calls that follow.
 * CELT-only packets carry no FEC: with the flag on, a packet that parses is concealed over the whole frame_size like a lost one
 * (opus/src/opus_decoder.c:744-750); frame_size must then be a multiple of 2.5 ms.  SILK / hybrid packets stay OB_UNIMPLEMENTED. */
int32_t ob_decoder_set_decode_fec(ObDecoder *dec, int32_t decode_fec);
int32_t ob_decoder_get_decode_fec(ObDecoder *dec, int32_t *value);

/* Introspection for benchmarks: number of streams / channels, device-event time in ms of the three kernels of
 * the most recent call (symbols, bands, synthesis), kernel launches issued so far. */
int32_t ob_decoder_streams(const ObDecoder *dec);
int32_t ob_decoder_channels(const ObDecoder *dec);
int32_t ob_decoder_sample_rate(const ObDecoder *dec);       /* OPUS_GET_SAMPLE_RATE (Decoder::get_sample_rate, src/decoder.rs) */
int32_t ob_decoder_kernel_ms(ObDecoder *dec, float ms[3]);
int64_t ob_decoder_launches(const ObDecoder *dec);
void *ob_decoder_cuda_stream(ObDecoder *dec);

/* ------------------------------------------------------------------------------------------------------------------
 * Encoder.  Replaces, for a batch, the reference's opus_encoder_create / opus_encode_float / opus_encoder_ctl /
 * opus_encoder_destroy (src/bindings.rs:293-346) as used by src/encoder.rs.
 * Scope of this version: Fs = 48000, OPUS_APPLICATION_RESTRICTED_LOWDELAY (2051; the application that forces
 * MODE_CELT_ONLY, opus/src/opus_encoder.c:1330-1332), frames of 120/240/480/960 samples.  The CELT encoder is complete
 * (pitch pre-filter, transient/tf/spread/dynalloc/trim/stereo decisions, two-pass energy, PVQ search, theta RDO, CBR/VBR/
 * CVBR), and so is the Opus-layer tonality analysis the reference runs at complexity >= 7 (opus/src/analysis.c: FFT, band
 * statistics, bandwidth detector, speech/music network): packets are bit-identical to the reference's pure-C build at
 * every complexity.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct ObEncoder ObEncoder;

/* n x opus_encoder_create(Fs, channels, application, &err) (src/bindings.rs:297-305; Encoder::new src/encoder.rs:40-73).
 * Defaults as opus_encoder_init: VBR on, constrained, bitrate AUTO, complexity 9.
 * application: 2051 (RESTRICTED_LOWDELAY), or 2049 (AUDIO) / 2048 (VOIP): these add the 4 ms delay compensation, the VOIP high-pass and
 * libopus' SILK / hybrid / CELT mode decision (opus_encoder.c:1333-1392); a frame that decision does not give to CELT returns
 * OB_UNIMPLEMENTED in lens_out for that stream (every packet before it is what libopus produces).  fs: 48000. */
ObEncoder *ob_encoder_create(int32_t n_streams, int32_t fs, int32_t channels, int32_t application, int32_t device,
                             int32_t max_frames, int32_t *error);
/* opus_encoder_destroy (src/bindings.rs:335-338). */
void ob_encoder_destroy(ObEncoder *enc);

/* n x opus_encode_float(st, pcm, frame_size, data, max_data_bytes) (src/bindings.rs:325-334; Encoder::encode_float
 * src/encoder.rs:215-247): one frame per stream.  pcm: host, [n_streams][frame_size*channels] interleaved floats in [-1,1];
 * out: host, [n_streams][max_bytes]; lens_out[s]: packet length (TOC included) or a negative OPUS_* code.
 * frame_size (samples per channel at 48 kHz): 120, 240, 480, 960, or 1920 / 2880 / 3840 / 4800 / 5760 (FrameSize::Ms40 ... 120 ms,
 * src/types.rs): the long ones are coded as 20 ms CELT frames and merged into one code-1/2/3 packet like
 * opus_encode_native's multi-frame branch (opus/src/opus_encoder.c:1649-1795, repacketizer.c:138-330). */
int32_t ob_encode_float(ObEncoder *enc, const float *pcm, int32_t frame_size, uint8_t *out, int32_t max_bytes, int32_t *lens_out);
/* n_frames consecutive frames per stream: pcm [n_streams][n_frames][frame_size*channels], out [n_streams][n_frames][max_bytes],
 * lens_out / ranges_out [n_streams][n_frames] (ranges_out optional: OPUS_GET_FINAL_RANGE after each frame). */
int32_t ob_encode_float_multi(ObEncoder *enc, int32_t n_frames, const float *pcm, int32_t frame_size, uint8_t *out,
                              int32_t max_bytes, int32_t *lens_out, uint32_t *ranges_out);
/* n x opus_encode(st, pcm, frame_size, data, max_data_bytes) (src/bindings.rs:313-324; Encoder::encode src/encoder.rs:80-127): int16 PCM in.
 * As in the reference's float build the samples are scaled by 1/32768 and encoded at min(16, lsb_depth) bits of depth. */
int32_t ob_encode(ObEncoder *enc, const int16_t *pcm, int32_t frame_size, uint8_t *out, int32_t max_bytes, int32_t *lens_out);
int32_t ob_encode_multi(ObEncoder *enc, int32_t n_frames, const int16_t *pcm, int32_t frame_size, uint8_t *out,
                        int32_t max_bytes, int32_t *lens_out, uint32_t *ranges_out);
/* Same with DEVICE pointers; asynchronous on the encoder's stream unless sync != 0. */
int32_t ob_encode_float_device(ObEncoder *enc, int32_t n_frames, const float *d_pcm, int32_t frame_size, uint8_t *d_out,
                               int32_t max_bytes, int32_t *d_lens_out, uint32_t *d_ranges_out, int32_t sync);

/* CTLs: opus_encoder_ctl(st, OPUS_SET_*_REQUEST, v) (src/bindings.rs:339-346) as wrapped by src/encoder.rs:545-652
 * (set_bitrate, set_complexity, set_vbr, set_vbr_constraint, set_max_bandwidth, set_bandwidth, set_force_channels,
 * set_packet_loss_perc, set_lsb_depth); one value for the whole batch.  bitrate: bits/s, -1000 = OPUS_AUTO, -1 = OPUS_BITRATE_MAX. */
int32_t ob_encoder_set_bitrate(ObEncoder *enc, int32_t bitrate);
/* Batch-only knob (no libopus counterpart): how the batch is laid onto the GPU.  OB_ENC_MAP_WARP: one warp per stream, cooperative stages --
 * the low-latency mapping (a frame step of <= ~2 400 streams takes ~5.5 ms), packets in the warp's summation order.  OB_ENC_MAP_THREAD: one
 * lane per stream, 32 streams per instruction -- the bulk mapping, packets bit-identical to the reference's C build.  OB_ENC_MAP_AUTO (default):
 * WARP below OB_ENC_MAP_CROSSOVER streams, THREAD from there up; ob_encoder_get_split reports how the last call was laid out (streams
 * [0, n_warp) one warp per stream, the rest one lane per stream).  A stream keeps its mapping for the life of the encoder.  Both are the same
 * encoder source; every packet of either decodes with the reference decoder. */
#define OB_ENC_MAP_AUTO 0
#define OB_ENC_MAP_WARP 1
#define OB_ENC_MAP_THREAD 2
#define OB_ENC_MAP_CROSSOVER 12288
int32_t ob_encoder_set_mapping(ObEncoder *enc, int32_t mapping);
int32_t ob_encoder_get_mapping(ObEncoder *enc, int32_t *value);
int32_t ob_encoder_get_split(ObEncoder *enc, int32_t *n_warp_streams);
int32_t ob_encoder_get_bitrate(ObEncoder *enc, int32_t *value);
int32_t ob_encoder_set_complexity(ObEncoder *enc, int32_t complexity);
int32_t ob_encoder_get_complexity(ObEncoder *enc, int32_t *value);
int32_t ob_encoder_set_vbr(ObEncoder *enc, int32_t vbr);
int32_t ob_encoder_get_vbr(ObEncoder *enc, int32_t *value);
int32_t ob_encoder_set_vbr_constraint(ObEncoder *enc, int32_t cvbr);
int32_t ob_encoder_get_vbr_constraint(ObEncoder *enc, int32_t *value);
int32_t ob_encoder_set_max_bandwidth(ObEncoder *enc, int32_t bandwidth);
int32_t ob_encoder_set_bandwidth(ObEncoder *enc, int32_t bandwidth);
int32_t ob_encoder_set_force_channels(ObEncoder *enc, int32_t channels);
int32_t ob_encoder_set_packet_loss_perc(ObEncoder *enc, int32_t percent);
int32_t ob_encoder_set_lsb_depth(ObEncoder *enc, int32_t depth);
int32_t ob_encoder_get_max_bandwidth(ObEncoder *enc, int32_t *value);
int32_t ob_encoder_get_force_channels(ObEncoder *enc, int32_t *value);
int32_t ob_encoder_get_packet_loss_perc(ObEncoder *enc, int32_t *value);
int32_t ob_encoder_get_lsb_depth(ObEncoder *enc, int32_t *value);
/* OPUS_SET/GET_SIGNAL (-1000 auto, 3001 voice, 3002 music; Encoder::set_signal src/encoder.rs), _PREDICTION_DISABLED, _PHASE_INVERSION_DISABLED,
 * _DTX (generalised DTX: after 200 ms without activity a packet is its TOC byte alone, opus_encoder.c:2364-2378; needs complexity >= 7 or
 * digital silence), _INBAND_FEC (0/1/2: a SILK feature -- on this path it only enters the mode decision of the AUDIO / VOIP applications),
 * _EXPERT_FRAME_DURATION (5000 = from the call's frame_size, 5001..5009 = 2.5..120 ms: must equal the call's frame_size, a shorter one is
 * OB_UNIMPLEMENTED), OPUS_GET_LOOKAHEAD (120, or 312 for AUDIO / VOIP), OPUS_GET_IN_DTX (one flag per stream). */
int32_t ob_encoder_set_signal(ObEncoder *enc, int32_t signal);
int32_t ob_encoder_get_signal(ObEncoder *enc, int32_t *value);
int32_t ob_encoder_set_prediction_disabled(ObEncoder *enc, int32_t disabled);
int32_t ob_encoder_get_prediction_disabled(ObEncoder *enc, int32_t *value);
int32_t ob_encoder_set_phase_inversion_disabled(ObEncoder *enc, int32_t disabled);
int32_t ob_encoder_get_phase_inversion_disabled(ObEncoder *enc, int32_t *value);
int32_t ob_encoder_set_dtx(ObEncoder *enc, int32_t enabled);
int32_t ob_encoder_get_dtx(ObEncoder *enc, int32_t *value);
int32_t ob_encoder_set_inband_fec(ObEncoder *enc, int32_t mode);
int32_t ob_encoder_get_inband_fec(ObEncoder *enc, int32_t *value);
int32_t ob_encoder_set_expert_frame_duration(ObEncoder *enc, int32_t duration);
int32_t ob_encoder_get_expert_frame_duration(ObEncoder *enc, int32_t *value);
int32_t ob_encoder_get_lookahead(ObEncoder *enc, int32_t *value);
int32_t ob_encoder_in_dtx(ObEncoder *enc, int32_t *out);
/* OPUS_GET_BANDWIDTH (Encoder::bandwidth, src/encoder.rs): the bandwidth (1101..1105) each stream's last packet was coded with; 1105 before the first. */
int32_t ob_encoder_get_bandwidth(ObEncoder *enc, int32_t *out);
/* OPUS_GET_FINAL_RANGE (Encoder::final_range src/encoder.rs:411-419) and OPUS_RESET_STATE (Encoder::reset :689-698). */
int32_t ob_encoder_final_range(ObEncoder *enc, uint32_t *out);
int32_t ob_encoder_reset(ObEncoder *enc, const int32_t *idx, int32_t n);
int32_t ob_encoder_streams(const ObEncoder *enc);
int32_t ob_encoder_channels(const ObEncoder *enc);
int32_t ob_encoder_sample_rate(const ObEncoder *enc);       /* OPUS_GET_SAMPLE_RATE (Encoder::sample_rate) */
int32_t ob_encoder_kernel_ms(ObEncoder *enc, float *ms);
int64_t ob_encoder_launches(const ObEncoder *enc);
void *ob_encoder_cuda_stream(ObEncoder *enc);

/* Static helpers mirroring src/packet.rs (packet_get_* on the TOC byte; opus/src/opus_decoder.c:1083-1129). */
int32_t ob_packet_get_nb_channels(const uint8_t *packet);
int32_t ob_packet_get_samples_per_frame(const uint8_t *packet, int32_t fs);
int32_t ob_packet_get_bandwidth(const uint8_t *packet);
int32_t ob_packet_get_nb_frames(const uint8_t *packet, int32_t len);
/* opus_packet_get_nb_samples / opus_packet_has_lbrr (packet_nb_samples / packet_has_lbrr, src/packet.rs:72-120; opus/src/opus_decoder.c:1119-1162). */
int32_t ob_packet_get_nb_samples(const uint8_t *packet, int32_t len, int32_t fs);
int32_t ob_packet_has_lbrr(const uint8_t *packet, int32_t len);

/* opus_packet_parse (src/bindings.rs; packet_parse src/packet.rs:162-215; opus/src/opus.c:194-360): returns the frame count (1..48) or an
 * OPUS_* code; sizes[i] / frame_offsets[i] (optional; offsets from `packet`) describe frame i; *payload_offset = offset of frame 0. */
int32_t ob_packet_parse(const uint8_t *packet, int32_t len, uint8_t *out_toc, int32_t *frame_offsets, int16_t *sizes, int32_t *payload_offset);
/* opus_packet_pad / opus_packet_unpad (packet_pad / packet_unpad, src/packet.rs:220-248; opus/src/repacketizer.c:283-353), in place.
 * pad: OB_OK or an error; unpad: the new length.  Padding extensions (opus/src/extensions.c) are kept by pad and dropped by unpad. */
int32_t ob_packet_pad(uint8_t *packet, int32_t len, int32_t new_len);
int32_t ob_packet_unpad(uint8_t *packet, int32_t len);
/* opus_multistream_packet_pad / _unpad (multistream_packet_pad / _unpad, src/packet.rs:253-290; opus/src/repacketizer.c:355-464): the same for a
 * multistream packet of nb_streams concatenated streams (all but the last in the self-delimited framing).  Byte work only: multistream
 * coding itself is out of scope. */
int32_t ob_multistream_packet_pad(uint8_t *packet, int32_t len, int32_t new_len, int32_t nb_streams);
int32_t ob_multistream_packet_unpad(uint8_t *packet, int32_t len, int32_t nb_streams);

/* n x opus_pcm_soft_clip(pcm, frame_size, channels, softclip_mem) (soft_clip, src/packet.rs:123-155; opus/src/opus.c:39-144) on the GPU, in
 * place: pcm host [n_streams][frame_size*channels], softclip_mem host [n_streams][channels] (zeros for a new stream). */
int32_t ob_pcm_soft_clip_batch(int32_t device, int32_t n_streams, float *pcm, int32_t frame_size, int32_t channels, float *softclip_mem);

/* The repacketizer object: opus_repacketizer_create / _destroy / _init / _cat / _get_nb_frames / _out_range / _out
 * (Repacketizer::new / drop / reset / push / frames / out_range / out, src/repacketizer.rs:18-100; opus/src/repacketizer.c:37-280).
 * As in libopus the packets given to _cat are referenced, not copied: they must stay valid until the last _out call. */
typedef struct ObRepacketizer ObRepacketizer;
ObRepacketizer *ob_repacketizer_create(void);
void ob_repacketizer_destroy(ObRepacketizer *rp);
void ob_repacketizer_init(ObRepacketizer *rp);
int32_t ob_repacketizer_cat(ObRepacketizer *rp, const uint8_t *packet, int32_t len);
int32_t ob_repacketizer_get_nb_frames(ObRepacketizer *rp);
int32_t ob_repacketizer_out_range(ObRepacketizer *rp, int32_t begin, int32_t end, uint8_t *out, int32_t maxlen);
int32_t ob_repacketizer_out(ObRepacketizer *rp, uint8_t *out, int32_t maxlen);
/* The same for a batch, on the GPU (a warp per output packet, payload bytes copied coalesced): every `group` consecutive packets of
 * each stream are merged as  init; cat x group; out  would.  packets / offsets / lens as for ob_decode_float_multi
 * ([n_streams][n_in]); out: [n_streams][ceil(n_in / group)][max_bytes]; lens_out: the packet length or the OPUS_* code of the first
 * failing cat / of out.  pad_to > 0: every output is padded to exactly pad_to bytes (opus_packet_pad), else left as short as possible.
 * Up to 128 padding extensions per output packet (OB_UNIMPLEMENTED beyond; the host object has no such limit).
 * The _device form takes device pointers and a cudaStream_t and does not synchronise. */
int32_t ob_repacketize_batch(int32_t device, int32_t n_streams, int32_t n_in, const uint8_t *packets, const int32_t *offsets, const int32_t *lens,
                             int32_t group, int32_t pad_to, uint8_t *out, int32_t max_bytes, int32_t *lens_out);
int32_t ob_repacketize_batch_device(int32_t n_streams, int32_t n_in, const uint8_t *d_packets, const int32_t *d_offsets, const int32_t *d_lens,
                                    int32_t group, int32_t pad_to, uint8_t *d_out, int32_t max_bytes, int32_t *d_lens_out, void *cuda_stream);

/* Test / debug access to the integer intermediate representation of the LAST decode call (csrc/ob_ir.h: coarse energy indices, tf, allocation,
 * fine bits, collapse masks, leaves, pulse vectors) -- what BASELINE's "decoded energy indices and pulse vectors must match exactly" is checked
 * on (tests/test_gpu_decode.py).  ob_debug_ir_layout: {sizeof(ObFrameIR), sizeof(ObFrameHdr), offsetof bands, leaves, iy, sizeof(ObLeaf),
 * sizeof(ObBand), OB_MAX_LEAVES}.  ob_decoder_debug_read_ir: copies the record of (stream, frame slot) to out (>= sizeof(ObFrameIR) bytes).
 * No libopus counterpart (the reference exposes these values only through the test taps of oracle/ref_shim.c). */
int32_t ob_debug_ir_layout(int32_t *out, int32_t n);
int32_t ob_decoder_debug_read_ir(ObDecoder *dec, int32_t stream, int32_t slot, void *out, int32_t nbytes);

/* "1.5.2-b200.<abi>" : bitstream compatibility level + ABI version (cf. version() src/lib.rs:52-54). */
const char *ob_version(void);
const char *ob_strerror(int32_t error);

#ifdef __cplusplus
}
#endif
#endif
