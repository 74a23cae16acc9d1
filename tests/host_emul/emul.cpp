// tests/host_emul/emul.cpp -- TEST INFRASTRUCTURE.  Compiles the per-thread / per-warp DEVICE code of the
// product (opus_codec_b200/csrc/*.cuh) as plain C++ with one emulated lane, so its logic can be checked
// against the oracle on the CPU-only build box before spending GPU time.  Never shipped, never a fallback:
// the product library contains no host implementation of these functions.
#include <cstring>
#include "../../opus_codec_b200/csrc/dec_symbols.cuh"

extern "C" {
// the framing pass alone: status (0 decode, > 0 conceal, < 0 error) and packet index of every slot; returns the slot count
int emul_frame_packets(const uint8_t *pkts, const int *offs, const int *lens, int F, int frame_size, int cap, int *status, int *pkt)
{
    ObSlot *sl = new ObSlot[cap > 0 ? cap : 1];
    const int n = ob_frame_packets(pkts, offs, lens, F, frame_size, sl, cap);
    for (int i = 0; i < n; i++) { status[i] = sl[i].status; pkt[i] = sl[i].pkt; }
    delete[] sl;
    return n;
}
// mismatches between the rectangular PVQ table and the reference's triangular one over every pair the reference's table holds
int emul_pvq_table_check()
{
    int bad = 0, seen = 0;
    for (int a = 0; a < 15; a++) {
        const int end = a < 14 ? OB_PVQ_U_ROW[a + 1] + a + 1 : 1272;
        for (int b = a; b < 177 && OB_PVQ_U_ROW[a] + b < end; b++) { seen++; bad += ob_pvq_u(a, b) != ob_pvq_u_tri(a, b) || ob_pvq_u(b, a) != ob_pvq_u_tri(a, b); }
    }
    return seen == 1272 ? bad : -1;
}
int emul_ir_size() { return (int)sizeof(ObFrameIR); }
int emul_hdr_size() { return (int)sizeof(ObFrameHdr); }
void emul_decode_symbols(const uint8_t *pkt, int len, int dec_channels, int max_frame, ObFrameIR *ir)
{
    // one code-0 packet through the framing pass and the symbol pass
    memset(ir, 0, sizeof(*ir));
    ObSlot sl[48];
    const int32_t off = 0, ln = len;
    const int n = ob_frame_packets(pkt, &off, &ln, 1, max_frame, sl, 48);
    if (n < 1) { ir->hdr.status = OB_INTERNAL_ERROR; return; }
    if (sl[0].status < 0) { ir->hdr.status = sl[0].status; return; }
    ob_decode_symbols(pkt + sl[0].off, sl[0].len, sl[0].toc, sl[0].status, dec_channels, ir);
}
}

#include "../../opus_codec_b200/csrc/dec_bands.cuh"
#include "../../opus_codec_b200/csrc/dec_synth.cuh"
#include <cstdlib>

extern "C" {
// Whole-stream decode through the three device stages with ONE emulated lane each.
// Xtap (optional): nframes * 1920 floats, normalised spectrum after band reconstruction.
int emul_decode_stream(const uint8_t *pkts, const int *lens, int stride, int nframes, int max_frame, int dec_channels,
                       float *pcm_out, uint32_t *ranges, int *samples, float *Xtap)
{
    ObSynthShared *sh = (ObSynthShared *)calloc(1, sizeof(ObSynthShared));
    ObFrameIR *ir = (ObFrameIR *)calloc(1, sizeof(ObFrameIR));
    float *X = (float *)calloc(2 * OB_MAX_N, sizeof(float));
    ObBandsShared *bsh = (ObBandsShared *)calloc(1, sizeof(ObBandsShared));
    for (int i = 0; i < 2 * OB_NB; i++) sh->oldLogE[i] = sh->oldLogE2[i] = -28.f;
    ObSolo g;
    ob_synth_init(g, *sh);
    sh->ring = (float *)calloc(2 * OB_RING, sizeof(float));
    sh->decode_gain = 1.f; sh->ds = 1;
    ObPlanState plan = {0u, 0, 1, 0, OB_SHORT};                     // OPUS_RESET_STATE: skip_plc = 1 (celt_decoder.c:1527)
    uint32_t final_range = 0;
    // framing pass over the whole stream, then slot by slot through the symbol, plan, band and synthesis stages
    const int cap = nframes * 48;
    ObSlot *slots = (ObSlot *)calloc((size_t)cap, sizeof(ObSlot));
    int32_t *offs = (int32_t *)calloc((size_t)nframes, sizeof(int32_t));
    for (int f = 0; f < nframes; f++) offs[f] = f * stride;
    const int ns = ob_frame_packets(pkts, offs, lens, nframes, max_frame, slots, cap);
    int acc = 0;
    for (int j = 0; j < ns; j++) {
        const ObSlot sl = slots[j];
        memset(ir, 0, sizeof(*ir));
        if (sl.status < 0) ir->hdr.status = sl.status;
        else ob_decode_symbols(pkts + sl.off, sl.len, sl.toc, sl.status, dec_channels, ir);
        ob_plan_frame(plan, ir->hdr, dec_channels);
        int n = ir->hdr.status;
        if (n > 0) {
            if (!(ir->hdr.flags & OB_F_LOST)) {
                for (int k = 0; k < 2 * OB_MAX_N; k++) X[k] = __builtin_nanf("");     // bands >= end are never written
                ob_reconstruct_bands(g, ir, ir->hdr.seed_in, *bsh, X);
                if (Xtap && (sl.flags & OB_SLOT_FIRST)) memcpy(Xtap + (size_t)sl.pkt * 1920, X, sizeof(float) * ir->hdr.C * n);
                final_range = ir->hdr.final_range;
            } else if (ir->hdr.end_in != 0) final_range = 0;
            n = ob_synth_frame(g, *sh, ir, X, pcm_out + ((size_t)sl.pkt * max_frame + sl.sample_off) * dec_channels, dec_channels);
        }
        if (sl.flags & OB_SLOT_FIRST) acc = 0;
        if (n > 0) { if (acc >= 0) acc += n; } else acc = n;
        if (sl.flags & OB_SLOT_LAST) { samples[sl.pkt] = acc; ranges[sl.pkt] = final_range; }
    }
    free(slots); free(offs);
    free(sh->ring); free(sh); free(ir); free(X); free(bsh);
    return 0;
}
}

extern "C" {
// the int16 back end of the synthesis kernel: soft clip over one packet + rounding, one emulated lane
void emul_packet_to_int16(float *x, int16_t *out, int n, int channels, float *softclip_mem)
{
    ObSolo g;
    ob_packet_to_int16(g, x, out, n, channels, softclip_mem);
}
}

// ---- encoder -------------------------------------------------------------------------------------------------------------
// The warp-per-stream encoder source, compiled for ONE host lane in two summation orders (ob_coop.cuh): ObSolo = the reference's
// order (bit-identical to the reference's C build), ObSoloW = the order of the 32-lane warp (what the GPU must produce).
#include "../../opus_codec_b200/csrc/enc_frame.cuh"
static int g_warp_order = 0;
extern "C" void emul_set_warp_order(int on) { g_warp_order = on; }

struct EmulEnc {
    ObEncState st; ObEncHist hist; ObEncShared sh; ObEncWork wk;
};
static void emul_poison(EmulEnc *e)
{
    // no stage may depend on what an earlier frame left in the work area: everything but the state carried between frames
    if (!getenv("OB_EMUL_POISON")) return;
    memset(e->wk.in, 0xFF, sizeof(e->wk.in)); memset(e->wk.pre, 0xFF, sizeof(e->wk.pre)); memset(e->wk.freq, 0xFF, sizeof(e->wk.freq));
    memset(e->wk.X, 0xFF, sizeof(e->wk.X)); memset(e->wk.pcm_hp, 0xFF, sizeof(e->wk.pcm_hp)); memset(&e->wk.bw, 0xFF, sizeof(e->wk.bw));
    memset(e->wk.multi_tmp, 0xFF, sizeof(e->wk.multi_tmp)); memset(e->wk.env, 0xFF, sizeof(e->wk.env));
    memset(e->sh.A, 0xFF, sizeof(e->sh.A)); memset(&e->sh.u, 0xFF, sizeof(e->sh.u)); memset(e->sh.bandE, 0xFF, sizeof(e->sh.bandE));
    memset(e->sh.bandLogE, 0xFF, sizeof(e->sh.bandLogE)); memset(e->sh.bandLogE2, 0xFF, sizeof(e->sh.bandLogE2)); memset(e->sh.error, 0xFF, sizeof(e->sh.error));
    memset(e->sh.fine_quant, 0xFF, 8 * sizeof(e->sh.fine_quant)); memset(e->sh.bytes, 0xFF, sizeof(e->sh.bytes));
}

template <class G>
static int emul_celt_encode_T(const float *pcm, int nframes, int frame_size, int channels, int bitrate, int vbr, int complexity, int nbytes,
                              unsigned char *out, int max_bytes, int *lens, uint32_t *ranges)
{
    const G g;
    EmulEnc *e = (EmulEnc *)calloc(1, sizeof(EmulEnc));
    ObEncState *st = &e->st;
    st->channels = st->stream_channels = channels; st->complexity = complexity; st->vbr = vbr != 0; st->constrained_vbr = vbr == 2;
    st->bitrate = vbr ? bitrate : OB_BITRATE_MAX; st->lsb_depth = 24; st->end = 21; st->clip = 1; st->disable_inv = 0;
    ob_enc_reset(*st);
    ob_enc_hist_reset(e->hist);
    ob_enc_load_hist(g, e->sh, e->wk, e->hist);
    if (nbytes > max_bytes) nbytes = max_bytes;
    int rc = 0;
    for (int f = 0; f < nframes; f++) {
        emul_poison(e);
        ObRangeEnc enc;
        enc.init(out + (size_t)f * max_bytes, (uint32_t)nbytes);
        const int n = ob_celt_encode(g, *st, e->sh, e->wk, pcm + (size_t)f * frame_size * channels, frame_size, nbytes, enc);
        if (n < 0) { rc = n; break; }
        lens[f] = n;
        ranges[f] = st->rng;
    }
    free(e);
    return rc;
}
extern "C" {
// CELT-level encode of one stream through the product's per-stream device code (one emulated lane); mirrors
// ref_celt_encode_stream() in oracle/ref_shim.c (no TOC byte, coder created by the caller over `nbytes`).
int emul_celt_encode_stream(const float *pcm, int nframes, int frame_size, int channels, int bitrate, int vbr, int complexity, int nbytes,
                            unsigned char *out, int max_bytes, int *lens, uint32_t *ranges)
{
    return g_warp_order ? emul_celt_encode_T<ObSoloW>(pcm, nframes, frame_size, channels, bitrate, vbr, complexity, nbytes, out, max_bytes, lens, ranges)
                        : emul_celt_encode_T<ObSolo>(pcm, nframes, frame_size, channels, bitrate, vbr, complexity, nbytes, out, max_bytes, lens, ranges);
}
}
static int g_signal = 0, g_pred_disabled = 0, g_phase_inv_disabled = 0, g_dtx = 0, g_fec = 0, g_loss = 0, g_last_in_dtx = 0, g_lsb_depth = 24;
extern "C" void emul_set_lsb_depth(int d) { g_lsb_depth = d; }
template <class G>
static int emul_opus_encode_T(const float *pcm, int nframes, int frame_size, int channels, int application, int bitrate, int vbr, int complexity,
                              unsigned char *out, int max_bytes, int *lens, uint32_t *ranges)
{
    const G g;
    EmulEnc *e = (EmulEnc *)calloc(1, sizeof(EmulEnc));
    ObEncState *st = &e->st;
    ObOpusEncCfg cfg = {bitrate, complexity, vbr != 0, vbr == 2, 1105, 0, 0, g_loss, g_lsb_depth, application, g_signal, g_pred_disabled, g_phase_inv_disabled, g_dtx, g_fec, 0};
    ObOpusEncState *osp = (ObOpusEncState *)calloc(1, sizeof(ObOpusEncState));
    ObOpusEncState &os = *osp;
    os.stream_channels = channels; os.first = 1; os.auto_bandwidth = 0; os.bandwidth = 1105; os.hybrid_stereo_width_Q14 = 1 << 14; os.voice_ratio = -1;
    os.tonal = (ObTonalState *)calloc(1, sizeof(ObTonalState));
    os.delay = application != 2051 ? (float *)calloc(OB_ENC_BUFFER * channels, sizeof(float)) : nullptr;
    st->channels = st->stream_channels = channels; st->end = 21; st->clip = 1;
    ob_enc_reset(*st);
    ob_enc_hist_reset(e->hist);
    ob_enc_load_hist(g, e->sh, e->wk, e->hist);
    int rc = 0;
    for (int f = 0; f < nframes; f++) {
        emul_poison(e);
        const int n = ob_opus_encode(g, cfg, os, *st, e->sh, e->wk, pcm + (size_t)f * frame_size * channels, frame_size, out + (size_t)f * max_bytes, max_bytes);
        if (n < 0) { rc = n; break; }
        lens[f] = n;
        ranges[f] = st->final_range;
    }
    g_last_in_dtx = cfg.use_dtx && os.nb_no_activity_ms_Q1 >= 10 * 20 * 2;
    free(os.tonal); free(os.delay); free(e); free(osp);
    return rc;
}
extern "C" {
// Opus-level encode (TOC byte included) of one stream; mirrors ref_encode_stream().  application: 2048 / 2049 / 2051.
int emul_opus_encode_stream_app(const float *pcm, int nframes, int frame_size, int channels, int application, int bitrate, int vbr, int complexity,
                                unsigned char *out, int max_bytes, int *lens, uint32_t *ranges)
{
    return g_warp_order ? emul_opus_encode_T<ObSoloW>(pcm, nframes, frame_size, channels, application, bitrate, vbr, complexity, out, max_bytes, lens, ranges)
                        : emul_opus_encode_T<ObSolo>(pcm, nframes, frame_size, channels, application, bitrate, vbr, complexity, out, max_bytes, lens, ranges);
}
int emul_opus_encode_stream(const float *pcm, int nframes, int frame_size, int channels, int bitrate, int vbr, int complexity,
                            unsigned char *out, int max_bytes, int *lens, uint32_t *ranges)
{
    return emul_opus_encode_stream_app(pcm, nframes, frame_size, channels, 2051, bitrate, vbr, complexity, out, max_bytes, lens, ranges);
}
void emul_set_encoder_extras(int signal, int pred_disabled, int phase_inv_disabled, int dtx, int fec, int loss)
{
    g_signal = signal; g_pred_disabled = pred_disabled; g_phase_inv_disabled = phase_inv_disabled; g_dtx = dtx; g_fec = fec; g_loss = loss;
}
int emul_last_in_dtx(void) { return g_last_in_dtx; }
}

// ---- self-delimited framing (opus_packet_parse_impl / opus_repacketizer_out_range_impl with self_delimited = 1): the form multistream
// packets use for all but their last stream.  Not reachable through the public ABI; exercised here against the reference's internals. ----
extern "C" {
// Parses one (possibly self-delimited) packet at data.  sizes: 48 entries; offs: frame offsets from data.  Returns the frame count or an error.
int emul_parse_packet(const unsigned char *data, int len, int self_delimited, unsigned char *toc, int *offs, short *sizes, int *payload_offset,
                      int *packet_offset, int *padding_len)
{
    const uint8_t *frames[48];
    const uint8_t *padding = nullptr;
    int16_t sz[48];
    const int n = ob_rp_parse(data, len, self_delimited, toc, frames, sz, payload_offset, packet_offset, &padding, padding_len);
    for (int i = 0; i < n; i++) { sizes[i] = sz[i]; offs[i] = (int)(frames[i] - data); }
    return n;
}
// cat (plain framing) of n packets, then out_range in the self-delimited form.
int emul_repacketize_self_delimited(const unsigned char *packets, const int *offsets, const int *lens, int n, int begin, int end, unsigned char *out,
                                    int maxlen, int pad)
{
    ObRepack rp;
    ob_repack_init(&rp);
    for (int i = 0; i < n; i++) { const int r = ob_repack_cat(&rp, packets + offsets[i], lens[i], 0); if (r != OB_OK) return r; }
    const int ne = ob_repack_count_ext(&rp, begin < 0 ? 0 : begin, end > rp.nb_frames ? rp.nb_frames : end);
    ObExt *ext = (ObExt *)calloc((size_t)(ne > 0 ? ne : 1), sizeof(ObExt));
    const int r = ob_repack_out_range(ObRpLanes1(), &rp, begin, end, out, maxlen, 1, pad, ext, ne > 0 ? ne : 0);
    free(ext);
    return r;
}
}
