// opus_b200.cu -- kernels + C ABI (include/opus_b200.h) of the batched CELT-only Opus decoder for sm_100a.
//
// Pipeline per call (S streams x F consecutive frames):
//   ob_k_symbols  one THREAD per (stream, frame): range decoder + all integer decisions -> ObFrameIR   (dec_symbols.cuh)
//   ob_k_bands    one WARP   per (stream, frame): IR -> normalised spectrum X                          (dec_bands.cuh)
//   ob_k_synth    one BLOCK  per stream, frames in order: energies, anti-collapse, denormalise, IMDCT,
//                 post-filter, de-emphasis; per-stream float state stays in shared memory across frames  (dec_synth.cuh)
// Streams are independent: multi-GPU use is one ObDecoder per device over disjoint stream ranges, no collective.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>
#include <new>

#include "../../include/opus_b200.h"
#include "dec_symbols.cuh"
#include "dec_bands.cuh"
#include "dec_synth.cuh"

#define OB_X_STRIDE (2 * OB_MAX_N)
#define OB_HIST_LEN (OB_HISTK + OB_OVERLAP)

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
// Framing pass, one thread per stream: the stream's F packets -> frame slots (code-0 packets: one slot each; codes 1-3: one per
// coded frame; lost packets and DTX frames: concealment slots; anything off this path: an error slot).
__global__ void ob_k_frame(const uint8_t *__restrict__ packets, const int32_t *__restrict__ offsets, const int32_t *__restrict__ lens,
                           ObSlot *__restrict__ slots, int32_t *__restrict__ nslots, int S, int F, int frame_size, int cap, int32_t *__restrict__ multi, int ds, int fec)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const int n = ob_frame_packets(packets, offsets + (size_t)s * F, lens + (size_t)s * F, F, frame_size, slots + (size_t)s * cap, cap, ds, fec);
    nslots[s] = n;
    // slot j <-> packet j unless some packet holds several frames (or did not fit): the host pipelines the call in slot windows
    // only in the one-to-one case
    int one_to_one = n == F;
    for (int j = 0; j < n && one_to_one; j++) one_to_one = slots[(size_t)s * cap + j].pkt == j;
    if (!one_to_one) *(volatile int32_t *)multi = 1;             // `multi` is mapped host memory: no copy on the stream to read it back
}

#ifndef OB_SYM_THREADS
#define OB_SYM_THREADS 128
#endif
// Resident blocks per SM = the register budget of the symbol kernel.  Measured after the round-2 changes (ms per 819 200 mono frames / per 163 840 stereo
// frames): 6 blocks (80 registers) 17.77 / 7.56, 7 (72) 17.06 / 8.07, 8 (64) 16.74 / 8.39 -- mono packets want the occupancy, stereo packets (more state
// per thread, more divergence) the registers: one instantiation per decoder channel count.
template <int BLOCKS>
__global__ void __launch_bounds__(OB_SYM_THREADS, BLOCKS)
ob_k_symbols(const uint8_t *__restrict__ packets, const ObSlot *__restrict__ slots, const int32_t *__restrict__ nslots,
             ObFrameIR *__restrict__ ir, int total, int dec_channels, int cap, int f0, int Fc, int phase_inv_disabled)
{
    // a launch covers the slot window [f0, f0+Fc) of every stream: k -> (stream, slot)
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    const int s = k / Fc, j = f0 + k % Fc;
    if (j >= nslots[s]) return;
    const size_t t = (size_t)s * cap + j;
    const ObSlot sl = slots[t];
    if (sl.status < 0) { ir[t].hdr.status = sl.status; ir[t].hdr.flags = 0; return; }
    ob_decode_symbols(packets + sl.off, sl.len, sl.toc, sl.status, dec_channels, ir + t, phase_inv_disabled);
}

// Plan pass: walks a stream's slot window in order through the integer loss state machine (which frames are concealed, noise- or
// pitch-based, how far the noise seed advances) and stamps every frame header with the state it starts from.  This is what lets
// the band kernel stay frame-parallel even though a lost packet changes the seed of the frames after it.
__global__ void ob_k_plan(ObFrameIR *__restrict__ ir, ObDecState *__restrict__ st, const int32_t *__restrict__ nslots, int S, int cap, int f0, int Fc, int CC)
{
    // One WARP per stream: the lanes fetch 32 headers at once (one 32-byte sector each, 12.6 KB apart), every lane then replays the
    // serial state machine from registers and keeps the state in front of its own frame -- the memory latency is paid once per
    // 32 frames instead of once per frame.
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (s >= S) return;
    ObPlanState p;
    p.rng = st[s].rng; p.loss_duration = st[s].loss_duration; p.skip_plc = st[s].skip_plc; p.plc_end = st[s].plc_end; p.last_fs = st[s].last_frame_size;
    const int hi = min(f0 + Fc, nslots[s]);
    for (int base = f0; base < hi; base += 32) {
        const int f = base + lane, cnt = min(32, hi - base);
        ObFrameHdr *h = f < hi ? &ir[(size_t)s * cap + f].hdr : nullptr;
        const int status = h ? h->status : 0, flags = h ? h->flags : 0, end = h ? h->end : 0, LM = h ? h->LM : 0;
        const uint32_t fr = h ? h->final_range : 0u;
        ObPlanState mine = p;
        for (int j = 0; j < cnt; j++) {
            if (lane == j) mine = p;
            ob_plan_step(p, __shfl_sync(0xffffffffu, status, j), __shfl_sync(0xffffffffu, flags, j), __shfl_sync(0xffffffffu, fr, j),
                         __shfl_sync(0xffffffffu, end, j), CC, __shfl_sync(0xffffffffu, LM, j));
        }
        if (h && status > 0) {
            h->seed_in = mine.rng; h->loss_in = mine.loss_duration; h->skip_in = (uint8_t)mine.skip_plc; h->end_in = (uint8_t)mine.plc_end;
            h->lastfs_in = (uint16_t)mine.last_fs;
        }
    }
    if (lane == 0) { st[s].rng = p.rng; st[s].loss_duration = p.loss_duration; st[s].skip_plc = p.skip_plc; st[s].plc_end = p.plc_end; st[s].last_frame_size = p.last_fs; }
}

#define OB_BANDS_WARPS 6
#define OB_BANDS_WARPS_MONO 4      // measured, 204 800 mono frames: 2 warps / block 5.99 ms, 3: 6.10, 4: 6.08, 5: 6.23, 6: 6.24 (stereo-sized: 7.64)
#define OB_BANDS_SMEM_PER_WARP ((int)sizeof(ObBandsShared))
#define OB_BANDS_SMEM_PER_WARP_MONO ((int)sizeof(ObBandsSharedT<1>))
// CH = 2: any frame.  CH = 1 (decoders created with one channel): shared memory for mono frames only -- more resident warps; a stereo frame
// (legal: a mono decoder down-mixes it) is put on the straggler list instead and reconstructed by ob_k_bands_stragglers right after.
#ifndef OB_BANDS_BLOCKS_MONO
#define OB_BANDS_BLOCKS_MONO 8      // 64 registers; 9 blocks (56 registers) and 10 (48) spill and measured 3 % slower
#endif
template <int CH>
__global__ void __launch_bounds__(CH == 1 ? OB_BANDS_WARPS_MONO * 32 : OB_BANDS_WARPS * 32, CH == 1 ? OB_BANDS_BLOCKS_MONO : 0)      // 0 = no minimum for the stereo-sized variant, as before
ob_k_bands(const ObFrameIR *__restrict__ ir, const int32_t *__restrict__ nslots, float *__restrict__ Xg, int S, int cap, int f0, int Fc,
           int32_t *__restrict__ strag_list, int32_t *__restrict__ strag_count)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5;
    const int k = blockIdx.x * (blockDim.x >> 5) + warp;
    if (k >= S * Fc) return;
    const int s = k / Fc, f = f0 + k % Fc;
    if (f >= nslots[s]) return;
    const size_t w = (size_t)s * cap + f;
    const ObFrameIR *fr = ir + w;
    if (fr->hdr.status <= 0 || (fr->hdr.flags & OB_F_LOST)) return;
    if (CH == 1 && fr->hdr.C == 2) {
        if ((threadIdx.x & 31) == 0) strag_list[atomicAdd(strag_count, 1)] = (int32_t)w;
        return;
    }
    const uint32_t seed = fr->hdr.seed_in;       // st->rng before this frame, stamped by the plan pass
    ObBandsSharedT<CH> &sh = *reinterpret_cast<ObBandsSharedT<CH> *>(smem_raw + (size_t)warp * sizeof(ObBandsSharedT<CH>));
    ObWarp g;
    ob_reconstruct_bands(g, fr, seed, sh, Xg + (size_t)w * OB_X_STRIDE);
}

__global__ void __launch_bounds__(OB_BANDS_WARPS * 32)
ob_k_bands_stragglers(const ObFrameIR *__restrict__ ir, float *__restrict__ Xg, const int32_t *__restrict__ strag_list, const int32_t *__restrict__ strag_count)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, n = *strag_count;
    ObBandsShared &sh = *reinterpret_cast<ObBandsShared *>(smem_raw + (size_t)warp * OB_BANDS_SMEM_PER_WARP);
    ObWarp g;
    for (int i = blockIdx.x * OB_BANDS_WARPS + warp; i < n; i += gridDim.x * OB_BANDS_WARPS) {
        const size_t w = (size_t)strag_list[i];
        const ObFrameIR *fr = ir + w;
        ob_reconstruct_bands(g, fr, fr->hdr.seed_in, sh, Xg + w * OB_X_STRIDE);
        __syncwarp();
    }
}

// CH = 2: any decoder.  CH = 1: mono decoders while every frame of the launch is mono (mono-sized shared memory: 10 blocks of 64 threads per SM
// instead of 7 of 96; block shapes measured in dec_synth.cuh ObSynthSharedT).  `stereo_frames` is the straggler counter the mono band kernel has just filled: the <1> instantiation runs when
// it is zero, the <2> instantiation launched behind it when it is not -- the choice needs no host round trip.
template <int CH>
__global__ void __launch_bounds__(ObSynthSharedT<CH>::synth_threads, ObSynthSharedT<CH>::synth_blocks)
ob_k_synth(const int32_t *__restrict__ stereo_frames, const ObFrameIR *__restrict__ ir, const ObSlot *__restrict__ slots, const int32_t *__restrict__ nslots, const float *__restrict__ Xg,
           ObDecState *__restrict__ st, float *__restrict__ hist, float *__restrict__ ring,
           float *__restrict__ pcm, int16_t *__restrict__ pcm16, int32_t *__restrict__ samples, uint32_t *__restrict__ ranges, int S, int F, int cap,
           int CC, int frame_size, int f0, int Fc, float decode_gain, int ds)
{
    __shared__ ObSynthSharedT<CH> sh;
    const int s = blockIdx.x;
    if (s >= S) return;
    if (stereo_frames && (*stereo_frames != 0) == (CH == 1)) return;
    ObBlock g(sh.red);
    ObDecState *state = st + s;
    ob_synth_init(g, sh);
    // ---- state: global -> shared ----
    for (int i = g.lane; i < 2 * OB_NB; i += g.n) {
        sh.oldBandE[i] = state->oldBandE[i]; sh.oldLogE[i] = state->oldLogE[i];
        sh.oldLogE2[i] = state->oldLogE2[i]; sh.backgroundLogE[i] = state->backgroundLogE[i];
    }
    for (int c = 0; c < CC; c++) {
        const float *h = hist + ((size_t)s * CC + c) * OB_HIST_LEN;
        for (int i = g.lane; i < OB_HIST_LEN; i += g.n) sh.buf[c][i] = h[i];
    }
    if (g.lane == 0) {
        sh.pf_period = state->pf_period; sh.pf_period_old = state->pf_period_old;
        sh.pf_tapset = state->pf_tapset; sh.pf_tapset_old = state->pf_tapset_old;
        sh.pf_gain = state->pf_gain; sh.pf_gain_old = state->pf_gain_old;
        sh.preemph_mem[0] = state->preemph_mem[0]; sh.preemph_mem[1] = state->preemph_mem[1];
        sh.last_pitch_index = state->last_pitch_index; sh.paf = state->prefilter_and_fold;
        sh.ring_pos = state->ring_pos; sh.ring = ring + (size_t)s * CC * OB_RING; sh.decode_gain = decode_gain; sh.ds = ds;
        sh.softclip_mem[0] = state->softclip_mem[0]; sh.softclip_mem[1] = state->softclip_mem[1];
    }
    for (int i = g.lane; i < 2 * 24; i += g.n) sh.lpc[i / 24][i % 24] = state->lpc[i / 24][i % 24];
    uint32_t final_range = state->final_range;
    int last_dur = state->last_packet_duration;
    g.sync();
    int acc = state->pkt_samples;                                   // samples / error of the packet whose frames are being decoded
    const int hi = min(f0 + Fc, nslots[s]);
    for (int f = f0; f < hi; f++) {
        const size_t w = (size_t)s * cap + f;
        const ObSlot sl = slots[w];
        const size_t pk = (size_t)s * F + sl.pkt;                   // the caller's packet slot
        const int n = ob_synth_frame(g, sh, ir + w, Xg + w * OB_X_STRIDE, pcm + (pk * (size_t)frame_size + sl.sample_off) * CC, CC);
        if (sl.flags & OB_SLOT_FIRST) acc = 0;
        if (n > 0) {
            if (acc >= 0) acc += n;
            if (!(sh.hdr.flags & OB_F_LOST)) final_range = sh.hdr.final_range;
            else if (sh.hdr.end_in != 0) final_range = 0;          // concealed frame: rangeFinal = 0 (opus_decoder.c:651-652)
        } else acc = n;                                             // the packet fails with its first failing frame (opus_decoder.c:786-787)
        if (sl.flags & OB_SLOT_LAST) {
            if (acc > 0) last_dur = acc;
            if (g.lane == 0) { samples[pk] = acc; if (ranges) ranges[pk] = final_range; }
            if (acc > 0) {                                          // opus_decoder.c:803-807: soft clip for the int16 API, else forget its state
                // a whole-packet concealment (lost packet, or decode_fec on a CELT packet) leaves opus_decode_native through its len == 0
                // branch (:714-729), before the soft clip and before softclip_mem is touched: saturate only, keep the clipper's state
                const int concealed = sl.status > 0 && sl.toc == 0;
                if (pcm16) ob_packet_to_int16(g, pcm + pk * (size_t)frame_size * CC, pcm16 + pk * (size_t)frame_size * CC, acc, CC, sh.softclip_mem, !concealed);
                else if (g.lane == 0 && !concealed) sh.softclip_mem[0] = sh.softclip_mem[1] = 0.f;
            }
        }
        g.sync();
    }
    // ---- state: shared -> global ----
    for (int i = g.lane; i < 2 * OB_NB; i += g.n) {
        state->oldBandE[i] = sh.oldBandE[i]; state->oldLogE[i] = sh.oldLogE[i];
        state->oldLogE2[i] = sh.oldLogE2[i]; state->backgroundLogE[i] = sh.backgroundLogE[i];
    }
    for (int c = 0; c < CC; c++) {
        float *h = hist + ((size_t)s * CC + c) * OB_HIST_LEN;
        for (int i = g.lane; i < OB_HIST_LEN; i += g.n) h[i] = sh.buf[c][i];
    }
    if (g.lane == 0) {
        state->pf_period = sh.pf_period; state->pf_period_old = sh.pf_period_old;
        state->pf_tapset = sh.pf_tapset; state->pf_tapset_old = sh.pf_tapset_old;
        state->pf_gain = sh.pf_gain; state->pf_gain_old = sh.pf_gain_old;
        state->preemph_mem[0] = sh.preemph_mem[0]; state->preemph_mem[1] = sh.preemph_mem[1];
        state->final_range = final_range; state->last_packet_duration = last_dur;
        state->last_pitch_index = sh.last_pitch_index; state->prefilter_and_fold = sh.paf; state->ring_pos = sh.ring_pos;
        state->pkt_samples = acc;
        state->softclip_mem[0] = sh.softclip_mem[0]; state->softclip_mem[1] = sh.softclip_mem[1];
    }
    for (int i = g.lane; i < 2 * 24; i += g.n) state->lpc[i / 24][i % 24] = sh.lpc[i / 24][i % 24];
}

// OPUS_RESET_STATE (celt_decoder.c:1514-1529): zero everything, oldLogE = oldLogE2 = -28.
__global__ void ob_k_reset(ObDecState *st, float *hist, float *ring, const int32_t *idx, int n, int S, int CC)
{
    const int k = blockIdx.x;
    if (k >= n) return;
    const int s = idx ? idx[k] : k;
    if (s < 0 || s >= S) return;
    ObDecState *state = st + s;
    for (int i = threadIdx.x; i < (int)(sizeof(ObDecState) / 4); i += blockDim.x) ((uint32_t *)state)[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * OB_NB; i += blockDim.x) { state->oldLogE[i] = -28.f; state->oldLogE2[i] = -28.f; }
    if (threadIdx.x == 0) { state->skip_plc = 1; state->last_frame_size = OB_SHORT; }      // celt_decoder.c:1527; opus_decoder.c:161 (frame_size = Fs/400)
    float *h = hist + (size_t)s * CC * OB_HIST_LEN;
    for (int i = threadIdx.x; i < CC * OB_HIST_LEN; i += blockDim.x) h[i] = 0.f;
    float *r = ring + (size_t)s * CC * OB_RING;
    for (int i = threadIdx.x; i < CC * OB_RING; i += blockDim.x) r[i] = 0.f;
}

__global__ void ob_k_gather_state(const ObDecState *st, uint32_t *ranges, int32_t *durations, int S, int pitch)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    if (ranges) ranges[s] = st[s].final_range;
    if (durations) durations[s] = pitch ? st[s].pf_period : st[s].last_packet_duration;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
#ifndef OB_MAX_CHUNKS
#define OB_MAX_CHUNKS 16
#endif
struct ObDecoder {
    int S, CC, device, max_frames;
    int ds;                               // 48000 / output sample rate
    int gain_q8, phase_inv_disabled;      // OPUS_SET_GAIN (Q8 dB), OPUS_SET_PHASE_INVERSION_DISABLED: one value for the batch
    int decode_fec;                       // the decode_fec argument of opus_decode / opus_decode_float for the calls that follow
    float gain_linear;
    cudaStream_t stream, copy_stream, aux_stream;
    cudaEvent_t ev[4], chunk_ev[OB_MAX_CHUNKS], copy_done, h2d_done;
    bool timed;
    ObDecState *d_state;
    float *d_hist, *d_ring;
    ObSlot *d_slots; int32_t *d_nslots;   // [S][max_frames] frame slots of the current call, [S] their counts
    int16_t *d_pcm16_2[2], *cur_pcm16;    // int16 API: device-side int16 output (per buffer pair), the one of the call being enqueued
    int32_t *d_multi, *h_multi;           // "some packet is not one slot": a word of mapped pinned host memory (host pointer, device alias)
    cudaEvent_t framed;
    int32_t *d_strag_list, *d_strag_count; int n_sm;   // mono decoders: stereo frames met by the mono-sized band kernel (ob_k_bands<1>)
    cudaStream_t in_stream;               // host->device staging of a call's packets: runs beside the previous call's band / synthesis kernels
    cudaEvent_t syms_done, in_ready;      // the symbol kernel has read the staged packets / the next call's packets are staged
    ObFrameIR *d_ir;
    float *d_X;
    // staging for the host-pointer entry points
    uint8_t *d_packets; size_t packets_cap;
    int32_t *d_offsets, *d_lens, *d_samples2[2]; uint32_t *d_ranges2[2];
    float *d_pcm2[2]; size_t pcm_cap2[2];
    int32_t *d_gather;                 // S words: scratch of the per-stream getters
    cudaEvent_t out_done[2], aux_done; // the device->host copies that last read output buffer p have finished / the auxiliary stream is idle
    uint32_t call_seq;                 // output buffers alternate between consecutive host-pointer calls (pipelined callers)
    int64_t launches;
};

#define OB_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "opus_b200: %s failed: %s\n", #x, cudaGetErrorString(e_)); return OB_INTERNAL_ERROR; } } while (0)

// Debugging aid (OB_DEBUG_SYNC=1 in the environment): wait for every kernel right after its launch and name the one that faults.
static const bool ob_debug_sync = getenv("OB_DEBUG_SYNC") != nullptr;
#define OB_KCHECK(name) do { if (ob_debug_sync) { cudaError_t e_ = cudaStreamSynchronize(stream); if (e_ == cudaSuccess) e_ = cudaGetLastError(); \
        if (e_ != cudaSuccess) { fprintf(stderr, "opus_b200: kernel %s failed: %s\n", name, cudaGetErrorString(e_)); return OB_INTERNAL_ERROR; } } } while (0)

// Launches the three kernels for streams [s0, s0+Sc).  All per-stream arrays are indexed by stream, so a sub-range is
// just a pointer offset; timed != 0 brackets the kernels with the handle's events.
static int ob_launch(ObDecoder *d, int s0, int Sc, int F, const uint8_t *d_packets, const int32_t *d_offsets, const int32_t *d_lens,
                     float *d_pcm, int frame_size, int32_t *d_samples, uint32_t *d_ranges, int timed, cudaStream_t stream,
                     int f0 = 0, int Fc = -1, int which = 7)
{
    // which: bit 2 = framing kernel, bit 0 = symbol kernel, bit 1 = plan + band + synthesis kernels, for the SLOT window [f0, f0+Fc) of streams
    // [s0, s0+Sc).  Per-packet arrays (offsets, lens, pcm, samples, ranges) have stride F; per-slot arrays (slots, IR, X) have
    // stride cap = max_frames.  Fc < 0: every slot the decoder has room for (packets may hold several frames).
    const int cap = d->max_frames;
    if (Fc < 0) Fc = cap;
    const int total = Sc * Fc;
    const size_t w0 = (size_t)s0 * F, c0 = (size_t)s0 * cap;
    ObFrameIR *ir = d->d_ir + c0;
    float *X = d->d_X + c0 * OB_X_STRIDE;
    ObSlot *slots = d->d_slots + c0;
    int32_t *nslots = d->d_nslots + s0;
    ObDecState *st = d->d_state + s0;
    float *hist = d->d_hist + (size_t)s0 * d->CC * OB_HIST_LEN;
    float *ring = d->d_ring + (size_t)s0 * d->CC * OB_RING;
    if (timed) OB_CUDA(cudaEventRecord(d->ev[0], stream));
    if (which & 4) {
        ob_k_frame<<<(Sc + 127) / 128, 128, 0, stream>>>(d_packets, d_offsets + w0, d_lens + w0, slots, nslots, Sc, F, frame_size, cap, d->d_multi, d->ds, d->decode_fec);
        d->launches += 1;
        OB_KCHECK("ob_k_frame");
    }
    if (which & 1) {
        if (d->CC == 1)
            ob_k_symbols<8><<<(total + OB_SYM_THREADS - 1) / OB_SYM_THREADS, OB_SYM_THREADS, 0, stream>>>(
                d_packets, slots, nslots, ir, total, d->CC, cap, f0, Fc, d->phase_inv_disabled);
        else
            ob_k_symbols<6><<<(total + OB_SYM_THREADS - 1) / OB_SYM_THREADS, OB_SYM_THREADS, 0, stream>>>(
                d_packets, slots, nslots, ir, total, d->CC, cap, f0, Fc, d->phase_inv_disabled);
        d->launches += 1;
        OB_KCHECK("ob_k_symbols");
    }
    if (timed) OB_CUDA(cudaEventRecord(d->ev[1], stream));
    if (which & 2) {
        ob_k_plan<<<(Sc + 3) / 4, 128, 0, stream>>>(ir, st, nslots, Sc, cap, f0, Fc, d->CC);
        OB_KCHECK("ob_k_plan");
        if (d->CC == 1) {
            // one counter per compute stream (launches on a stream are ordered), one list region per stream range
            int32_t *cnt = d->d_strag_count + (stream == d->aux_stream ? 1 : 0), *list = d->d_strag_list + c0;
            OB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t), stream));
            const int mw = OB_BANDS_WARPS_MONO;
            ob_k_bands<1><<<(total + mw - 1) / mw, mw * 32, mw * OB_BANDS_SMEM_PER_WARP_MONO, stream>>>(ir, nslots, X, Sc, cap, f0, Fc, list, cnt);
            ob_k_bands_stragglers<<<d->n_sm, OB_BANDS_WARPS * 32, OB_BANDS_WARPS * OB_BANDS_SMEM_PER_WARP, stream>>>(ir, X, list, cnt);
            d->launches += 1;
        } else      // stereo-sized: 2 / 3 / 4 / 6 warps per block measured alike (30.0 ms per 163 840 stereo frames)
            ob_k_bands<2><<<(total + OB_BANDS_WARPS - 1) / OB_BANDS_WARPS, OB_BANDS_WARPS * 32, OB_BANDS_WARPS * OB_BANDS_SMEM_PER_WARP, stream>>>(
                ir, nslots, X, Sc, cap, f0, Fc, nullptr, nullptr);
        OB_KCHECK("ob_k_bands");
        if (timed) OB_CUDA(cudaEventRecord(d->ev[2], stream));
        float *pcm_w = d_pcm + w0 * (size_t)frame_size * d->CC;
        int16_t *pcm16_w = d->cur_pcm16 ? d->cur_pcm16 + w0 * (size_t)frame_size * d->CC : nullptr;
        uint32_t *ranges_w = d_ranges ? d_ranges + w0 : nullptr;
        if (d->CC == 1) {
            const int32_t *cnt = d->d_strag_count + (stream == d->aux_stream ? 1 : 0);
            ob_k_synth<1><<<Sc, ObSynthSharedT<1>::synth_threads, 0, stream>>>(cnt, ir, slots, nslots, X, st, hist, ring, pcm_w, pcm16_w, d_samples + w0, ranges_w, Sc, F, cap, d->CC,
                                                                frame_size, f0, Fc, d->gain_linear, d->ds);
            ob_k_synth<2><<<Sc, ObSynthSharedT<2>::synth_threads, 0, stream>>>(cnt, ir, slots, nslots, X, st, hist, ring, pcm_w, pcm16_w, d_samples + w0, ranges_w, Sc, F, cap, d->CC,
                                                                frame_size, f0, Fc, d->gain_linear, d->ds);
            d->launches += 1;
        } else
            ob_k_synth<2><<<Sc, ObSynthSharedT<2>::synth_threads, 0, stream>>>(nullptr, ir, slots, nslots, X, st, hist, ring, pcm_w, pcm16_w, d_samples + w0, ranges_w, Sc, F, cap, d->CC,
                                                                frame_size, f0, Fc, d->gain_linear, d->ds);
        d->launches += 3;
        OB_KCHECK("ob_k_synth");
    }
    if (timed) OB_CUDA(cudaEventRecord(d->ev[3], stream));
    OB_CUDA(cudaGetLastError());
    if (timed) d->timed = true;
    return OB_OK;
}

extern "C" {

ObDecoder *ob_decoder_create(int32_t n_streams, int32_t fs, int32_t channels, int32_t device, int32_t max_frames, int32_t *error)
{
    int err = OB_OK;
    ObDecoder *d = nullptr;
    int ndev = 0;
    if (n_streams <= 0 || (channels != 1 && channels != 2) || max_frames <= 0 || max_frames > 65535 || (size_t)n_streams * (size_t)max_frames > 0x7fffffffu) err = OB_BAD_ARG;   // ObSlot.pkt is 16 bits; slot indices are 32 bits
    else if (fs != 48000 && fs != 24000 && fs != 16000 && fs != 12000 && fs != 8000) err = OB_BAD_ARG;     // opus_decoder_init, opus_decoder.c:130-131
    else if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        fprintf(stderr, "opus_b200: no usable CUDA device (count=%d, requested=%d); there is no CPU fallback\n", ndev, device);
        err = OB_INTERNAL_ERROR;
    }
    if (err == OB_OK) {
        d = new (std::nothrow) ObDecoder();
        if (!d) err = OB_ALLOC_FAIL;
    }
    if (err == OB_OK) {
        memset(d, 0, sizeof(*d));
        d->S = n_streams; d->CC = channels; d->device = device; d->max_frames = max_frames;
        d->gain_q8 = 0; d->gain_linear = 1.f; d->phase_inv_disabled = 0; d->decode_fec = 0; d->ds = 48000 / fs;
        const size_t total = (size_t)n_streams * max_frames;
        bool ok = cudaSetDevice(device) == cudaSuccess;
        ok = ok && cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
        for (int i = 0; i < 4 && ok; i++) ok = cudaEventCreate(&d->ev[i]) == cudaSuccess;
        for (int i = 0; i < OB_MAX_CHUNKS && ok; i++) ok = cudaEventCreateWithFlags(&d->chunk_ev[i], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&d->copy_done, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&d->h2d_done, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaStreamCreateWithFlags(&d->aux_stream, cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaMalloc(&d->d_state, sizeof(ObDecState) * n_streams) == cudaSuccess;
        ok = ok && cudaMalloc(&d->d_hist, sizeof(float) * (size_t)n_streams * channels * OB_HIST_LEN) == cudaSuccess;
        ok = ok && cudaMalloc(&d->d_ring, sizeof(float) * (size_t)n_streams * channels * OB_RING) == cudaSuccess;
        ok = ok && cudaHostAlloc(&d->h_multi, sizeof(int32_t), cudaHostAllocMapped) == cudaSuccess
                && cudaHostGetDevicePointer(&d->d_multi, d->h_multi, 0) == cudaSuccess
                && cudaEventCreateWithFlags(&d->framed, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaMalloc(&d->d_slots, sizeof(ObSlot) * total) == cudaSuccess && cudaMalloc(&d->d_nslots, sizeof(int32_t) * n_streams) == cudaSuccess;
        ok = ok && cudaMalloc(&d->d_ir, sizeof(ObFrameIR) * total) == cudaSuccess;
        ok = ok && cudaMalloc(&d->d_X, sizeof(float) * OB_X_STRIDE * total) == cudaSuccess;
        ok = ok && cudaMalloc(&d->d_offsets, sizeof(int32_t) * total) == cudaSuccess;
        ok = ok && cudaMalloc(&d->d_lens, sizeof(int32_t) * total) == cudaSuccess;
        for (int p = 0; p < 2 && ok; p++) {
            ok = cudaMalloc(&d->d_samples2[p], sizeof(int32_t) * total) == cudaSuccess && cudaMalloc(&d->d_ranges2[p], sizeof(uint32_t) * total) == cudaSuccess
                 && cudaEventCreateWithFlags(&d->out_done[p], cudaEventDisableTiming) == cudaSuccess;
        }
        ok = ok && cudaEventCreateWithFlags(&d->aux_done, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaStreamCreateWithFlags(&d->in_stream, cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&d->syms_done, cudaEventDisableTiming) == cudaSuccess
                && cudaEventCreateWithFlags(&d->in_ready, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaMalloc(&d->d_gather, sizeof(int32_t) * n_streams) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(ob_k_bands<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, OB_BANDS_WARPS * OB_BANDS_SMEM_PER_WARP) == cudaSuccess;
        ok = ok && cudaFuncSetAttribute(ob_k_bands_stragglers, cudaFuncAttributeMaxDynamicSharedMemorySize, OB_BANDS_WARPS * OB_BANDS_SMEM_PER_WARP) == cudaSuccess;
        if (channels == 1) {
            cudaDeviceProp prop;
            ok = ok && cudaGetDeviceProperties(&prop, device) == cudaSuccess;
            d->n_sm = ok ? prop.multiProcessorCount : 1;
            ok = ok && cudaMalloc(&d->d_strag_list, sizeof(int32_t) * (size_t)n_streams * max_frames) == cudaSuccess
                    && cudaMalloc(&d->d_strag_count, 2 * sizeof(int32_t)) == cudaSuccess;
        }
        if (!ok) {
            fprintf(stderr, "opus_b200: device allocation failed: %s\n", cudaGetErrorString(cudaGetLastError()));
            ob_decoder_destroy(d);
            d = nullptr;
            err = OB_ALLOC_FAIL;
        } else if (ob_decoder_reset(d, nullptr, 0) != OB_OK) {
            ob_decoder_destroy(d);
            d = nullptr;
            err = OB_INTERNAL_ERROR;
        }
    }
    if (error) *error = err;
    return d;
}

void ob_decoder_destroy(ObDecoder *d)
{
    if (!d) return;
    cudaSetDevice(d->device);
    // every stream of this decoder drains first: an asynchronous call may still be copying into the caller's buffers
    if (d->in_stream) cudaStreamSynchronize(d->in_stream);
    if (d->stream) cudaStreamSynchronize(d->stream);
    if (d->aux_stream) cudaStreamSynchronize(d->aux_stream);
    if (d->copy_stream) cudaStreamSynchronize(d->copy_stream);
    cudaFree(d->d_state); cudaFree(d->d_hist); cudaFree(d->d_ring); cudaFree(d->d_slots); cudaFree(d->d_nslots); if (d->h_multi) cudaFreeHost(d->h_multi); if (d->framed) cudaEventDestroy(d->framed); cudaFree(d->d_ir); cudaFree(d->d_X); cudaFree(d->d_packets); cudaFree(d->d_strag_list); cudaFree(d->d_strag_count);
    cudaFree(d->d_offsets); cudaFree(d->d_lens); cudaFree(d->d_gather);
    for (int p = 0; p < 2; p++) { cudaFree(d->d_samples2[p]); cudaFree(d->d_ranges2[p]); cudaFree(d->d_pcm2[p]); cudaFree(d->d_pcm16_2[p]); if (d->out_done[p]) cudaEventDestroy(d->out_done[p]); }
    if (d->aux_done) cudaEventDestroy(d->aux_done);
    if (d->syms_done) cudaEventDestroy(d->syms_done);
    if (d->in_ready) cudaEventDestroy(d->in_ready);
    if (d->in_stream) cudaStreamDestroy(d->in_stream);
    for (int i = 0; i < 4; i++) if (d->ev[i]) cudaEventDestroy(d->ev[i]);
    for (int i = 0; i < OB_MAX_CHUNKS; i++) if (d->chunk_ev[i]) cudaEventDestroy(d->chunk_ev[i]);
    if (d->copy_done) cudaEventDestroy(d->copy_done);
    if (d->h2d_done) cudaEventDestroy(d->h2d_done);
    if (d->aux_stream) cudaStreamDestroy(d->aux_stream);
    if (d->copy_stream) cudaStreamDestroy(d->copy_stream);
    if (d->stream) cudaStreamDestroy(d->stream);
    delete d;
}

int32_t ob_decoder_reset(ObDecoder *d, const int32_t *idx, int32_t n)
{
    if (!d || n < 0) return OB_BAD_ARG;
    OB_CUDA(cudaSetDevice(d->device));
    // Legal with asynchronous calls in flight: they are drained first (their synthesis may still be running on the auxiliary stream
    // over the very state this call clears).
    { const int r = ob_decoder_wait(d, 0); if (r != OB_OK && r != OB_BAD_ARG) return r; }
    if (d->aux_stream) OB_CUDA(cudaStreamSynchronize(d->aux_stream));
    int32_t *d_idx = nullptr;
    int count = d->S;
    if (idx) {
        if (n == 0) return OB_OK;
        count = n;
        OB_CUDA(cudaMalloc(&d_idx, sizeof(int32_t) * n));
    }
    cudaError_t err = d_idx ? cudaMemcpyAsync(d_idx, idx, sizeof(int32_t) * n, cudaMemcpyHostToDevice, d->stream) : cudaSuccess;
    if (err == cudaSuccess) {
        ob_k_reset<<<count, 128, 0, d->stream>>>(d->d_state, d->d_hist, d->d_ring, d_idx, count, d->S, d->CC);
        d->launches += 1;
        err = cudaStreamSynchronize(d->stream);
    }
    if (d_idx) cudaFree(d_idx);                                    // on every exit path
    if (err != cudaSuccess) { fprintf(stderr, "opus_b200: decoder reset failed: %s\n", cudaGetErrorString(err)); return OB_INTERNAL_ERROR; }
    return OB_OK;
}

int32_t ob_decode_float_device(ObDecoder *d, int32_t n_frames, const uint8_t *d_packets, const int32_t *d_offsets,
                               const int32_t *d_lens, float *d_pcm_out, int32_t frame_size, int32_t *d_samples_out,
                               uint32_t *d_ranges_out, int32_t sync)
{
    if (!d || !d_packets || !d_offsets || !d_lens || !d_pcm_out || !d_samples_out) return OB_BAD_ARG;
    if (n_frames <= 0 || n_frames > d->max_frames || frame_size <= 0) return OB_BAD_ARG;
    OB_CUDA(cudaSetDevice(d->device));
    d->cur_pcm16 = nullptr;
    const int r = ob_launch(d, 0, d->S, n_frames, d_packets, d_offsets, d_lens, d_pcm_out, frame_size, d_samples_out, d_ranges_out, 1, d->stream);
    if (r != OB_OK) return r;
    if (sync) OB_CUDA(cudaStreamSynchronize(d->stream));
    return OB_OK;
}

// Enqueues one host-pointer call: packets up, kernels, PCM / samples / ranges down.  wait != 0: returns when the outputs are in
// the caller's buffers.  wait == 0: returns as soon as everything is enqueued (ob_decoder_wait completes it); the next call's
// kernels then overlap this call's device->host copies, which is why the device-side output buffers alternate between calls.
static int32_t ob_decode_submit(ObDecoder *d, int32_t n_frames, const uint8_t *packets, const int32_t *offsets, const int32_t *lens,
                                float *pcm_out, int16_t *pcm16_out, int32_t frame_size, int32_t *samples_out, uint32_t *ranges_out, int wait)
{
    if (!d || !packets || !offsets || !lens || (!pcm_out && !pcm16_out) || !samples_out) return OB_BAD_ARG;
    if (n_frames <= 0 || n_frames > d->max_frames || frame_size <= 0) return OB_BAD_ARG;
    OB_CUDA(cudaSetDevice(d->device));
    const size_t total = (size_t)d->S * n_frames;
    size_t nbytes = 0;
    for (size_t i = 0; i < total; i++) {
        if (lens[i] < 0 || offsets[i] < 0) return OB_BAD_ARG;
        const size_t e = (size_t)offsets[i] + (size_t)lens[i];
        if (lens[i] > 0 && e > nbytes) nbytes = e;
    }
    if (nbytes == 0) nbytes = 1;
    if (nbytes > d->packets_cap) {
        cudaFree(d->d_packets); d->d_packets = nullptr; d->packets_cap = 0;
        OB_CUDA(cudaMalloc(&d->d_packets, nbytes + nbytes / 4));
        d->packets_cap = nbytes + nbytes / 4;
    }
    const int par = (int)(d->call_seq++ & 1u);
    const size_t pcm_floats = total * (size_t)frame_size * d->CC;
    OB_CUDA(cudaStreamWaitEvent(d->stream, d->out_done[par], 0));      // the copies of the call before last (same buffers) are done
    OB_CUDA(cudaStreamWaitEvent(d->stream, d->aux_done, 0));           // nobody still reads the staged packets of the previous call
    if (pcm_floats > d->pcm_cap2[par]) {
        OB_CUDA(cudaEventSynchronize(d->out_done[par]));
        cudaFree(d->d_pcm2[par]); d->d_pcm2[par] = nullptr; d->pcm_cap2[par] = 0;
        OB_CUDA(cudaMalloc(&d->d_pcm2[par], pcm_floats * sizeof(float)));
        cudaFree(d->d_pcm16_2[par]); d->d_pcm16_2[par] = nullptr;
        OB_CUDA(cudaMalloc(&d->d_pcm16_2[par], pcm_floats * sizeof(int16_t)));
        d->pcm_cap2[par] = pcm_floats;
    }
    d->cur_pcm16 = pcm16_out ? d->d_pcm16_2[par] : nullptr;
    // what travels back to the host: the float PCM, or -- int16 API -- the soft-clipped, rounded int16 PCM (half the bytes)
    const size_t es = pcm16_out ? sizeof(int16_t) : sizeof(float);
    const char *const d_out = pcm16_out ? (const char *)d->d_pcm16_2[par] : (const char *)d->d_pcm2[par];
    char *const h_out = pcm16_out ? (char *)pcm16_out : (char *)pcm_out;
    float *const d_pcm = d->d_pcm2[par];
    int32_t *const d_samples = d->d_samples2[par];
    uint32_t *const d_ranges = d->d_ranges2[par];
    // Only the framing and symbol kernels read the staged packets: once the previous call's symbol kernel is done the staging buffers are
    // free, and this call's packets come up on their own stream while that call's band / synthesis windows are still running.
    OB_CUDA(cudaStreamWaitEvent(d->in_stream, d->syms_done, 0));
    OB_CUDA(cudaStreamWaitEvent(d->in_stream, d->aux_done, 0));
    OB_CUDA(cudaMemcpyAsync(d->d_packets, packets, nbytes, cudaMemcpyHostToDevice, d->in_stream));
    OB_CUDA(cudaMemcpyAsync(d->d_offsets, offsets, total * sizeof(int32_t), cudaMemcpyHostToDevice, d->in_stream));
    OB_CUDA(cudaMemcpyAsync(d->d_lens, lens, total * sizeof(int32_t), cudaMemcpyHostToDevice, d->in_stream));
    OB_CUDA(cudaEventRecord(d->in_ready, d->in_stream));
    OB_CUDA(cudaStreamWaitEvent(d->stream, d->in_ready, 0));
    // Framing first: it tells whether slot j is packet j for every stream (always true for code-0 packets).  The symbol kernel is
    // launched behind it right away; the host only waits for the one-word answer of the framing kernel.
    *d->h_multi = 0;                                 // the previous call's framing kernel has completed (we waited for it)
    {
        const int r = ob_launch(d, 0, d->S, n_frames, d->d_packets, d->d_offsets, d->d_lens, d_pcm, frame_size, d_samples, d_ranges, 0, d->stream, 0, -1, 4);
        if (r != OB_OK) return r;
    }
    OB_CUDA(cudaEventRecord(d->framed, d->stream));
    {
        const int r = ob_launch(d, 0, d->S, n_frames, d->d_packets, d->d_offsets, d->d_lens, d_pcm, frame_size, d_samples, d_ranges, 0, d->stream, 0, -1, 1);
        if (r != OB_OK) return r;
    }
    OB_CUDA(cudaEventRecord(d->syms_done, d->stream));
    OB_CUDA(cudaEventSynchronize(d->framed));
    const int multi = *(volatile int32_t *)d->h_multi;
    // The call is processed in chunks so that the device->host copy of chunk k overlaps the kernels of chunk k+1 (kernels on
    // d->stream, copies on d->copy_stream, one event per chunk).  With several frames per stream the chunks are FRAME windows
    // of all streams: every launch keeps the full stream-level parallelism the synthesis kernel needs (one block per stream),
    // and the per-stream state simply carries over from launch to launch.  Single-frame calls are split by stream ranges.
    // measured: 819 200 mono frames (F = 200), blocking call: 3 chunks 96.8 ms, 6 chunks 89.4 ms, chunks as stream ranges (contiguous copies instead of
    // 2-D ones) 97.9 ms; 163 840 stereo frames (F = 10), two calls in flight: 2 / 3 / 5 / 10 chunks -> 120 / 123 / 126 / 96 k audio-s/s;
    // 204 800 mono frames (F = 50), two calls in flight: 1: 202 k, 2: 208 k, 3: 209 k, 5: 209 k (int16: 205 / 206 / 206 / 200 k), 8: 202 k
    int nchunks = total >= 400000 ? 6 : (total >= 131072 ? 5 : (total >= 65536 ? 3 : (total >= 16384 ? 2 : 1)));
    if (n_frames < nchunks) {
        // few frames per stream (the live case, F = 1): chunks are stream ranges, kernels alternate between two compute streams, and the copies of
        // chunk k hide behind the kernels of chunk k+1.  Measured blocking latency of one call, mono: 131 072 streams 17.6 ms with 3 chunks, 14.5 with
        // 12, 14.0 with 16 (24 / 32: no further gain); 65 536 streams 9.5 -> 7.6 ms; 163 840 streams 22.9 -> 18.4 ms.
        const size_t by_size = total / 8192;
        if (by_size > (size_t)nchunks) nchunks = by_size > OB_MAX_CHUNKS ? OB_MAX_CHUNKS : (int)by_size;
    }
    if (const char *v = getenv("OB_DEC_CHUNKS")) { const int t = atoi(v); if (t >= 1) nchunks = t; }      // tuning aid
    if (nchunks > OB_MAX_CHUNKS) nchunks = OB_MAX_CHUNKS;
    const size_t pf = (size_t)frame_size * d->CC;
    if (multi) {
        // Packets with several frames: the slot <-> packet mapping is data dependent, so the call runs as one window over every
        // slot the decoder has room for and the outputs come down in one piece.
        const int r = ob_launch(d, 0, d->S, n_frames, d->d_packets, d->d_offsets, d->d_lens, d_pcm, frame_size, d_samples, d_ranges, 0, d->stream, 0, -1, 2);
        if (r != OB_OK) return r;
        OB_CUDA(cudaEventRecord(d->chunk_ev[0], d->stream));
        OB_CUDA(cudaStreamWaitEvent(d->copy_stream, d->chunk_ev[0], 0));
        OB_CUDA(cudaMemcpyAsync(h_out, d_out, pcm_floats * es, cudaMemcpyDeviceToHost, d->copy_stream));
        OB_CUDA(cudaMemcpyAsync(samples_out, d_samples, total * sizeof(int32_t), cudaMemcpyDeviceToHost, d->copy_stream));
        if (ranges_out) OB_CUDA(cudaMemcpyAsync(ranges_out, d_ranges, total * sizeof(uint32_t), cudaMemcpyDeviceToHost, d->copy_stream));
    } else if (n_frames >= nchunks && (size_t)(n_frames / nchunks) * pf * es >= 32768 && !getenv("OB_DEC_STREAM_CHUNKS")) {
        // Frame windows, while a window's row (one stream's share of the 2-D copy) stays large: 163 840 stereo frames at F = 10 would copy 15 KB rows,
        // and measured 126 k audio-s/s end to end against 132.7 k with the same call cut into stream ranges (contiguous copies) below.
        // The symbol kernel is one thread per frame and latency bound (a launch takes >= 1.6 ms however few frames it covers:
        // measured), so it runs once over the whole call; only the band + synthesis kernels are windowed.
        const int per = (n_frames + nchunks - 1) / nchunks;
        for (int k = 0, f0 = 0; f0 < n_frames; k++, f0 += per) {
            const int Fc = n_frames - f0 < per ? n_frames - f0 : per;
            const int r = ob_launch(d, 0, d->S, n_frames, d->d_packets, d->d_offsets, d->d_lens, d_pcm, frame_size, d_samples, d_ranges, 0, d->stream, f0, Fc, 2);
            if (r != OB_OK) return r;
            OB_CUDA(cudaEventRecord(d->chunk_ev[k], d->stream));
            OB_CUDA(cudaStreamWaitEvent(d->copy_stream, d->chunk_ev[k], 0));
            OB_CUDA(cudaMemcpy2DAsync(h_out + f0 * pf * es, n_frames * pf * es, d_out + f0 * pf * es, n_frames * pf * es,
                                      Fc * pf * es, d->S, cudaMemcpyDeviceToHost, d->copy_stream));
        }
        OB_CUDA(cudaMemcpyAsync(samples_out, d_samples, total * sizeof(int32_t), cudaMemcpyDeviceToHost, d->copy_stream));
        if (ranges_out) OB_CUDA(cudaMemcpyAsync(ranges_out, d_ranges, total * sizeof(uint32_t), cudaMemcpyDeviceToHost, d->copy_stream));
    } else {
        const int per = (d->S + nchunks - 1) / nchunks;
        OB_CUDA(cudaEventRecord(d->h2d_done, d->stream));
        OB_CUDA(cudaStreamWaitEvent(d->aux_stream, d->h2d_done, 0));
        for (int k = 0, s0 = 0; s0 < d->S; k++, s0 += per) {
            const int Sc = d->S - s0 < per ? d->S - s0 : per;
            const size_t w0 = (size_t)s0 * n_frames, cnt = (size_t)Sc * n_frames;
            cudaStream_t cs = (k & 1) ? d->aux_stream : d->stream;       // alternate compute streams: kernels of neighbouring chunks overlap
            const int r = ob_launch(d, s0, Sc, n_frames, d->d_packets, d->d_offsets, d->d_lens, d_pcm, frame_size, d_samples, d_ranges, 0, cs, 0, -1, 2);
            if (r != OB_OK) return r;
            OB_CUDA(cudaEventRecord(d->chunk_ev[k], cs));
            OB_CUDA(cudaStreamWaitEvent(d->copy_stream, d->chunk_ev[k], 0));
            OB_CUDA(cudaMemcpyAsync(h_out + w0 * pf * es, d_out + w0 * pf * es, cnt * pf * es, cudaMemcpyDeviceToHost, d->copy_stream));
            OB_CUDA(cudaMemcpyAsync(samples_out + w0, d_samples + w0, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, d->copy_stream));
            if (ranges_out) OB_CUDA(cudaMemcpyAsync(ranges_out + w0, d_ranges + w0, cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost, d->copy_stream));
        }
    }
    OB_CUDA(cudaEventRecord(d->aux_done, d->aux_stream));
    OB_CUDA(cudaEventRecord(d->out_done[par], d->copy_stream));
    if (wait) OB_CUDA(cudaEventSynchronize(d->out_done[par]));
    return OB_OK;
}

int32_t ob_decode_float_multi(ObDecoder *d, int32_t n_frames, const uint8_t *packets, const int32_t *offsets, const int32_t *lens,
                              float *pcm_out, int32_t frame_size, int32_t *samples_out, uint32_t *ranges_out)
{
    return ob_decode_submit(d, n_frames, packets, offsets, lens, pcm_out, nullptr, frame_size, samples_out, ranges_out, 1);
}

// opus_decode (int16 PCM): Decoder::decode (src/decoder.rs:75-127).
int32_t ob_decode_multi(ObDecoder *d, int32_t n_frames, const uint8_t *packets, const int32_t *offsets, const int32_t *lens,
                        int16_t *pcm_out, int32_t frame_size, int32_t *samples_out, uint32_t *ranges_out)
{
    return ob_decode_submit(d, n_frames, packets, offsets, lens, nullptr, pcm_out, frame_size, samples_out, ranges_out, 1);
}
int32_t ob_decode_multi_async(ObDecoder *d, int32_t n_frames, const uint8_t *packets, const int32_t *offsets, const int32_t *lens,
                              int16_t *pcm_out, int32_t frame_size, int32_t *samples_out, uint32_t *ranges_out)
{
    return ob_decode_submit(d, n_frames, packets, offsets, lens, nullptr, pcm_out, frame_size, samples_out, ranges_out, 0);
}
int32_t ob_decode(ObDecoder *d, const uint8_t *packets, const int32_t *offsets, const int32_t *lens, int16_t *pcm_out, int32_t frame_size, int32_t *samples_out)
{
    return ob_decode_submit(d, 1, packets, offsets, lens, nullptr, pcm_out, frame_size, samples_out, nullptr, 1);
}

int32_t ob_decode_float_multi_async(ObDecoder *d, int32_t n_frames, const uint8_t *packets, const int32_t *offsets, const int32_t *lens,
                                    float *pcm_out, int32_t frame_size, int32_t *samples_out, uint32_t *ranges_out)
{
    return ob_decode_submit(d, n_frames, packets, offsets, lens, pcm_out, nullptr, frame_size, samples_out, ranges_out, 0);
}

// Debug / test access to the integer intermediate representation of the LAST call (ob_ir.h): the record the symbol kernel wrote for
// (stream, slot).  Lets a test compare the decoded energy indices, allocation and pulse vectors with the reference's, exactly.
int32_t ob_debug_ir_layout(int32_t *out, int32_t n)
{
    const int32_t v[8] = {(int32_t)sizeof(ObFrameIR), (int32_t)sizeof(ObFrameHdr), (int32_t)offsetof(ObFrameIR, bands), (int32_t)offsetof(ObFrameIR, leaves),
                          (int32_t)offsetof(ObFrameIR, iy), (int32_t)sizeof(ObLeaf), (int32_t)sizeof(ObBand), OB_MAX_LEAVES};
    if (!out || n < 8) return OB_BAD_ARG;
    for (int i = 0; i < 8; i++) out[i] = v[i];
    return OB_OK;
}
int32_t ob_decoder_debug_read_ir(ObDecoder *d, int32_t stream, int32_t slot, void *out, int32_t nbytes)
{
    if (!d || !out || stream < 0 || stream >= d->S || slot < 0 || slot >= d->max_frames || nbytes < (int32_t)sizeof(ObFrameIR)) return OB_BAD_ARG;
    OB_CUDA(cudaSetDevice(d->device));
    { const int r = ob_decoder_wait(d, 0); if (r != OB_OK) return r; }
    OB_CUDA(cudaMemcpy(out, d->d_ir + (size_t)stream * d->max_frames + slot, sizeof(ObFrameIR), cudaMemcpyDeviceToHost));
    return OB_OK;
}

int32_t ob_decoder_wait(ObDecoder *d, int32_t keep_in_flight)
{
    if (!d || keep_in_flight < 0 || keep_in_flight > 1) return OB_BAD_ARG;
    OB_CUDA(cudaSetDevice(d->device));
    if (keep_in_flight) {              // everything but the most recent call: its outputs were the other buffer pair
        if (d->call_seq >= 2) OB_CUDA(cudaEventSynchronize(d->out_done[d->call_seq & 1u]));
        return OB_OK;
    }
    OB_CUDA(cudaStreamSynchronize(d->stream));
    OB_CUDA(cudaStreamSynchronize(d->aux_stream));
    OB_CUDA(cudaStreamSynchronize(d->copy_stream));
    return OB_OK;
}

int32_t ob_decode_float(ObDecoder *d, const uint8_t *packets, const int32_t *offsets, const int32_t *lens, float *pcm_out,
                        int32_t frame_size, int32_t *samples_out)
{
    return ob_decode_float_multi(d, 1, packets, offsets, lens, pcm_out, frame_size, samples_out, nullptr);
}

static int32_t ob_gather(ObDecoder *d, uint32_t *ranges, int32_t *durations, int pitch = 0)
{
    OB_CUDA(cudaSetDevice(d->device));
    ob_k_gather_state<<<(d->S + 127) / 128, 128, 0, d->stream>>>(d->d_state, ranges ? (uint32_t *)d->d_gather : nullptr, durations ? d->d_gather : nullptr, d->S, pitch);
    d->launches += 1;
    if (ranges) OB_CUDA(cudaMemcpyAsync(ranges, d->d_gather, sizeof(uint32_t) * d->S, cudaMemcpyDeviceToHost, d->stream));
    if (durations) OB_CUDA(cudaMemcpyAsync(durations, d->d_gather, sizeof(int32_t) * d->S, cudaMemcpyDeviceToHost, d->stream));
    OB_CUDA(cudaStreamSynchronize(d->stream));
    return OB_OK;
}
int32_t ob_decoder_final_range(ObDecoder *d, uint32_t *out) { return (!d || !out) ? OB_BAD_ARG : ob_gather(d, out, nullptr); }
int32_t ob_decoder_last_packet_duration(ObDecoder *d, int32_t *out) { return (!d || !out) ? OB_BAD_ARG : ob_gather(d, nullptr, out); }
// OPUS_GET_PITCH (opus_decoder.c:987-999 -> celt_decoder.c OPUS_GET_PITCH: st->postfilter_period; 0 before the first packet)
int32_t ob_decoder_get_pitch(ObDecoder *d, int32_t *out) { return (!d || !out) ? OB_BAD_ARG : ob_gather(d, nullptr, out, 1); }

// OPUS_SET_GAIN / OPUS_GET_GAIN (opus_decoder.c:985-1004): Q8 dB, applied as pcm *= celt_exp2(6.48814081e-4 * gain) (:639-649).
int32_t ob_decoder_set_gain(ObDecoder *d, int32_t gain_q8)
{
    if (!d || gain_q8 < -32768 || gain_q8 > 32767) return OB_BAD_ARG;
    d->gain_q8 = gain_q8;
    const float x = 6.48814081e-4f * (float)gain_q8;
    d->gain_linear = gain_q8 == 0 ? 1.f : (float)exp(0.6931471805599453094 * x);
    return OB_OK;
}
int32_t ob_decoder_get_gain(ObDecoder *d, int32_t *v) { if (!d || !v) return OB_BAD_ARG; *v = d->gain_q8; return OB_OK; }
// OPUS_SET/GET_PHASE_INVERSION_DISABLED (celt_decoder.c:1560-1579)
int32_t ob_decoder_set_phase_inversion_disabled(ObDecoder *d, int32_t v) { if (!d || v < 0 || v > 1) return OB_BAD_ARG; d->phase_inv_disabled = v; return OB_OK; }
int32_t ob_decoder_get_phase_inversion_disabled(ObDecoder *d, int32_t *v) { if (!d || !v) return OB_BAD_ARG; *v = d->phase_inv_disabled; return OB_OK; }
int32_t ob_decoder_set_decode_fec(ObDecoder *d, int32_t v) { if (!d || v < 0 || v > 1) return OB_BAD_ARG; d->decode_fec = v; return OB_OK; }
int32_t ob_decoder_get_decode_fec(ObDecoder *d, int32_t *v) { if (!d || !v) return OB_BAD_ARG; *v = d->decode_fec; return OB_OK; }

int32_t ob_decoder_streams(const ObDecoder *d) { return d ? d->S : OB_BAD_ARG; }
int32_t ob_decoder_channels(const ObDecoder *d) { return d ? d->CC : OB_BAD_ARG; }
int32_t ob_decoder_sample_rate(const ObDecoder *d) { return d ? 48000 / d->ds : OB_BAD_ARG; }          // OPUS_GET_SAMPLE_RATE
int64_t ob_decoder_launches(const ObDecoder *d) { return d ? d->launches : 0; }
void *ob_decoder_cuda_stream(ObDecoder *d) { return d ? (void *)d->stream : nullptr; }
int32_t ob_decoder_kernel_ms(ObDecoder *d, float ms[3])
{
    if (!d || !ms || !d->timed) return OB_BAD_ARG;
    OB_CUDA(cudaSetDevice(d->device));
    OB_CUDA(cudaEventSynchronize(d->ev[3]));
    for (int i = 0; i < 3; i++) OB_CUDA(cudaEventElapsedTime(&ms[i], d->ev[i], d->ev[i + 1]));
    return OB_OK;
}

// ---- opus_pcm_soft_clip for a batch (soft_clip, src/packet.rs:123-155; opus/src/opus.c:39-144): a warp per stream, one lane per channel for the
// serial part.  pcm / softclip_mem: host, in place. ----
}  // extern "C"
__global__ void ob_k_soft_clip(float *pcm, float *mem, int S, int n, int C)
{
    const int w = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= S) return;
    float *x = pcm + (size_t)w * n * C;
    for (int t = lane; t < n * C; t += 32) x[t] = fmaxf(-2.f, fminf(2.f, x[t]));
    __syncwarp();
    if (lane < C) mem[(size_t)w * C + lane] = ob_soft_clip_channel(x + lane, n, C, mem[(size_t)w * C + lane]);
}
extern "C" {
int32_t ob_pcm_soft_clip_batch(int32_t device, int32_t n_streams, float *pcm, int32_t frame_size, int32_t channels, float *softclip_mem)
{
    if (n_streams <= 0 || frame_size <= 0 || channels <= 0 || channels > 32 || !pcm || !softclip_mem) return OB_BAD_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return OB_INTERNAL_ERROR;       // no CPU fallback
    if (device < 0 || device >= ndev || cudaSetDevice(device) != cudaSuccess) return OB_BAD_ARG;
    const size_t nb = (size_t)n_streams * frame_size * channels * sizeof(float), mb = (size_t)n_streams * channels * sizeof(float);
    float *d_x = nullptr, *d_m = nullptr;
    int32_t rc = OB_ALLOC_FAIL;
    if (cudaMalloc(&d_x, nb) == cudaSuccess && cudaMalloc(&d_m, mb) == cudaSuccess) {
        rc = OB_INTERNAL_ERROR;
        if (cudaMemcpy(d_x, pcm, nb, cudaMemcpyHostToDevice) == cudaSuccess && cudaMemcpy(d_m, softclip_mem, mb, cudaMemcpyHostToDevice) == cudaSuccess) {
            ob_k_soft_clip<<<(n_streams + 3) / 4, 128>>>(d_x, d_m, n_streams, frame_size, channels);
            if (cudaGetLastError() == cudaSuccess && cudaMemcpy(pcm, d_x, nb, cudaMemcpyDeviceToHost) == cudaSuccess &&
                cudaMemcpy(softclip_mem, d_m, mb, cudaMemcpyDeviceToHost) == cudaSuccess) rc = OB_OK;
        }
    }
    cudaFree(d_x); cudaFree(d_m);
    return rc;
}

// ---- TOC helpers (opus/src/opus_decoder.c:1083-1129, opus/src/opus.c:173-192) ----
int32_t ob_packet_get_nb_channels(const uint8_t *p) { return (p[0] & 0x4) ? 2 : 1; }
int32_t ob_packet_get_samples_per_frame(const uint8_t *p, int32_t fs)
{
    if (p[0] & 0x80) return (fs << ((p[0] >> 3) & 0x3)) / 400;
    if ((p[0] & 0x60) == 0x60) return (p[0] & 0x08) ? fs / 50 : fs / 100;
    const int a = (p[0] >> 3) & 0x3;
    return a == 3 ? fs * 60 / 1000 : (fs << a) / 100;
}
int32_t ob_packet_get_bandwidth(const uint8_t *p)
{
    int bw;
    if (p[0] & 0x80) { bw = 1102 + ((p[0] >> 5) & 0x3); if (bw == 1102) bw = 1101; }
    else if ((p[0] & 0x60) == 0x60) bw = (p[0] & 0x10) ? 1105 : 1104;
    else bw = 1101 + ((p[0] >> 5) & 0x3);
    return bw;
}
int32_t ob_packet_get_nb_frames(const uint8_t *p, int32_t len)
{
    if (len < 1) return OB_BAD_ARG;
    const int c = p[0] & 0x3;
    if (c == 0) return 1;
    if (c != 3) return 2;
    if (len < 2) return OB_INVALID_PACKET;
    return p[1] & 0x3F;
}

const char *ob_version(void) { return "1.5.2-b200.7"; }
const char *ob_strerror(int32_t e)
{
    static const char *const s[8] = {"success", "invalid argument", "buffer too small", "internal error", "corrupted stream",
                                     "request not implemented", "invalid state", "memory allocation failed"};
    return (e > 0 || e < -7) ? "unknown error" : s[-e];
}

}  // extern "C"
