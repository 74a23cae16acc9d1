// opus_b200_enc.cu -- batched CELT-only Opus ENCODER for sm_100a: kernel + C ABI (include/opus_b200.h, ob_encoder_*).
//
// GPU mapping: ONE WARP PER STREAM (enc_frame.cuh, enc_analysis.cuh, enc_quant.cuh).  The warp runs the stream's frames in order --
// every encoder decision feeds the range coder and the inter-frame float state, so there is no stateless pass to split off as in
// the decoder -- with warp-uniform control flow: scalars and the coder registers are replicated in every lane, vectors are strided
// over the lanes, cross-lane work goes through ob_coop.cuh (reductions, arg-max, recurrence scans).  Resident warps pull streams
// from a work counter, so the long work vectors need one slot per RESIDENT warp (they stay in L2) and not one per stream.
// This translation unit is compiled with -fmad=false: the float decision heuristics then round like the host emulation of the same
// source (tests/host_emul), which is what the packets are checked against.  Streams are independent: no collective, one ObEncoder
// per GPU over disjoint stream ranges.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <new>

#include "../../include/opus_b200.h"
#include "enc_frame.cuh"

struct ObEncStream {           // the scalar state one stream owns on the device (its vectors: ObEncHist)
    ObOpusEncState os;
    ObEncState st;
};

#define OB_ENC_THREADS 32             // ob_k_analysis: one thread per stream
#ifndef OB_ENC_BLOCKS
#define OB_ENC_BLOCKS 1                // resident blocks per SM the register allocation aims at (OB_ENC_WARPS x OB_ENC_BLOCKS x 32 x regs <= 64 K)
#endif
#ifndef OB_ENC_WARPS
#define OB_ENC_WARPS 16               // ob_k_encode: warps (= streams in flight) per block; 16 x 13.2 KB of shared memory = one block per SM
#endif
struct ObEncBlockShared {
    int progress[32];                 // ObWarpPaced: how far each warp of the block has come
    int base_stream, pad[31];
    ObEncShared sh[OB_ENC_WARPS];     // as many as the launch has warps (dynamic shared memory: 256 bytes + warps x sizeof(ObEncShared))
};
__global__ void __launch_bounds__(32 * OB_ENC_WARPS, OB_ENC_BLOCKS)
ob_k_encode(const float *__restrict__ pcm, uint8_t *__restrict__ out, int32_t *__restrict__ lens, uint32_t *__restrict__ ranges,
            ObEncStream *__restrict__ streams, ObEncHist *__restrict__ hist, ObEncWork *__restrict__ work, int *__restrict__ counter,
            const ObAnalysisInfo *__restrict__ info, ObTonalState *__restrict__ tonal, float *__restrict__ delay,
            ObOpusEncCfg cfg, int S, int F, int frame_size, int max_bytes, int f0, int Fc, int paced, int s_lo)
{
    extern __shared__ __align__(16) unsigned char ob_enc_smem[];
    ObEncBlockShared &bs = *reinterpret_cast<ObEncBlockShared *>(ob_enc_smem);
    const int nw = (int)(blockDim.x >> 5), w = (int)(threadIdx.x >> 5);
    const ObWarpPaced g(bs.progress, nw, paced & 15, paced >> 4);
    ObEncShared &sh = bs.sh[w];
    ObEncWork &wk = work[blockIdx.x * nw + w];
    for (;;) {
        // the block pulls nw consecutive streams, one per warp; its warps then walk the frames side by side (ObWarpPaced)
        __syncthreads();
        if (threadIdx.x == 0) bs.base_stream = atomicAdd(counter, nw);
        if (g.lane == 0) bs.progress[w] = 0;
        __syncthreads();
        const int s = s_lo + bs.base_stream + w;                 // this launch codes the streams [s_lo, S)
        if (s_lo + bs.base_stream >= S) break;
        if (s >= S) { g.publish(0x7fffffff); continue; }          // a tail block: this warp has nothing to do and must hold nobody up
        ObEncStream es = streams[s];
        es.os.delay = delay ? delay + (size_t)s * OB_ENC_BUFFER * es.st.channels : nullptr;      // AUDIO / VOIP: the 4 ms delay compensation, state in global memory
        es.os.tonal = tonal ? tonal + s : nullptr;                 // packets longer than 20 ms run the analysis inline (lane 0), on the state in global memory
        const int CC = es.st.channels;
        ob_enc_load_hist(g, sh, wk, hist[s]);
        for (int f = f0; f < f0 + Fc; f++) {                       // this launch covers the frame window [f0, f0+Fc) of a [S][F] batch
            const size_t wi = (size_t)s * F + f;
            ObAnalysisInfo an;                                      // computed ahead of this kernel by ob_k_analysis (complexity >= 7)
            if (info) an = info[wi];
            const int n = ob_opus_encode(g, cfg, es.os, es.st, sh, wk, pcm + wi * (size_t)frame_size * CC, frame_size, out + wi * (size_t)max_bytes, max_bytes,
                                         info ? &an : nullptr, (f - f0 + 1) * 16384);
            if (g.lane == 0) {
                lens[wi] = n;
                if (ranges) ranges[wi] = n > 0 ? es.st.final_range : 0;
            }
        }
        g.publish(0x7fffffff);                                      // done with this stream: wait for nobody, hold nobody up
        ob_enc_store_hist(g, sh, wk, hist[s]);
        if (g.lane == 0) streams[s] = es;
    }
}

// The SAME encoder source with ONE LANE PER STREAM (G = ObSolo): 32 streams share every instruction a warp issues, which is the better
// trade for bulk batches -- the encoder's control flow is ~45 k instructions of mostly scalar code, and a warp-per-stream mapping pays
// one instruction stream per stream for it (measured on B200, 16 384 stereo complexity-10 streams: 48 ms per frame step warp-per-stream,
// ~32 ms thread-per-stream; at <= 2368 streams the order reverses, 7 ms against 25 ms, because one lane per stream leaves the SMs with
// 3 warps each).  The per-stream working set lives in LOCAL memory here: CUDA interleaves it across the lanes of a warp, so lanes that
// walk their arrays in step coalesce.  Packets are those of the reference's summation order (bit-identical to its C build).
__global__ void __launch_bounds__(32)
ob_k_encode_thread(const float *__restrict__ pcm, uint8_t *__restrict__ out, int32_t *__restrict__ lens, uint32_t *__restrict__ ranges,
                   ObEncStream *__restrict__ streams, ObEncHist *__restrict__ hist,
                   const ObAnalysisInfo *__restrict__ info, ObTonalState *__restrict__ tonal, float *__restrict__ delay,
                   ObOpusEncCfg cfg, int S, int F, int frame_size, int max_bytes, int f0, int Fc, int s_lo)
{
    const int s = s_lo + blockIdx.x * 32 + threadIdx.x;          // this launch codes the streams [s_lo, S)
    if (s >= S) return;
    const ObSolo g;
    ObEncShared sh;       // not initialised: no stage reads what it (or an earlier frame) has not written (tests/test_host_emul.py poisons it)
    ObEncWork wk;
    ObEncStream es = streams[s];
    es.os.delay = delay ? delay + (size_t)s * OB_ENC_BUFFER * es.st.channels : nullptr;
    es.os.tonal = tonal ? tonal + s : nullptr;
    const int CC = es.st.channels;
    ob_enc_load_hist(g, sh, wk, hist[s]);
    // NOTE: the frame counter is deliberately volatile.  The loop body is one huge divergent region; with a plain `int f` nvcc 12.9 keeps f
    // in a UNIFORM register, and lanes that fall behind re-execute the shared increment (found on B200 in round 1).
    for (volatile int f = f0; f < f0 + Fc; f++) {
        const size_t wi = (size_t)s * F + f;
        ObAnalysisInfo an;
        if (info) an = info[wi];
        const int n = ob_opus_encode(g, cfg, es.os, es.st, sh, wk, pcm + wi * (size_t)frame_size * CC, frame_size, out + wi * (size_t)max_bytes, max_bytes,
                                     info ? &an : nullptr, 0);
        lens[wi] = n;
        if (ranges) ranges[wi] = n > 0 ? es.st.final_range : 0;
    }
    ob_enc_store_hist(g, sh, wk, hist[s]);
    streams[s] = es;
}

// The Opus-layer signal analysis (enc_tonal.cuh) depends on nothing but the input PCM and its own state, so it does not have to sit
// inside the encoder's serial chain: it runs in this kernel, on its own CUDA stream, one frame window ahead of ob_k_encode.  Both
// kernels are one thread per stream and latency bound at ~3.4 warps per SM; running them side by side doubles the warps in
// flight and hides the analysis (16 % of the fused kernel's time) almost completely.
__global__ void __launch_bounds__(OB_ENC_THREADS)
ob_k_analysis(const float *__restrict__ pcm, ObTonalState *__restrict__ tonal, ObAnalysisInfo *__restrict__ info, int S, int F, int frame_size,
              int channels, int lsb_depth, int f0, int Fc, int s_lo)
{
    const int s = s_lo + blockIdx.x * blockDim.x + threadIdx.x;   // this launch covers the streams [s_lo, S)
    if (s >= S) return;
    float work[3920];                                           // lane-interleaved local memory, like the encoder's scratch: 2720 work + 1200 FFT / resampler
    const ObSolo g;
    ObTonalState t = tonal[s];
    for (volatile int f = f0; f < f0 + Fc; f++) {
        const size_t w = (size_t)s * F + f;
        ObAnalysisInfo a;
        a.valid = 0;
        ob_run_analysis(g, t, pcm + w * (size_t)frame_size * channels, frame_size, channels, lsb_depth, a, work, work + 2720);
        info[w] = a;
    }
    tonal[s] = t;
}
__global__ void ob_k_tonal_reset(ObTonalState *tonal, const int32_t *idx, int n, int S)
{
    const int k = blockIdx.x;
    if (k >= n) return;
    const int s = idx ? idx[k] : k;
    if (s < 0 || s >= S) return;
    uint32_t *z = reinterpret_cast<uint32_t *>(tonal + s);
    for (int i = threadIdx.x; i < (int)(sizeof(ObTonalState) / 4); i += blockDim.x) z[i] = 0;
}

// int16 API (opus_encode, opus_encoder.c:2346-2376, float build): in[i] = (1/32768) * pcm[i], then the float path at 16-bit depth.
__global__ void ob_k_i16_to_f32(const int16_t *__restrict__ in, float *__restrict__ out, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (1.0f / 32768) * (float)in[i];
}

__global__ void ob_k_enc_reset(ObEncStream *streams, ObEncHist *hist, float *delay, const int32_t *idx, int n, int S, int channels)
{
    const int k = blockIdx.x;                                      // one block per stream to reset
    if (k >= n) return;
    const int s = idx ? idx[k] : k;
    if (s < 0 || s >= S) return;
    if (threadIdx.x == 0) {
        ObEncStream &es = streams[s];
        es.st.channels = es.st.stream_channels = channels; es.st.end = 21; es.st.clip = 1; es.st.force_intra = 0; es.st.disable_inv = 0; es.st.disable_pf = 0;
        ob_enc_reset(es.st);
        es.os.stream_channels = channels; es.os.first = 1; es.os.auto_bandwidth = 0; es.os.bandwidth = 1105; es.os.hybrid_stereo_width_Q14 = 1 << 14;
        es.os.voice_ratio = -1; es.os.detected_bandwidth = 0; es.os.tonal = nullptr;
        es.os.prev_mode = 0; es.os.width_mem = ObStereoWidth{0, 0, 0, 0, 0}; es.os.delay = nullptr;
        es.os.nb_no_activity_ms_Q1 = 0; es.os.peak_signal_energy = 0;
    }
    ObEncHist &h = hist[s];
    for (int i = threadIdx.x; i < 2 * OB_OVERLAP; i += blockDim.x) h.in_mem[i] = 0;
    for (int i = threadIdx.x; i < 2 * OB_MAXPERIOD; i += blockDim.x) h.prefilter_mem[i] = 0;
    for (int i = threadIdx.x; i < 2 * OB_NB; i += blockDim.x) { h.oldBandE[i] = 0; h.oldLogE[i] = h.oldLogE2[i] = -28.f; h.energyError[i] = 0; }
    if (delay) for (int i = threadIdx.x; i < OB_ENC_BUFFER * channels; i += blockDim.x) delay[(size_t)s * OB_ENC_BUFFER * channels + i] = 0;
}

__global__ void ob_k_enc_gather(const ObEncStream *streams, uint32_t *ranges, int S)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < S) ranges[s] = streams[s].st.final_range;
}

__global__ void ob_k_enc_gather_dtx(const ObEncStream *streams, uint32_t *out, int S, int what)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    if (what == 0) out[s] = streams[s].os.nb_no_activity_ms_Q1 >= 10 * 20 * 2;       // OPUS_GET_IN_DTX, "DTX determined by Opus" (opus_encoder.c:3038-3042)
    else out[s] = (uint32_t)streams[s].os.bandwidth;                                  // OPUS_GET_BANDWIDTH: st->bandwidth of the last packet (:2700-2708)
}

struct ObEncoder {
    int mapping;                       // OB_ENC_MAP_*: which instantiation of the encoder source codes the batch
    int S, CC, device, max_frames, slots;      // slots: resident blocks of ob_k_encode (OB_ENC_WARPS streams each)
    ObOpusEncCfg cfg;
    cudaStream_t stream, copy_stream, an_stream, split_stream;   // split_stream: the warp-per-stream part of a split batch
    cudaEvent_t split_go, split_done;
    int n_warp;                        // streams [0, n_warp) are coded one warp per stream, [n_warp, S) one lane per stream (ob_enc_launch)
    cudaEvent_t ev[2], win_ev[4], an_ev[4], enc_done;
    bool timed, tonal_dirty;
    ObEncStream *d_streams;
    ObEncHist *d_hist;                 // [S] the streams' vectors (pre-filter memory, overlap, band energies)
    ObEncWork *d_work;                 // [slots] long work vectors, one per RESIDENT warp of ob_k_encode
    int *d_counter;                    // work counter the resident warps pull stream indices from
    ObTonalState *d_tonal;             // [S] state of the signal analysis (complexity >= 7)
    float *d_delay;                    // [S][OB_ENC_BUFFER * channels] delay buffers of the AUDIO / VOIP applications (null for RESTRICTED_LOWDELAY)
    ObAnalysisInfo *d_info;            // [S][max_frames] its per-frame result, consumed by ob_k_encode
    float *d_pcm; size_t pcm_cap;
    int16_t *d_pcm16; size_t pcm16_cap;
    uint8_t *d_out; size_t out_cap;
    int32_t *d_lens; uint32_t *d_ranges;
    int64_t launches;
};

#define OB_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "opus_b200: %s failed: %s\n", #x, cudaGetErrorString(e_)); return OB_INTERNAL_ERROR; } } while (0)

extern "C" {

ObEncoder *ob_encoder_create(int32_t n_streams, int32_t fs, int32_t channels, int32_t application, int32_t device, int32_t max_frames, int32_t *error)
{
    int err = OB_OK, ndev = 0;
    ObEncoder *e = nullptr;
    if (n_streams <= 0 || (channels != 1 && channels != 2) || max_frames <= 0) err = OB_BAD_ARG;
    else if (application != 2048 && application != 2049 && application != 2051) err = OB_BAD_ARG;
    else if (fs != 48000) err = OB_UNIMPLEMENTED;
    else if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        fprintf(stderr, "opus_b200: no usable CUDA device (count=%d, requested=%d); there is no CPU fallback\n", ndev, device);
        err = OB_INTERNAL_ERROR;
    }
    if (err == OB_OK) { e = new (std::nothrow) ObEncoder(); if (!e) err = OB_ALLOC_FAIL; }
    if (err == OB_OK) {
        memset(e, 0, sizeof(*e));
        e->S = n_streams; e->CC = channels; e->device = device; e->max_frames = max_frames;
        // defaults of opus_encoder_init (opus_encoder.c:202-297): VBR on, constrained, bitrate AUTO, complexity 9, 24-bit depth
        e->cfg.bitrate = -1000; e->cfg.complexity = 9; e->cfg.vbr = 1; e->cfg.vbr_constraint = 1; e->cfg.max_bandwidth = 1105;
        e->cfg.user_bandwidth = 0; e->cfg.force_channels = 0; e->cfg.packet_loss = 0; e->cfg.lsb_depth = 24; e->cfg.application = application;
        const size_t total = (size_t)n_streams * max_frames;
        bool ok = cudaSetDevice(device) == cudaSuccess;
        ok = ok && cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaEventCreate(&e->ev[0]) == cudaSuccess && cudaEventCreate(&e->ev[1]) == cudaSuccess;
        ok = ok && cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
        for (int i = 0; i < 4 && ok; i++) ok = cudaEventCreateWithFlags(&e->win_ev[i], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaMalloc(&e->d_streams, sizeof(ObEncStream) * n_streams) == cudaSuccess;
        ok = ok && cudaMalloc(&e->d_hist, sizeof(ObEncHist) * n_streams) == cudaSuccess;
        if (ok) {   // one work slot per warp that can be resident at once (a persistent grid; streams are pulled from d_counter)
            int per_sm = 0, sms = 0;
            ok = cudaFuncSetAttribute(ob_k_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ObEncBlockShared)) == cudaSuccess
                 && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ob_k_encode, 32 * OB_ENC_WARPS, sizeof(ObEncBlockShared)) == cudaSuccess
                 && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && per_sm > 0 && sms > 0;
            e->slots = per_sm * sms;                               // resident blocks
            const int need = (n_streams + OB_ENC_WARPS - 1) / OB_ENC_WARPS;
            if (e->slots > need) e->slots = need;
        }
        ok = ok && cudaMalloc(&e->d_work, sizeof(ObEncWork) * e->slots * OB_ENC_WARPS) == cudaSuccess;
        ok = ok && cudaMalloc(&e->d_counter, sizeof(int)) == cudaSuccess;
        ok = ok && cudaStreamCreateWithFlags(&e->split_stream, cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&e->split_go, cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&e->split_done, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaStreamCreateWithFlags(&e->an_stream, cudaStreamNonBlocking) == cudaSuccess;
        for (int i = 0; i < 4 && ok; i++) ok = cudaEventCreateWithFlags(&e->an_ev[i], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&e->enc_done, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaMalloc(&e->d_tonal, sizeof(ObTonalState) * n_streams) == cudaSuccess;
        ok = ok && cudaMemset(e->d_tonal, 0, sizeof(ObTonalState) * n_streams) == cudaSuccess;
        ok = ok && cudaMalloc(&e->d_info, sizeof(ObAnalysisInfo) * total) == cudaSuccess;
        if (application != 2051) ok = ok && cudaMalloc(&e->d_delay, sizeof(float) * OB_ENC_BUFFER * channels * n_streams) == cudaSuccess;
        ok = ok && cudaMalloc(&e->d_lens, sizeof(int32_t) * total) == cudaSuccess;
        ok = ok && cudaMalloc(&e->d_ranges, sizeof(uint32_t) * total) == cudaSuccess;
        if (!ok) {
            fprintf(stderr, "opus_b200: device allocation failed: %s\n", cudaGetErrorString(cudaGetLastError()));
            ob_encoder_destroy(e); e = nullptr; err = OB_ALLOC_FAIL;
        } else if (ob_encoder_reset(e, nullptr, 0) != OB_OK) { ob_encoder_destroy(e); e = nullptr; err = OB_INTERNAL_ERROR; }
    }
    if (error) *error = err;
    return e;
}

void ob_encoder_destroy(ObEncoder *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->copy_stream) cudaStreamSynchronize(e->copy_stream);
    if (e->an_stream) cudaStreamSynchronize(e->an_stream);
    if (e->split_stream) cudaStreamSynchronize(e->split_stream);
    if (e->stream) cudaStreamSynchronize(e->stream);
    cudaFree(e->d_streams); cudaFree(e->d_hist); cudaFree(e->d_work); cudaFree(e->d_counter); cudaFree(e->d_pcm); cudaFree(e->d_pcm16); cudaFree(e->d_out); cudaFree(e->d_lens); cudaFree(e->d_ranges);
    for (int i = 0; i < 2; i++) if (e->ev[i]) cudaEventDestroy(e->ev[i]);
    for (int i = 0; i < 4; i++) { if (e->win_ev[i]) cudaEventDestroy(e->win_ev[i]); if (e->an_ev[i]) cudaEventDestroy(e->an_ev[i]); }
    if (e->enc_done) cudaEventDestroy(e->enc_done);
    if (e->an_stream) cudaStreamDestroy(e->an_stream);
    if (e->split_stream) cudaStreamDestroy(e->split_stream);
    if (e->split_go) cudaEventDestroy(e->split_go);
    if (e->split_done) cudaEventDestroy(e->split_done);
    cudaFree(e->d_tonal); cudaFree(e->d_info); cudaFree(e->d_delay);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

int32_t ob_encoder_reset(ObEncoder *e, const int32_t *idx, int32_t n)
{
    if (!e || n < 0) return OB_BAD_ARG;
    OB_CUDA(cudaSetDevice(e->device));
    // Legal with calls in flight: the reset is ordered on e->stream behind every encode kernel, and waits for the analysis stream's last kernel
    if (e->an_stream) OB_CUDA(cudaStreamSynchronize(e->an_stream));
    int32_t *d_idx = nullptr;
    int count = e->S;
    if (idx) {
        if (n == 0) return OB_OK;
        count = n;
        OB_CUDA(cudaMalloc(&d_idx, sizeof(int32_t) * n));
    }
    cudaError_t err = d_idx ? cudaMemcpyAsync(d_idx, idx, sizeof(int32_t) * n, cudaMemcpyHostToDevice, e->stream) : cudaSuccess;
    if (err == cudaSuccess) {
        ob_k_enc_reset<<<count, 128, 0, e->stream>>>(e->d_streams, e->d_hist, e->d_delay, d_idx, count, e->S, e->CC);
        ob_k_tonal_reset<<<count, 128, 0, e->stream>>>(e->d_tonal, d_idx, count, e->S);
        e->launches += 2;
        err = cudaStreamSynchronize(e->stream);
    }
    if (d_idx) cudaFree(d_idx);                                    // on every exit path
    if (err != cudaSuccess) { fprintf(stderr, "opus_b200: encoder reset failed: %s\n", cudaGetErrorString(err)); return OB_INTERNAL_ERROR; }
    return OB_OK;
}

// CTLs (src/encoder.rs:545-652): one value for the whole batch.
int32_t ob_encoder_set_bitrate(ObEncoder *e, int32_t bitrate)
{
    if (!e) return OB_BAD_ARG;
    if (bitrate != -1000 && bitrate != -1) {                   // opus_encoder.c OPUS_SET_BITRATE_REQUEST: clamp to [500, 300000*channels]
        if (bitrate <= 0) return OB_BAD_ARG;
        if (bitrate <= 500) bitrate = 500;
        else if (bitrate > 300000 * e->CC) bitrate = 300000 * e->CC;
    }
    e->cfg.bitrate = bitrate;
    return OB_OK;
}
int32_t ob_encoder_get_bitrate(ObEncoder *e, int32_t *v) { if (!e || !v) return OB_BAD_ARG; *v = e->cfg.bitrate; return OB_OK; }
// which instantiation of the encoder source codes the batch (see ob_k_encode / ob_k_encode_thread)
int32_t ob_encoder_set_mapping(ObEncoder *e, int32_t m) { if (!e || m < OB_ENC_MAP_AUTO || m > OB_ENC_MAP_THREAD) return OB_BAD_ARG; e->mapping = m; return OB_OK; }
int32_t ob_encoder_get_mapping(ObEncoder *e, int32_t *v) { if (!e || !v) return OB_BAD_ARG; *v = e->mapping; return OB_OK; }
// how the last call was laid out: streams [0, *n_warp) one warp per stream, the rest one lane per stream
int32_t ob_encoder_get_split(ObEncoder *e, int32_t *n_warp) { if (!e || !n_warp) return OB_BAD_ARG; *n_warp = e->n_warp; return OB_OK; }
int32_t ob_encoder_set_complexity(ObEncoder *e, int32_t c) { if (!e || c < 0 || c > 10) return OB_BAD_ARG; e->cfg.complexity = c; return OB_OK; }
int32_t ob_encoder_get_complexity(ObEncoder *e, int32_t *v) { if (!e || !v) return OB_BAD_ARG; *v = e->cfg.complexity; return OB_OK; }
int32_t ob_encoder_set_vbr(ObEncoder *e, int32_t v) { if (!e || v < 0 || v > 1) return OB_BAD_ARG; e->cfg.vbr = v; return OB_OK; }
int32_t ob_encoder_get_vbr(ObEncoder *e, int32_t *v) { if (!e || !v) return OB_BAD_ARG; *v = e->cfg.vbr; return OB_OK; }
int32_t ob_encoder_set_vbr_constraint(ObEncoder *e, int32_t v) { if (!e || v < 0 || v > 1) return OB_BAD_ARG; e->cfg.vbr_constraint = v; return OB_OK; }
int32_t ob_encoder_get_vbr_constraint(ObEncoder *e, int32_t *v) { if (!e || !v) return OB_BAD_ARG; *v = e->cfg.vbr_constraint; return OB_OK; }
int32_t ob_encoder_set_max_bandwidth(ObEncoder *e, int32_t bw) { if (!e || bw < 1101 || bw > 1105) return OB_BAD_ARG; e->cfg.max_bandwidth = bw; return OB_OK; }
int32_t ob_encoder_set_bandwidth(ObEncoder *e, int32_t bw) { if (!e || (bw != -1000 && (bw < 1101 || bw > 1105))) return OB_BAD_ARG; e->cfg.user_bandwidth = bw == -1000 ? 0 : bw; return OB_OK; }
int32_t ob_encoder_set_force_channels(ObEncoder *e, int32_t ch) { if (!e || (ch != -1000 && (ch < 1 || ch > e->CC))) return OB_BAD_ARG; e->cfg.force_channels = ch == -1000 ? 0 : ch; return OB_OK; }
int32_t ob_encoder_set_packet_loss_perc(ObEncoder *e, int32_t p) { if (!e || p < 0 || p > 100) return OB_BAD_ARG; e->cfg.packet_loss = p; return OB_OK; }
int32_t ob_encoder_set_lsb_depth(ObEncoder *e, int32_t d) { if (!e || d < 8 || d > 24) return OB_BAD_ARG; e->cfg.lsb_depth = d; return OB_OK; }
int32_t ob_encoder_get_max_bandwidth(ObEncoder *e, int32_t *v) { if (!e || !v) return OB_BAD_ARG; *v = e->cfg.max_bandwidth; return OB_OK; }
int32_t ob_encoder_get_force_channels(ObEncoder *e, int32_t *v) { if (!e || !v) return OB_BAD_ARG; *v = e->cfg.force_channels ? e->cfg.force_channels : -1000; return OB_OK; }
int32_t ob_encoder_get_packet_loss_perc(ObEncoder *e, int32_t *v) { if (!e || !v) return OB_BAD_ARG; *v = e->cfg.packet_loss; return OB_OK; }
int32_t ob_encoder_get_lsb_depth(ObEncoder *e, int32_t *v) { if (!e || !v) return OB_BAD_ARG; *v = e->cfg.lsb_depth; return OB_OK; }
// OPUS_SET_SIGNAL / _PREDICTION_DISABLED / _PHASE_INVERSION_DISABLED / _DTX / _INBAND_FEC / _EXPERT_FRAME_DURATION (opus_encoder.c:2815-2940)
int32_t ob_encoder_set_signal(ObEncoder *e, int32_t v) { if (!e || (v != -1000 && v != 3001 && v != 3002)) return OB_BAD_ARG; e->cfg.signal_type = v == -1000 ? 0 : v; return OB_OK; }
int32_t ob_encoder_get_signal(ObEncoder *e, int32_t *v) { if (!e || !v) return OB_BAD_ARG; *v = e->cfg.signal_type ? e->cfg.signal_type : -1000; return OB_OK; }
int32_t ob_encoder_set_prediction_disabled(ObEncoder *e, int32_t v) { if (!e || v < 0 || v > 1) return OB_BAD_ARG; e->cfg.prediction_disabled = v; return OB_OK; }
int32_t ob_encoder_get_prediction_disabled(ObEncoder *e, int32_t *v) { if (!e || !v) return OB_BAD_ARG; *v = e->cfg.prediction_disabled; return OB_OK; }
int32_t ob_encoder_set_phase_inversion_disabled(ObEncoder *e, int32_t v) { if (!e || v < 0 || v > 1) return OB_BAD_ARG; e->cfg.phase_inversion_disabled = v; return OB_OK; }
int32_t ob_encoder_get_phase_inversion_disabled(ObEncoder *e, int32_t *v) { if (!e || !v) return OB_BAD_ARG; *v = e->cfg.phase_inversion_disabled; return OB_OK; }
int32_t ob_encoder_set_dtx(ObEncoder *e, int32_t v) { if (!e || v < 0 || v > 1) return OB_BAD_ARG; e->cfg.use_dtx = v; return OB_OK; }
int32_t ob_encoder_get_dtx(ObEncoder *e, int32_t *v) { if (!e || !v) return OB_BAD_ARG; *v = e->cfg.use_dtx; return OB_OK; }
int32_t ob_encoder_set_inband_fec(ObEncoder *e, int32_t v) { if (!e || v < 0 || v > 2) return OB_BAD_ARG; e->cfg.inband_fec = v; return OB_OK; }
int32_t ob_encoder_get_inband_fec(ObEncoder *e, int32_t *v) { if (!e || !v) return OB_BAD_ARG; *v = e->cfg.inband_fec; return OB_OK; }
int32_t ob_encoder_set_expert_frame_duration(ObEncoder *e, int32_t v) { if (!e || v < 5000 || v > 5009) return OB_BAD_ARG; e->cfg.variable_duration = v; return OB_OK; }
int32_t ob_encoder_get_expert_frame_duration(ObEncoder *e, int32_t *v) { if (!e || !v) return OB_BAD_ARG; *v = e->cfg.variable_duration ? e->cfg.variable_duration : 5000; return OB_OK; }
// OPUS_GET_LOOKAHEAD (:2835-2846)
int32_t ob_encoder_get_lookahead(ObEncoder *e, int32_t *v) { if (!e || !v) return OB_BAD_ARG; *v = 48000 / 400 + (e->cfg.application != 2051 ? OB_ENC_DELAY : 0); return OB_OK; }

// frame_size_select (opus_encoder.c:704-728)
static int ob_frame_size_select(int frame_size, int variable_duration, int Fs)
{
    int new_size;
    if (frame_size < Fs / 400) return -1;
    if (variable_duration == 0 || variable_duration == 5000) new_size = frame_size;
    else if (variable_duration >= 5001 && variable_duration <= 5009) new_size = variable_duration <= 5005 ? (Fs / 400) << (variable_duration - 5001) : (variable_duration - 5001 - 2) * Fs / 50;
    else return -1;
    if (new_size > frame_size) return -1;
    if (400 * new_size != Fs && 200 * new_size != Fs && 100 * new_size != Fs && 50 * new_size != Fs && 25 * new_size != Fs && 50 * new_size != 3 * Fs &&
        50 * new_size != 4 * Fs && 50 * new_size != 5 * Fs && 50 * new_size != 6 * Fs) return -1;
    return new_size;
}

// One frame window [f0, f0+Fc) of every stream.  At complexity >= 7 the analysis kernel of the window runs on its own stream, right
// away (it needs only the PCM), and the encode kernel of the window waits for it; while the encoder works on window k the analysis
// of window k+1 is already running beside it.  analysis_ahead: the caller has already enqueued the analysis of this window.
// Which streams go where (see ob_enc_launch): streams [0, n_warp) one warp per stream, the rest one lane per stream.
static int ob_enc_plan(ObEncoder *e)
{
    int n_warp = e->mapping == OB_ENC_MAP_WARP ? e->S : 0;
    if (e->mapping == OB_ENC_MAP_AUTO) {
        if (e->S < OB_ENC_MAP_CROSSOVER) n_warp = e->S;
        else {
            double frac = 0.0;
            if (const char *v = getenv("OB_ENC_SPLIT")) frac = atof(v);                                     // tuning aid (see ob_enc_launch)
            n_warp = (int)(e->S * frac) / 12 * 12;
        }
    }
    e->n_warp = n_warp;
    return n_warp;
}

static int ob_enc_analysis(ObEncoder *e, int F, const float *d_pcm, int frame_size, int f0, int Fc, int k, cudaEvent_t pcm_ready)
{
    if (e->cfg.complexity < 7 || frame_size > 960) return OB_OK;   // long packets interleave analysis reads with their 20 ms frames: done inline
    const int n_warp = ob_enc_plan(e);
    e->tonal_dirty = true;
    if (n_warp >= e->S) return OB_OK;                              // the warp-per-stream kernel runs the analysis itself (cooperatively, in front of every frame)
    if (pcm_ready) OB_CUDA(cudaStreamWaitEvent(e->an_stream, pcm_ready, 0));
    if (f0 == 0) OB_CUDA(cudaStreamWaitEvent(e->an_stream, e->enc_done, 0));       // d_info of the previous call has been consumed
    ob_k_analysis<<<(e->S - n_warp + OB_ENC_THREADS - 1) / OB_ENC_THREADS, OB_ENC_THREADS, 0, e->an_stream>>>(d_pcm, e->d_tonal, e->d_info, e->S, F, frame_size, e->CC,
                                                                                                              e->cfg.lsb_depth, f0, Fc, n_warp);
    OB_CUDA(cudaEventRecord(e->an_ev[k & 3], e->an_stream));
    e->launches += 1;
    return OB_OK;
}

static int ob_enc_launch(ObEncoder *e, int F, const float *d_pcm, int frame_size, uint8_t *d_out, int max_bytes, int32_t *d_lens, uint32_t *d_ranges,
                         int f0 = 0, int Fc = -1, int k = 0)
{
    if (Fc < 0) Fc = F;
    const bool an = e->cfg.complexity >= 7 && frame_size <= 960, an_inline = e->cfg.complexity >= 7 && frame_size > 960;
    if (an_inline) e->tonal_dirty = true;
    if (e->cfg.complexity < 7 && e->tonal_dirty) {                               // "else if (st->analysis.initialized) tonality_analysis_reset()" (opus_encoder.c:1131-1133)
        ob_k_tonal_reset<<<e->S, 128, 0, e->stream>>>(e->d_tonal, nullptr, e->S, e->S);
        e->tonal_dirty = false;
        e->launches += 1;
    }
    if (f0 == 0) OB_CUDA(cudaEventRecord(e->ev[0], e->stream));
    if (an && ob_enc_plan(e) < e->S) OB_CUDA(cudaStreamWaitEvent(e->stream, e->an_ev[k & 3], 0));
    OB_CUDA(cudaMemsetAsync(e->d_counter, 0, sizeof(int), e->stream));
    // pace level (0 none, 1 stages, 2 + bands, 3 + leaves) | slack << 4.  Measured, ms per call of 4 frames, 2 368 / 16 384 streams, after the band
    // loop's code was cut to ~11 k instructions: 0: 41.9 / 238.5, 1: 31.4 / 199.3, 2: 23.6 / 161.6, 3: 24.6 / 168.3; a slack of 1-4 ids: no gain.
    // (With the 17 k-instruction band loop of the first version level 3 was the best: 50.7 against 53.7 ms per frame step of 16 384 streams.)
    int paced = 2;
    if (const char *v = getenv("OB_ENC_PACED")) paced = atoi(v);                                            // tuning aid
    // Which streams go where.  WARP / THREAD: all of them.  AUTO: below the crossover one warp per stream (latency), from the crossover up one
    // lane per stream (throughput).  A SPLIT of a bulk batch between the two mappings (first n_warp streams on a second CUDA stream, 12 warps
    // per block so that both kernels fit an SM's registers) was measured on B200 and rejected: 16 384 stereo complexity-10 streams x 4 frames
    // take 143 ms all lane-per-stream, 167 / 183 / 197 ms with 20 / 30 / 40 % of the streams warp-per-stream -- the two kernels do not share an
    // SM (one wants the maximal shared-memory carve-out, the other its L1; forcing the same carve-out on both: 244 ms).  OB_ENC_SPLIT keeps
    // the experiment reproducible.  A stream keeps its mapping for the life of the encoder.
    const int n_warp = ob_enc_plan(e);
    const bool split = n_warp > 0 && n_warp < e->S;
    if (n_warp > 0) {
        cudaStream_t ws = split ? e->split_stream : e->stream;
        const int warps = split ? 12 : OB_ENC_WARPS;
        if (split) { OB_CUDA(cudaEventRecord(e->split_go, e->stream)); OB_CUDA(cudaStreamWaitEvent(ws, e->split_go, 0)); }
        OB_CUDA(cudaMemsetAsync(e->d_counter, 0, sizeof(int), ws));
        int blocks = (n_warp + warps - 1) / warps;
        if (blocks > e->slots) blocks = e->slots;
        ob_k_encode<<<blocks, 32 * warps, sizeof(ObEncShared) * warps + 256, ws>>>(d_pcm, d_out, d_lens, d_ranges, e->d_streams, e->d_hist, e->d_work, e->d_counter,
                nullptr, (an || an_inline) ? e->d_tonal : nullptr, e->d_delay, e->cfg, n_warp, F, frame_size, max_bytes, f0, Fc, paced, 0);
        if (split) OB_CUDA(cudaEventRecord(e->split_done, ws));
    }
    if (n_warp < e->S) {
        ob_k_encode_thread<<<(e->S - n_warp + 31) / 32, 32, 0, e->stream>>>(d_pcm, d_out, d_lens, d_ranges, e->d_streams, e->d_hist, an ? e->d_info : nullptr,
                                                                            an_inline ? e->d_tonal : nullptr, e->d_delay, e->cfg, e->S, F, frame_size, max_bytes, f0, Fc, n_warp);
        if (split) OB_CUDA(cudaStreamWaitEvent(e->stream, e->split_done, 0));
    }
    if (f0 + Fc == F) { OB_CUDA(cudaEventRecord(e->ev[1], e->stream)); OB_CUDA(cudaEventRecord(e->enc_done, e->stream)); }
    OB_CUDA(cudaGetLastError());
    e->launches += 1;
    e->timed = true;
    return OB_OK;
}

int32_t ob_encode_float_device(ObEncoder *e, int32_t n_frames, const float *d_pcm, int32_t frame_size, uint8_t *d_out, int32_t max_bytes,
                               int32_t *d_lens_out, uint32_t *d_ranges_out, int32_t sync)
{
    if (!e || !d_pcm || !d_out || !d_lens_out || n_frames <= 0 || n_frames > e->max_frames || max_bytes <= 0) return OB_BAD_ARG;
    if (e->cfg.variable_duration > 5000) {                          // OPUS_SET_EXPERT_FRAME_DURATION: the duration must be the one the caller's buffers have
        const int sel = ob_frame_size_select(frame_size, e->cfg.variable_duration, 48000);
        if (sel <= 0) return OB_BAD_ARG;
        if (sel != frame_size) return OB_UNIMPLEMENTED;              // libopus would code the first `sel` samples of each longer buffer
    }
    OB_CUDA(cudaSetDevice(e->device));
    // the caller's PCM is ready in stream order of e->stream; with the analysis on (complexity >= 7) the call runs in frame windows so
    // that the analysis of window k+1 overlaps the encoder of window k
    const int nwin = e->cfg.complexity >= 7 ? (n_frames >= 4 ? 4 : (n_frames >= 2 ? 2 : 1)) : 1;
    const int per = (n_frames + nwin - 1) / nwin;
    OB_CUDA(cudaEventRecord(e->win_ev[0], e->stream));
    for (int k = 0, f0 = 0; f0 < n_frames; k++, f0 += per) {
        const int Fc = n_frames - f0 < per ? n_frames - f0 : per;
        const int r = ob_enc_analysis(e, n_frames, d_pcm, frame_size, f0, Fc, k, k == 0 ? e->win_ev[0] : nullptr);
        if (r != OB_OK) return r;
    }
    for (int k = 0, f0 = 0; f0 < n_frames; k++, f0 += per) {
        const int Fc = n_frames - f0 < per ? n_frames - f0 : per;
        const int r = ob_enc_launch(e, n_frames, d_pcm, frame_size, d_out, max_bytes, d_lens_out, d_ranges_out, f0, Fc, k);
        if (r != OB_OK) return r;
    }
    if (sync) OB_CUDA(cudaStreamSynchronize(e->stream));
    return OB_OK;
}

static int32_t ob_encode_submit(ObEncoder *e, int32_t n_frames, const float *pcm, const int16_t *pcm16, int32_t frame_size, uint8_t *out,
                                int32_t max_bytes, int32_t *lens_out, uint32_t *ranges_out)
{
    if (!e || (!pcm && !pcm16) || !out || !lens_out || n_frames <= 0 || n_frames > e->max_frames || max_bytes <= 0 || frame_size <= 0) return OB_BAD_ARG;
    if (e->cfg.variable_duration > 5000) {                          // OPUS_SET_EXPERT_FRAME_DURATION: the duration must be the one the caller's buffers have
        const int sel = ob_frame_size_select(frame_size, e->cfg.variable_duration, 48000);
        if (sel <= 0) return OB_BAD_ARG;
        if (sel != frame_size) return OB_UNIMPLEMENTED;              // libopus would code the first `sel` samples of each longer buffer
    }
    OB_CUDA(cudaSetDevice(e->device));
    const size_t total = (size_t)e->S * n_frames, pcm_floats = total * (size_t)frame_size * e->CC, out_bytes = total * (size_t)max_bytes;
    if (pcm_floats > e->pcm_cap) { cudaFree(e->d_pcm); e->d_pcm = nullptr; e->pcm_cap = 0; OB_CUDA(cudaMalloc(&e->d_pcm, pcm_floats * sizeof(float))); e->pcm_cap = pcm_floats; }
    if (out_bytes > e->out_cap) { cudaFree(e->d_out); e->d_out = nullptr; e->out_cap = 0; OB_CUDA(cudaMalloc(&e->d_out, out_bytes)); e->out_cap = out_bytes; }
    // Frame windows: the upload of window k+1 (2-D copy out of the [S][F] host array, on the copy stream) overlaps the kernel of
    // window k; the per-stream state simply carries over from launch to launch.
    const int nwin = (total >= 32768 && n_frames >= 4) ? 4 : (total >= 8192 && n_frames >= 2 ? 2 : 1);
    const int per = (n_frames + nwin - 1) / nwin;
    const size_t pf = (size_t)frame_size * e->CC;
    const int lsb_saved = e->cfg.lsb_depth;
    if (pcm16) {                                       // half the upload; widened on the device; lsb_depth = min(16, user) (opus_encoder.c:1136)
        if (pcm_floats > e->pcm16_cap) { cudaFree(e->d_pcm16); e->d_pcm16 = nullptr; e->pcm16_cap = 0; OB_CUDA(cudaMalloc(&e->d_pcm16, pcm_floats * sizeof(int16_t))); e->pcm16_cap = pcm_floats; }
        OB_CUDA(cudaMemcpyAsync(e->d_pcm16, pcm16, pcm_floats * sizeof(int16_t), cudaMemcpyHostToDevice, e->stream));
        ob_k_i16_to_f32<<<(unsigned)((pcm_floats + 255) / 256), 256, 0, e->stream>>>(e->d_pcm16, e->d_pcm, pcm_floats);
        e->launches += 1;
        if (e->cfg.lsb_depth > 16) e->cfg.lsb_depth = 16;
    }
    for (int k = 0, f0 = 0; f0 < n_frames; k++, f0 += per) {
        const int Fc = n_frames - f0 < per ? n_frames - f0 : per;
        if (!pcm16) OB_CUDA(cudaMemcpy2DAsync(e->d_pcm + f0 * pf, n_frames * pf * sizeof(float), pcm + f0 * pf, n_frames * pf * sizeof(float),
                                  Fc * pf * sizeof(float), e->S, cudaMemcpyHostToDevice, e->copy_stream));
        OB_CUDA(cudaEventRecord(e->win_ev[k], pcm16 ? e->stream : e->copy_stream));      // this window's PCM is on the device
        OB_CUDA(cudaStreamWaitEvent(e->stream, e->win_ev[k], 0));
        int r = ob_enc_analysis(e, n_frames, e->d_pcm, frame_size, f0, Fc, k, e->win_ev[k]);
        if (r == OB_OK) r = ob_enc_launch(e, n_frames, e->d_pcm, frame_size, e->d_out, max_bytes, e->d_lens, e->d_ranges, f0, Fc, k);
        if (r != OB_OK) { e->cfg.lsb_depth = lsb_saved; return r; }
    }
    e->cfg.lsb_depth = lsb_saved;
    OB_CUDA(cudaMemcpyAsync(out, e->d_out, out_bytes, cudaMemcpyDeviceToHost, e->stream));
    OB_CUDA(cudaMemcpyAsync(lens_out, e->d_lens, total * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    if (ranges_out) OB_CUDA(cudaMemcpyAsync(ranges_out, e->d_ranges, total * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    OB_CUDA(cudaStreamSynchronize(e->stream));
    return OB_OK;
}

int32_t ob_encode_float_multi(ObEncoder *e, int32_t n_frames, const float *pcm, int32_t frame_size, uint8_t *out, int32_t max_bytes,
                              int32_t *lens_out, uint32_t *ranges_out)
{
    return ob_encode_submit(e, n_frames, pcm, nullptr, frame_size, out, max_bytes, lens_out, ranges_out);
}

// opus_encode (int16 PCM): Encoder::encode (src/encoder.rs:80-127).
int32_t ob_encode_multi(ObEncoder *e, int32_t n_frames, const int16_t *pcm, int32_t frame_size, uint8_t *out, int32_t max_bytes,
                        int32_t *lens_out, uint32_t *ranges_out)
{
    return ob_encode_submit(e, n_frames, nullptr, pcm, frame_size, out, max_bytes, lens_out, ranges_out);
}
int32_t ob_encode(ObEncoder *e, const int16_t *pcm, int32_t frame_size, uint8_t *out, int32_t max_bytes, int32_t *lens_out)
{
    return ob_encode_submit(e, 1, nullptr, pcm, frame_size, out, max_bytes, lens_out, nullptr);
}

int32_t ob_encode_float(ObEncoder *e, const float *pcm, int32_t frame_size, uint8_t *out, int32_t max_bytes, int32_t *lens_out)
{
    return ob_encode_float_multi(e, 1, pcm, frame_size, out, max_bytes, lens_out, nullptr);
}

int32_t ob_encoder_final_range(ObEncoder *e, uint32_t *out)
{
    if (!e || !out) return OB_BAD_ARG;
    OB_CUDA(cudaSetDevice(e->device));
    ob_k_enc_gather<<<(e->S + 127) / 128, 128, 0, e->stream>>>(e->d_streams, e->d_ranges, e->S);
    e->launches += 1;
    OB_CUDA(cudaMemcpyAsync(out, e->d_ranges, sizeof(uint32_t) * e->S, cudaMemcpyDeviceToHost, e->stream));
    OB_CUDA(cudaStreamSynchronize(e->stream));
    return OB_OK;
}

int32_t ob_encoder_in_dtx(ObEncoder *e, int32_t *out)
{
    if (!e || !out) return OB_BAD_ARG;
    OB_CUDA(cudaSetDevice(e->device));
    if (!e->cfg.use_dtx) { for (int s = 0; s < e->S; s++) out[s] = 0; return OB_OK; }
    ob_k_enc_gather_dtx<<<(e->S + 127) / 128, 128, 0, e->stream>>>(e->d_streams, e->d_ranges, e->S, 0);
    e->launches += 1;
    OB_CUDA(cudaMemcpyAsync(out, e->d_ranges, sizeof(uint32_t) * e->S, cudaMemcpyDeviceToHost, e->stream));
    OB_CUDA(cudaStreamSynchronize(e->stream));
    return OB_OK;
}

int32_t ob_encoder_get_bandwidth(ObEncoder *e, int32_t *out)
{
    if (!e || !out) return OB_BAD_ARG;
    OB_CUDA(cudaSetDevice(e->device));
    ob_k_enc_gather_dtx<<<(e->S + 127) / 128, 128, 0, e->stream>>>(e->d_streams, e->d_ranges, e->S, 1);
    e->launches += 1;
    OB_CUDA(cudaMemcpyAsync(out, e->d_ranges, sizeof(uint32_t) * e->S, cudaMemcpyDeviceToHost, e->stream));
    OB_CUDA(cudaStreamSynchronize(e->stream));
    return OB_OK;
}

int32_t ob_encoder_streams(const ObEncoder *e) { return e ? e->S : OB_BAD_ARG; }
int32_t ob_encoder_channels(const ObEncoder *e) { return e ? e->CC : OB_BAD_ARG; }
int32_t ob_encoder_sample_rate(const ObEncoder *e) { return e ? 48000 : OB_BAD_ARG; }                   // OPUS_GET_SAMPLE_RATE
int64_t ob_encoder_launches(const ObEncoder *e) { return e ? e->launches : 0; }
void *ob_encoder_cuda_stream(ObEncoder *e) { return e ? (void *)e->stream : nullptr; }
int32_t ob_encoder_kernel_ms(ObEncoder *e, float *ms)
{
    if (!e || !ms || !e->timed) return OB_BAD_ARG;
    OB_CUDA(cudaSetDevice(e->device));
    OB_CUDA(cudaEventSynchronize(e->ev[1]));
    OB_CUDA(cudaEventElapsedTime(ms, e->ev[0], e->ev[1]));
    return OB_OK;
}

}  // extern "C"
