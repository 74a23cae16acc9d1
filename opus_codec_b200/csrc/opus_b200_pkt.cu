// Packet-level byte work of the C ABI: opus_packet_parse / pad / unpad, the repacketizer object (one packet at a time, host) and
// the batched repacketizer kernel (a warp per output packet).  All of it is the restatement in repacketizer.cuh.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <vector>
#include <new>

#include "ob_ir.h"
#include "repacketizer.cuh"
#include "../../include/opus_b200.h"

struct ObRepacketizer { ObRepack rp; };

extern "C" {

// ---- opus_repacketizer_* (src/bindings.rs; Repacketizer, src/repacketizer.rs:11-100) ----
ObRepacketizer *ob_repacketizer_create(void)
{
    ObRepacketizer *r = new (std::nothrow) ObRepacketizer;
    if (r) ob_repack_init(&r->rp);
    return r;
}
void ob_repacketizer_destroy(ObRepacketizer *r) { delete r; }
void ob_repacketizer_init(ObRepacketizer *r) { if (r) ob_repack_init(&r->rp); }
int32_t ob_repacketizer_cat(ObRepacketizer *r, const uint8_t *data, int32_t len)
{
    if (!r || !data) return OB_BAD_ARG;
    return ob_repack_cat(&r->rp, data, len, 0);
}
int32_t ob_repacketizer_get_nb_frames(ObRepacketizer *r) { return r ? r->rp.nb_frames : OB_BAD_ARG; }

static int32_t ob_out_range_host(const ObRepack *rp, int begin, int end, uint8_t *out, int32_t maxlen, int pad)
{
    if (begin < 0 || begin >= end || end > rp->nb_frames) return OB_BAD_ARG;
    const int n = ob_repack_count_ext(rp, begin, end);
    std::vector<ObExt> ext((size_t)(n > 0 ? n : 1));
    return ob_repack_out_range(ObRpLanes1(), rp, begin, end, out, maxlen, 0, pad, ext.data(), n > 0 ? n : 0);
}
int32_t ob_repacketizer_out_range(ObRepacketizer *r, int32_t begin, int32_t end, uint8_t *out, int32_t maxlen)
{
    if (!r || !out) return OB_BAD_ARG;
    return ob_out_range_host(&r->rp, begin, end, out, maxlen, 0);
}
int32_t ob_repacketizer_out(ObRepacketizer *r, uint8_t *out, int32_t maxlen)
{
    if (!r || !out) return OB_BAD_ARG;
    return ob_out_range_host(&r->rp, 0, r->rp.nb_frames, out, maxlen, 0);
}

// ---- opus_packet_pad / _unpad / _parse (src/packet.rs:162-248) ----
int32_t ob_packet_pad(uint8_t *data, int32_t len, int32_t new_len)
{
    if (!data || len < 1) return OB_BAD_ARG;
    if (len == new_len) return OB_OK;
    if (len > new_len) return OB_BAD_ARG;
    std::vector<uint8_t> copy(data, data + len);
    ObRepack rp;
    ob_repack_init(&rp);
    int ret = ob_repack_cat(&rp, copy.data(), len, 0);
    if (ret != OB_OK) return ret;
    ret = ob_out_range_host(&rp, 0, rp.nb_frames, data, new_len, 1);
    return ret > 0 ? OB_OK : ret;
}
int32_t ob_packet_unpad(uint8_t *data, int32_t len)
{
    if (!data || len < 1) return OB_BAD_ARG;
    ObRepack rp;
    ob_repack_init(&rp);
    const int ret = ob_repack_cat(&rp, data, len, 0);
    if (ret < 0) return ret;
    for (int i = 0; i < rp.nb_frames; i++) { rp.padding_len[i] = 0; rp.paddings[i] = nullptr; }
    ObExt none;
    return ob_repack_out_range(ObRpLanes1(), &rp, 0, rp.nb_frames, data, len, 0, 0, &none, 0);     // in place: every byte moves towards the front
}
int32_t ob_packet_parse(const uint8_t *data, int32_t len, uint8_t *out_toc, int32_t *frame_offsets, int16_t *sizes, int32_t *payload_offset)
{
    if (!data && len > 0) return OB_BAD_ARG;
    const uint8_t *frames[48];
    int16_t sz[48];
    int po = 0;
    const int n = ob_rp_parse(data, len, 0, out_toc, frames, sizes ? sz : nullptr, &po, nullptr, nullptr, nullptr);
    if (n < 0) return n;
    for (int i = 0; i < n; i++) { sizes[i] = sz[i]; if (frame_offsets) frame_offsets[i] = (int32_t)(frames[i] - data); }
    if (payload_offset) *payload_offset = po;
    return n;
}

// ---- opus_packet_get_nb_samples / opus_packet_has_lbrr (packet_nb_samples / packet_has_lbrr, src/packet.rs:72-120; opus_decoder.c:1119-1162) ----
int32_t ob_packet_get_nb_samples(const uint8_t *packet, int32_t len, int32_t fs)
{
    if (!packet || len < 1) return OB_BAD_ARG;
    const int code = packet[0] & 3;
    const int count = code == 0 ? 1 : code != 3 ? 2 : (len < 2 ? OB_INVALID_PACKET : (packet[1] & 0x3F));
    if (count < 0) return count;
    const int samples = count * ob_rp_samples_per_frame(packet[0], fs);
    return samples * 25 > fs * 3 ? OB_INVALID_PACKET : samples;             // more than 120 ms
}
int32_t ob_packet_has_lbrr(const uint8_t *packet, int32_t len)
{
    if (!packet || len < 1) return OB_BAD_ARG;
    if (packet[0] & 0x80) return 0;                                           // CELT-only packets carry no LBRR
    const int frame_size = ob_rp_samples_per_frame(packet[0], 48000), nb_frames = frame_size > 960 ? frame_size / 960 : 1;
    const uint8_t *frames[48];
    int16_t size[48];
    const int ret = ob_rp_parse(packet, len, 0, nullptr, frames, size, nullptr, nullptr, nullptr, nullptr);
    if (ret <= 0) return ret;
    if (frames[0] >= packet + len) return 0;                                  // an empty first frame at the very end of the buffer: nothing to look at
    int lbrr = (frames[0][0] >> (7 - nb_frames)) & 1;
    if (packet[0] & 4) lbrr = lbrr || ((frames[0][0] >> (6 - 2 * nb_frames)) & 1);
    return lbrr;
}

// ---- opus_multistream_packet_pad / _unpad (multistream_packet_pad / _unpad, src/packet.rs:253-290; repacketizer.c:355-464) ----
int32_t ob_multistream_packet_pad(uint8_t *data, int32_t len, int32_t new_len, int32_t nb_streams)
{
    if (!data || len < 1) return OB_BAD_ARG;
    if (len == new_len) return OB_OK;
    if (len > new_len) return OB_BAD_ARG;
    const int32_t amount = new_len - len;
    for (int s = 0; s < nb_streams - 1; s++) {               // seek to the last stream: the others are self-delimited
        if (len <= 0) return OB_INVALID_PACKET;
        uint8_t toc;
        int16_t size[48];
        int packet_offset = 0;
        const int count = ob_rp_parse(data, len, 1, &toc, nullptr, size, nullptr, &packet_offset, nullptr, nullptr);
        if (count < 0) return count;
        data += packet_offset;
        len -= packet_offset;
    }
    return ob_packet_pad(data, len, len + amount);
}
int32_t ob_multistream_packet_unpad(uint8_t *data, int32_t len, int32_t nb_streams)
{
    if (!data || len < 1) return OB_BAD_ARG;
    uint8_t *dst = data;
    int32_t dst_len = 0;
    for (int s = 0; s < nb_streams; s++) {
        const int self_delimited = s != nb_streams - 1;
        if (len <= 0) return OB_INVALID_PACKET;
        uint8_t toc;
        int16_t size[48];
        int packet_offset = 0;
        int ret = ob_rp_parse(data, len, self_delimited, &toc, nullptr, size, nullptr, &packet_offset, nullptr, nullptr);
        if (ret < 0) return ret;
        ObRepack rp;
        ob_repack_init(&rp);
        ret = ob_repack_cat(&rp, data, packet_offset, self_delimited);
        if (ret < 0) return ret;
        for (int i = 0; i < rp.nb_frames; i++) { rp.padding_len[i] = 0; rp.paddings[i] = nullptr; }       // padding and extensions are dropped
        ObExt none;
        ret = ob_repack_out_range(ObRpLanes1(), &rp, 0, rp.nb_frames, dst, len, self_delimited, 0, &none, 0);
        if (ret < 0) return ret;
        dst_len += ret;
        dst += ret;
        data += packet_offset;
        len -= packet_offset;
    }
    return dst_len;
}

}  // extern "C"

// ---- batched: each `group` consecutive packets of a stream -> one packet ----------------------------------------------------------
#define OB_RP_WARPS 4
struct ObRpWarp { int lane; static constexpr int n = 32;  __device__ __forceinline__ void sync() const { __syncwarp(); } };

__global__ void __launch_bounds__(OB_RP_WARPS * 32)
ob_k_repacketize(const uint8_t *__restrict__ packets, const int32_t *__restrict__ offsets, const int32_t *__restrict__ lens, int S, int n_in, int group,
                 int pad_to, uint8_t *__restrict__ out, int max_bytes, int32_t *__restrict__ lens_out)
{
    __shared__ ObRepack s_rp[OB_RP_WARPS];
    __shared__ ObExt s_ext[OB_RP_WARPS][OB_RP_MAX_EXT];
    __shared__ int s_ret[OB_RP_WARPS];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = (n_in + group - 1) / group;
    const long long w = (long long)blockIdx.x * OB_RP_WARPS + wib;
    if (w >= (long long)S * G) return;
    const int s = (int)(w / G), gi = (int)(w % G);
    ObRepack *rp = &s_rp[wib];
    if (lane == 0) {
        ob_repack_init(rp);
        int ret = OB_OK;
        const int f1 = min(n_in, (gi + 1) * group);
        for (int f = gi * group; f < f1 && ret == OB_OK; f++) {
            const size_t k = (size_t)s * n_in + f;
            ret = ob_repack_cat(rp, packets + offsets[k], lens[k], 0);
        }
        s_ret[wib] = ret;
    }
    __syncwarp();
    int ret = s_ret[wib];
    ObRpWarp g{lane};
    uint8_t *dst = out + (size_t)w * max_bytes;
    if (ret == OB_OK) {
        const int maxlen = pad_to > 0 ? min(pad_to, max_bytes) : max_bytes;
        ret = ob_repack_out_range(g, rp, 0, rp->nb_frames, dst, maxlen, 0, pad_to > 0, s_ext[wib], OB_RP_MAX_EXT);
    }
    if (lane == 0) lens_out[w] = ret;
}

extern "C" {

int32_t ob_repacketize_batch_device(int32_t n_streams, int32_t n_in, const uint8_t *d_packets, const int32_t *d_offsets, const int32_t *d_lens, int32_t group,
                                    int32_t pad_to, uint8_t *d_out, int32_t max_bytes, int32_t *d_lens_out, void *stream)
{
    if (n_streams <= 0 || n_in <= 0 || group <= 0 || group > 48 || max_bytes <= 0 || pad_to < 0 || !d_packets || !d_offsets || !d_lens || !d_out || !d_lens_out)
        return OB_BAD_ARG;
    const long long G = (n_in + group - 1) / group, warps = (long long)n_streams * G;
    const long long blocks = (warps + OB_RP_WARPS - 1) / OB_RP_WARPS;
    if (blocks > 0x7fffffffLL) return OB_BAD_ARG;
    ob_k_repacketize<<<(unsigned)blocks, OB_RP_WARPS * 32, 0, (cudaStream_t)stream>>>(d_packets, d_offsets, d_lens, n_streams, n_in, group, pad_to, d_out, max_bytes,
                                                                                    d_lens_out);
    return cudaGetLastError() == cudaSuccess ? OB_OK : OB_INTERNAL_ERROR;
}

int32_t ob_repacketize_batch(int32_t device, int32_t n_streams, int32_t n_in, const uint8_t *packets, const int32_t *offsets, const int32_t *lens, int32_t group,
                             int32_t pad_to, uint8_t *out, int32_t max_bytes, int32_t *lens_out)
{
    if (n_streams <= 0 || n_in <= 0 || group <= 0 || group > 48 || max_bytes <= 0 || pad_to < 0 || !packets || !offsets || !lens || !out || !lens_out) return OB_BAD_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return OB_INTERNAL_ERROR;       // no CPU fallback
    if (device < 0 || device >= ndev || cudaSetDevice(device) != cudaSuccess) return OB_BAD_ARG;
    const size_t n = (size_t)n_streams * n_in, G = (size_t)(n_in + group - 1) / group, n_out = (size_t)n_streams * G;
    size_t bytes = 0;
    for (size_t k = 0; k < n; k++) {
        if (lens[k] < 0 || offsets[k] < 0) return OB_BAD_ARG;
        const size_t e = (size_t)offsets[k] + (size_t)lens[k];
        if (e > bytes) bytes = e;
    }
    uint8_t *d_pk = nullptr, *d_out = nullptr;
    int32_t *d_meta = nullptr;
    int32_t rc = OB_ALLOC_FAIL;
    if (cudaMalloc(&d_pk, bytes ? bytes : 1) == cudaSuccess && cudaMalloc(&d_out, n_out * (size_t)max_bytes) == cudaSuccess &&
        cudaMalloc(&d_meta, (2 * n + n_out) * sizeof(int32_t)) == cudaSuccess) {
        rc = OB_INTERNAL_ERROR;
        if (cudaMemcpy(d_pk, packets, bytes, cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMemcpy(d_meta, offsets, n * sizeof(int32_t), cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMemcpy(d_meta + n, lens, n * sizeof(int32_t), cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMemset(d_out, 0, n_out * (size_t)max_bytes) == cudaSuccess &&
            ob_repacketize_batch_device(n_streams, n_in, d_pk, d_meta, d_meta + n, group, pad_to, d_out, max_bytes, d_meta + 2 * n, nullptr) == OB_OK &&
            cudaMemcpy(out, d_out, n_out * (size_t)max_bytes, cudaMemcpyDeviceToHost) == cudaSuccess &&
            cudaMemcpy(lens_out, d_meta + 2 * n, n_out * sizeof(int32_t), cudaMemcpyDeviceToHost) == cudaSuccess)
            rc = OB_OK;
    }
    cudaFree(d_pk); cudaFree(d_out); cudaFree(d_meta);
    return rc;
}

}  // extern "C"
