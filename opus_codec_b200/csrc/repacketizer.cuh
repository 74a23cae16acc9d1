// Opus framing as byte work: the full packet parser (self-delimited form, padding), padding extensions, and the repacketizer
// (merge / split / pad / unpad).  One restatement compiled three ways: for the host entry points of the C ABI (one packet at a time,
// like the crate's Repacketizer, src/repacketizer.rs:11-100), for the batched kernel ob_k_repacketize (a warp per output packet:
// every lane runs the same header logic, the payload bytes are copied lane-strided, i.e. coalesced), and by g++ for the tests.
//
// Follows: opus_packet_parse_impl  opus/src/opus.c:194-353;  parse_size / encode_size  opus.c:146-192;
//          skip_extension, opus_packet_extensions_count / _parse / _generate  opus/src/extensions.c:36-276;
//          opus_repacketizer_cat_impl / _out_range_impl, opus_packet_pad_impl / _unpad  opus/src/repacketizer.c:56-353.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define OB_HD __host__ __device__ inline
#else
#define OB_HD static inline
#endif

#ifndef OB_OK
#define OB_OK 0
#define OB_BAD_ARG (-1)
#define OB_BUFFER_TOO_SMALL (-2)
#define OB_INTERNAL_ERROR (-3)
#define OB_INVALID_PACKET (-4)
#define OB_UNIMPLEMENTED (-5)
#endif

#define OB_RP_MAX_EXT 128                 // extensions one output packet may carry on the DEVICE path (the host path sizes its list exactly)

#ifdef __CUDACC__
#define OB_RP_MEM __host__ __device__
#else
#define OB_RP_MEM
#endif
struct ObRpLanes1 { static constexpr int lane = 0, n = 1; OB_RP_MEM void sync() const {} };     // one thread does everything (host, tests)

struct ObExt { int id, frame; const uint8_t *data; int len; };

struct ObRepack {                          // OpusRepacketizer (opus/src/opus_private.h:39-47)
    uint8_t toc;
    int nb_frames;
    const uint8_t *frames[48];
    int16_t len[48];
    int framesize;                         // samples per frame at 8 kHz (the 120 ms test of cat)
    const uint8_t *paddings[48];
    int padding_len[48];
};

OB_HD int ob_rp_parse_size(const uint8_t *data, int len, int16_t *size)
{
    if (len < 1) { *size = -1; return -1; }
    if (data[0] < 252) { *size = data[0]; return 1; }
    if (len < 2) { *size = -1; return -1; }
    *size = (int16_t)(4 * data[1] + data[0]);
    return 2;
}

OB_HD int ob_rp_encode_size(int size, uint8_t *data, bool write)
{
    if (size < 252) { if (write) data[0] = (uint8_t)size; return 1; }
    if (write) { data[0] = (uint8_t)(252 + (size & 3)); data[1] = (uint8_t)((size - (int)(252 + (size & 3))) >> 2); }
    return 2;
}

OB_HD int ob_rp_samples_per_frame(int toc, int fs)     // opus_packet_get_samples_per_frame, opus.c:173-192
{
    if (toc & 0x80) return (fs << ((toc >> 3) & 3)) / 400;
    if ((toc & 0x60) == 0x60) return (toc & 8) ? fs / 50 : fs / 100;
    const int a = (toc >> 3) & 3;
    return a == 3 ? fs * 60 / 1000 : (fs << a) / 100;
}

// Returns the frame count or an OPUS_* error.  frames / payload_offset / packet_offset / padding may be null.
OB_HD int ob_rp_parse(const uint8_t *data, int len, int self_delimited, uint8_t *out_toc, const uint8_t **frames, int16_t *size, int *payload_offset,
                      int *packet_offset, const uint8_t **padding, int *padding_len)
{
    if (!size || len < 0) return OB_BAD_ARG;
    if (len == 0) return OB_INVALID_PACKET;
    const uint8_t *data0 = data;
    const int framesize = ob_rp_samples_per_frame(data[0], 48000);
    int cbr = 0, count, bytes, pad = 0;
    const uint8_t toc = *data++;
    len--;
    int last_size = len;
    switch (toc & 3) {
    case 0: count = 1; break;
    case 1:
        count = 2; cbr = 1;
        if (!self_delimited) {
            if (len & 1) return OB_INVALID_PACKET;
            last_size = len / 2;
            size[0] = (int16_t)last_size;
        }
        break;
    case 2:
        count = 2;
        bytes = ob_rp_parse_size(data, len, size);
        len -= bytes;
        if (size[0] < 0 || size[0] > len) return OB_INVALID_PACKET;
        data += bytes;
        last_size = len - size[0];
        break;
    default: {
        if (len < 1) return OB_INVALID_PACKET;
        const uint8_t ch = *data++;
        count = ch & 0x3F;
        if (count <= 0 || framesize * count > 5760) return OB_INVALID_PACKET;
        len--;
        if (ch & 0x40) {
            int p;
            do {
                if (len <= 0) return OB_INVALID_PACKET;
                p = *data++;
                len--;
                const int tmp = p == 255 ? 254 : p;
                len -= tmp;
                pad += tmp;
            } while (p == 255);
        }
        if (len < 0) return OB_INVALID_PACKET;
        cbr = !(ch & 0x80);
        if (!cbr) {
            last_size = len;
            for (int i = 0; i < count - 1; i++) {
                bytes = ob_rp_parse_size(data, len, size + i);
                len -= bytes;
                if (size[i] < 0 || size[i] > len) return OB_INVALID_PACKET;
                data += bytes;
                last_size -= bytes + size[i];
            }
            if (last_size < 0) return OB_INVALID_PACKET;
        } else if (!self_delimited) {
            last_size = len / count;
            if (last_size * count != len) return OB_INVALID_PACKET;
            for (int i = 0; i < count - 1; i++) size[i] = (int16_t)last_size;
        }
    } }
    if (self_delimited) {                                    // one more size, for the last frame
        bytes = ob_rp_parse_size(data, len, size + count - 1);
        len -= bytes;
        if (size[count - 1] < 0 || size[count - 1] > len) return OB_INVALID_PACKET;
        data += bytes;
        if (cbr) {
            if (size[count - 1] * count > len) return OB_INVALID_PACKET;
            for (int i = 0; i < count - 1; i++) size[i] = size[count - 1];
        } else if (bytes + size[count - 1] > last_size) return OB_INVALID_PACKET;
    } else {
        if (last_size > 1275) return OB_INVALID_PACKET;
        size[count - 1] = (int16_t)last_size;
    }
    if (payload_offset) *payload_offset = (int)(data - data0);
    for (int i = 0; i < count; i++) {
        if (frames) frames[i] = data;
        data += size[i];
    }
    if (padding) { *padding = data; *padding_len = pad; }
    if (packet_offset) *packet_offset = pad + (int)(data - data0);
    if (out_toc) *out_toc = toc;
    return count;
}

// ---- padding extensions (extensions.c) ----
OB_HD int ob_ext_skip(const uint8_t **data, int len, int *header_size)
{
    if (len == 0) return 0;
    const int id = **data >> 1, L = **data & 1;
    if (id == 0 && L == 1) { *header_size = 1; (*data)++; return len - 1; }
    if (id > 0 && id < 32) {
        if (len < 1 + L) return -1;
        *data += 1 + L;
        *header_size = 1;
        return len - (1 + L);
    }
    if (L == 0) { *data += len; *header_size = 1; return 0; }
    int bytes = 0;
    *header_size = 1;
    do {
        (*data)++;
        len--;
        if (len == 0) return -1;
        bytes += **data;
        (*header_size)++;
    } while (**data == 255);
    (*data)++;
    len--;
    if (bytes > len) return -1;
    *data += bytes;
    return len - bytes;
}

OB_HD int ob_ext_count(const uint8_t *data, int len)
{
    int count = 0;
    while (len > 0) {
        int hs;
        const int id = *data >> 1;
        len = ob_ext_skip(&data, len, &hs);
        if (len < 0) return OB_INVALID_PACKET;
        if (id > 1) count++;
    }
    return count;
}

OB_HD int ob_ext_parse(const uint8_t *data, int len, ObExt *ext, int *nb)
{
    int curr_frame = 0, count = 0;
    while (len > 0) {
        int hs;
        ObExt e;
        e.id = 0; e.frame = 0; e.data = data; e.len = 0;
        const int id = *data >> 1;
        if (id > 1) { e.id = id; e.frame = curr_frame; }
        else if (id == 1) {
            if ((*data & 1) == 0) curr_frame++;
            else if (len >= 2) curr_frame += data[1];
            if (curr_frame >= 48) { *nb = count; return OB_INVALID_PACKET; }
        }
        len = ob_ext_skip(&data, len, &hs);
        if (len < 0) { *nb = count; return OB_INVALID_PACKET; }
        if (id > 1) {
            if (count == *nb) return OB_BUFFER_TOO_SMALL;
            e.len = (int)(data - e.data) - hs;
            e.data += hs;
            ext[count++] = e;
        }
    }
    *nb = count;
    return OB_OK;
}

// data == null: sizes only.  Lanes: the id / length bytes are written by lane 0, payload bytes lane-strided.
template <class G>
OB_HD int ob_ext_generate(const G &g, uint8_t *data, int len, const ObExt *ext, int nb, int pad)
{
    int max_frame = 0, curr_frame = 0, pos = 0, written = 0;
    for (int i = 0; i < nb; i++) {
        if (ext[i].frame > max_frame) max_frame = ext[i].frame;
        if (ext[i].id < 2 || ext[i].id > 127) return OB_BAD_ARG;
    }
    if (max_frame >= 48) return OB_BAD_ARG;
    const bool w0 = data && g.lane == 0;
    for (int frame = 0; frame <= max_frame; frame++) for (int i = 0; i < nb; i++) {
        if (ext[i].frame != frame) continue;
        if (frame != curr_frame) {                          // separator
            const int diff = frame - curr_frame;
            if (len - pos < 2) return OB_BUFFER_TOO_SMALL;
            if (diff == 1) { if (w0) data[pos] = 0x02; pos++; }
            else { if (w0) { data[pos] = 0x03; data[pos + 1] = (uint8_t)diff; } pos += 2; }
            curr_frame = frame;
        }
        if (ext[i].id < 32) {
            if (ext[i].len < 0 || ext[i].len > 1) return OB_BAD_ARG;
            if (len - pos < ext[i].len + 1) return OB_BUFFER_TOO_SMALL;
            if (w0) data[pos] = (uint8_t)((ext[i].id << 1) + ext[i].len);
            pos++;
            if (ext[i].len > 0) { if (w0) data[pos] = ext[i].data[0]; pos++; }
        } else {
            if (ext[i].len < 0) return OB_BAD_ARG;
            const int last = written == nb - 1;
            const int length_bytes = last ? 0 : 1 + ext[i].len / 255;
            if (len - pos < 1 + length_bytes + ext[i].len) return OB_BUFFER_TOO_SMALL;
            if (w0) data[pos] = (uint8_t)((ext[i].id << 1) + !last);
            pos++;
            if (!last) {
                for (int j = 0; j < ext[i].len / 255; j++) { if (w0) data[pos] = 255; pos++; }
                if (w0) data[pos] = (uint8_t)(ext[i].len % 255);
                pos++;
            }
            if (data) for (int j = g.lane; j < ext[i].len; j += g.n) data[pos + j] = ext[i].data[j];
            pos += ext[i].len;
        }
        written++;
    }
    if (pad && pos < len) {                                 // not reached from the repacketizer (pad = 0 there); single-lane form
        const int padding = len - pos;
        if (w0) {
            for (int j = pos - 1; j >= 0; j--) data[padding + j] = data[j];
            for (int j = 0; j < padding; j++) data[j] = 0x01;
        }
        pos += padding;
    }
    return pos;
}

// ---- repacketizer ----
OB_HD void ob_repack_init(ObRepack *rp) { rp->nb_frames = 0; }

OB_HD int ob_repack_cat(ObRepack *rp, const uint8_t *data, int len, int self_delimited)
{
    if (len < 1) return OB_INVALID_PACKET;
    if (rp->nb_frames == 0) { rp->toc = data[0]; rp->framesize = ob_rp_samples_per_frame(data[0], 8000); }
    else if ((rp->toc & 0xFC) != (data[0] & 0xFC)) return OB_INVALID_PACKET;
    const int code = data[0] & 3;
    int curr = code == 0 ? 1 : code != 3 ? 2 : (len < 2 ? OB_INVALID_PACKET : (data[1] & 0x3F));       // opus_packet_get_nb_frames
    if (curr < 1) return OB_INVALID_PACKET;
    if ((curr + rp->nb_frames) * rp->framesize > 960) return OB_INVALID_PACKET;                        // 120 ms
    uint8_t tmp_toc;
    const int ret = ob_rp_parse(data, len, self_delimited, &tmp_toc, &rp->frames[rp->nb_frames], &rp->len[rp->nb_frames], nullptr, nullptr,
                                &rp->paddings[rp->nb_frames], &rp->padding_len[rp->nb_frames]);
    if (ret < 1) return ret;
    while (curr > 1) {                                      // the padding belongs to the packet's first frame
        rp->nb_frames++;
        rp->padding_len[rp->nb_frames] = 0;
        rp->paddings[rp->nb_frames] = nullptr;
        curr--;
    }
    rp->nb_frames++;
    return OB_OK;
}

// all_ext: room for max_ext entries (the caller counts with ob_repack_count_ext when it wants an exact size).
OB_HD int ob_repack_count_ext(const ObRepack *rp, int begin, int end)
{
    int total = 0;
    for (int i = begin; i < end; i++) {
        const int n = ob_ext_count(rp->paddings[i], rp->padding_len[i]);
        if (n > 0) total += n;
    }
    return total;
}

// data may overlap the frames when every byte moves towards the front (unpad in place): the copy runs forwards.
template <class G>
OB_HD int ob_repack_out_range(const G &g, const ObRepack *rp, int begin, int end, uint8_t *data, int maxlen, int self_delimited, int pad, ObExt *all_ext,
                              int max_ext)
{
    if (begin < 0 || begin >= end || end > rp->nb_frames) return OB_BAD_ARG;
    const int count = end - begin;
    const int16_t *len = rp->len + begin;
    const uint8_t *const *frames = rp->frames + begin;
    const bool w0 = g.lane == 0;
    int tot_size = self_delimited ? 1 + (len[count - 1] >= 252) : 0;
    int ext_count = 0, ext_len = 0, ext_begin = 0, ones_begin = 0, ones_end = 0;
    for (int i = begin; i < end; i++) {
        int n = max_ext - ext_count;
        const int ret = ob_ext_parse(rp->paddings[i], rp->padding_len[i], all_ext + ext_count, &n);
        if (ret == OB_BUFFER_TOO_SMALL) return OB_UNIMPLEMENTED;          // more extensions than this build keeps (device path only)
        if (ret < 0) return OB_INTERNAL_ERROR;
        // every lane of the group has just parsed the same records into the shared all_ext (identical stores); the frame fix-up is a
        // read-modify-write, so exactly one lane does it, between two group barriers
        g.sync();
        if (w0) for (int j = 0; j < n; j++) all_ext[ext_count + j].frame += i - begin;
        g.sync();
        ext_count += n;
    }
    uint8_t *ptr = data;
    if (count == 1) {
        tot_size += len[0] + 1;
        if (tot_size > maxlen) return OB_BUFFER_TOO_SMALL;
        if (w0) *ptr = (uint8_t)(rp->toc & 0xFC);
        ptr++;
    } else if (count == 2) {
        if (len[1] == len[0]) {
            tot_size += 2 * len[0] + 1;
            if (tot_size > maxlen) return OB_BUFFER_TOO_SMALL;
            if (w0) *ptr = (uint8_t)((rp->toc & 0xFC) | 1);
            ptr++;
        } else {
            tot_size += len[0] + len[1] + 2 + (len[0] >= 252);
            if (tot_size > maxlen) return OB_BUFFER_TOO_SMALL;
            if (w0) *ptr = (uint8_t)((rp->toc & 0xFC) | 2);
            ptr++;
            ptr += ob_rp_encode_size(len[0], ptr, w0);
        }
    }
    if (count > 2 || (pad && tot_size < maxlen) || ext_count > 0) {       // code 3
        ptr = data;
        tot_size = self_delimited ? 1 + (len[count - 1] >= 252) : 0;
        int vbr = 0;
        for (int i = 1; i < count; i++) if (len[i] != len[0]) { vbr = 1; break; }
        int byte1;
        if (vbr) {
            tot_size += 2;
            for (int i = 0; i < count - 1; i++) tot_size += 1 + (len[i] >= 252) + len[i];
            tot_size += len[count - 1];
            if (tot_size > maxlen) return OB_BUFFER_TOO_SMALL;
            byte1 = count | 0x80;
        } else {
            tot_size += count * len[0] + 2;
            if (tot_size > maxlen) return OB_BUFFER_TOO_SMALL;
            byte1 = count;
        }
        ptr += 2;
        int pad_amount = pad ? maxlen - tot_size : 0;
        if (ext_count > 0) {
            ext_len = ob_ext_generate(g, (uint8_t *)nullptr, maxlen - tot_size, all_ext, ext_count, 0);
            if (ext_len < 0) return ext_len;
            if (!pad) pad_amount = ext_len + ext_len / 254 + 1;
        }
        if (pad_amount != 0) {
            byte1 |= 0x40;
            const int nb_255s = (pad_amount - 1) / 255;
            if (tot_size + ext_len + nb_255s + 1 > maxlen) return OB_BUFFER_TOO_SMALL;
            ext_begin = tot_size + pad_amount - ext_len;
            ones_begin = tot_size + nb_255s + 1;
            ones_end = tot_size + pad_amount - ext_len;
            if (w0) { for (int i = 0; i < nb_255s; i++) ptr[i] = 255; ptr[nb_255s] = (uint8_t)(pad_amount - 255 * nb_255s - 1); }
            ptr += nb_255s + 1;
            tot_size += pad_amount;
        }
        if (w0) { data[0] = (uint8_t)((rp->toc & 0xFC) | 3); data[1] = (uint8_t)byte1; }
        if (vbr) for (int i = 0; i < count - 1; i++) ptr += ob_rp_encode_size(len[i], ptr, w0);
    }
    if (self_delimited) ptr += ob_rp_encode_size(len[count - 1], ptr, w0);
    for (int i = 0; i < count; i++) {
        const uint8_t *src = frames[i];
        if (ptr != src) for (int k = g.lane; k < len[i]; k += g.n) ptr[k] = src[k];
        ptr += len[i];
    }
    if (ext_len > 0) ob_ext_generate(g, data + ext_begin, ext_len, all_ext, ext_count, 0);
    for (int i = ones_begin + g.lane; i < ones_end; i += g.n) data[i] = 0x01;
    if (pad && ext_count == 0) for (int i = (int)(ptr - data) + g.lane; i < maxlen; i += g.n) data[i] = 0;
    return tot_size;
}
