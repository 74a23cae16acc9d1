"""Packet utilities of the crate over the C ABI: `packet_parse` / `packet_pad` / `packet_unpad` (src/packet.rs:162-248), the
`Repacketizer` object (src/repacketizer.rs:11-100) and its batched GPU form `repacketize_batch` (no counterpart: one call merges
every `group` consecutive packets of every stream).  Error behaviour follows the crate: negative OPUS_* codes raise `OpusError`."""
import ctypes as C

import numpy as np

from . import _lib
from .batch import OpusError, BAD_ARG, pack_packets


def _buf(b):
    return (C.c_uint8 * max(1, len(b))).from_buffer_copy(bytes(b) if len(b) else b"\0")


def packet_parse(packet):
    """-> (toc, payload_offset, [frame bytes]) like packet_parse (src/packet.rs:162-215): empty frames are skipped."""
    if len(packet) == 0:
        raise OpusError(BAD_ARG)
    L = _lib.lib()
    data = _buf(packet)
    toc = C.c_uint8(0); po = C.c_int32(0)
    offs = (C.c_int32 * 48)(); sizes = (C.c_int16 * 48)()
    n = L.ob_packet_parse(data, len(packet), C.byref(toc), offs, sizes, C.byref(po))
    if n < 0:
        raise OpusError(n)
    raw = bytes(packet)
    return toc.value, po.value, [raw[offs[i]:offs[i] + sizes[i]] for i in range(n) if sizes[i] > 0]


def packet_parse_raw(packet):
    """-> (count or error code, toc, payload_offset, offsets, sizes): the unfiltered result (tests)."""
    L = _lib.lib()
    data = _buf(packet)
    toc = C.c_uint8(0); po = C.c_int32(0)
    offs = (C.c_int32 * 48)(); sizes = (C.c_int16 * 48)()
    n = L.ob_packet_parse(data, len(packet), C.byref(toc), offs, sizes, C.byref(po))
    return n, toc.value, po.value, list(offs[:max(n, 0)]), list(sizes[:max(n, 0)])


def packet_pad(packet, new_len):
    """-> the packet padded to new_len bytes (packet_pad, src/packet.rs:220-232)."""
    if new_len < len(packet):
        raise OpusError(BAD_ARG)
    L = _lib.lib()
    data = (C.c_uint8 * max(1, new_len))()
    C.memmove(data, bytes(packet), len(packet))
    r = L.ob_packet_pad(data, len(packet), new_len)
    if r != 0:
        raise OpusError(r)
    return bytes(data[:new_len])


def packet_unpad(packet):
    """-> the packet without padding (packet_unpad, src/packet.rs:237-248)."""
    L = _lib.lib()
    data = _buf(packet)
    n = L.ob_packet_unpad(data, len(packet))
    if n < 0:
        raise OpusError(n)
    return bytes(data[:n])


def multistream_packet_pad(packet, new_len, nb_streams):
    """-> the multistream packet padded to new_len bytes (multistream_packet_pad, src/packet.rs:253-272)."""
    if new_len < len(packet):
        raise OpusError(BAD_ARG)
    data = (C.c_uint8 * max(1, new_len))()
    C.memmove(data, bytes(packet), len(packet))
    r = _lib.lib().ob_multistream_packet_pad(data, len(packet), new_len, nb_streams)
    if r != 0:
        raise OpusError(r)
    return bytes(data[:new_len])


def multistream_packet_unpad(packet, nb_streams):
    """-> the multistream packet without padding (multistream_packet_unpad, src/packet.rs:277-290)."""
    data = _buf(packet)
    n = _lib.lib().ob_multistream_packet_unpad(data, len(packet), nb_streams)
    if n < 0:
        raise OpusError(n)
    return bytes(data[:n])


class Repacketizer:
    """Repacketizer of the crate (src/repacketizer.rs): push() packets of one configuration, out()/out_range() merged packets.
    Pushed packets are kept alive by this object (libopus references them)."""

    def __init__(self):
        self._L = _lib.lib()
        self._h = self._L.ob_repacketizer_create()
        if not self._h:
            raise OpusError(-7)
        self._keep = []

    def close(self):
        if self._h:
            self._L.ob_repacketizer_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def reset(self):
        self._L.ob_repacketizer_init(self._h)
        self._keep = []

    def push(self, packet):
        if len(packet) == 0:
            raise OpusError(BAD_ARG)
        b = _buf(packet)
        r = self._L.ob_repacketizer_cat(self._h, b, len(packet))
        if r != 0:
            raise OpusError(r)
        self._keep.append(b)

    def frames(self):
        return self._L.ob_repacketizer_get_nb_frames(self._h)

    def out_range(self, begin, end, max_len=1277 * 6):
        if max_len <= 0 or begin < 0 or end <= begin:
            raise OpusError(BAD_ARG)
        out = (C.c_uint8 * max_len)()
        n = self._L.ob_repacketizer_out_range(self._h, begin, end, out, max_len)
        if n < 0:
            raise OpusError(n)
        return bytes(out[:n])

    def out(self, max_len=1277 * 6):
        if max_len <= 0:
            raise OpusError(BAD_ARG)
        out = (C.c_uint8 * max_len)()
        n = self._L.ob_repacketizer_out(self._h, out, max_len)
        if n < 0:
            raise OpusError(n)
        return bytes(out[:n])


def repacketize_batch(packets, group, pad_to=0, max_bytes=None, device=0):
    """packets: [S][F] byte strings.  Every `group` consecutive packets of each stream become one packet on the GPU.
    -> (out u8 [S, ceil(F/group), max_bytes], lens i32 [S, ceil(F/group)]); lens < 0 are OPUS_* codes."""
    L = _lib.lib()
    S, F = len(packets), len(packets[0])
    flat, offsets, lens = pack_packets(packets)
    G = (F + group - 1) // group
    if max_bytes is None:
        max_bytes = pad_to if pad_to > 0 else int(2 + 2 * group + sum(sorted((len(p) for row in packets for p in row), reverse=True)[:group]))
    out = np.zeros((S, G, max_bytes), np.uint8)
    lens_out = np.zeros((S, G), np.int32)
    r = L.ob_repacketize_batch(device, S, F, flat.ctypes.data, offsets.ctypes.data, lens.ctypes.data, group, pad_to, out.ctypes.data, max_bytes,
                               lens_out.ctypes.data)
    if r != 0:
        raise OpusError(r)
    return out, lens_out


def soft_clip_batch(pcm, channels, softclip_mem, device=0):
    """soft_clip (src/packet.rs:123-155) for a batch, in place on the GPU: pcm f32 [S, frame_size*channels], softclip_mem f32 [S, channels]."""
    if pcm.dtype != np.float32 or softclip_mem.dtype != np.float32 or pcm.ndim != 2 or softclip_mem.shape != (pcm.shape[0], channels) \
            or pcm.shape[1] % channels or not pcm.flags.c_contiguous or not softclip_mem.flags.c_contiguous:
        raise OpusError(BAD_ARG)
    r = _lib.lib().ob_pcm_soft_clip_batch(device, pcm.shape[0], pcm.ctypes.data, pcm.shape[1] // channels, channels, softclip_mem.ctypes.data)
    if r != 0:
        raise OpusError(r)
