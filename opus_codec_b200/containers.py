"""Wire / file formats around the batched codec -- pure host byte work that turns captures into the
`[stream][packet]` lists `BatchDecoder` takes and back:

* `.bit`  -- the container of the reference's own tool, opus/src/opus_demo.c: per packet a 4-byte big-endian length, the
             4-byte big-endian encoder final range, the payload (`int_to_char`/`char_to_int` :145-160, reader :843-866, writer :946-962).
* RTP     -- RFC 7587 (opus/doc/draft-ietf-payload-rtp-opus.xml): one Opus packet per RTP payload, 48 kHz timestamp clock whatever the
             coded bandwidth; `rtp_to_packets` orders datagrams by sequence number and turns gaps into lost packets (b"") of the right
             duration, which is what the decoder's concealment wants.
* Ogg     -- RFC 7845 (opus/doc/draft-ietf-codec-oggopus.xml): OpusHead / OpusTags, lacing, continued packets, granule positions, page CRC.

Nothing here touches the GPU; the reference crate has no counterpart (its users bring their own transport)."""
import struct

from . import _lib


# ---- opus_demo .bit -----------------------------------------------------------------------------------------------------------------
def read_bit(data):
    """bytes of a .bit file -> (packets [bytes], enc_final_ranges [int]).  A zero length is a lost packet (opus_demo -d treats it so)."""
    packets, ranges, pos = [], [], 0
    while pos + 8 <= len(data):
        n, rng = struct.unpack_from(">II", data, pos)
        pos += 8
        if n > 1500 * 6 or pos + n > len(data):                # opus_demo: "Invalid payload length" / truncated file
            raise ValueError("invalid payload length %d at byte %d" % (n, pos - 8))
        packets.append(bytes(data[pos:pos + n]))
        ranges.append(rng)
        pos += n
    return packets, ranges


def write_bit(packets, ranges):
    """-> bytes of a .bit file; ranges[i] = OPUS_GET_FINAL_RANGE after packet i (opus_demo -d checks it against its decoder)."""
    out = bytearray()
    for p, r in zip(packets, ranges):
        out += struct.pack(">II", len(p), int(r) & 0xFFFFFFFF) + bytes(p)
    return bytes(out)


# ---- RTP (RFC 3550 header, RFC 7587 payload) ------------------------------------------------------------------------------------------
def rtp_pack(payload, seq, timestamp, ssrc, payload_type=111, marker=False):
    """One Opus packet -> one RTP datagram (no CSRCs, no header extension, no padding: RFC 7587 section 4.2)."""
    return struct.pack(">BBHII", 0x80, (0x80 if marker else 0) | (payload_type & 0x7F), seq & 0xFFFF, timestamp & 0xFFFFFFFF, ssrc & 0xFFFFFFFF) + bytes(payload)


def rtp_unpack(datagram):
    """-> dict(version, marker, payload_type, seq, timestamp, ssrc, csrc, extension, payload).  Honours CSRC count, the header extension
    (X) and trailing padding (P).  Raises ValueError on a malformed datagram."""
    if len(datagram) < 12:
        raise ValueError("RTP datagram shorter than its fixed header")
    b0, b1, seq, ts, ssrc = struct.unpack_from(">BBHII", datagram, 0)
    if b0 >> 6 != 2:
        raise ValueError("RTP version %d" % (b0 >> 6))
    cc, pos = b0 & 0x0F, 12
    if len(datagram) < pos + 4 * cc:
        raise ValueError("truncated CSRC list")
    csrc = list(struct.unpack_from(">%dI" % cc, datagram, pos))
    pos += 4 * cc
    ext = None
    if b0 & 0x10:
        if len(datagram) < pos + 4:
            raise ValueError("truncated header extension")
        prof, words = struct.unpack_from(">HH", datagram, pos)
        if len(datagram) < pos + 4 + 4 * words:
            raise ValueError("truncated header extension")
        ext = (prof, bytes(datagram[pos + 4:pos + 4 + 4 * words]))
        pos += 4 + 4 * words
    end = len(datagram)
    if b0 & 0x20:
        pad = datagram[-1]
        if pad == 0 or pad > end - pos:
            raise ValueError("bad RTP padding")
        end -= pad
    return dict(version=2, marker=bool(b1 & 0x80), payload_type=b1 & 0x7F, seq=seq, timestamp=ts, ssrc=ssrc, csrc=csrc, extension=ext,
                payload=bytes(datagram[pos:end]))


def packet_duration_48k(packet):
    """Samples at 48 kHz one Opus packet covers (opus_packet_get_nb_samples); 0 for an empty payload."""
    if len(packet) == 0:
        return 0
    L = _lib.lib()
    n = L.ob_packet_get_nb_frames(bytes(packet), len(packet))
    if n < 0:
        raise ValueError("invalid Opus packet (%d)" % n)
    return n * L.ob_packet_get_samples_per_frame(bytes(packet), 48000)


def rtp_to_packets(datagrams, frame_samples=None, payload_type=None):
    """Datagrams of ONE SSRC, in arrival order (duplicates and reordering allowed) -> the packet list the decoder takes: ordered by
    extended sequence number, duplicates dropped, every missing stretch replaced by lost packets (b"") covering the timestamp gap in
    units of `frame_samples` (default: the duration of the packet before the gap).  Returns (packets, timestamps)."""
    pk = {}
    base = None
    for d in datagrams:
        h = rtp_unpack(d)
        if payload_type is not None and h["payload_type"] != payload_type:
            continue
        if base is None:
            base = h["seq"]
        ext = ((h["seq"] - base + 0x8000) & 0xFFFF) - 0x8000        # signed distance from the first datagram: handles wrap-around
        pk.setdefault(ext, h)
    out, stamps, prev_end, prev_dur = [], [], None, frame_samples or 960
    for k in sorted(pk):
        h = pk[k]
        if prev_end is not None:
            gap = (h["timestamp"] - prev_end) & 0xFFFFFFFF
            if gap >= 0x80000000:
                gap = 0                                          # overlapping timestamps: nothing to conceal
            step = frame_samples or prev_dur
            while gap >= step > 0:
                out.append(b""); stamps.append(prev_end)
                prev_end = (prev_end + step) & 0xFFFFFFFF
                gap -= step
        dur = packet_duration_48k(h["payload"])
        out.append(h["payload"]); stamps.append(h["timestamp"])
        prev_end = (h["timestamp"] + dur) & 0xFFFFFFFF
        prev_dur = dur or prev_dur
    return out, stamps


# ---- Ogg Opus (RFC 3533 pages, RFC 7845 mapping) ------------------------------------------------------------------------------------
def _crc_table():
    t = []
    for i in range(256):
        r = i << 24
        for _ in range(8):
            r = ((r << 1) ^ 0x04C11DB7) & 0xFFFFFFFF if r & 0x80000000 else (r << 1) & 0xFFFFFFFF
        t.append(r)
    return t


_CRC = _crc_table()


def ogg_crc(page):
    """The Ogg page checksum: CRC-32 with polynomial 0x04c11db7, initial value 0, no reflection, no final xor (RFC 3533 section 6)."""
    c = 0
    for b in page:
        c = ((c << 8) & 0xFFFFFFFF) ^ _CRC[((c >> 24) ^ b) & 0xFF]
    return c


def _ogg_page(flags, granule, serial, seqno, segments, body):
    hdr = struct.pack("<4sBBqIIIB", b"OggS", 0, flags, granule, serial, seqno, 0, len(segments)) + bytes(segments)
    crc = ogg_crc(hdr + body)
    return hdr[:22] + struct.pack("<I", crc) + hdr[26:] + body


def _lacing(n):
    return [255] * (n // 255) + [n % 255]


def opus_head(channels, pre_skip=312, input_sample_rate=48000, output_gain_q8=0):
    """The identification header (RFC 7845 section 5.1), channel mapping family 0 (mono / stereo)."""
    if channels not in (1, 2):
        raise ValueError("mapping family 0 carries 1 or 2 channels")
    return struct.pack("<8sBBHIhB", b"OpusHead", 1, channels, pre_skip, input_sample_rate, output_gain_q8, 0)


def opus_tags(vendor="opus_codec_b200", comments=()):
    v = vendor.encode()
    out = b"OpusTags" + struct.pack("<I", len(v)) + v + struct.pack("<I", len(comments))
    for c in comments:
        cb = c.encode()
        out += struct.pack("<I", len(cb)) + cb
    return out


def ogg_write(packets, channels, pre_skip=312, input_sample_rate=48000, output_gain_q8=0, serial=0x42323030, packets_per_page=50, vendor="opus_codec_b200",
              comments=()):
    """Opus packets of one stream -> the bytes of an .opus file.  Granule positions count 48 kHz samples (pre-skip included) at the end
    of the last packet that completes on each page; the last page carries the end-of-stream flag."""
    out = bytearray()
    out += _ogg_page(2, 0, serial, 0, _lacing(19), opus_head(channels, pre_skip, input_sample_rate, output_gain_q8))
    tags = opus_tags(vendor, comments)
    seqno = 1
    segs = _lacing(len(tags))
    pos, first = 0, True
    while segs:                                                  # the comment header may span pages
        take = segs[:255]
        segs = segs[255:]
        n = sum(take)
        out += _ogg_page(0 if first else 1, 0 if not segs else -1, serial, seqno, take, tags[pos:pos + n])
        pos += n; seqno += 1; first = False
    granule, i = 0, 0
    while i < len(packets):
        segs, body, j = [], bytearray(), i
        while j < len(packets) and j - i < packets_per_page and len(segs) + len(_lacing(len(packets[j]))) <= 255:
            segs += _lacing(len(packets[j]))
            body += packets[j]
            granule += packet_duration_48k(packets[j])
            j += 1
        if j == i:
            raise ValueError("packet too large for one page")  # > 255 * 255 bytes: not an Opus packet
        out += _ogg_page(4 if j == len(packets) else 0, granule, serial, seqno, segs, bytes(body))
        seqno += 1
        i = j
    return bytes(out)


def ogg_read(data, check_crc=True):
    """bytes of an .opus file (one logical stream) -> dict(channels, pre_skip, input_sample_rate, output_gain_q8, mapping_family, vendor,
    comments, packets, granules, eos).  granules[k] = granule position of the page packet k completed on (-1 if it is not the last packet
    completing there).  Raises ValueError on bad capture pattern, CRC, page sequence or headers."""
    pos, partial, packets, granules, serial, expect, eos = 0, bytearray(), [], [], None, 0, False
    while pos < len(data):
        if data[pos:pos + 4] != b"OggS" or pos + 27 > len(data):
            raise ValueError("no OggS capture pattern at byte %d" % pos)
        _, ver, flags, granule, ser, seqno, crc, nseg = struct.unpack_from("<4sBBqIIIB", data, pos)
        segs = data[pos + 27:pos + 27 + nseg]
        body_len = sum(segs)
        end = pos + 27 + nseg + body_len
        if ver != 0 or len(segs) != nseg or end > len(data):
            raise ValueError("truncated Ogg page at byte %d" % pos)
        if check_crc and ogg_crc(bytes(data[pos:pos + 22]) + b"\0\0\0\0" + bytes(data[pos + 26:end])) != crc:
            raise ValueError("Ogg page checksum mismatch at byte %d" % pos)
        if serial is None:
            serial = ser
            if not flags & 2:
                raise ValueError("first page is not a beginning-of-stream page")
        elif ser != serial:
            raise ValueError("multiplexed / chained Ogg streams are not handled")
        if seqno != expect:
            raise ValueError("page %d missing" % expect)
        expect += 1
        if bool(flags & 1) != bool(partial):
            raise ValueError("continuation flag does not match the packet state")
        b, done = pos + 27 + nseg, []
        for s in segs:
            partial += data[b:b + s]
            b += s
            if s < 255:
                done.append(bytes(partial))
                partial = bytearray()
        for k, p in enumerate(done):
            packets.append(p)
            granules.append(granule if k == len(done) - 1 else -1)
        eos = bool(flags & 4)
        pos = end
    if len(packets) < 2 or packets[0][:8] != b"OpusHead" or len(packets[0]) < 19 or packets[1][:8] != b"OpusTags":
        raise ValueError("not an Ogg Opus stream")
    _, ver, ch, pre_skip, rate, gain, family = struct.unpack_from("<8sBBHIhB", packets[0], 0)
    if ver >> 4 != 0:
        raise ValueError("OpusHead version %d" % ver)
    t = packets[1]
    vl = struct.unpack_from("<I", t, 8)[0]
    vendor = t[12:12 + vl].decode(errors="replace")
    nc = struct.unpack_from("<I", t, 12 + vl)[0]
    comments, q = [], 16 + vl
    for _ in range(nc):
        cl = struct.unpack_from("<I", t, q)[0]
        comments.append(t[q + 4:q + 4 + cl].decode(errors="replace"))
        q += 4 + cl
    return dict(channels=ch, pre_skip=pre_skip, input_sample_rate=rate, output_gain_q8=gain, mapping_family=family, vendor=vendor, comments=comments,
                packets=packets[2:], granules=granules[2:], eos=eos)
