"""Host-side mirror of the crate's per-stream wrappers for a BATCH of streams.

`BatchDecoder` keeps the semantics of `Decoder` (reference src/decoder.rs:35-346) for every stream of the
batch -- `decode_float` returns samples per channel, PCM is interleaved, `final_range()` follows each call,
`reset()` is OPUS_RESET_STATE -- and calls the CUDA library through the C ABI of include/opus_b200.h.
Error values are the crate's (src/error.rs:36-62).
"""
import ctypes as C

import numpy as np

from . import _lib

OK, BAD_ARG, BUFFER_TOO_SMALL, INTERNAL_ERROR, INVALID_PACKET, UNIMPLEMENTED, INVALID_STATE, ALLOC_FAIL = 0, -1, -2, -3, -4, -5, -6, -7
MAX_FRAME_SAMPLES_48KHZ = 5760          # src/constants.rs:8


class OpusError(Exception):
    """Mirror of opus_codec::Error (src/error.rs:8-34)."""
    NAMES = {-1: "BadArg", -2: "BufferTooSmall", -3: "InternalError", -4: "InvalidPacket", -5: "Unimplemented",
             -6: "InvalidState", -7: "AllocFail"}

    def __init__(self, code):
        self.code = int(code)
        super().__init__("%s (%d): %s" % (self.NAMES.get(self.code, "Unknown"), self.code,
                                          _lib.lib().ob_strerror(self.code).decode()))


def _check(code):
    if code != OK:
        raise OpusError(code)


def _vp(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def pack_packets(packets):
    """list (streams) of list (frames) of bytes -> (buf u8, offsets i32 [S,F], lens i32 [S,F])."""
    S = len(packets)
    F = len(packets[0])
    lens = np.array([[len(p) for p in row] for row in packets], np.int32).reshape(S, F)
    offsets = np.zeros(S * F, np.int64)
    np.cumsum(lens.reshape(-1)[:-1], out=offsets[1:])
    buf = np.frombuffer(b"".join(b"".join(row) for row in packets), np.uint8).copy() if lens.sum() else np.zeros(1, np.uint8)
    return buf, offsets.astype(np.int32).reshape(S, F), lens


class BatchDecoder:
    """n_streams independent 48 kHz CELT-only Opus decoders on one B200 (mirror of Decoder, src/decoder.rs)."""

    def __init__(self, n_streams, sample_rate=48000, channels=1, device=0, max_frames=1):
        self._L = _lib.lib()
        err = C.c_int32(0)
        self._h = self._L.ob_decoder_create(n_streams, sample_rate, channels, device, max_frames, C.byref(err))
        if not self._h:
            raise OpusError(err.value)
        self.n_streams, self.channels, self.sample_rate, self.max_frames, self.device = n_streams, channels, sample_rate, max_frames, device

    def close(self):
        if getattr(self, "_h", None):
            self._L.ob_decoder_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- decode -------------------------------------------------------------------------------------------
    @staticmethod
    def _check_packet_table(packets, offsets, lens):
        """The C side copies lens[i] bytes from packets + offsets[i]: every (offset, length) pair must lie inside the packet buffer."""
        pos = lens > 0
        if (lens < 0).any() or (offsets[pos] < 0).any():
            raise OpusError(BAD_ARG)
        if pos.any() and int((offsets[pos].astype(np.int64) + lens[pos]).max()) > packets.size:
            raise OpusError(BAD_ARG)

    def decode_float_multi(self, packets, offsets, lens, frame_size, out=None):
        """packets: u8 buffer; offsets/lens: i32 [S, F].  Returns (pcm f32 [S, F, frame_size*channels], samples i32 [S, F],
        final ranges u32 [S, F]).  frame_size is the per-channel capacity of each slot (output.len()/channels in
        Decoder::decode_float, src/decoder.rs:149)."""
        offsets = np.ascontiguousarray(offsets, np.int32)
        lens = np.ascontiguousarray(lens, np.int32)
        packets = np.ascontiguousarray(packets, np.uint8)
        if offsets.shape != lens.shape or offsets.ndim != 2 or offsets.shape[0] != self.n_streams:
            raise OpusError(BAD_ARG)
        if frame_size <= 0 or frame_size > MAX_FRAME_SAMPLES_48KHZ:
            raise OpusError(BAD_ARG)
        S, F = offsets.shape
        self._check_packet_table(packets, offsets, lens)
        if out is not None and (not isinstance(out, np.ndarray) or out.dtype != np.float32 or not out.flags["C_CONTIGUOUS"] or not out.flags["WRITEABLE"]
                                or out.size < S * F * frame_size * self.channels):
            raise OpusError(BAD_ARG)
        pcm = out if out is not None else np.zeros((S, F, frame_size * self.channels), np.float32)
        samples = np.zeros((S, F), np.int32)
        ranges = np.zeros((S, F), np.uint32)
        _check(self._L.ob_decode_float_multi(self._h, F, _vp(packets), _vp(offsets), _vp(lens), _vp(pcm), frame_size,
                                             _vp(samples), _vp(ranges)))
        return pcm, samples, ranges

    def decode_multi(self, packets, offsets, lens, frame_size):
        """int16 form (ob_decode_multi; Decoder::decode, src/decoder.rs:75-127): soft-clipped, rounded PCM i16 [S, F, frame_size*channels]."""
        offsets = np.ascontiguousarray(offsets, np.int32)
        lens = np.ascontiguousarray(lens, np.int32)
        packets = np.ascontiguousarray(packets, np.uint8)
        if offsets.shape != lens.shape or offsets.ndim != 2 or offsets.shape[0] != self.n_streams or frame_size <= 0 or frame_size > MAX_FRAME_SAMPLES_48KHZ:
            raise OpusError(BAD_ARG)
        S, F = offsets.shape
        self._check_packet_table(packets, offsets, lens)
        pcm = np.zeros((S, F, frame_size * self.channels), np.int16)
        samples = np.zeros((S, F), np.int32)
        ranges = np.zeros((S, F), np.uint32)
        _check(self._L.ob_decode_multi(self._h, F, _vp(packets), _vp(offsets), _vp(lens), _vp(pcm), frame_size, _vp(samples), _vp(ranges)))
        return pcm, samples, ranges

    def decode_float_multi_async(self, packets, offsets, lens, frame_size, pcm, samples, ranges):
        """Pipelined form (ob_decode_float_multi_async): enqueue and return; all arrays are caller-owned, C-contiguous numpy arrays
        (ideally over pinned memory) that must stay alive and untouched until wait() says the call has completed."""
        S, F = offsets.shape
        if offsets.dtype != np.int32 or lens.dtype != np.int32 or packets.dtype != np.uint8 or pcm.dtype != np.float32 \
                or samples.dtype != np.int32 or ranges.dtype != np.uint32 or S != self.n_streams or pcm.size < S * F * frame_size * self.channels \
                or lens.shape != offsets.shape or samples.size < S * F or ranges.size < S * F \
                or not all(a.flags["C_CONTIGUOUS"] for a in (packets, offsets, lens, pcm, samples, ranges)):
            raise OpusError(BAD_ARG)
        self._check_packet_table(packets, offsets, lens)
        _check(self._L.ob_decode_float_multi_async(self._h, F, _vp(packets), _vp(offsets), _vp(lens), _vp(pcm), frame_size,
                                                   _vp(samples), _vp(ranges)))

    def wait(self, keep_in_flight=0):
        """ob_decoder_wait: 0 = every outstanding call has completed; 1 = every call but the most recent one."""
        _check(self._L.ob_decoder_wait(self._h, int(keep_in_flight)))

    def decode_float(self, packets, frame_size):
        """packets: one bytes object per stream.  Returns (pcm [S, frame_size*channels], samples [S])."""
        if len(packets) != self.n_streams:
            raise OpusError(BAD_ARG)
        buf, offsets, lens = pack_packets([[p] for p in packets])
        pcm, samples, _ = self.decode_float_multi(buf, offsets, lens, frame_size)
        return pcm[:, 0], samples[:, 0]

    # -- CTLs ------------------------------------------------------------------------------------------------
    def final_range(self):
        out = np.zeros(self.n_streams, np.uint32)
        _check(self._L.ob_decoder_final_range(self._h, _vp(out)))
        return out

    def last_packet_duration(self):
        out = np.zeros(self.n_streams, np.int32)
        _check(self._L.ob_decoder_last_packet_duration(self._h, _vp(out)))
        return out

    def set_gain(self, q8_db):
        """Decoder::set_gain (src/decoder.rs:318-320): Q8 dB, whole batch."""
        _check(self._L.ob_decoder_set_gain(self._h, int(q8_db)))

    def gain(self):
        v = C.c_int32(0)
        _check(self._L.ob_decoder_get_gain(self._h, C.byref(v)))
        return v.value

    def pitch(self):
        """OPUS_GET_PITCH per stream (Decoder::get_pitch)."""
        out = np.zeros(self.n_streams, np.int32)
        _check(self._L.ob_decoder_get_pitch(self._h, _vp(out)))
        return out

    def set_decode_fec(self, on):
        """decode_fec of the calls that follow (Decoder::decode(.., fec)): CELT-only packets are then concealed like lost ones."""
        _check(self._L.ob_decoder_set_decode_fec(self._h, int(bool(on))))

    def set_phase_inversion_disabled(self, disabled):
        """Decoder::set_phase_inversion_disabled (src/decoder.rs:341-346)."""
        _check(self._L.ob_decoder_set_phase_inversion_disabled(self._h, 1 if disabled else 0))

    def phase_inversion_disabled(self):
        v = C.c_int32(0)
        _check(self._L.ob_decoder_get_phase_inversion_disabled(self._h, C.byref(v)))
        return bool(v.value)

    def reset(self, streams=None):
        if streams is None:
            _check(self._L.ob_decoder_reset(self._h, None, 0))
        else:
            idx = np.ascontiguousarray(streams, np.int32)
            _check(self._L.ob_decoder_reset(self._h, _vp(idx), idx.size))

    def kernel_ms(self):
        ms = (C.c_float * 3)()
        _check(self._L.ob_decoder_kernel_ms(self._h, ms))
        return [float(v) for v in ms]

    def launches(self):
        return int(self._L.ob_decoder_launches(self._h))

    @property
    def handle(self):
        return self._h


def version():
    return _lib.lib().ob_version().decode()


APPLICATION_VOIP, APPLICATION_AUDIO, APPLICATION_RESTRICTED_LOWDELAY = 2048, 2049, 2051      # src/types.rs Application
BITRATE_AUTO, BITRATE_MAX = -1000, -1                                                       # src/types.rs Bitrate
BANDWIDTH_NARROWBAND, BANDWIDTH_WIDEBAND, BANDWIDTH_SUPERWIDEBAND, BANDWIDTH_FULLBAND = 1101, 1103, 1104, 1105


MAP_AUTO, MAP_WARP, MAP_THREAD = 0, 1, 2


class BatchEncoder:
    """n_streams independent 48 kHz CELT-only Opus encoders on one B200 (mirror of Encoder, reference src/encoder.rs:40-699).
    CTLs apply to the whole batch."""

    def __init__(self, n_streams, sample_rate=48000, channels=1, application=APPLICATION_RESTRICTED_LOWDELAY, device=0, max_frames=1):
        self._L = _lib.lib()
        err = C.c_int32(0)
        self._h = self._L.ob_encoder_create(n_streams, sample_rate, channels, application, device, max_frames, C.byref(err))
        if not self._h:
            raise OpusError(err.value)
        self.n_streams, self.channels, self.sample_rate, self.max_frames, self.device = n_streams, channels, sample_rate, max_frames, device

    def close(self):
        if getattr(self, "_h", None):
            self._L.ob_encoder_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- CTLs (src/encoder.rs:545-652) ------------------------------------------------------------------------------
    def set_bitrate(self, bitrate):
        _check(self._L.ob_encoder_set_bitrate(self._h, int(bitrate)))

    def bitrate(self):
        v = C.c_int32(0); _check(self._L.ob_encoder_get_bitrate(self._h, C.byref(v))); return v.value

    def set_complexity(self, c):
        _check(self._L.ob_encoder_set_complexity(self._h, int(c)))

    def complexity(self):
        v = C.c_int32(0); _check(self._L.ob_encoder_get_complexity(self._h, C.byref(v))); return v.value

    def set_vbr(self, on):
        _check(self._L.ob_encoder_set_vbr(self._h, int(bool(on))))

    def vbr(self):
        v = C.c_int32(0); _check(self._L.ob_encoder_get_vbr(self._h, C.byref(v))); return bool(v.value)

    def set_vbr_constraint(self, on):
        _check(self._L.ob_encoder_set_vbr_constraint(self._h, int(bool(on))))

    def vbr_constraint(self):
        v = C.c_int32(0); _check(self._L.ob_encoder_get_vbr_constraint(self._h, C.byref(v))); return bool(v.value)

    def set_max_bandwidth(self, bw):
        _check(self._L.ob_encoder_set_max_bandwidth(self._h, int(bw)))

    def set_bandwidth(self, bw):
        _check(self._L.ob_encoder_set_bandwidth(self._h, int(bw)))

    def set_force_channels(self, ch):
        _check(self._L.ob_encoder_set_force_channels(self._h, int(ch)))

    def set_packet_loss_perc(self, p):
        _check(self._L.ob_encoder_set_packet_loss_perc(self._h, int(p)))

    def set_lsb_depth(self, d):
        _check(self._L.ob_encoder_set_lsb_depth(self._h, int(d)))

    def set_mapping(self, m):
        """MAP_AUTO (0), MAP_WARP (1: one warp per stream, low latency), MAP_THREAD (2: one lane per stream, bulk) -- include/opus_b200.h."""
        _check(self._L.ob_encoder_set_mapping(self._h, int(m)))

    def mapping(self): return self._get(self._L.ob_encoder_get_mapping)

    def split(self):
        """Streams [0, n) of the last call were coded one warp per stream, the rest one lane per stream (ob_encoder_get_split)."""
        return self._get(self._L.ob_encoder_get_split)

    def _get(self, fn):
        v = C.c_int32(0)
        _check(fn(self._h, C.byref(v)))
        return v.value

    # the remaining CTLs of Encoder (src/encoder.rs): same names, same value ranges, BadArg outside them
    def max_bandwidth(self): return self._get(self._L.ob_encoder_get_max_bandwidth)
    def force_channels(self): return self._get(self._L.ob_encoder_get_force_channels)
    def packet_loss_perc(self): return self._get(self._L.ob_encoder_get_packet_loss_perc)
    def lsb_depth(self): return self._get(self._L.ob_encoder_get_lsb_depth)
    def set_signal(self, v): _check(self._L.ob_encoder_set_signal(self._h, int(v)))
    def signal(self): return self._get(self._L.ob_encoder_get_signal)
    def set_prediction_disabled(self, on): _check(self._L.ob_encoder_set_prediction_disabled(self._h, int(bool(on))))
    def prediction_disabled(self): return bool(self._get(self._L.ob_encoder_get_prediction_disabled))
    def set_phase_inversion_disabled(self, on): _check(self._L.ob_encoder_set_phase_inversion_disabled(self._h, int(bool(on))))
    def phase_inversion_disabled(self): return bool(self._get(self._L.ob_encoder_get_phase_inversion_disabled))
    def set_dtx(self, on): _check(self._L.ob_encoder_set_dtx(self._h, int(bool(on))))
    def dtx(self): return bool(self._get(self._L.ob_encoder_get_dtx))
    def set_inband_fec(self, mode): _check(self._L.ob_encoder_set_inband_fec(self._h, int(mode)))
    def inband_fec(self): return self._get(self._L.ob_encoder_get_inband_fec)
    def set_expert_frame_duration(self, v): _check(self._L.ob_encoder_set_expert_frame_duration(self._h, int(v)))
    def expert_frame_duration(self): return self._get(self._L.ob_encoder_get_expert_frame_duration)
    def lookahead(self): return self._get(self._L.ob_encoder_get_lookahead)

    def bandwidth(self):
        """OPUS_GET_BANDWIDTH per stream (Encoder::bandwidth): the bandwidth of each stream's last packet."""
        out = np.zeros(self.n_streams, np.int32)
        _check(self._L.ob_encoder_get_bandwidth(self._h, _vp(out)))
        return out

    def in_dtx(self):
        out = np.zeros(self.n_streams, np.int32)
        _check(self._L.ob_encoder_in_dtx(self._h, _vp(out)))
        return out.astype(bool)

    # -- encode ------------------------------------------------------------------------------------------------------
    def encode_float_multi(self, pcm, frame_size, max_bytes=1276):
        """pcm: f32 [S, F, frame_size*channels] in [-1,1].  Returns (packets u8 [S, F, max_bytes], lens i32 [S, F], ranges u32 [S, F]);
        lens < 0 are OPUS_* codes for that (stream, frame)."""
        pcm = np.ascontiguousarray(pcm, np.float32)
        if pcm.ndim != 3 or pcm.shape[0] != self.n_streams or pcm.shape[2] != frame_size * self.channels:
            raise OpusError(BAD_ARG)
        S, F = pcm.shape[:2]
        out = np.zeros((S, F, max_bytes), np.uint8)
        lens = np.zeros((S, F), np.int32)
        ranges = np.zeros((S, F), np.uint32)
        _check(self._L.ob_encode_float_multi(self._h, F, _vp(pcm), frame_size, _vp(out), max_bytes, _vp(lens), _vp(ranges)))
        return out, lens, ranges

    def encode_multi(self, pcm, frame_size, max_bytes=1276):
        """int16 form (ob_encode_multi; Encoder::encode, src/encoder.rs:80-127): pcm i16 [S, F, frame_size*channels]."""
        pcm = np.ascontiguousarray(pcm, np.int16)
        if pcm.ndim != 3 or pcm.shape[0] != self.n_streams or pcm.shape[2] != frame_size * self.channels:
            raise OpusError(BAD_ARG)
        S, F = pcm.shape[:2]
        out = np.zeros((S, F, max_bytes), np.uint8)
        lens = np.zeros((S, F), np.int32)
        ranges = np.zeros((S, F), np.uint32)
        _check(self._L.ob_encode_multi(self._h, F, _vp(pcm), frame_size, _vp(out), max_bytes, _vp(lens), _vp(ranges)))
        return out, lens, ranges

    def encode_float(self, pcm, max_bytes=1276):
        """pcm: f32 [S, frame_size*channels]; frame_size = input.len()/channels as in Encoder::encode_float (src/encoder.rs:215-247).
        Returns a list of bytes objects (one packet per stream) and the i32 length/status array."""
        pcm = np.ascontiguousarray(pcm, np.float32)
        if pcm.ndim != 2 or pcm.shape[1] % self.channels:
            raise OpusError(BAD_ARG)
        frame_size = pcm.shape[1] // self.channels
        out, lens, _ = self.encode_float_multi(pcm[:, None, :], frame_size, max_bytes)
        return [bytes(out[s, 0, :max(0, lens[s, 0])]) for s in range(self.n_streams)], lens[:, 0]

    def final_range(self):
        out = np.zeros(self.n_streams, np.uint32)
        _check(self._L.ob_encoder_final_range(self._h, _vp(out)))
        return out

    def reset(self, streams=None):
        if streams is None:
            _check(self._L.ob_encoder_reset(self._h, None, 0))
        else:
            idx = np.ascontiguousarray(streams, np.int32)
            _check(self._L.ob_encoder_reset(self._h, _vp(idx), idx.size))

    def kernel_ms(self):
        ms = C.c_float(0)
        _check(self._L.ob_encoder_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def launches(self):
        return int(self._L.ob_encoder_launches(self._h))

    @property
    def handle(self):
        return self._h
