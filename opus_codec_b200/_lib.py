"""Loads libopus_b200.so (the CUDA sm_100a library behind include/opus_b200.h) with ctypes.

There is no CPU implementation behind this package: if the library is missing or no CUDA device is
usable, every decode entry point fails loudly.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.environ.get("OB_LIB") or os.path.join(HERE, "libopus_b200.so")      # OB_LIB: tuning aid, points at an alternative build of the same library
SRC = os.path.join(HERE, "csrc", "opus_b200.cu")
SRC_ENC = os.path.join(HERE, "csrc", "opus_b200_enc.cu")       # compiled with -fmad=false (see the file header)
SRC_PKT = os.path.join(HERE, "csrc", "opus_b200_pkt.cu")       # packet parse / pad / unpad, repacketizer (host objects + batched kernel)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]

# Every symbol include/opus_b200.h declares (tests check that the library exports all of them).
SYMBOLS = [
    "ob_decoder_create", "ob_decoder_destroy", "ob_decode_float", "ob_decode_float_multi", "ob_decode_float_device",
    "ob_decoder_final_range", "ob_decoder_reset", "ob_decoder_last_packet_duration", "ob_decoder_streams",
    "ob_decoder_channels", "ob_decoder_sample_rate", "ob_encoder_sample_rate", "ob_decoder_kernel_ms", "ob_decoder_launches", "ob_decoder_cuda_stream",
    "ob_decode_float_multi_async", "ob_decoder_wait", "ob_decode", "ob_decode_multi", "ob_decode_multi_async", "ob_encode", "ob_encode_multi", "ob_decoder_set_gain", "ob_decoder_get_gain", "ob_decoder_set_phase_inversion_disabled", "ob_decoder_get_phase_inversion_disabled", "ob_decoder_get_pitch", "ob_decoder_set_decode_fec", "ob_decoder_get_decode_fec",
    "ob_packet_get_nb_channels", "ob_packet_get_samples_per_frame", "ob_packet_get_bandwidth", "ob_packet_get_nb_frames",
    "ob_version", "ob_strerror", "ob_debug_ir_layout", "ob_decoder_debug_read_ir",
    "ob_packet_get_nb_samples", "ob_packet_has_lbrr", "ob_packet_parse", "ob_packet_pad", "ob_packet_unpad", "ob_multistream_packet_pad", "ob_multistream_packet_unpad", "ob_repacketizer_create", "ob_repacketizer_destroy", "ob_repacketizer_init", "ob_repacketizer_cat",
    "ob_repacketizer_get_nb_frames", "ob_repacketizer_out_range", "ob_repacketizer_out", "ob_repacketize_batch", "ob_repacketize_batch_device", "ob_pcm_soft_clip_batch",
    "ob_encoder_create", "ob_encoder_destroy", "ob_encode_float", "ob_encode_float_multi", "ob_encode_float_device",
    "ob_encoder_set_bitrate", "ob_encoder_get_bitrate", "ob_encoder_set_mapping", "ob_encoder_get_mapping", "ob_encoder_get_split", "ob_encoder_set_complexity", "ob_encoder_get_complexity",
    "ob_encoder_set_vbr", "ob_encoder_get_vbr", "ob_encoder_set_vbr_constraint", "ob_encoder_get_vbr_constraint",
    "ob_encoder_set_max_bandwidth", "ob_encoder_set_bandwidth", "ob_encoder_set_force_channels",
    "ob_encoder_set_packet_loss_perc", "ob_encoder_set_lsb_depth", "ob_encoder_final_range", "ob_encoder_reset",
    "ob_encoder_get_max_bandwidth", "ob_encoder_get_force_channels", "ob_encoder_get_packet_loss_perc", "ob_encoder_get_lsb_depth", "ob_encoder_set_signal", "ob_encoder_get_signal", "ob_encoder_set_prediction_disabled", "ob_encoder_get_prediction_disabled", "ob_encoder_set_phase_inversion_disabled", "ob_encoder_get_phase_inversion_disabled", "ob_encoder_set_dtx", "ob_encoder_get_dtx", "ob_encoder_set_inband_fec", "ob_encoder_get_inband_fec", "ob_encoder_set_expert_frame_duration", "ob_encoder_get_expert_frame_duration", "ob_encoder_get_lookahead", "ob_encoder_in_dtx", "ob_encoder_get_bandwidth",
    "ob_encoder_streams", "ob_encoder_channels", "ob_encoder_kernel_ms", "ob_encoder_launches", "ob_encoder_cuda_stream",
]

_lib = None


def build(verbose=False):
    """nvcc cross-compiles for sm_100a without a GPU; the .so is built in-tree so it travels to the GPU box."""
    srcs = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))]
    if os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(s) for s in srcs):
        return SO
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = ["-Xptxas", "-v"] if verbose else []
    o_dec, o_enc, o_pkt = os.path.join(HERE, "opus_b200.o"), os.path.join(HERE, "opus_b200_enc.o"), os.path.join(HERE, "opus_b200_pkt.o")
    # Decoder: global stores are evict-first (-dscm=cs).  Everything a decoder kernel stores (IR, spectrum tile, PCM, state) is read next by a LATER launch, after
    # gigabytes of other stores at bulk sizes, so keeping it in L2 buys nothing, while the symbol kernel's per-thread spill lines do get re-read.  Measured on B200:
    # symbol kernel 16.09 -> 15.83 ms per 819 200 mono frames, 7.1 -> 6.86 ms per 163 840 stereo frames (stereo decode 157.5 k -> 160.1 k audio-s/s); the opposite
    # experiment, global loads past L1 (-dlcm=cg), makes the symbol kernel 3.4x slower: its table loads live in L1.
    procs = [subprocess.Popen([nvcc] + NVCC_FLAGS + extra + ["-Xptxas", "-dscm=cs", "-c", "-o", o_dec, SRC]),
             subprocess.Popen([nvcc] + NVCC_FLAGS + extra + ["-fmad=false", "-c", "-o", o_enc, SRC_ENC]),
             subprocess.Popen([nvcc] + NVCC_FLAGS + extra + ["-c", "-o", o_pkt, SRC_PKT])]
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed")
    subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO, o_dec, o_enc, o_pkt], check=True)
    return SO


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO):
        raise ImportError("opus_codec_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). There is no CPU fallback." % SO)
    L = C.CDLL(SO)
    vp, i32, u8p, i32p, u32p, f32p = C.c_void_p, C.c_int32, C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_float)
    L.ob_decoder_create.argtypes = [i32, i32, i32, i32, i32, i32p]; L.ob_decoder_create.restype = vp
    L.ob_decoder_destroy.argtypes = [vp]; L.ob_decoder_destroy.restype = None
    L.ob_decode_float.argtypes = [vp, vp, vp, vp, vp, i32, vp]; L.ob_decode_float.restype = i32
    L.ob_decode_float_multi.argtypes = [vp, i32, vp, vp, vp, vp, i32, vp, vp]; L.ob_decode_float_multi.restype = i32
    L.ob_decode_float_multi_async.argtypes = [vp, i32, vp, vp, vp, vp, i32, vp, vp]; L.ob_decode_float_multi_async.restype = i32
    L.ob_decoder_wait.argtypes = [vp, i32]; L.ob_decoder_wait.restype = i32
    L.ob_encode.argtypes = [vp, vp, i32, vp, i32, vp]; L.ob_encode.restype = i32
    L.ob_encode_multi.argtypes = [vp, i32, vp, i32, vp, i32, vp, vp]; L.ob_encode_multi.restype = i32
    L.ob_decode.argtypes = [vp, vp, vp, vp, vp, i32, vp]; L.ob_decode.restype = i32
    L.ob_decode_multi.argtypes = [vp, i32, vp, vp, vp, vp, i32, vp, vp]; L.ob_decode_multi.restype = i32
    L.ob_decode_multi_async.argtypes = [vp, i32, vp, vp, vp, vp, i32, vp, vp]; L.ob_decode_multi_async.restype = i32
    L.ob_decode_float_device.argtypes = [vp, i32, vp, vp, vp, vp, i32, vp, vp, i32]; L.ob_decode_float_device.restype = i32
    L.ob_decoder_final_range.argtypes = [vp, vp]; L.ob_decoder_final_range.restype = i32
    L.ob_debug_ir_layout.argtypes = [vp, i32]; L.ob_debug_ir_layout.restype = i32
    L.ob_decoder_debug_read_ir.argtypes = [vp, i32, i32, vp, i32]; L.ob_decoder_debug_read_ir.restype = i32
    L.ob_decoder_reset.argtypes = [vp, vp, i32]; L.ob_decoder_reset.restype = i32
    L.ob_decoder_last_packet_duration.argtypes = [vp, vp]; L.ob_decoder_last_packet_duration.restype = i32
    for n in ("ob_decoder_set_gain", "ob_decoder_set_phase_inversion_disabled"):
        getattr(L, n).argtypes = [vp, i32]; getattr(L, n).restype = i32
    for n in ("ob_decoder_get_gain", "ob_decoder_get_phase_inversion_disabled"):
        getattr(L, n).argtypes = [vp, i32p]; getattr(L, n).restype = i32
    L.ob_decoder_streams.argtypes = [vp]; L.ob_decoder_streams.restype = i32
    L.ob_decoder_channels.argtypes = [vp]; L.ob_decoder_channels.restype = i32
    L.ob_decoder_kernel_ms.argtypes = [vp, f32p]; L.ob_decoder_kernel_ms.restype = i32
    L.ob_decoder_launches.argtypes = [vp]; L.ob_decoder_launches.restype = C.c_int64
    L.ob_decoder_cuda_stream.argtypes = [vp]; L.ob_decoder_cuda_stream.restype = vp
    for n in ("ob_packet_get_nb_channels", "ob_packet_get_bandwidth"):
        getattr(L, n).argtypes = [vp]; getattr(L, n).restype = i32
    L.ob_packet_get_samples_per_frame.argtypes = [vp, i32]; L.ob_packet_get_samples_per_frame.restype = i32
    L.ob_packet_get_nb_frames.argtypes = [vp, i32]; L.ob_packet_get_nb_frames.restype = i32
    L.ob_encoder_get_max_bandwidth.argtypes = [vp, vp]; L.ob_encoder_get_max_bandwidth.restype = i32
    L.ob_encoder_get_force_channels.argtypes = [vp, vp]; L.ob_encoder_get_force_channels.restype = i32
    L.ob_encoder_get_packet_loss_perc.argtypes = [vp, vp]; L.ob_encoder_get_packet_loss_perc.restype = i32
    L.ob_encoder_get_lsb_depth.argtypes = [vp, vp]; L.ob_encoder_get_lsb_depth.restype = i32
    L.ob_encoder_set_mapping.argtypes = [vp, i32]; L.ob_encoder_set_mapping.restype = i32
    L.ob_encoder_get_mapping.argtypes = [vp, vp]; L.ob_encoder_get_mapping.restype = i32
    L.ob_encoder_get_split.argtypes = [vp, vp]; L.ob_encoder_get_split.restype = i32
    L.ob_encoder_set_signal.argtypes = [vp, i32]; L.ob_encoder_set_signal.restype = i32
    L.ob_encoder_get_signal.argtypes = [vp, vp]; L.ob_encoder_get_signal.restype = i32
    L.ob_encoder_set_prediction_disabled.argtypes = [vp, i32]; L.ob_encoder_set_prediction_disabled.restype = i32
    L.ob_encoder_get_prediction_disabled.argtypes = [vp, vp]; L.ob_encoder_get_prediction_disabled.restype = i32
    L.ob_encoder_set_phase_inversion_disabled.argtypes = [vp, i32]; L.ob_encoder_set_phase_inversion_disabled.restype = i32
    L.ob_encoder_get_phase_inversion_disabled.argtypes = [vp, vp]; L.ob_encoder_get_phase_inversion_disabled.restype = i32
    L.ob_encoder_set_dtx.argtypes = [vp, i32]; L.ob_encoder_set_dtx.restype = i32
    L.ob_encoder_get_dtx.argtypes = [vp, vp]; L.ob_encoder_get_dtx.restype = i32
    L.ob_encoder_set_inband_fec.argtypes = [vp, i32]; L.ob_encoder_set_inband_fec.restype = i32
    L.ob_encoder_get_inband_fec.argtypes = [vp, vp]; L.ob_encoder_get_inband_fec.restype = i32
    L.ob_encoder_set_expert_frame_duration.argtypes = [vp, i32]; L.ob_encoder_set_expert_frame_duration.restype = i32
    L.ob_encoder_get_expert_frame_duration.argtypes = [vp, vp]; L.ob_encoder_get_expert_frame_duration.restype = i32
    L.ob_encoder_get_lookahead.argtypes = [vp, vp]; L.ob_encoder_get_lookahead.restype = i32
    L.ob_encoder_in_dtx.argtypes = [vp, vp]; L.ob_encoder_in_dtx.restype = i32
    L.ob_decoder_get_pitch.argtypes = [vp, vp]; L.ob_decoder_get_pitch.restype = i32
    L.ob_decoder_set_decode_fec.argtypes = [vp, i32]; L.ob_decoder_set_decode_fec.restype = i32
    L.ob_decoder_get_decode_fec.argtypes = [vp, vp]; L.ob_decoder_get_decode_fec.restype = i32
    L.ob_encoder_get_bandwidth.argtypes = [vp, vp]; L.ob_encoder_get_bandwidth.restype = i32
    L.ob_decoder_sample_rate.argtypes = [vp]; L.ob_decoder_sample_rate.restype = i32
    L.ob_encoder_sample_rate.argtypes = [vp]; L.ob_encoder_sample_rate.restype = i32
    L.ob_packet_get_nb_samples.argtypes = [vp, i32, i32]; L.ob_packet_get_nb_samples.restype = i32
    L.ob_packet_has_lbrr.argtypes = [vp, i32]; L.ob_packet_has_lbrr.restype = i32
    L.ob_packet_parse.argtypes = [vp, i32, vp, vp, vp, vp]; L.ob_packet_parse.restype = i32
    L.ob_packet_pad.argtypes = [vp, i32, i32]; L.ob_packet_pad.restype = i32
    L.ob_packet_unpad.argtypes = [vp, i32]; L.ob_packet_unpad.restype = i32
    L.ob_multistream_packet_pad.argtypes = [vp, i32, i32, i32]; L.ob_multistream_packet_pad.restype = i32
    L.ob_multistream_packet_unpad.argtypes = [vp, i32, i32]; L.ob_multistream_packet_unpad.restype = i32
    L.ob_repacketizer_create.argtypes = []; L.ob_repacketizer_create.restype = vp
    L.ob_repacketizer_destroy.argtypes = [vp]; L.ob_repacketizer_destroy.restype = None
    L.ob_repacketizer_init.argtypes = [vp]; L.ob_repacketizer_init.restype = None
    L.ob_repacketizer_cat.argtypes = [vp, vp, i32]; L.ob_repacketizer_cat.restype = i32
    L.ob_repacketizer_get_nb_frames.argtypes = [vp]; L.ob_repacketizer_get_nb_frames.restype = i32
    L.ob_repacketizer_out_range.argtypes = [vp, i32, i32, vp, i32]; L.ob_repacketizer_out_range.restype = i32
    L.ob_repacketizer_out.argtypes = [vp, vp, i32]; L.ob_repacketizer_out.restype = i32
    L.ob_repacketize_batch.argtypes = [i32, i32, i32, vp, vp, vp, i32, i32, vp, i32, vp]; L.ob_repacketize_batch.restype = i32
    L.ob_pcm_soft_clip_batch.argtypes = [i32, i32, vp, i32, i32, vp]; L.ob_pcm_soft_clip_batch.restype = i32
    L.ob_repacketize_batch_device.argtypes = [i32, i32, vp, vp, vp, i32, i32, vp, i32, vp, vp]; L.ob_repacketize_batch_device.restype = i32
    L.ob_encoder_create.argtypes = [i32, i32, i32, i32, i32, i32, i32p]; L.ob_encoder_create.restype = vp
    L.ob_encoder_destroy.argtypes = [vp]; L.ob_encoder_destroy.restype = None
    L.ob_encode_float.argtypes = [vp, vp, i32, vp, i32, vp]; L.ob_encode_float.restype = i32
    L.ob_encode_float_multi.argtypes = [vp, i32, vp, i32, vp, i32, vp, vp]; L.ob_encode_float_multi.restype = i32
    L.ob_encode_float_device.argtypes = [vp, i32, vp, i32, vp, i32, vp, vp, i32]; L.ob_encode_float_device.restype = i32
    for n in ("bitrate", "complexity", "vbr", "vbr_constraint"):
        getattr(L, "ob_encoder_set_" + n).argtypes = [vp, i32]; getattr(L, "ob_encoder_set_" + n).restype = i32
        getattr(L, "ob_encoder_get_" + n).argtypes = [vp, i32p]; getattr(L, "ob_encoder_get_" + n).restype = i32
    for n in ("max_bandwidth", "bandwidth", "force_channels", "packet_loss_perc", "lsb_depth"):
        getattr(L, "ob_encoder_set_" + n).argtypes = [vp, i32]; getattr(L, "ob_encoder_set_" + n).restype = i32
    L.ob_encoder_final_range.argtypes = [vp, vp]; L.ob_encoder_final_range.restype = i32
    L.ob_encoder_reset.argtypes = [vp, vp, i32]; L.ob_encoder_reset.restype = i32
    L.ob_encoder_streams.argtypes = [vp]; L.ob_encoder_streams.restype = i32
    L.ob_encoder_channels.argtypes = [vp]; L.ob_encoder_channels.restype = i32
    L.ob_encoder_kernel_ms.argtypes = [vp, f32p]; L.ob_encoder_kernel_ms.restype = i32
    L.ob_encoder_launches.argtypes = [vp]; L.ob_encoder_launches.restype = C.c_int64
    L.ob_encoder_cuda_stream.argtypes = [vp]; L.ob_encoder_cuda_stream.restype = vp
    L.ob_version.restype = C.c_char_p
    L.ob_strerror.argtypes = [i32]; L.ob_strerror.restype = C.c_char_p
    _lib = L
    return L
