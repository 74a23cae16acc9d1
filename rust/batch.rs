//! `BatchDecoder` / `BatchEncoder`: safe wrappers over the C ABI of `libopus_b200.so` (include/opus_b200.h).
//!
//! REVIEW-ONLY in this repository: the build image has no `cargo`/`rustc`, so this file is not compiled here.
//! It is the file a maintainer of the `opus-codec` crate would add as `src/batch.rs` (next to `src/decoder.rs`),
//! with the `extern "C"` block below added beside the bindgen output in `src/bindings.rs`.
//! It reuses the crate's own `Error`, `Result`, `Channels` and `SampleRate` types (src/error.rs, src/types.rs).

use crate::error::{Error, Result};
use crate::types::{Application, Bitrate, Channels, Complexity, SampleRate};
use std::ptr::NonNull;

#[repr(C)]
pub struct ObDecoder {
    _unused: [u8; 0],
}

unsafe extern "C" {
    pub fn ob_decoder_create(n_streams: i32, fs: i32, channels: i32, device: i32, max_frames: i32, error: *mut i32) -> *mut ObDecoder;
    pub fn ob_decoder_destroy(dec: *mut ObDecoder);
    pub fn ob_decode_float(dec: *mut ObDecoder, packets: *const u8, offsets: *const i32, lens: *const i32,
                           pcm_out: *mut f32, frame_size: i32, samples_out: *mut i32) -> i32;
    pub fn ob_decode_float_multi(dec: *mut ObDecoder, n_frames: i32, packets: *const u8, offsets: *const i32, lens: *const i32,
                                 pcm_out: *mut f32, frame_size: i32, samples_out: *mut i32, ranges_out: *mut u32) -> i32;
    pub fn ob_decoder_final_range(dec: *mut ObDecoder, out: *mut u32) -> i32;
    pub fn ob_decoder_reset(dec: *mut ObDecoder, idx: *const i32, n: i32) -> i32;
    pub fn ob_decoder_last_packet_duration(dec: *mut ObDecoder, out: *mut i32) -> i32;
    pub fn ob_decoder_set_gain(dec: *mut ObDecoder, gain_q8: i32) -> i32;
    pub fn ob_decoder_set_phase_inversion_disabled(dec: *mut ObDecoder, disabled: i32) -> i32;
    pub fn ob_decoder_set_decode_fec(dec: *mut ObDecoder, decode_fec: i32) -> i32;

    pub fn ob_encoder_create(n_streams: i32, fs: i32, channels: i32, application: i32, device: i32, max_frames: i32,
                             error: *mut i32) -> *mut ObEncoder;
    pub fn ob_encoder_destroy(enc: *mut ObEncoder);
    pub fn ob_encode_float(enc: *mut ObEncoder, pcm: *const f32, frame_size: i32, out: *mut u8, max_bytes: i32,
                           lens_out: *mut i32) -> i32;
    pub fn ob_encode_float_multi(enc: *mut ObEncoder, n_frames: i32, pcm: *const f32, frame_size: i32, out: *mut u8,
                                 max_bytes: i32, lens_out: *mut i32, ranges_out: *mut u32) -> i32;
    pub fn ob_encoder_set_bitrate(enc: *mut ObEncoder, bitrate: i32) -> i32;
    pub fn ob_encoder_set_mapping(enc: *mut ObEncoder, mapping: i32) -> i32;      // 0 auto, 1 one warp per stream, 2 one lane per stream
    pub fn ob_encoder_get_mapping(enc: *mut ObEncoder, value: *mut i32) -> i32;
    pub fn ob_encoder_get_split(enc: *mut ObEncoder, n_warp_streams: *mut i32) -> i32;
    pub fn ob_encoder_get_bitrate(enc: *mut ObEncoder, value: *mut i32) -> i32;
    pub fn ob_encoder_set_complexity(enc: *mut ObEncoder, complexity: i32) -> i32;
    pub fn ob_encoder_set_vbr(enc: *mut ObEncoder, vbr: i32) -> i32;
    pub fn ob_encoder_set_vbr_constraint(enc: *mut ObEncoder, cvbr: i32) -> i32;
    pub fn ob_encoder_set_signal(enc: *mut ObEncoder, signal: i32) -> i32;
    pub fn ob_encoder_set_prediction_disabled(enc: *mut ObEncoder, disabled: i32) -> i32;
    pub fn ob_encoder_set_phase_inversion_disabled(enc: *mut ObEncoder, disabled: i32) -> i32;
    pub fn ob_encoder_set_dtx(enc: *mut ObEncoder, enabled: i32) -> i32;
    pub fn ob_encoder_in_dtx(enc: *mut ObEncoder, out: *mut i32) -> i32;
    pub fn ob_encoder_set_inband_fec(enc: *mut ObEncoder, mode: i32) -> i32;
    pub fn ob_encoder_set_expert_frame_duration(enc: *mut ObEncoder, duration: i32) -> i32;
    pub fn ob_encoder_get_lookahead(enc: *mut ObEncoder, value: *mut i32) -> i32;
    pub fn ob_encoder_final_range(enc: *mut ObEncoder, out: *mut u32) -> i32;
    pub fn ob_encoder_reset(enc: *mut ObEncoder, idx: *const i32, n: i32) -> i32;

    pub fn ob_repacketize_batch(device: i32, n_streams: i32, n_in: i32, packets: *const u8, offsets: *const i32, lens: *const i32,
                                group: i32, pad_to: i32, out: *mut u8, max_bytes: i32, lens_out: *mut i32) -> i32;
}

#[repr(C)]
pub struct ObEncoder {
    _unused: [u8; 0],
}

/// `n_streams` independent CELT-only Opus decoders living on one B200.
/// Mirrors `Decoder` (src/decoder.rs:18-28): owns the raw handle, frees it in `Drop`, all methods take `&mut self`.
pub struct BatchDecoder {
    raw: NonNull<ObDecoder>,
    n_streams: usize,
    channels: Channels,
}

unsafe impl Send for BatchDecoder {}

impl BatchDecoder {
    /// Cf. `Decoder::new` (src/decoder.rs:35-63). Only `SampleRate::Hz48000` is supported by the CUDA path.
    pub fn new(n_streams: usize, sample_rate: SampleRate, channels: Channels, device: i32, max_frames: usize) -> Result<Self> {
        let mut err = 0i32;
        let raw = unsafe {
            ob_decoder_create(n_streams as i32, sample_rate as i32, channels as i32, device, max_frames as i32, &mut err)
        };
        match NonNull::new(raw) {
            Some(raw) if err == 0 => Ok(Self { raw, n_streams, channels }),
            _ => Err(Error::from_code(err)),
        }
    }

    /// One packet per stream. `packets[s]` may be empty: a lost packet, concealed for `frame_size` samples exactly like
    /// `Decoder::decode_float(&[], out, false)`.
    /// `output` is `n_streams * frame_size * channels` interleaved floats; `frame_size = output.len() / n_streams / channels`
    /// exactly like `Decoder::decode_float` derives it (src/decoder.rs:149).
    /// `fec` is the flag of `Decoder::decode_float` (src/decoder.rs:134): CELT-only packets carry no FEC, libopus (and this path)
    /// then conceals the frame like a lost packet.
    /// Returns per-stream `Ok(samples_per_channel)` / `Err(code)`.
    pub fn decode_float(&mut self, packets: &[&[u8]], output: &mut [f32], fec: bool) -> Result<Vec<Result<usize>>> {
        if packets.len() != self.n_streams || output.len() % (self.n_streams * self.channels as usize) != 0 {
            return Err(Error::BadArg);
        }
        let rc = unsafe { ob_decoder_set_decode_fec(self.raw.as_ptr(), fec as i32) };
        if rc != 0 {
            return Err(Error::from_code(rc));
        }
        let frame_size = output.len() / self.n_streams / self.channels as usize;
        let mut flat = Vec::with_capacity(packets.iter().map(|p| p.len()).sum());
        let (mut offsets, mut lens) = (Vec::with_capacity(self.n_streams), Vec::with_capacity(self.n_streams));
        for p in packets {
            offsets.push(flat.len() as i32);
            lens.push(p.len() as i32);
            flat.extend_from_slice(p);
        }
        if flat.is_empty() {
            flat.push(0);
        }
        let mut samples = vec![0i32; self.n_streams];
        let rc = unsafe {
            ob_decode_float(self.raw.as_ptr(), flat.as_ptr(), offsets.as_ptr(), lens.as_ptr(), output.as_mut_ptr(),
                            frame_size as i32, samples.as_mut_ptr())
        };
        if rc != 0 {
            return Err(Error::from_code(rc));
        }
        Ok(samples.into_iter().map(|n| if n >= 0 { Ok(n as usize) } else { Err(Error::from_code(n)) }).collect())
    }

    /// Cf. `Decoder::final_range` (src/decoder.rs:302-312), for every stream.
    pub fn final_range(&mut self) -> Result<Vec<u32>> {
        let mut out = vec![0u32; self.n_streams];
        let rc = unsafe { ob_decoder_final_range(self.raw.as_ptr(), out.as_mut_ptr()) };
        if rc != 0 { Err(Error::from_code(rc)) } else { Ok(out) }
    }

    /// Cf. `Decoder::set_gain` (src/decoder.rs:318-320): Q8 dB for the whole batch.
    pub fn set_gain(&mut self, q8_db: i32) -> Result<()> {
        let rc = unsafe { ob_decoder_set_gain(self.raw.as_ptr(), q8_db) };
        if rc != 0 { Err(Error::from_code(rc)) } else { Ok(()) }
    }

    /// Cf. `Decoder::set_phase_inversion_disabled` (src/decoder.rs:341-346).
    pub fn set_phase_inversion_disabled(&mut self, disabled: bool) -> Result<()> {
        let rc = unsafe { ob_decoder_set_phase_inversion_disabled(self.raw.as_ptr(), disabled as i32) };
        if rc != 0 { Err(Error::from_code(rc)) } else { Ok(()) }
    }

    /// Cf. `Decoder::reset` (src/decoder.rs:241-255); `None` resets every stream.
    pub fn reset(&mut self, streams: Option<&[i32]>) -> Result<()> {
        let rc = unsafe {
            match streams {
                Some(s) => ob_decoder_reset(self.raw.as_ptr(), s.as_ptr(), s.len() as i32),
                None => ob_decoder_reset(self.raw.as_ptr(), std::ptr::null(), 0),
            }
        };
        if rc != 0 { Err(Error::from_code(rc)) } else { Ok(()) }
    }
}

impl Drop for BatchDecoder {
    fn drop(&mut self) {
        unsafe { ob_decoder_destroy(self.raw.as_ptr()) }
    }
}

/// `n_streams` independent CELT-only (`Application::RestrictedLowDelay`) Opus encoders living on one B200.
/// Mirrors `Encoder` (src/encoder.rs:26-73): owns the raw handle, frees it in `Drop`, all methods take `&mut self`.
pub struct BatchEncoder {
    raw: NonNull<ObEncoder>,
    n_streams: usize,
    channels: Channels,
}

unsafe impl Send for BatchEncoder {}

impl BatchEncoder {
    /// Cf. `Encoder::new` (src/encoder.rs:40-73). The CUDA path supports `Hz48000` + `Application::RestrictedLowDelay`
    /// (OPUS_APPLICATION_RESTRICTED_LOWDELAY, the CELT-only application); anything else is `Error::Unimplemented`.
    pub fn new(n_streams: usize, sample_rate: SampleRate, channels: Channels, application: Application, device: i32,
               max_frames: usize) -> Result<Self> {
        let mut err = 0i32;
        let raw = unsafe {
            ob_encoder_create(n_streams as i32, sample_rate as i32, channels as i32, application as i32, device,
                              max_frames as i32, &mut err)
        };
        match NonNull::new(raw) {
            Some(raw) if err == 0 => Ok(Self { raw, n_streams, channels }),
            _ => Err(Error::from_code(err)),
        }
    }

    /// One frame per stream. `input` is `n_streams * frame_size * channels` interleaved floats (frame size derived from
    /// the slice length like `Encoder::encode_float`, src/encoder.rs:215-247); `output` is `n_streams * max_bytes`.
    /// Returns per-stream `Ok(packet_len)` / `Err(code)`; packet `s` is `output[s*max_bytes .. s*max_bytes + len]`.
    pub fn encode_float(&mut self, input: &[f32], output: &mut [u8]) -> Result<Vec<Result<usize>>> {
        let per = self.n_streams * self.channels as usize;
        if input.is_empty() || input.len() % per != 0 || output.len() % self.n_streams != 0 {
            return Err(Error::BadArg);
        }
        let (frame_size, max_bytes) = (input.len() / per, output.len() / self.n_streams);
        let mut lens = vec![0i32; self.n_streams];
        let rc = unsafe {
            ob_encode_float(self.raw.as_ptr(), input.as_ptr(), frame_size as i32, output.as_mut_ptr(), max_bytes as i32,
                            lens.as_mut_ptr())
        };
        if rc != 0 {
            return Err(Error::from_code(rc));
        }
        Ok(lens.into_iter().map(|n| if n >= 0 { Ok(n as usize) } else { Err(Error::from_code(n)) }).collect())
    }

    /// Cf. `Encoder::set_bitrate` (src/encoder.rs:545-562): one value for the whole batch.
    pub fn set_bitrate(&mut self, bitrate: Bitrate) -> Result<()> {
        let v = match bitrate { Bitrate::Auto => -1000, Bitrate::Max => -1, Bitrate::Custom(b) => b };
        let rc = unsafe { ob_encoder_set_bitrate(self.raw.as_ptr(), v) };
        if rc != 0 { Err(Error::from_code(rc)) } else { Ok(()) }
    }

    /// Cf. `Encoder::set_complexity` (src/encoder.rs:588-610).
    pub fn set_complexity(&mut self, complexity: Complexity) -> Result<()> {
        let rc = unsafe { ob_encoder_set_complexity(self.raw.as_ptr(), complexity.value() as i32) };
        if rc != 0 { Err(Error::from_code(rc)) } else { Ok(()) }
    }

    /// Cf. `Encoder::set_vbr` / `set_vbr_constraint` (src/encoder.rs:639-656, :311-319).
    pub fn set_vbr(&mut self, enabled: bool, constrained: bool) -> Result<()> {
        let rc = unsafe { ob_encoder_set_vbr(self.raw.as_ptr(), enabled as i32) };
        if rc != 0 { return Err(Error::from_code(rc)); }
        let rc = unsafe { ob_encoder_set_vbr_constraint(self.raw.as_ptr(), constrained as i32) };
        if rc != 0 { Err(Error::from_code(rc)) } else { Ok(()) }
    }

    /// Cf. `Encoder::final_range` (src/encoder.rs:411-419), for every stream.
    pub fn final_range(&mut self) -> Result<Vec<u32>> {
        let mut out = vec![0u32; self.n_streams];
        let rc = unsafe { ob_encoder_final_range(self.raw.as_ptr(), out.as_mut_ptr()) };
        if rc != 0 { Err(Error::from_code(rc)) } else { Ok(out) }
    }

    /// Cf. `Encoder::reset` (src/encoder.rs:689-698); `None` resets every stream.
    pub fn reset(&mut self, streams: Option<&[i32]>) -> Result<()> {
        let rc = unsafe {
            match streams {
                Some(s) => ob_encoder_reset(self.raw.as_ptr(), s.as_ptr(), s.len() as i32),
                None => ob_encoder_reset(self.raw.as_ptr(), std::ptr::null(), 0),
            }
        };
        if rc != 0 { Err(Error::from_code(rc)) } else { Ok(()) }
    }
}

impl Drop for BatchEncoder {
    fn drop(&mut self) {
        unsafe { ob_encoder_destroy(self.raw.as_ptr()) }
    }
}

/// Batched form of `Repacketizer` (src/repacketizer.rs:11-100): every `group` consecutive packets of each stream are merged into
/// one packet (`reset(); push() x group; out()`), on the GPU. `packets` / `offsets` / `lens` are laid out as for
/// `BatchDecoder::decode_float_multi` (`[n_streams][n_in]`). Returns one `Result<Vec<u8>>` per (stream, group).
pub fn repacketize_batch(device: i32, n_streams: usize, n_in: usize, packets: &[u8], offsets: &[i32], lens: &[i32], group: usize,
                         pad_to: usize, max_bytes: usize) -> Result<Vec<Result<Vec<u8>>>> {
    if offsets.len() != n_streams * n_in || lens.len() != offsets.len() || group == 0 || max_bytes == 0 {
        return Err(Error::BadArg);
    }
    // the C side reads lens[i] bytes at packets + offsets[i]: a safe wrapper must not let either run past the slice
    for (&o, &l) in offsets.iter().zip(lens.iter()) {
        if l < 0 || (l > 0 && (o < 0 || (o as usize) + (l as usize) > packets.len())) {
            return Err(Error::BadArg);
        }
    }
    let n_out = n_streams * n_in.div_ceil(group);
    let mut out = vec![0u8; n_out * max_bytes];
    let mut lens_out = vec![0i32; n_out];
    let rc = unsafe {
        ob_repacketize_batch(device, n_streams as i32, n_in as i32, packets.as_ptr(), offsets.as_ptr(), lens.as_ptr(), group as i32,
                             pad_to as i32, out.as_mut_ptr(), max_bytes as i32, lens_out.as_mut_ptr())
    };
    if rc != 0 { return Err(Error::from_code(rc)); }
    Ok(lens_out.iter().enumerate().map(|(k, &n)| {
        if n < 0 { Err(Error::from_code(n)) } else { Ok(out[k * max_bytes..k * max_bytes + n as usize].to_vec()) }
    }).collect())
}
