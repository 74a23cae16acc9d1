#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched CELT decode path (BASELINE.json configs[1]):
4096 mono 48 kHz CELT-only 20 ms @ 64 kb/s streams per GPU, decoded with final-range verification.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl ours|reference]

A "step" is one pass of the hot path over one batch: every one of the 4096 streams advances by F consecutive
20 ms frames (default F=200 = 4 s of audio per stream per step, ~60 ms of GPU time: the timed region of a default run is ~2 s).  Packets were pre-encoded by the reference
(tests/golden/make_golden.py -> tests/golden/pool_cfg2_mono_20ms_64k_cbr.npz: 256 distinct streams, tiled).

  value  = audio-seconds decoded per second (= number of real-time streams one GPU sustains), whole job over all
           N GPUs, inputs resident in HBM when the timed region starts (device-pointer C-ABI entry point);
  e2e    = the same metric through the host-pointer C-ABI call (ob_decode_float_multi) with pinned host buffers:
           H2D of packets/offsets/lengths and D2H of PCM/samples/final ranges are inside the timed region;
  roofline = HBM roofline of the dominant kernel, from SURVEY.md 8(d)'s algorithmic bytes per frame;
  cpu_baseline = the UNMODIFIED reference (oracle/_ref, libopus 1.5.2) on this box's host cores, one stream per
           thread, on a bounded sample of the same workload.

--impl reference prints the same line for the reference's own CPU implementation (all host threads).
Multi-GPU (torchrun, one rank per GPU): streams are sharded, there is no collective on the data path; "weak" scaling
(4096 streams per GPU).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STREAMS_PER_GPU = 4096
FRAME = 960
POOL = os.path.join(ROOT, "tests", "golden", "pool_cfg2_mono_20ms_64k_cbr.npz")
WORKLOAD = "4096 mono 48 kHz CELT-only 20 ms @64 kb/s CBR decode streams per GPU (BASELINE configs[1]), final range verified"


def workload_config(F):
    """The `config` object of the JSON line: the same for both arms (`--impl ours` / `--impl reference`), so that the driver's same-config check compares
    like with like; what is specific to an arm goes into its `arm_note`."""
    S = STREAMS_PER_GPU
    return {"workload": WORKLOAD, "streams_per_gpu": S, "frames_per_stream_per_step": F, "frame_ms": 20, "bitrate": 64000, "packet_bytes": 160,
            "l2": "GPU arm: per-step working set (IR+spectrum+PCM, %.1f GB) exceeds the 126 MB L2, no flush needed" % ((S * F * (12.6e3 + 7.7e3 + 3.84e3)) / 1e9),
            "sharding": "streams split by rank, no collective"}


def algorithmic_bytes_per_frame(F, C=1, N=FRAME, P=160):
    """B_dec = P + 4*C*N + S_dec/F, S_dec = per-channel (4576 read + (N+120)*4 written) + energies
    1344 + scalars 128 (read and written once per launch)."""
    s_dec = C * (4576 + (N + 120) * 4) + 1344 + 128
    return P + 4 * C * N + s_dec / F


def load_pool(nstreams, F):
    z = np.load(POOL)
    pk, ln, rng = z["packets"], z["lens"], z["dec_rng"]
    P = pk.shape[0]
    reps = (F + pk.shape[1] - 1) // pk.shape[1]
    if reps > 1:            # longer runs replay the pool's frames (final range is stateless, PCM is not compared here)
        pk, ln, rng = np.tile(pk, (1, reps, 1)), np.tile(ln, (1, reps)), np.tile(rng, (1, reps))
    idx = np.arange(nstreams) % P
    return np.ascontiguousarray(pk[idx, :F]), np.ascontiguousarray(ln[idx, :F]), np.ascontiguousarray(rng[idx, :F])


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = []
        for k, name in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap")):
            if any(len(r) >= 7 and r[k].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_reference_run(nthreads, target_audio_s, F=50):
    """Times the unmodified reference (oracle/_ref) decoding a bounded sample of the workload with a pthread pool,
    one stream per thread at a time.  Returns (audio-s/s, description)."""
    from oracle import refpy
    ns = max(nthreads, int(round(target_audio_s / (F * 0.02))))
    pk, ln, rng = load_pool(ns, F)
    secs, _, got = refpy.decode_pool(pk, ln, FRAME, 1, nthreads, want_ranges=True)
    assert (got == rng).all(), "reference decoder final range != stored encoder range?!"
    audio = ns * F * 0.02
    return audio / secs, "%d streams x %d frames (%.0f audio-s) of the same packet pool, %d threads, %.2f s wall" % (ns, F, audio, nthreads, secs)


ENC_STREAMS_PER_GPU = 16384
ENC_WORKLOAD = "16384 stereo 48 kHz CELT-only encode streams per GPU, complexity 10, 20 ms @96 kb/s CBR (BASELINE configs[2])"


def enc_algorithmic_bytes_per_frame(F, C=2, N=FRAME, P=240):
    """B_enc = 4*C*N + P + S_enc/F; S_enc = per channel (4576 read + min(N,1024)*4+480 written) + energies 1344 + scalars 256."""
    s_enc = C * (4576 + min(N, 1024) * 4 + 480) + 1344 + 256
    return 4 * C * N + P + s_enc / F


_ENC_POOL = {}


def enc_pcm(nstreams, F):
    from opus_codec_b200 import synth
    if F not in _ENC_POOL:
        _ENC_POOL[F] = np.stack([synth.stream_pcm(s, FRAME * F, 2, base_seed=31337) for s in range(64)])      # 64 distinct streams, tiled
    return np.ascontiguousarray(_ENC_POOL[F][np.arange(nstreams) % 64]).reshape(nstreams, F, FRAME * 2)


def cpu_reference_encode(nthreads, nstreams, F):
    from oracle import refpy
    pcm = enc_pcm(nstreams, F).reshape(nstreams, -1)
    secs, _, lens, _ = refpy.encode_pool(pcm, FRAME, 2, 96000, 0, 10, nthreads, stride=256, want_packets=False)
    assert (lens == 240).all()
    audio = nstreams * F * 0.02
    return audio / secs, "%d streams x %d frames (%.0f audio-s), %d threads, %.2f s wall" % (nstreams, F, audio, nthreads, secs)


def run_encode_leg(args, L, local, world, rank, dev, barrier):
    """Secondary measurement (BASELINE metric names both directions): batched encode, BASELINE configs[2]."""
    import torch
    from opus_codec_b200.batch import BatchEncoder
    from opus_codec_b200.shard import max_over_ranks
    S, F = ENC_STREAMS_PER_GPU, args.enc_frames
    pcm = enc_pcm(S, F)
    enc = BatchEncoder(S, 48000, 2, device=local, max_frames=F)
    enc.set_bitrate(96000); enc.set_complexity(10); enc.set_vbr(False)
    ext = torch.cuda.ExternalStream(L.ob_encoder_cuda_stream(enc.handle), device=local)
    d_pcm = torch.from_numpy(pcm.reshape(-1)).to(dev)
    d_out = torch.zeros(S * F * 256, dtype=torch.uint8, device=dev)
    d_len = torch.zeros(S * F, dtype=torch.int32, device=dev)
    d_rng = torch.zeros(S * F, dtype=torch.int32, device=dev)

    def step():
        r = L.ob_encode_float_device(enc.handle, F, d_pcm.data_ptr(), FRAME, d_out.data_ptr(), 256, d_len.data_ptr(), d_rng.data_ptr(), 0)
        assert r == 0, r

    step()
    barrier()
    steps = max(3, min(args.steps, 6))                           # ~0.35 s of GPU time per step: >= 1 s timed
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    kms = 0.0
    for _ in range(steps):
        step()
        kms += enc.kernel_ms()
    e1.record(ext)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1), dev)
    assert (d_len.cpu().numpy() == 240).all()
    # end to end: pinned host PCM in, packets + lengths + final ranges out
    h_pcm = torch.from_numpy(pcm.reshape(-1).copy()).pin_memory()
    h_out = torch.empty(S * F * 256, dtype=torch.uint8).pin_memory()
    h_len = torch.empty(S * F, dtype=torch.int32).pin_memory()
    h_rng = torch.empty(S * F, dtype=torch.int32).pin_memory()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(ext)
    r = L.ob_encode_float_multi(enc.handle, F, h_pcm.data_ptr(), FRAME, h_out.data_ptr(), 256, h_len.data_ptr(), h_rng.data_ptr())
    assert r == 0, r
    f1.record(ext)
    barrier()
    ms_e2e = max_over_ranks(f0.elapsed_time(f1), dev)
    launches = enc.launches()
    enc.close()
    verified = None
    if rank == 0:                                               # checker only: the reference decoder takes a sample of the packets just produced
        try:
            from oracle import refpy
            o = h_out.numpy().reshape(S, F, 256); ln = h_len.numpy().reshape(S, F); rg = h_rng.numpy().view(np.uint32).reshape(S, F)
            idx = list(range(0, S, max(1, S // 16)))[:16]
            for s_ in idx:
                _, dr, smp = refpy.decode_stream(o[s_], ln[s_], FRAME, 2)
                assert (smp == FRAME).all() and (dr == rg[s_]).all(), "reference decoder final range != GPU encoder final range (stream %d)" % s_
            verified = "reference decoder accepts the packets of %d sampled streams, final ranges equal" % len(idx)
        except ImportError:
            verified = "not checked (oracle/_ref missing)"
    audio = world * S * F * 0.02
    res = {"workload": ENC_WORKLOAD, "frames_per_stream_per_step": F, "steps": steps,
           "value": audio * steps / (ms / 1000.0), "unit": "audio-s/s", "ms_per_step": ms / steps,
           "e2e": {"value": audio / (ms_e2e / 1000.0), "unit": "audio-s/s", "h2d_bytes_per_step": int(4 * h_pcm.numel()),
                   "d2h_bytes_per_step": int(h_out.numel() + 8 * h_len.numel())},
           "gpu_launches": int(launches), "verified": verified}
    if rank == 0:
        peak = 6549.1
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        k = kms / steps
        ach = enc_algorithmic_bytes_per_frame(F) * S * F / (k / 1000.0) / 1e9
        traffic, tsrc = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_dram_traffic.json")))["encode"]
            traffic, tsrc = tj["dram_bytes_per_stream_frame"] * S * F, tj["source"]
        except Exception:
            pass
        res["roofline"] = {"bound": "hbm", "kernel": "ob_k_encode_thread (OB_ENC_MAP_AUTO at %d streams: one lane per stream)" % S, "achieved": ach, "peak": peak,
                           "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                           "traffic_unit": "bytes per step of F frames (dram__bytes_read.sum + dram__bytes_write.sum per stream-frame x S x F)", "traffic_source": tsrc,
                           "kernel_ms": k, "algorithmic_bytes_per_frame": enc_algorithmic_bytes_per_frame(F)}
        if not args.kernels_only:
            cores = os.cpu_count() or 1
            try:
                v, sample = cpu_reference_encode(cores, cores * 16, 600)      # >= 16 streams per core x 12 s of audio: >= 1 s of wall time (0.86 s with 8 s of audio on a 16-core box)
                res["cpu_baseline"] = {"value": v, "unit": "audio-s/s", "cores": cores, "kind": "reference", "sample": sample}
            except Exception as ex:
                res["cpu_baseline"] = {"value": None, "sample": "unavailable: %r" % (ex,)}
    return res


def run_transcode_leg(args, L, local, world, rank, dev, barrier):
    """BASELINE configs[4] per GPU: 16384 stereo streams decoded and re-encoded (96 kb/s CBR, complexity 10) without leaving the device --
    the decoder's PCM buffer is the encoder's input (keep decoded PCM on-device) -- and the decode half alone, which is BASELINE's
    target case (stereo CELT decode), kernel-resident and end to end, every final range verified."""
    import torch
    from opus_codec_b200.batch import BatchDecoder, BatchEncoder
    from opus_codec_b200.shard import max_over_ranks
    S, F = ENC_STREAMS_PER_GPU, args.enc_frames
    z = np.load(os.path.join(ROOT, "tests", "golden", "cfg3_stereo_20ms_96k_cbr.npz"))
    pk0, ln0 = z["packets"], z["lens"]                         # [6, 50, stride]: tiled over streams, the first F frames of each
    idx = np.arange(S) % pk0.shape[0]
    pk = np.ascontiguousarray(pk0[idx, :F]); ln = np.ascontiguousarray(ln0[idx, :F]).astype(np.int32)
    rng_expect = np.ascontiguousarray(z["dec_rng"][idx, :F])
    stride = pk.shape[2]
    offsets = (np.arange(S * F, dtype=np.int32) * stride).reshape(S, F)
    dec = BatchDecoder(S, 48000, 2, device=local, max_frames=F)
    enc = BatchEncoder(S, 48000, 2, device=local, max_frames=F)
    enc.set_bitrate(96000); enc.set_complexity(10); enc.set_vbr(False)
    dext = torch.cuda.ExternalStream(L.ob_decoder_cuda_stream(dec.handle), device=local)
    eext = torch.cuda.ExternalStream(L.ob_encoder_cuda_stream(enc.handle), device=local)
    d_pk = torch.from_numpy(pk.reshape(-1)).to(dev); d_off = torch.from_numpy(offsets.reshape(-1)).to(dev); d_len = torch.from_numpy(ln.reshape(-1)).to(dev)
    d_pcm = torch.empty(S * F * FRAME * 2, dtype=torch.float32, device=dev)
    d_smp = torch.empty(S * F, dtype=torch.int32, device=dev); d_rng = torch.empty(S * F, dtype=torch.int32, device=dev)
    d_out = torch.zeros(S * F * 256, dtype=torch.uint8, device=dev)
    d_olen = torch.zeros(S * F, dtype=torch.int32, device=dev); d_orng = torch.zeros(S * F, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ready, consumed = torch.cuda.Event(), torch.cuda.Event()

    def step():
        dext.wait_event(consumed)                               # the encoder has read the previous step's PCM
        r = L.ob_decode_float_device(dec.handle, F, d_pk.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), d_pcm.data_ptr(), FRAME, d_smp.data_ptr(), d_rng.data_ptr(), 0)
        assert r == 0, r
        ready.record(dext)
        eext.wait_event(ready)
        r = L.ob_encode_float_device(enc.handle, F, d_pcm.data_ptr(), FRAME, d_out.data_ptr(), 256, d_olen.data_ptr(), d_orng.data_ptr(), 0)
        assert r == 0, r
        consumed.record(eext)

    consumed.record(eext)
    step()
    barrier()
    steps = max(3, min(args.steps, 6))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(dext)
    for _ in range(steps):
        step()
    e1.record(eext)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1), dev)
    assert (d_smp.cpu().numpy() == FRAME).all() and (d_olen.cpu().numpy() == 240).all()
    assert (d_rng.cpu().numpy().view(np.uint32).reshape(S, F) == rng_expect).all(), "transcode leg: decoder final range != reference"
    verified = "decoder final ranges equal the reference's on all %d frames" % (S * F)
    if rank == 0:                                               # checker only: the reference decoder takes a sample of the re-encoded packets
        try:
            from oracle import refpy
            o = d_out.cpu().numpy().reshape(S, F, 256); ol = d_olen.cpu().numpy().reshape(S, F); og = d_orng.cpu().numpy().view(np.uint32).reshape(S, F)
            # the streams restart every step from the encoder's carried state: decode the LAST step's packets with a decoder primed by nothing --
            # final ranges are a function of the packet alone, so they must match whatever the decoder's history is
            sidx = list(range(0, S, max(1, S // 16)))[:16]
            for s_ in sidx:
                _, dr, smp = refpy.decode_stream(o[s_], ol[s_], FRAME, 2)
                assert (smp == FRAME).all() and (dr == og[s_]).all(), "transcode leg: reference decoder final range != GPU encoder final range (stream %d)" % s_
            verified += "; the reference decoder accepts the re-encoded packets of %d sampled streams with equal final ranges" % len(sidx)
        except ImportError:
            verified += "; re-encoded packets not checked (oracle/_ref missing)"
    launches = dec.launches() + enc.launches()
    enc.close()
    # ---- the decode half alone: BASELINE's target is stated for STEREO streams (>= 4096 real-time 48 kHz stereo CELT decode streams per B200) ----
    torch.cuda.synchronize()
    barrier()
    dsteps = max(4, min(args.steps, 24))
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(dext)
    kms = np.zeros(3)
    for _ in range(dsteps):
        r = L.ob_decode_float_device(dec.handle, F, d_pk.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), d_pcm.data_ptr(), FRAME, d_smp.data_ptr(), d_rng.data_ptr(), 0)
        assert r == 0, r
        kms += np.array(dec.kernel_ms())
    g1.record(dext)
    barrier()
    ms_dec = max_over_ranks(g0.elapsed_time(g1), dev)
    kms /= dsteps
    assert (d_rng.cpu().numpy().view(np.uint32).reshape(S, F) == rng_expect).all(), "stereo decode: final range mismatch"
    # end to end: host-pointer ABI, two calls in flight, pinned buffers (1.26 GB of float PCM per step come down)
    del d_pcm, d_out
    h_pk = torch.from_numpy(pk.reshape(-1).copy()).pin_memory(); h_off = torch.from_numpy(offsets.reshape(-1).copy()).pin_memory()
    h_len = torch.from_numpy(ln.reshape(-1).copy()).pin_memory()
    h_pcm = [torch.empty(S * F * FRAME * 2, dtype=torch.float32).pin_memory() for _ in range(2)]
    h_smp = [torch.empty(S * F, dtype=torch.int32).pin_memory() for _ in range(2)]
    h_rng = [torch.empty(S * F, dtype=torch.int32).pin_memory() for _ in range(2)]

    def submit(n):
        p = n & 1
        r = L.ob_decode_float_multi_async(dec.handle, F, h_pk.data_ptr(), h_off.data_ptr(), h_len.data_ptr(), h_pcm[p].data_ptr(), FRAME,
                                          h_smp[p].data_ptr(), h_rng[p].data_ptr())
        assert r == 0, r
    for n in range(3):
        submit(n)
        assert L.ob_decoder_wait(dec.handle, 0) == 0
    barrier()
    esteps = dsteps                                       # as many as the device-timed leg: the pipeline's fill / drain is one step's worth
    t0 = time.perf_counter()
    for n in range(esteps):
        submit(n)
        if n:
            assert L.ob_decoder_wait(dec.handle, 1) == 0
    assert L.ob_decoder_wait(dec.handle, 0) == 0
    torch.cuda.synchronize()
    ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3, dev)
    for p in range(2):
        assert (h_rng[p].numpy().view(np.uint32).reshape(S, F) == rng_expect).all() and (h_smp[p].numpy() == FRAME).all()
    dec.close()
    audio = world * S * F * 0.02
    bytes_pf = algorithmic_bytes_per_frame(F, C=2, P=240)
    dom = int(np.argmax(kms))
    names = ["ob_k_symbols", "ob_k_bands", "ob_k_synth"]
    stereo = {"workload": "%d stereo 48 kHz CELT-only 20 ms @96 kb/s decode streams per GPU (BASELINE's target case), %d frames per stream per step" % (S, F),
              "value": audio * dsteps / (ms_dec / 1000.0), "unit": "audio-s/s", "ms_per_step": ms_dec / dsteps, "steps": dsteps,
              "streams_equivalent_per_gpu": audio * dsteps / (ms_dec / 1000.0) / world, "target_rt_streams": 4096,
              "final_range_verified": "all %d frames, kernel-resident and end-to-end legs" % (S * F),
              "e2e": {"value": audio * esteps / (ms_e2e / 1000.0), "unit": "audio-s/s", "steps": esteps,
                      "h2d_bytes_per_step": int(h_pk.numel() + 8 * h_off.numel()), "d2h_bytes_per_step": int(4 * h_pcm[0].numel() + 8 * h_smp[0].numel())},
              "kernel_ms": dict(zip(names, [float(v) for v in kms])),
              "roofline": {"bound": "hbm", "kernel": names[dom], "algorithmic_bytes_per_frame": bytes_pf,
                           "achieved": bytes_pf * S * F / (float(kms[dom]) / 1e3) / 1e9, "unit": "GB/s",
                           "pipeline_achieved_GBps": bytes_pf * S * F / (float(kms.sum()) / 1e3) / 1e9}}
    return {"decode_stereo": stereo, "workload": "%d stereo streams per GPU decoded (96 kb/s CBR packets) and re-encoded at 96 kb/s CBR, complexity 10, PCM stays on the device "
                        "(BASELINE configs[4] per-GPU share)" % S,
            "frames_per_stream_per_step": F, "steps": steps, "value": audio * steps / (ms / 1000.0), "unit": "audio-s/s", "ms_per_step": ms / steps,
            "gpu_launches": int(launches), "verified": verified}


def host_d2h_ceiling(dev, barrier, seconds=1.0, nbytes=512 << 20):
    """What this host can absorb: every rank copies device -> pinned host memory at the same time (one stream, 512 MB copies back to back)
    for `seconds`; returns this rank's GB/s.  With N ranks on one box the sum over ranks is the box's host-bound ceiling -- the number the
    float end-to-end leg (788 MB of PCM per step and GPU at F=50) runs into at N = 8 (tools/host_d2h_ceiling.py is the stand-alone form)."""
    import torch
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    h.copy_(d, non_blocking=True); torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, t0 = 0, time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < seconds:
        h.copy_(d, non_blocking=True); n += 1
        if n % 4 == 0:
            torch.cuda.current_stream().synchronize()
    e1.record(); torch.cuda.synchronize()
    barrier()
    return n * nbytes / (e0.elapsed_time(e1) / 1e3) / 1e9


def run_live_leg(args, L, local, world, rank, dev, barrier):
    """The LIVE case (SURVEY 8d, F = 1): one packet per stream per call, as Decoder::decode_float is called (src/decoder.rs:134-182), through the
    host-pointer ABI with pinned buffers -- upload, kernels and download inside the timed call.  A frame is 20 ms of audio, so the batch is
    real time while a call takes < 20 ms: rt_stream_capacity is the largest swept batch that does (per GPU).  Same for the encoder."""
    import torch
    from opus_codec_b200.batch import BatchDecoder, BatchEncoder
    from opus_codec_b200.shard import max_over_ranks
    out = {"definition": "one 20 ms frame per stream per call (F = 1), host-pointer C ABI, pinned host buffers; real time while a call takes < 20 ms",
           "frame_ms": 20.0}
    z3 = np.load(os.path.join(ROOT, "tests", "golden", "cfg3_stereo_20ms_96k_cbr.npz"))
    zp = np.load(POOL)
    NSTEP, WARM = 6, 3
    for key, CC, pk0, ln0, rg0, sizes in (("decode_mono_64k", 1, zp["packets"], zp["lens"], zp["dec_rng"], (4096, 16384, 65536, 131072, 163840, 180224, 196608)),
                                          ("decode_stereo_96k", 2, z3["packets"], z3["lens"], z3["dec_rng"], (4096, 16384, 65536, 81920, 90112, 98304, 114688))):
        rows, cap = [], 0
        stride = pk0.shape[2]
        for S in sizes:
            idx = np.arange(S) % pk0.shape[0]
            nfr = WARM + NSTEP
            h_pk = [torch.from_numpy(np.ascontiguousarray(pk0[idx, f]).reshape(-1)).pin_memory() for f in range(nfr)]
            h_ln = [torch.from_numpy(np.ascontiguousarray(ln0[idx, f]).astype(np.int32)).pin_memory() for f in range(nfr)]
            h_off = torch.from_numpy(np.arange(S, dtype=np.int32) * stride).pin_memory()
            h_pcm = [torch.empty(S * FRAME * CC, dtype=torch.float32).pin_memory() for _ in range(2)]
            h_smp = [torch.empty(S, dtype=torch.int32).pin_memory() for _ in range(2)]
            h_rng = [torch.empty(S, dtype=torch.int32).pin_memory() for _ in range(2)]
            dec = BatchDecoder(S, 48000, CC, device=local, max_frames=1)

            def call(n, wait):
                p = n & 1
                r = L.ob_decode_float_multi_async(dec.handle, 1, h_pk[n].data_ptr(), h_off.data_ptr(), h_ln[n].data_ptr(), h_pcm[p].data_ptr(), FRAME,
                                                  h_smp[p].data_ptr(), h_rng[p].data_ptr())
                assert r == 0, r
                if wait:
                    assert L.ob_decoder_wait(dec.handle, 0) == 0
            for n in range(WARM):
                call(n, True)
            barrier()
            lat = []
            for n in range(WARM, nfr):                                  # blocking: the latency of one call, nothing overlapped
                t0 = time.perf_counter()
                call(n, True)
                lat.append((time.perf_counter() - t0) * 1e3)
                assert (h_rng[n & 1].numpy().view(np.uint32) == rg0[idx, n]).all(), "final range mismatch (live leg)"
                assert (h_smp[n & 1].numpy() == FRAME).all()
            dec.close()
            ms = max_over_ranks(float(np.median(lat)), dev)
            rows.append({"streams_per_gpu": S, "ms_per_call": ms, "realtime": bool(ms < 20.0)})
            if ms < 20.0:
                cap = S
            del h_pk, h_ln, h_pcm, h_smp, h_rng
            if ms > 24.0:
                break
        P = 160 if CC == 1 else 240
        bytes_f1 = algorithmic_bytes_per_frame(1, C=CC, P=P)
        best = [r for r in rows if r["realtime"]]
        out[key] = {"sweep": rows, "rt_stream_capacity_per_gpu": cap, "final_range_verified": True, "algorithmic_bytes_per_frame_F1": bytes_f1,
                    "achieved_GBps_at_capacity": (bytes_f1 * best[-1]["streams_per_gpu"] / (best[-1]["ms_per_call"] / 1e3) / 1e9) if best else None}
    # live encode: one warp per stream is the low-latency mapping (include/opus_b200.h OB_ENC_MAP_WARP)
    rows, cap = [], 0
    for S in (1184, 2368, 4736, 5920, 7104):
        pcm = enc_pcm(S, WARM + NSTEP)
        h_in = [torch.from_numpy(np.ascontiguousarray(pcm[:, f]).reshape(-1)).pin_memory() for f in range(WARM + NSTEP)]
        h_out = torch.empty(S * 256, dtype=torch.uint8).pin_memory()
        h_len = torch.empty(S, dtype=torch.int32).pin_memory(); h_rg = torch.empty(S, dtype=torch.int32).pin_memory()
        enc = BatchEncoder(S, 48000, 2, device=local, max_frames=1)
        enc.set_mapping(1); enc.set_bitrate(96000); enc.set_complexity(10); enc.set_vbr(False)
        lat = []
        for n in range(WARM + NSTEP):
            t0 = time.perf_counter()
            r = L.ob_encode_float_multi(enc.handle, 1, h_in[n].data_ptr(), FRAME, h_out.data_ptr(), 256, h_len.data_ptr(), h_rg.data_ptr())
            assert r == 0, r
            if n >= WARM:
                lat.append((time.perf_counter() - t0) * 1e3)
            assert (h_len.numpy() == 240).all()
        enc.close()
        ms = max_over_ranks(float(np.median(lat)), dev)
        rows.append({"streams_per_gpu": S, "ms_per_call": ms, "realtime": bool(ms < 20.0)})
        if ms < 20.0:
            cap = S
        if ms > 24.0:
            break
    out["encode_stereo_96k_c10"] = {"sweep": rows, "rt_stream_capacity_per_gpu": cap, "mapping": "one warp per stream (OB_ENC_MAP_WARP)"}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # bounded sample per step: ~1.5 s of wall time on all cores
    vals = []
    target = 450.0 * cores * 1.5
    for i in range(args.warmup):
        cpu_reference_run(cores, target / 4, args.frames)
    t0 = time.perf_counter()
    for i in range(args.steps):
        v, sample = cpu_reference_run(cores, target, args.frames)
        vals.append(v)
    wall = time.perf_counter() - t0
    val = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "decode_audio_seconds_per_second", "value": val, "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * wall / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32+u32", "data": "synthetic",
        "config": workload_config(args.frames),
        "arm_note": "reference libopus 1.5.2 (oracle/_ref) on all host cores; each step decodes a bounded sample of the workload's streams (cpu_baseline.sample)",
        "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def bind_to_gpu_numa(local):
    """Pin this process to the CPUs next to its GPU before any pinned buffer is allocated: a NUMA-remote staging buffer halves the
    PCIe rate and with it the end-to-end number (seen on this pool: 111 k vs 196 k audio-s/s for the same build)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        dev = "/sys/bus/pci/devices/%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open(dev + "/numa_node").read())
        cpus = open(dev + "/local_cpulist").read().strip()
        if node < 0 or not cpus:
            return {"node": node, "bound": False}
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
        return {"node": node, "bound": bool(ids), "cpus": cpus}
    except Exception as ex:      # no sysfs entry, container restrictions: run unbound
        return {"node": None, "bound": False, "why": repr(ex)[:80]}


def pcie_d2h_gbs(dev, nbytes=256 << 20):
    import torch
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    h.copy_(d, non_blocking=True); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    return 3 * nbytes / (e0.elapsed_time(e1) / 1e3) / 1e9


def run_ours(args):
    import torch
    import torch.distributed as dist
    from opus_codec_b200 import _lib
    from opus_codec_b200.batch import BatchDecoder

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # NCCL would print its version banner on stdout, in front of the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))   # plumbing only: barrier + max-reduce of timings
    L = _lib.lib()
    S, F = STREAMS_PER_GPU, args.frames
    pk, ln, rng_expect = load_pool(S, F)        # rank-independent content; streams are sharded by rank (disjoint index ranges)
    stride = pk.shape[2]
    offsets = (np.arange(S * F, dtype=np.int32) * stride).reshape(S, F)

    dec = BatchDecoder(S, 48000, 1, device=local, max_frames=F)
    ext = torch.cuda.ExternalStream(L.ob_decoder_cuda_stream(dec.handle), device=local)
    dev = torch.device("cuda", local)

    # ---- resident inputs / outputs (for `value`) ----
    d_pk = torch.from_numpy(pk.reshape(-1)).to(dev)
    d_off = torch.from_numpy(offsets.reshape(-1)).to(dev)
    d_len = torch.from_numpy(ln.reshape(-1)).to(dev)
    d_pcm = torch.empty(S * F * FRAME, dtype=torch.float32, device=dev)
    d_smp = torch.empty(S * F, dtype=torch.int32, device=dev)
    d_rng = torch.empty(S * F, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    def step_device():
        r = L.ob_decode_float_device(dec.handle, F, d_pk.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), d_pcm.data_ptr(), FRAME,
                                     d_smp.data_ptr(), d_rng.data_ptr(), 0)
        assert r == 0, r

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step_device()
    barrier()
    kms = np.zeros(3)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = dec.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(args.steps):
        step_device()
        kms += np.array(dec.kernel_ms())        # per-kernel CUDA-event times on the launching stream (waits for the step)
    e1.record(ext)
    barrier()
    ms_dev = e0.elapsed_time(e1)
    launches = dec.launches() - launches0
    kms /= args.steps
    # verify what the timed steps produced
    assert (d_smp.cpu().numpy() == FRAME).all()
    assert (d_rng.cpu().numpy().view(np.uint32).reshape(S, F) == rng_expect).all(), "final range mismatch"

    # ---- end to end through the host-pointer C ABI (pinned buffers) ----
    pcie = pcie_d2h_gbs(dev) if not args.kernels_only else None        # the ceiling of this leg: 788 MB of PCM per step cross PCIe
    h_pk = torch.from_numpy(pk.reshape(-1).copy()).pin_memory()
    h_off = torch.from_numpy(offsets.reshape(-1).copy()).pin_memory()
    h_len = torch.from_numpy(ln.reshape(-1).copy()).pin_memory()
    # Two calls in flight, as a server draining jitter buffers would run it: while call n's PCM travels to the host, call n+1's
    # kernels run (ob_decode_float_multi_async / ob_decoder_wait).  Every step uploads its packets from pinned memory and
    # downloads PCM + samples + final ranges into pinned memory; a step's result is consumed (its final ranges read) once
    # ob_decoder_wait says it is complete.
    h_pcm = [torch.empty(S * F * FRAME, dtype=torch.float32).pin_memory() for _ in range(2)]
    h_smp = [torch.empty(S * F, dtype=torch.int32).pin_memory() for _ in range(2)]
    h_rng = [torch.empty(S * F, dtype=torch.int32).pin_memory() for _ in range(2)]
    consumed = []

    def submit(n):
        p = n & 1
        r = L.ob_decode_float_multi_async(dec.handle, F, h_pk.data_ptr(), h_off.data_ptr(), h_len.data_ptr(), h_pcm[p].data_ptr(), FRAME,
                                          h_smp[p].data_ptr(), h_rng[p].data_ptr())
        assert r == 0, r

    def consume(n):
        consumed.append(int(h_rng[n & 1][-1]))            # the read of the step's result

    # With two calls in flight the pipeline's fill (the first call's kernels run with nothing to copy) and drain (the last call's copies with nothing to compute) cost one
    # step's worth of time however many steps are timed: at least 32 steps (~2 s), so that this one-off is ~3 % of the region (it was 10 % at 10 steps); `e2e.steps` says.
    e2e_steps = 0 if args.kernels_only else max(32, args.steps)
    if not args.kernels_only:
        for n in range(max(3, min(args.warmup, 4))):          # warm-up: both output buffers of the decoder get allocated here, not in the timed region
            submit(n)
            assert L.ob_decoder_wait(dec.handle, 0) == 0
    barrier()
    t0 = time.perf_counter()
    for n in range(e2e_steps):
        submit(n)
        if n:
            assert L.ob_decoder_wait(dec.handle, 1) == 0
            consume(n - 1)
    if e2e_steps:
        assert L.ob_decoder_wait(dec.handle, 0) == 0
        consume(e2e_steps - 1)
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3
    barrier()
    # the same loop through the int16 API (Decoder::decode): soft clip + rounding on the GPU, half the bytes over PCIe
    ms_i16 = None
    if e2e_steps:
        h_i16 = [torch.empty(S * F * FRAME, dtype=torch.int16).pin_memory() for _ in range(2)]

        def submit16(n):
            p = n & 1
            r = L.ob_decode_multi_async(dec.handle, F, h_pk.data_ptr(), h_off.data_ptr(), h_len.data_ptr(), h_i16[p].data_ptr(), FRAME,
                                        h_smp[p].data_ptr(), h_rng[p].data_ptr())
            assert r == 0, r
        for n in range(3):
            submit16(n)
            assert L.ob_decoder_wait(dec.handle, 0) == 0
        t1 = time.perf_counter()
        for n in range(e2e_steps):
            submit16(n)
            if n:
                assert L.ob_decoder_wait(dec.handle, 1) == 0
                consume(n - 1)
        assert L.ob_decoder_wait(dec.handle, 0) == 0
        consume(e2e_steps - 1)
        torch.cuda.synchronize()
        ms_i16 = (time.perf_counter() - t1) * 1e3
        barrier()
    clocks = sampler.stop() if rank == 0 else None
    if not args.kernels_only:
        for p in range(2):
            assert (h_rng[p].numpy().view(np.uint32).reshape(S, F) == rng_expect).all()
            assert (h_smp[p].numpy() == FRAME).all()
    h2d = int(h_pk.numel() + 4 * h_off.numel() + 4 * h_len.numel())
    d2h = int(4 * h_pcm[0].numel() + 4 * h_smp[0].numel() + 4 * h_rng[0].numel())

    from opus_codec_b200.shard import max_over_ranks
    ms_dev, ms_e2e = max_over_ranks(ms_dev, dev), max_over_ranks(ms_e2e, dev)     # slowest rank defines the job's time
    if ms_i16 is not None:
        ms_i16 = max_over_ranks(ms_i16, dev)
    audio_per_step = world * S * F * 0.02
    value = audio_per_step * args.steps / (ms_dev / 1000.0)
    e2e = audio_per_step * e2e_steps / (ms_e2e / 1000.0) if e2e_steps else None

    # what the host can absorb when every rank copies at once (the float leg's ceiling at N > 1), measured in this very run
    host_ceiling = host_d2h_ceiling(dev, barrier) if not args.kernels_only else None
    host_ceiling_sum = None
    if host_ceiling is not None:
        t = torch.tensor([host_ceiling], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        host_ceiling_sum = float(t.item())
    encode = transcode = live = None
    dec.close()
    def leg(fn):                                                 # a secondary leg that fails must not take the headline line with it
        try:
            return fn(args, L, local, world, rank, dev, barrier)
        except Exception as ex:
            import traceback
            traceback.print_exc()
            return {"error": repr(ex)[:300]}
    if not args.no_encode:
        encode = leg(run_encode_leg)
        transcode = leg(run_transcode_leg)
    if not args.no_live and not args.kernels_only:
        live = leg(run_live_leg)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        names = ["ob_k_symbols", "ob_k_bands", "ob_k_synth"]
        dom = int(np.argmax(kms))
        bytes_per_launch = algorithmic_bytes_per_frame(F) * S * F
        achieved = bytes_per_launch / (kms[dom] / 1000.0) / 1e9
        traffic, traffic_src = None, None
        try:      # DRAM bytes of one launch of that kernel at this shape, from the committed ncu --set full capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_dram_traffic.json")))["decode"]
            traffic, traffic_src = tj["kernels"][names[dom]]["dram_bytes_per_frame"] * S * F, tj["source"]      # captured per frame at tj["frames_per_launch"]
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)", "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": bytes_per_launch, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                "kernel_ms": dict(zip(names, [float(v) for v in kms])),
                "algorithmic_bytes_per_frame": algorithmic_bytes_per_frame(F),
                "pipeline_achieved_GBps": bytes_per_launch / (float(kms.sum()) / 1000.0) / 1e9}
        cpu = None
        if not args.kernels_only:
            cores = os.cpu_count() or 1
            try:
                v, sample = cpu_reference_run(cores, 450.0 * cores * 2.5, 50)
                cpu = {"value": v, "unit": "audio-s/s", "cores": cores, "kind": "reference", "sample": sample}
            except Exception as ex:   # the reference .so did not travel: report, do not fake
                cpu = {"value": None, "unit": "audio-s/s", "cores": cores, "kind": "reference", "sample": "unavailable: %r" % (ex,)}
        line = {
            "metric": "decode_audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32+u32", "data": "synthetic",
            "config": workload_config(F),
            "arm_note": "value is offline throughput (F frames per call); the real-time capacity (F = 1) is in `live`",
            "e2e": {"value": e2e, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "pcie_d2h_GBps": pcie, "d2h_achieved_GBps": (d2h * e2e_steps / (ms_e2e / 1000.0) / 1e9) if e2e_steps else None,
                    "host_ceiling_GBps": host_ceiling, "host_ceiling_all_ranks_GBps": host_ceiling_sum,
                    "host_ceiling_note": "device->pinned-host copy rate of this rank while ALL ranks copy at once (1 s, 512 MB copies): what the host can absorb; "
                                         "d2h_achieved_GBps / host_ceiling_GBps is how much of it the float leg uses",
                    "pcie_note": "the float leg needs d2h_bytes_per_step x steps/s of sustained host-bound traffic; pcie_d2h_GBps is a short 256 MB probe", "numa": numa,
                    "int16_api": {"value": audio_per_step * e2e_steps / (ms_i16 / 1000.0) if ms_i16 else None, "unit": "audio-s/s",
                                  "d2h_bytes_per_step": int(2 * h_pcm[0].numel() + 8 * h_smp[0].numel())},
                    "mode": "host wall clock; two host-pointer calls in flight (ob_decode_float_multi_async + ob_decoder_wait), pinned buffers"},
            "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "clocks": clocks, "encode": encode, "transcode": transcode,
            "decode_stereo": transcode.get("decode_stereo") if transcode else None, "live": live,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=200, help="consecutive 20 ms frames per stream per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--enc-frames", type=int, default=10, help="frames per stream per step of the encode leg")
    ap.add_argument("--no-encode", action="store_true", help="skip the secondary encode / transcode measurements")
    ap.add_argument("--no-live", action="store_true", help="skip the live-streaming (one frame per call) sweeps")
    ap.add_argument("--kernels-only", action="store_true", help="profiling aid: skip the e2e and cpu_baseline legs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
