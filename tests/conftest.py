import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names():
    g = os.path.join(ROOT, "tests", "golden")
    return sorted(f[:-4] for f in os.listdir(g) if f.endswith(".npz") and not f.startswith(("pool_", "plc_")))


def plc_golden_names():
    """Base names of the packet-loss fixtures (tests/golden/plc_<base>.npz, made by make_golden_plc.py)."""
    g = os.path.join(ROOT, "tests", "golden")
    return sorted(f[4:-4] for f in os.listdir(g) if f.endswith(".npz") and f.startswith("plc_"))


def load_plc_golden(base):
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "plc_" + base + ".npz"))
    return {k: z[k] for k in z.files}


def load_golden(name):
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    ch, fs, br, vbr, dec_ch = [int(v) for v in z["meta"]]
    return dict(packets=z["packets"], lens=z["lens"], enc_rng=z["enc_rng"], dec_rng=z["dec_rng"], pcm=z["pcm"],
                channels=ch, frame_size=fs, bitrate=br, vbr=vbr, dec_channels=dec_ch, name=name)


_EMUL = None


def emul_lib():
    """tests/host_emul/libemul.so: the product's device code compiled by g++ for one host lane (TEST INFRASTRUCTURE, see emul.cpp).
    Rebuilt when any source is newer.  The encoder entry points run in the reference's summation order by default and in the 32-lane
    warp's order after emul_set_warp_order(1)."""
    global _EMUL
    if _EMUL is not None:
        return _EMUL
    import ctypes
    import subprocess
    emu = os.path.join(ROOT, "tests", "host_emul")
    so, src = os.path.join(emu, "libemul.so"), os.path.join(emu, "emul.cpp")
    csrc = os.path.join(ROOT, "opus_codec_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-Wno-unknown-pragmas", "-o", so, src], check=True)
    _EMUL = ctypes.CDLL(so)
    return _EMUL


def emul_warp_encode(pcm, fs, ch, br, vbr, cx, app=2051, extras=None, lsb_depth=24):
    """One stream through the encoder's device code on the host, in the WARP's summation order: the packets the GPU must produce."""
    import ctypes as C
    import numpy as np
    L = emul_lib()
    pcm = np.ascontiguousarray(pcm, np.float32)
    nf = pcm.size // (fs * ch)
    out = np.zeros((nf, 1276), np.uint8); lens = np.zeros(nf, np.int32); rng = np.zeros(nf, np.uint32)
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    L.emul_set_warp_order(1)
    L.emul_set_lsb_depth(lsb_depth)
    if extras is not None:
        L.emul_set_encoder_extras(*extras)
    try:
        r = L.emul_opus_encode_stream_app(P(pcm, C.c_float), nf, fs, ch, app, br, vbr, cx, P(out, C.c_ubyte), 1276, P(lens, C.c_int), P(rng, C.c_uint32))
    finally:
        L.emul_set_warp_order(0)
        L.emul_set_lsb_depth(24)
        if extras is not None:
            L.emul_set_encoder_extras(0, 0, 0, 0, 0, 0)
    return out, lens, rng, r


def opus_compare(ref_pcm, test_pcm, channels):
    """oracle/_ref/opus_compare (opus/src/opus_compare.c) on two float PCM arrays; returns (passed, weighted_error, text).  The reference file is
    always read as stereo by the tool, so a mono reference is written with both channels equal."""
    import subprocess
    import tempfile
    import numpy as np
    import re
    exe = os.path.join(ROOT, "oracle", "_ref", "opus_compare")
    to16 = lambda x: np.clip(np.rint(np.asarray(x, np.float64).reshape(-1) * 32768), -32768, 32767).astype("<i2")
    with tempfile.TemporaryDirectory() as d:
        fa, fb = os.path.join(d, "ref.sw"), os.path.join(d, "test.sw")
        a = to16(ref_pcm)
        if channels == 1:
            a = np.repeat(a, 2)
        a.tofile(fa); to16(test_pcm).tofile(fb)
        r = subprocess.run([exe] + (["-s"] if channels == 2 else []) + [fa, fb], capture_output=True, text=True)
    text = (r.stdout + r.stderr).strip().splitlines()[-1]
    m = re.search(r"rror is ([0-9.eE+-]+)", text)
    return r.returncode == 0, float(m.group(1)) if m else float("nan"), text


@pytest.fixture(scope="session")
def have_ref():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libopus_ref.so"))


def repacketize(frames, code, pad=0):
    """Builds one Opus packet (RFC 6716 section 3.2) out of code-0 packets that share a TOC configuration.
    code 1: two equal-size frames; code 2: two frames, first size explicit; code 3: M frames, VBR sizes, `pad` padding bytes."""
    toc = frames[0][0] & 0xFC
    body = [bytes(f[1:]) for f in frames]
    assert all((f[0] & 0xFC) == toc for f in frames)

    def size_bytes(n):
        return bytes([n]) if n < 252 else bytes([252 + (n & 3), (n - (252 + (n & 3))) >> 2])
    if code == 1:
        assert len(body) == 2 and len(body[0]) == len(body[1])
        return bytes([toc | 1]) + body[0] + body[1]
    if code == 2:
        assert len(body) == 2
        return bytes([toc | 2]) + size_bytes(len(body[0])) + body[0] + body[1]
    assert code == 3
    out = bytes([toc | 3, 0x80 | (0x40 if pad else 0) | len(body)])
    if pad:
        p = pad
        while p > 254:
            out += bytes([255]); p -= 254
        out += bytes([p])
    for b in body[:-1]:
        out += size_bytes(len(b))
    out += b"".join(body)
    return out + bytes(pad)


def multiframe_stream(g, s, nframes_out):
    """Code-1/2/3 packets (incl. padding and a DTX frame inside a code-3 packet) built from stream s of a golden's code-0 packets."""
    pk, ln = g["packets"][s], g["lens"][s]
    fr = [bytes(pk[f, :ln[f]]) for f in range(pk.shape[0])]
    pkts, k = [], 0
    kinds = [("c3", 3, 0), ("c1", 2, 0), ("c2", 2, 0), ("c3", 2, 300), ("c0", 1, 0), ("c3", 3, 7), ("c3dtx", 3, 0)]
    while len(pkts) < nframes_out and k + 3 <= len(fr):
        kind, m, pad = kinds[len(pkts) % len(kinds)]
        grp = fr[k:k + m]
        k += m
        if kind == "c0":
            pkts.append(grp[0])
        elif kind == "c1":
            if len(grp[0]) != len(grp[1]):
                pkts.append(repacketize(grp, 2))
            else:
                pkts.append(repacketize(grp, 1))
        elif kind == "c2":
            pkts.append(repacketize(grp, 2))
        elif kind == "c3dtx":
            grp = [grp[0], grp[1][:1], grp[2]]            # middle frame: TOC only -> zero-length frame inside the packet -> concealed
            pkts.append(repacketize(grp, 3))
        else:
            pkts.append(repacketize(grp, 3, pad))
    stride = max(len(p) for p in pkts)
    import numpy as np
    out = np.zeros((len(pkts), stride), np.uint8)
    lens = np.zeros(len(pkts), np.int32)
    for i, p in enumerate(pkts):
        out[i, :len(p)] = np.frombuffer(p, np.uint8); lens[i] = len(p)
    return out, lens




def pathological_pcm(s, n, ch, rng):
    """Inputs an encoder meets in the wild: digital silence, DC, full-scale and beyond (clipping), denormal-sized noise, bursts,
    NaN / Inf samples (the reference zeroes such a frame after its DC-reject filter, opus_encoder.c:1966-1976)."""
    import numpy as np
    from opus_codec_b200 import synth
    x = synth.stream_pcm(s, n, ch, base_seed=31).reshape(n, ch).copy()
    kind = s % 8
    if kind == 0: x[:] = 0
    elif kind == 1: x[:] = 0.25
    elif kind == 2: x *= 6.0                                  # far beyond full scale
    elif kind == 3: x = (rng.standard_normal((n, ch)) * 1e-7).astype(np.float32)
    elif kind == 4: x[n // 3:n // 3 + 960] = 0; x[n // 2:n // 2 + 40] = 3.0
    elif kind == 5: x[n // 4, 0] = np.nan; x[n // 2 + 7, ch - 1] = np.inf
    elif kind == 6: x[::2] *= -1.0                             # energy at Nyquist
    else: x = np.sign(x) * np.float32(0.99)                     # square-ish wave at full scale
    return np.ascontiguousarray(x.reshape(-1), np.float32)


