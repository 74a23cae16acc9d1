"""Containers (opus_codec_b200/containers.py): the .bit format against the reference's own tool (oracle/_ref/opus_demo reads what we
write and checks every final range; we read what it writes), RTP and Ogg Opus by construction, round trip and malformed input."""
import os
import struct
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ROOT, load_golden

DEMO = os.path.join(ROOT, "oracle", "_ref", "opus_demo")


def golden_packets(name, s=0, n=40):
    g = load_golden(name)
    return g, [bytes(g["packets"][s][f, :g["lens"][s][f]]) for f in range(min(n, g["packets"].shape[1]))]


def test_bit_container_is_read_by_the_reference_tool_and_back():
    from opus_codec_b200 import containers, synth
    if not os.path.exists(DEMO):
        pytest.skip("oracle/_ref/opus_demo not built")
    g, pk = golden_packets("cfg3_stereo_20ms_96k_cbr")
    rng = [int(r) for r in g["enc_rng"][0][:len(pk)]]
    blob = containers.write_bit(pk, rng)
    assert containers.read_bit(blob) == (pk, rng)
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "a.bit"), "wb").write(blob)
        r = subprocess.run([DEMO, "-d", "48000", "2", os.path.join(d, "a.bit"), os.path.join(d, "a.pcm")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr                                  # a final-range mismatch makes opus_demo fail
        assert os.path.getsize(os.path.join(d, "a.pcm")) == len(pk) * 960 * 2 * 2
        bad = containers.write_bit(pk, rng[:5] + [rng[5] ^ 1] + rng[6:])          # (the tool skips the check on the first packet)
        open(os.path.join(d, "b.bit"), "wb").write(bad)
        r = subprocess.run([DEMO, "-d", "48000", "2", os.path.join(d, "b.bit"), os.path.join(d, "b.pcm")], capture_output=True, text=True)
        assert r.returncode != 0 and "mismatch" in (r.stderr + r.stdout).lower()
        # the other direction: the tool encodes, we read
        pcm = synth.stream_pcm(3, 48000, 2)
        (np.clip(np.rint(pcm * 32768), -32768, 32767).astype("<i2")).tofile(os.path.join(d, "in.pcm"))
        r = subprocess.run([DEMO, "-e", "restricted-lowdelay", "48000", "2", "96000", "-cbr", "-framesize", "10", os.path.join(d, "in.pcm"), os.path.join(d, "c.bit")],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        pk2, rng2 = containers.read_bit(open(os.path.join(d, "c.bit"), "rb").read())
        assert len(pk2) in (100, 101) and all(len(p) == 120 for p in pk2) and all((p[0] & 0x98) == 0x90 for p in pk2)          # CELT-only, 10 ms
    with pytest.raises(ValueError):
        containers.read_bit(struct.pack(">II", 50, 0) + bytes(10))


def test_rtp_pack_unpack_and_loss_reconstruction():
    from opus_codec_b200 import containers as c
    g, pk = golden_packets("cfg2_mono_20ms_64k_cbr", n=30)
    dg = [c.rtp_pack(p, 65530 + i, 1000 + 960 * i, 0xDEADBEEF, marker=(i == 0)) for i, p in enumerate(pk)]          # sequence numbers wrap
    h = c.rtp_unpack(dg[0])
    assert h["marker"] and h["payload_type"] == 111 and h["seq"] == 65530 and h["timestamp"] == 1000 and h["ssrc"] == 0xDEADBEEF and h["payload"] == pk[0]
    assert c.rtp_unpack(dg[7])["seq"] == 1
    # reordering, duplicates, a 3-packet hole and a single hole -> ordered list with b"" where packets are missing
    lost = {5, 6, 7, 20}
    arrive = [d for i, d in enumerate(dg) if i not in lost]
    arrive[2], arrive[3] = arrive[3], arrive[2]
    arrive.insert(10, arrive[9])
    out, stamps = c.rtp_to_packets(arrive)
    assert out == [b"" if i in lost else p for i, p in enumerate(pk)]
    assert stamps == [1000 + 960 * i for i in range(len(pk))]
    # CSRCs, header extension and padding are skipped
    raw = struct.pack(">BBHII", 0x80 | 0x20 | 0x10 | 2, 96, 7, 42, 1) + struct.pack(">II", 11, 12) + struct.pack(">HH", 0xBEDE, 1) + b"\x10\xaa\x00\x00" + pk[0] + b"\x00\x00\x03"
    h = c.rtp_unpack(raw)
    assert h["payload"] == pk[0] and h["csrc"] == [11, 12] and h["extension"] == (0xBEDE, b"\x10\xaa\x00\x00") and h["payload_type"] == 96
    for bad in (b"\x80" * 5, b"\x40" + bytes(11), struct.pack(">BBHII", 0xA0, 96, 0, 0, 0) + b"\x00", struct.pack(">BBHII", 0x8F, 96, 0, 0, 0)):
        with pytest.raises(ValueError):
            c.rtp_unpack(bad)
    # 60 ms packets: the gap is filled in units of the previous packet's duration
    from conftest import repacketize
    p60 = [repacketize(pk[3 * k:3 * k + 3], 3) for k in range(6)]
    dg = [c.rtp_pack(p, k, 2880 * k, 1) for k, p in enumerate(p60)]
    out, _ = c.rtp_to_packets([dg[0], dg[1], dg[4], dg[5]])
    assert out == [p60[0], p60[1], b"", b"", p60[4], p60[5]]
    assert c.packet_duration_48k(p60[0]) == 2880


def test_ogg_opus_write_read_round_trip_and_structure():
    from opus_codec_b200 import containers as c
    assert c.ogg_crc(b"123456789") == 0x765E7680 ^ 0xFFFFFFFF             # the CRC-32/CKSUM core (poly 04c11db7, init 0, not reflected)
    g, pk = golden_packets("stereo_20ms_510k_highrate", n=30)            # 1275-byte packets: 6 lacing values each, pages fill up
    g2, small = golden_packets("cfg4_mono_2p5ms_64k", n=200)
    for packets, ch, fs in ((pk, 2, 960), (small, 1, 120)):
        blob = c.ogg_write(packets, ch, pre_skip=312, output_gain_q8=-256, comments=["TITLE=sweep", "ENCODER=b200"])
        o = c.ogg_read(blob)
        assert o["packets"] == packets and o["channels"] == ch and o["pre_skip"] == 312 and o["output_gain_q8"] == -256 and o["mapping_family"] == 0
        assert o["comments"] == ["TITLE=sweep", "ENCODER=b200"] and o["vendor"] == "opus_codec_b200" and o["eos"]
        assert o["granules"][-1] == len(packets) * fs
        ends = [k for k, v in enumerate(o["granules"]) if v >= 0]
        assert all(o["granules"][k] == (k + 1) * fs for k in ends)
        # page structure: BOS page holds OpusHead only; every page verifies; sequence numbers count up
        assert blob[:4] == b"OggS" and blob[5] == 2 and blob[26] == 1 and blob[27] == 19 and blob[28:36] == b"OpusHead"
        pos, n = 0, 0
        while pos < len(blob):
            nseg = blob[pos + 26]
            assert struct.unpack_from("<I", blob, pos + 18)[0] == n and nseg <= 255
            pos += 27 + nseg + sum(blob[pos + 27:pos + 27 + nseg])
            n += 1
        assert pos == len(blob) and n == 2 + (len(packets) + 49) // 50
        bad = bytearray(blob); bad[len(blob) // 2] ^= 0x10
        with pytest.raises(ValueError, match="checksum"):
            c.ogg_read(bytes(bad))
        assert c.ogg_read(bytes(bad), check_crc=False)["channels"] == ch
    # a packet of exactly 255 bytes takes a terminating zero lacing value; a long comment header spans pages
    p255 = bytes([pk[0][0]]) + bytes(254)
    blob = c.ogg_write([p255, pk[0]], 2, comments=["X=" + "y" * 70000])
    o = c.ogg_read(blob)
    assert o["packets"] == [p255, pk[0]] and len(o["comments"][0]) == 70002
    with pytest.raises(ValueError):
        c.ogg_read(blob[300:])
    with pytest.raises(ValueError):
        c.ogg_read(blob[:len(blob) - 10])
