"""CPU test: the product's per-thread / per-warp / per-block DEVICE code (opus_codec_b200/csrc/*.cuh), compiled by g++
with one emulated lane (tests/host_emul/emul.cpp), must agree with the oracle.  This validates the kernels' logic on the
GPU-less build box; the GPU tests then only have to catch synchronisation and hardware-specific problems."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden_names, load_golden
from oracle import oraclepy

EMU = os.path.join(ROOT, "tests", "host_emul")


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(EMU, "libemul.so")
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", so, os.path.join(EMU, "emul.cpp")], check=True)
    return C.CDLL(so)


def P(a, t):
    return a.ctypes.data_as(C.POINTER(t))


@pytest.mark.parametrize("name", golden_names())
def test_device_code_single_lane_matches_oracle(emul, name):
    g = load_golden(name)
    fs, dc = g["frame_size"], g["dec_channels"]
    for s in range(min(2, g["packets"].shape[0])):
        pk = np.ascontiguousarray(g["packets"][s]); ln = np.ascontiguousarray(g["lens"][s]); nf = pk.shape[0]
        opcm, orng, osmp = oraclepy.decode_stream(pk, ln, fs, dc)
        pcm = np.zeros((nf, fs * dc), np.float32); rng = np.zeros(nf, np.uint32); smp = np.zeros(nf, np.int32)
        emul.emul_decode_stream(P(pk, C.c_ubyte), P(ln, C.c_int), pk.shape[1], nf, fs, dc, P(pcm, C.c_float), P(rng, C.c_uint32), P(smp, C.c_int), None)
        assert (smp == osmp).all() and (rng == orng).all()
        assert np.abs(pcm - opcm).max() <= 1e-6
