"""Seeded synthetic 48 kHz PCM for parity sets and throughput runs.

Three signal classes mixed 1:1:1 by stream index:
  0: log sine sweep 50 Hz -> 16 kHz at -9 dBFS, random start phase / sweep rate
  1: pink-ish filtered Gaussian noise at -20 dBFS
  2: transients: 1-sample and 5-ms-burst clicks (2-8 per second) over -40 dBFS noise
Stereo = two decorrelated instances mixed 0.7/0.3.  Output is float32 in [-1, 1], interleaved.
"""
import numpy as np
from scipy.signal import lfilter

FS = 48000


def _sweep(rng, n):
    t = np.arange(n) / FS
    dur = rng.uniform(2.0, 10.0)
    f0, f1 = 50.0, 16000.0
    k = np.log(f1 / f0) / dur
    tt = np.mod(t + rng.uniform(0, dur), dur)
    phase = 2 * np.pi * f0 * (np.exp(k * tt) - 1.0) / k + rng.uniform(0, 2 * np.pi)
    return (10 ** (-9 / 20)) * np.sin(phase)


def _pink(rng, n):
    w = rng.standard_normal(n + 64)
    acc = np.zeros_like(w)
    # pink-ish shaping: three one-pole low-passes y[n] = a*y[n-1] + (1-a)*w[n], variance-normalised and mixed
    for a, g in ((0.99, 0.6), (0.9, 0.3), (0.5, 0.1)):
        acc += g * lfilter([1 - a], [1, -a], w) / np.sqrt((1 - a) / (1 + a))
    y = acc[64:]
    y = y / (np.std(y) + 1e-12)
    return (10 ** (-20 / 20)) * y


def _transients(rng, n):
    y = (10 ** (-40 / 20)) * rng.standard_normal(n)
    nclicks = max(1, int(rng.uniform(2, 8) * n / FS))
    for _ in range(nclicks):
        p = int(rng.integers(0, n))
        if rng.random() < 0.5:
            y[p] += rng.choice([-1.0, 1.0]) * rng.uniform(0.3, 0.9)
        else:
            m = min(n - p, int(0.005 * FS))
            y[p:p + m] += rng.uniform(0.2, 0.7) * rng.standard_normal(m) * np.hanning(m + 2)[1:-1] if m > 2 else 0.0
    return np.clip(y, -1.0, 1.0)


_CLASSES = (_sweep, _pink, _transients)


def stream_pcm(stream, nsamples, channels=1, base_seed=1234, klass=None):
    """One stream's PCM: float32 [nsamples*channels] interleaved."""
    rng = np.random.default_rng(base_seed + stream)
    gen = _CLASSES[(stream if klass is None else klass) % 3]
    if channels == 1:
        return gen(rng, nsamples).astype(np.float32)
    a = gen(rng, nsamples)
    b = gen(rng, nsamples)
    out = np.empty((nsamples, 2), np.float64)
    out[:, 0] = 0.7 * a + 0.3 * b
    out[:, 1] = 0.3 * a + 0.7 * b
    return out.astype(np.float32).reshape(-1)


def batch_pcm(nstreams, nsamples, channels=1, base_seed=1234):
    """[nstreams, nsamples*channels] float32."""
    return np.stack([stream_pcm(s, nsamples, channels, base_seed) for s in range(nstreams)])
