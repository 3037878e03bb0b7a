"""Deterministic synthetic Iridium recordings.

There is no network and no recorded capture in this sandbox, so every parity
test and every bench number runs on signals made here: complex AWGN plus
planted DL bursts (16-symbol preamble + 12-symbol unique word + QPSK payload,
RRC alpha=0.4 shaped, 25 ksym/s) whose ground-truth DQPSK bit strings are known.
The recipe was validated against the unmodified reference in the survey and is
re-validated by tests/test_oracle_ref.py.

Symbol conventions follow the reference: quadrant -> symbol map
qpsk_demod.c:218-225, unique word iridium.h:30, DQPSK table qpsk_demod.c:46.
"""
from __future__ import annotations

import dataclasses
import math
from typing import List, Optional, Sequence

import numpy as np

SYMBOL_RATE = 25_000
UW_DL = [0, 2, 2, 2, 2, 0, 0, 0, 2, 0, 0, 2]
UW_UL = [2, 2, 0, 0, 0, 2, 0, 0, 2, 0, 2, 2]
_CONST = {0: 1 + 1j, 1: -1 + 1j, 2: -1 - 1j, 3: 1 - 1j}
_DQPSK = [0, 2, 3, 1]
ACCESS_DL = "001100000011000011110011"   # frame_decode.c:51-56
ACCESS_UL = "110011000011110011111100"


def rrc_taps(alpha: float = 0.4, sps: int = 10, ntaps: int = 81) -> np.ndarray:
    """Unit-energy root-raised-cosine (closed form as fir_filter.c:74-111, in double)."""
    c = ntaps // 2
    h = np.zeros(ntaps)
    for i in range(ntaps):
        t = (i - c) / sps
        if abs(t) < 1e-12:
            h[i] = 1.0 - alpha + 4.0 * alpha / math.pi
        elif abs(abs(t) - 1.0 / (4.0 * alpha)) < 1e-9:
            h[i] = alpha / math.sqrt(2.0) * (
                (1 + 2 / math.pi) * math.sin(math.pi / (4 * alpha))
                + (1 - 2 / math.pi) * math.cos(math.pi / (4 * alpha)))
        else:
            num = math.sin(math.pi * t * (1 - alpha)) + 4 * alpha * t * math.cos(math.pi * t * (1 + alpha))
            den = math.pi * t * (1 - (4 * alpha * t) ** 2)
            h[i] = num / den
    return h / math.sqrt(np.sum(h * h))


def expected_bits(symbols: Sequence[int]) -> str:
    """DQPSK bits the demodulator prints for UW+payload symbols (old symbol starts at 0)."""
    old = 0
    out = []
    for v in symbols:
        d = _DQPSK[(int(v) - old) % 4]
        old = int(v)
        out.append(f"{d >> 1}{d & 1}")
    return "".join(out)


def symbols_for_bits(bits: Sequence[int], uplink: bool = False) -> List[int]:
    """Payload symbols whose DQPSK bits (after the unique word) are `bits`: the inverse of expected_bits()."""
    inv = {d: k for k, d in enumerate(_DQPSK)}
    old = (UW_UL if uplink else UW_DL)[-1]
    out = []
    for i in range(0, len(bits) - 1, 2):
        old = (old + inv[(int(bits[i]) << 1) | int(bits[i + 1])]) % 4
        out.append(old)
    return out


def burst_waveform(rng: np.random.Generator, up: int, n_payload: int = 179,
                   uplink: bool = False, payload_bits: Optional[Sequence[int]] = None):
    """One burst at the capture rate (250 kHz * up).  Returns (waveform, bit string).  payload_bits: the bits
    the demodulator should print after the 24 access-code bits (default: random symbols)."""
    from scipy.signal import resample_poly

    pay = rng.integers(0, 4, n_payload) if payload_bits is None else symbols_for_bits(payload_bits, uplink)
    if uplink:
        pre = [2 if i % 2 == 0 else 0 for i in range(16)]
        uw = UW_UL
    else:
        pre = [0] * 16
        uw = UW_DL
    syms = list(pre) + list(uw) + [int(v) for v in pay]
    x = np.zeros(len(syms) * 10, dtype=np.complex128)
    x[::10] = [_CONST[s] / math.sqrt(2.0) for s in syms]
    y = np.convolve(x, rrc_taps()) * math.sqrt(10.0)
    hi = resample_poly(y, up, 1)
    return hi.astype(np.complex128), expected_bits(list(uw) + [int(v) for v in pay])


@dataclasses.dataclass
class PlantedBurst:
    start: int            # first sample of the waveform in the recording
    freq_hz: float        # offset from the capture centre
    amp: float
    snr_db: float
    bits: str
    n_payload: int
    uplink: bool = False


@dataclasses.dataclass
class Recording:
    iq: np.ndarray               # complex64 [n] (cf32) -- or int16/int8 [2n] for ci16/ci8
    fmt: str                     # "cf32" | "ci16" | "ci8"
    sample_rate: int
    center_freq: float
    truth: List[PlantedBurst]

    @property
    def n_samples(self) -> int:
        return self.iq.shape[0] if self.fmt == "cf32" else self.iq.shape[0] // 2


def default_fft_size(sample_rate: int) -> int:
    """burst_detect.c:181-186."""
    return 1 << int(round(math.log2(sample_rate / 1000.0)))


def make_recording(seed: int, sample_rate: int = 10_000_000, duration_s: float = 1.5,
                   n_bursts: int = 12, snr_db=(12.0, 25.0), fmt: str = "cf32",
                   center_freq: float = 1_622_000_000.0, sigma: float = 0.01,
                   channels: Optional[Sequence[float]] = None,
                   starts_s: Optional[Sequence[float]] = None,
                   n_payload: int = 179, uplink_fraction: float = 0.0,
                   waveform_pool: int = 0,
                   frame_bits: Optional[Sequence[Sequence[int]]] = None) -> Recording:
    """Noise + planted bursts.  All randomness from numpy default_rng(seed).

    channels: candidate centre offsets in Hz (default: 41.667 kHz raster inside
    +-(fs/2-200 kHz), |f|>=60 kHz).  starts_s: explicit burst start times; default
    spreads bursts uniformly after the 512-frame quiet lead-in and keeps same-channel
    reuse >= 27 ms apart (SURVEY.md 8d constraints).  frame_bits: whole frames as bit arrays (24 access-code
    bits first, e.g. from tests/frame_gen.py); burst k carries frame_bits[k % len] instead of random symbols.
    """
    rng = np.random.default_rng(seed)
    fs = int(sample_rate)
    up = fs // 250_000
    n = int(round(duration_s * fs))
    nfft = default_fft_size(fs)
    lead = 512 * nfft / fs + 0.02

    noise = rng.standard_normal((n, 2), dtype=np.float32)
    sig = (noise[:, 0] + 1j * noise[:, 1]).astype(np.complex64)
    sig *= np.float32(sigma)
    del noise

    if channels is None:
        raster = 1e6 / 24.0
        kmax = int((fs / 2 - 200e3) / raster)
        channels = [k * raster for k in range(-kmax, kmax + 1) if abs(k * raster) >= 60e3]
    channels = list(channels)

    pool = []
    if waveform_pool > 0:
        for _ in range(waveform_pool):
            pool.append(burst_waveform(rng, up, n_payload, False))

    if frame_bits:
        n_payload = max((len(fb) - 24) // 2 for fb in frame_bits)
    blen = (16 + 12 + n_payload) * 10 * up + 81 * up
    if starts_s is None:
        lo, hi = lead, duration_s - blen / fs - 0.03
        if hi <= lo:
            raise ValueError("recording too short for the 512-frame lead-in plus one burst")
        starts_s = np.sort(rng.uniform(lo, hi, n_bursts))
    truth: List[PlantedBurst] = []
    last_use = {}
    for t0 in starts_s:
        # choose a channel whose previous use ended >= 27 ms ago
        for _ in range(64):
            ch = channels[int(rng.integers(0, len(channels)))]
            if t0 - last_use.get(ch, -1.0) >= 0.027 + blen / fs:
                break
        else:
            continue
        last_use[ch] = float(t0)
        ul = bool(rng.random() < uplink_fraction)
        if frame_bits:
            fb = [int(b) for b in frame_bits[len(truth) % len(frame_bits)]]
            ul = "".join(map(str, fb[:24])) == ACCESS_UL
            hi_wave, bits = burst_waveform(rng, up, 0, ul, payload_bits=fb[24:])
        elif pool and not ul:
            hi_wave, bits = pool[int(rng.integers(0, len(pool)))]
        else:
            hi_wave, bits = burst_waveform(rng, up, n_payload, ul)
        snr = float(rng.uniform(snr_db[0], snr_db[1]))
        amp = sigma * 10.0 ** (snr / 20.0)
        f = ch + float(rng.uniform(-3e3, 3e3))
        ph = float(rng.uniform(0, 2 * math.pi))
        s0 = int(round(t0 * fs))
        m = min(len(hi_wave), n - s0)
        if m <= 0:
            continue
        k = np.arange(m, dtype=np.float64)
        rot = np.exp(1j * (2 * math.pi * f / fs * k + ph))
        sig[s0:s0 + m] += (amp * hi_wave[:m] * rot).astype(np.complex64)
        truth.append(PlantedBurst(s0, f, amp, snr, bits, (len(bits) - 24) // 2, ul))

    if fmt == "cf32":
        iq = sig
    elif fmt == "ci16":
        # the reference keeps only the upper byte (main.c:245-246): scale so it carries signal
        v = sig.view(np.float32) * np.float32(32768.0 * 4.0)
        iq = np.round(np.clip(v, -32767, 32767)).astype(np.int16)
    elif fmt == "ci8":
        v = sig.view(np.float32) * np.float32(128.0 * 4.0)
        iq = np.round(np.clip(v, -127, 127)).astype(np.int8)
    else:
        raise ValueError(fmt)
    return Recording(iq, fmt, fs, center_freq, truth)


def make_dense_recording(seed: int = 1234, sample_rate: int = 10_000_000) -> Recording:
    """BASELINE config 4: 672 bursts = 56 channels x 12 slots inside ~150 ms."""
    rng = np.random.default_rng(seed)
    fs = sample_rate
    up = fs // 250_000
    dur = 0.85
    n = int(dur * fs)
    sigma = 0.01
    noise = rng.standard_normal((n, 2), dtype=np.float32)
    sig = ((noise[:, 0] + 1j * noise[:, 1]) * np.float32(sigma)).astype(np.complex64)
    truth: List[PlantedBurst] = []
    raster = 1e6 / 24.0
    for slot in range(12):
        for c in range(56):
            ch = -4.0e6 + (3 * c + slot % 3) * raster
            t0 = 0.5 + slot * 0.010 + float(rng.uniform(0, 0.001))
            hi_wave, bits = burst_waveform(rng, up, 179, False)
            snr = float(rng.uniform(6.0, 14.0))
            amp = sigma * 10.0 ** (snr / 20.0)
            f = ch + float(rng.uniform(-3e3, 3e3))
            ph = float(rng.uniform(0, 2 * math.pi))
            s0 = int(round(t0 * fs))
            m = min(len(hi_wave), n - s0)
            k = np.arange(m, dtype=np.float64)
            sig[s0:s0 + m] += (amp * hi_wave[:m] * np.exp(1j * (2 * math.pi * f / fs * k + ph))).astype(np.complex64)
            truth.append(PlantedBurst(s0, f, amp, snr, bits, 179))
    return Recording(sig, "cf32", fs, 1_622_000_000.0, truth)


def make_tone_recording(seed: int, n_tones: int, dur_s: float, t0_s: float, total_s: float = 0.62,
                        snr_db: float = 20.0, sample_rate: int = 10_000_000, sigma: float = 0.01) -> np.ndarray:
    """Noise + n_tones unmodulated carriers switched on together for dur_s seconds from t0_s: the
    detector edge cases (more carriers than max_bursts -> squelch, burst_detect.c:593-631; one long
    carrier -> a burst exceeding max_burst_len and the forced baseline update of :498-517)."""
    rng = np.random.default_rng(seed)
    fs = sample_rate
    n = int(total_s * fs)
    x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) * sigma).astype(np.complex64)
    k = np.arange(int(dur_s * fs))
    s0 = int(t0_s * fs)
    amp = sigma * 10 ** (snr_db / 20)
    for i in range(n_tones):
        f = -4.7e6 + i * (9.4e6 / max(n_tones - 1, 1)) if n_tones > 1 else 1.0e6
        if abs(f) < 60e3:
            f += 120e3
        x[s0:s0 + len(k)] += (amp * np.exp(2j * np.pi * (f / fs) * k + 1j * rng.uniform(0, 6.28))).astype(np.complex64)
    return x
