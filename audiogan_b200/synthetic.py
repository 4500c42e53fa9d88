"""Synthetic 8 kHz telephone-band waveforms and step inputs (the Fisher corpus is not
available offline).  Mirrors what ``dataset.py`` hands the training loop:
zero-padded (B, L) float32 waveforms peak-normalised to max|x| = 1 (dataset.py:68-71) and
lengths rounded up to the generator frame (dataset.py:57).  SURVEY.md section 8(d).
"""
import numpy as np
import torch


def telephone_band_waveforms(batch, nsamples, seed=1234, lengths=None, rate=8000,
                             band=(300.0, 3400.0)):
    """White N(0,1) -> FFT band-pass 300-3400 Hz -> slow random envelope -> peak-normalise."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((batch, nsamples))
    spec = np.fft.rfft(x, axis=1)
    freqs = np.fft.rfftfreq(nsamples, d=1.0 / rate)
    spec[:, (freqs < band[0]) | (freqs > band[1])] = 0
    x = np.fft.irfft(spec, n=nsamples, axis=1)
    nknots = max(4, nsamples // 800)
    knots = rng.uniform(0.1, 1.0, size=(batch, nknots))
    env = np.stack([np.interp(np.linspace(0, nknots - 1, nsamples), np.arange(nknots), k) for k in knots])
    x = x * env
    if lengths is not None:
        mask = np.arange(nsamples)[None, :] < np.asarray(lengths)[:, None]
        x = x * mask
    x = x / np.abs(x).max(axis=1, keepdims=True)
    return torch.from_numpy(x.astype(np.float32))


def mixed_lengths(batch, nsamples, frame=200, seed=1234):
    """U{L/4..L} rounded up to the frame size; sample 0 keeps the full length."""
    rng = np.random.default_rng(seed + 7)
    ln = rng.integers(nsamples // 4, nsamples + 1, size=batch)
    ln = (ln + frame - 1) // frame * frame
    ln = np.minimum(ln, nsamples)
    ln[0] = nsamples
    return torch.from_numpy(ln.astype(np.int64))


def step_inputs(batch, nsamples, seed=1234, embed=100, noise_size=100, frame=200, noisescale=0.01,
                full_length=True):
    """Everything one core step (1 D-update + 1 G-update) consumes, as CPU fp32 tensors."""
    g = torch.Generator().manual_seed(seed)
    nfr = (nsamples + frame - 1) // frame
    ln = torch.full((batch,), nsamples, dtype=torch.int64) if full_length else mixed_lengths(batch, nsamples, frame, seed)
    r = lambda *s: torch.randn(*s, generator=g)
    return {
        "real": telephone_band_waveforms(batch, nsamples, seed, None if full_length else ln.numpy()),
        "real_len": ln,
        "c_real": r(batch, embed), "c_g": r(batch, embed), "c_d2": r(batch, embed),
        "z": r(batch, nfr, noise_size),
        "noise_real": r(batch, nsamples) * noisescale, "noise_fake": r(batch, nsamples) * noisescale,
        # G-update draws
        "g_c_g": r(batch, embed), "g_c_d": r(batch, embed), "g_z": r(batch, nfr, noise_size),
        "g_noise_fake": r(batch, nsamples) * noisescale,
    }
