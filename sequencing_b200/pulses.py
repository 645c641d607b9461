"""Sampled pulse waveforms: the *values* in the coefficient tables the engine streams.

Host-side numpy, mirroring the public surface of the reference's ``sequencing/pulses.py``
(``array_pulse :31-89``, ``gaussian_wave :92``, ``gaussian_deriv_wave :100``,
``gaussian_chop_norm :113``, ring-ups ``:120-173``, ``Pulse.__call__ :262-293`` and the
``Pulse`` subclasses ``:315-422``).  Sample grids are built exactly as the reference builds
them (``linspace(-L/2, L/2, int(L/dt))`` etc.) because those samples *are* the engine's input:
the waveform is later turned into a natural cubic spline by the solver.
On-device synthesis of the same waveforms is the "next" row N3 of SURVEY.md section 8(f).
"""
import inspect
import re
from functools import lru_cache

import attr
import numpy as np
from scipy.integrate import quad

from .parameters import (
    BoolParameter,
    FloatParameter,
    GigahertzParameter,
    IntParameter,
    NanosecondParameter,
    Parameterized,
    RadianParameter,
    StringParameter,
)


def powerlaw_psd_gaussian(exponent, size, rng=None):
    """Gaussian noise with PSD ~ 1/f**exponent (Timmer & Koenig 1995), unit variance.

    Stand-in for ``colorednoise.powerlaw_psd_gaussian`` (absent here), same call signature as
    used by the reference at ``pulses.py:77-79``.
    """
    rng = np.random.default_rng() if rng is None else rng
    size = [size] if np.isscalar(size) else list(size)
    n = size[-1]
    f = np.fft.rfftfreq(n)
    f[0] = f[1] if n > 1 else 1.0
    scale = f ** (-exponent / 2.0)
    w = scale[1:].copy()
    if n % 2 == 0:
        w[-1] *= 0.5  # the Nyquist term is real
    sigma = 2 * np.sqrt(np.sum(w ** 2)) / n
    shape = size[:-1] + [len(f)]
    re = rng.normal(scale=scale, size=shape)
    im = rng.normal(scale=scale, size=shape)
    im[..., 0] = 0
    re[..., 0] *= np.sqrt(2)
    if n % 2 == 0:
        im[..., -1] = 0
        re[..., -1] *= np.sqrt(2)
    return np.fft.irfft(re + 1j * im, n=n, axis=-1) / sigma


def array_pulse(i_wave, q_wave=None, amp=1, phase=0, detune=0, dt=1.0,
                noise_sigma=0, noise_alpha=0, scale_noise=False):
    """``amp * (I + iQ) * exp(-2 pi i detune t) * exp(i phase)`` (+ optional additive noise)."""
    wave = np.asarray(i_wave) + 1j * (0 if q_wave is None else np.asarray(q_wave))
    n = wave.size
    if detune:
        wave = wave * np.exp(-2j * np.pi * np.linspace(0, n * dt, n) * detune)
    noise = 0
    if noise_sigma:
        ni, nq = noise_sigma * powerlaw_psd_gaussian(noise_alpha, [2, n])
        noise = ni + 1j * nq
    wave = amp * (wave + noise) if scale_noise else amp * wave + noise
    return wave * np.exp(1j * phase)


def _centered_grid(length, dt):
    return np.linspace(-length / 2, length / 2, int(length / dt))


def gaussian_wave(sigma, chop=4, dt=1):
    ts = _centered_grid(chop * sigma, dt)
    g = np.exp(-(ts ** 2) / (2.0 * sigma ** 2))
    return (g - g[0]) / (1 - g[0])


def gaussian_deriv_wave(sigma, chop=4, dt=1):
    ts = _centered_grid(chop * sigma, dt)
    edge = np.exp(-ts[0] ** 2 / (2 * sigma ** 2))
    return (0.25 / sigma ** 2) * ts * np.exp(-(ts ** 2) / (2 * sigma ** 2)) / (1 - edge)


def gaussian_chop(t, sigma, t0):
    edge = np.exp(-(t0 ** 2) / (2.0 * sigma ** 2))
    return (np.exp(-(t ** 2) / (2.0 * sigma ** 2)) - edge) / (1 - edge)


@lru_cache()
def gaussian_chop_norm(sigma, chop):
    """Area under the chopped Gaussian (x2): the rotation-angle normalisation (modes.py:746)."""
    half = sigma * chop / 2
    area, _ = quad(gaussian_chop, -half, half, args=(sigma, -half))
    return 2 * area


def ring_up_cos(length, dt=1):
    return 0.5 * (1 - np.cos(np.linspace(0, np.pi, int(length / dt))))


def ring_up_tanh(length, dt=1):
    return (1 + np.tanh(np.linspace(-2, 2, int(length / dt)))) / 2


def ring_up_gaussian_flattop(length, sigma, ramp_offset=None, dt=1):
    off = 0 if ramp_offset is None else ramp_offset
    ts = np.linspace(-length + 1, 0, int(length / dt))
    shifted = np.where(np.abs(ts) < off, 0.0, ts - np.sign(ts) * off)
    p = np.exp(-(shifted ** 2) / (2.0 * sigma ** 2))
    return (p - p[0]) / (1 - p[0])


def ring_up_wave(length, reverse=False, shape="tanh", **kwargs):
    if shape == "cos":
        wave = ring_up_cos(length)
    elif shape == "tanh":
        wave = ring_up_tanh(length)
    elif shape == "gaussian":
        sigma = kwargs.pop("gaussian_sigma", 6)
        wave = ring_up_gaussian_flattop(length, sigma, **kwargs)
    else:
        raise ValueError(f"Shape must be 'cos' or 'tanh', or 'gaussian', not {shape}.")
    return wave[::-1] if reverse else wave


def smoothed_constant_wave(length, sigma, shape="tanh", **kwargs):
    dt = kwargs.get("dt", 1)
    if sigma == 0:
        return np.ones(int(length / dt))
    return np.concatenate([
        ring_up_wave(sigma, shape=shape, **kwargs),
        np.ones(int((length - 2 * sigma) / dt)),
        ring_up_wave(sigma, reverse=True, shape=shape, **kwargs),
    ])


def constant_pulse(length=None, dt=1):
    n = int(length / dt)
    return np.ones(n), np.zeros(n)


def gaussian_pulse(sigma=None, chop=4, drag=0, dt=1):
    return gaussian_wave(sigma, chop=chop, dt=dt), drag * gaussian_deriv_wave(sigma, chop=chop, dt=dt)


def smoothed_constant_pulse(length=None, sigma=None, shape="tanh", dt=1):
    i_wave = smoothed_constant_wave(length, sigma, shape=shape, dt=dt)
    return i_wave, np.zeros_like(i_wave)


def sech_wave(sigma, chop=4, dt=1):
    half = chop * sigma // 2
    ts = np.linspace(-half, half, int(chop * sigma / dt))
    p = 1 / np.cosh(np.pi / (2 * sigma) * ts)
    return (p - p[0]) / (1 - p[0])


def sech_deriv_wave(sigma, chop=4, dt=1):
    rho = np.pi / (2 * sigma)
    half = chop * sigma // 2
    ts = np.linspace(-half, half, int(chop * sigma / dt))
    edge = 1 / np.cosh(rho * ts[0])
    return (-np.sinh(rho * ts) / np.cosh(rho * ts) ** 2 - edge) / (1 - edge)


def sech_pulse(sigma=None, chop=4, drag=0, dt=1):
    i_wave = sech_wave(sigma, chop=chop, dt=dt)
    return i_wave, drag * np.gradient(i_wave) / dt


def slepian_pulse(tau=None, width=10, drag=0, dt=1):
    from scipy.signal.windows import dpss

    n = int(tau)
    i_wave = dpss(n, n * (width / tau / dt) / 2.0)
    i_wave = i_wave / i_wave.max()
    return i_wave, drag * np.gradient(i_wave) / dt


_PARAMETER_NAMES = {}


def _parameter_names(func):
    """Parameter names of a waveform function (``inspect.signature`` once per function: a sweep calls every pulse per point)."""
    key = getattr(func, "__func__", func)
    names = _PARAMETER_NAMES.get(key)
    if names is None:
        names = _PARAMETER_NAMES[key] = frozenset(inspect.signature(func).parameters)
    return names


@attr.s
class Pulse(Parameterized):
    """Parameterised complex waveform generator: ``pulse_func`` shapes, ``array_pulse`` modulates."""

    pulse_func = staticmethod(constant_pulse)
    amp = FloatParameter(1)
    detune = GigahertzParameter(0)
    phase = RadianParameter(0)
    dt = NanosecondParameter(1)
    noise_sigma = FloatParameter(0)
    noise_alpha = FloatParameter(0)
    scale_noise = BoolParameter(False)

    def __call__(self, **kwargs):
        shape_names = _parameter_names(self.pulse_func)
        mod_names = _parameter_names(array_pulse)
        params = self.as_dict()
        # parameters feed the shape function first (so ``dt`` goes there), overrides passed
        # as keyword arguments feed the modulation first -- same precedence as pulses.py:276-287
        shape_kw = {k: v for k, v in params.items() if k in shape_names}
        mod_kw = {k: v for k, v in params.items() if k in mod_names and k not in shape_names}
        for key in list(kwargs):
            if key in mod_names:
                mod_kw[key] = kwargs.pop(key)
            elif key in shape_names:
                shape_kw[key] = kwargs.pop(key)
        waves = self.pulse_func(**shape_kw)
        i_wave, q_wave = waves if len(waves) == 2 else (waves, None)
        return array_pulse(i_wave, q_wave=q_wave, **mod_kw)

    def plot(self, ax=None, grid=True, legend=True, **kwargs):
        import matplotlib.pyplot as plt

        if ax is None:
            _, ax = plt.subplots()
        wave = self(**kwargs)
        ts = np.linspace(0, len(wave) * kwargs.get("dt", self.dt), len(wave))
        (line,) = ax.plot(ts, wave.real, ls="-", label=self.name)
        ax.plot(ts, wave.imag, color=line.get_color(), ls="--")
        ax.grid(grid)
        if legend:
            ax.legend(loc="best")
        return ax


@attr.s
class ConstantPulse(Pulse):
    """Rectangular pulse."""


@attr.s
class SmoothedConstantPulse(Pulse):
    """Constant pulse with smoothed ring-up/-down of duration ``sigma``."""

    VALID_SHAPES = ["tanh", "cos", "gaussian"]
    pulse_func = staticmethod(smoothed_constant_pulse)
    length = NanosecondParameter(100)
    sigma = NanosecondParameter(0)
    shape = StringParameter("tanh", validator=attr.validators.in_(VALID_SHAPES))


@attr.s
class GaussianPulse(Pulse):
    """Gaussian chopped at +/- ``chop/2 * sigma`` with optional DRAG quadrature."""

    pulse_func = staticmethod(gaussian_pulse)
    sigma = NanosecondParameter(10)
    chop = IntParameter(4, unit="sigma")
    drag = FloatParameter(0)


@attr.s
class SechPulse(Pulse):
    """Hyperbolic-secant pulse chopped at +/- ``chop/2 * sigma``."""

    pulse_func = staticmethod(sech_pulse)
    sigma = NanosecondParameter(10)
    chop = IntParameter(4, unit="sigma")
    drag = FloatParameter(0)


@attr.s
class SlepianPulse(Pulse):
    """Slepian (DPSS) pulse of length ``tau``."""

    pulse_func = staticmethod(slepian_pulse)
    tau = NanosecondParameter(40)
    width = NanosecondParameter(10)
    drag = FloatParameter(0)


def pulse_factory(cls, name=None, **kwargs):
    """Zero-argument constructor for ``cls``; default name is the snake_case class name."""
    if name is None:
        name = "_".join(re.findall("[a-zA-Z][^A-Z]*", cls.__name__)).lower()
    return lambda: cls(name=name, **kwargs)
