"""Calibration sweeps: the batch source of the hot path.

Same entry points, arguments and return values ``((fig, ax), old, new)`` as the reference's
``sequencing/calibration.py`` (``tune_rabi :65``, ``tune_repeated_pi_pulses :172``,
``tune_repeated_pio2_pulses :279``, ``tune_drag :396``, ``tune_displacement :506``).  The
reference runs each sweep point as its own ``seq.run()`` -> ``qutip.mesolve`` call inside a serial
Python loop (``:120-125``, ``:452-466``, ``:560-565``); here the loop bodies are unchanged but run
inside ``solver.deferred_solves()``, so all points of a sweep are integrated by ONE batched
engine launch.  The ``tune_repeated_*`` sweeps chain the final state of one point into the next
(``:225-230``, ``:332-346``) and therefore stay sequential.

Fits use scipy (``curve_fit`` / ``polyfit``) with the reference's models, initial guesses and
bounds, because ``lmfit`` is not available here; plotting is skipped when ``matplotlib`` is absent.
"""
from math import factorial

import numpy as np
from scipy.optimize import curve_fit

from .sequencing import get_sequence, ket2dm, ops2dms, sync, tqdm
from .solver import deferred_solves, synth


class _Fit:
    """Minimal stand-in for an lmfit ``ModelResult``: ``.params[name]`` and ``.best_fit``."""

    class _Value(float):
        @property
        def value(self):
            return float(self)

    def __init__(self, names, values, best_fit):
        self.params = {n: self._Value(v) for n, v in zip(names, values)}
        self.best_fit = best_fit


def fit_line(xs, ys):
    slope, intercept = np.polyfit(np.asarray(xs, dtype=float), np.asarray(ys, dtype=float), 1)
    return _Fit(["slope", "intercept"], [slope, intercept], slope * np.asarray(xs) + intercept)


def fit_sine(xs, ys):
    """``ofs + amp sin(2 pi f0 x + phi)`` with an FFT-derived initial guess for ``f0`` and ``phi``."""
    xs = np.asarray(xs, dtype=float)
    ys = np.asarray(ys, dtype=float)

    def sine(x, amp, f0, phi, ofs):
        return ofs + amp * np.sin(2 * np.pi * f0 * x + phi)

    n_fft = 1000
    freqs = np.fft.rfftfreq(n_fft, xs[1] - xs[0])
    spectrum = np.fft.rfft(ys - ys.mean(), n_fft)
    peak = int(np.argmax(np.abs(spectrum)))
    f0 = freqs[peak]
    phi = np.mod(np.angle(spectrum[peak]), 2 * np.pi)
    span = np.ptp(ys)
    eps = 1e-12
    lo = [0.0, freqs.min(), 0.0, ys.min() - eps]
    hi = [span + eps, 2 * freqs.max(), 2 * np.pi, ys.max() + eps]
    p0 = np.clip([span / 2, f0, phi, ys.mean()], lo, hi)
    best, best_cost = None, np.inf
    # the phase guess from a zero-padded FFT can sit on the wrong side of a wrap: try both
    for phase in (p0[2], np.mod(p0[2] + np.pi / 2, 2 * np.pi), np.mod(p0[2] - np.pi / 2, 2 * np.pi)):
        start = p0.copy()
        start[2] = phase
        try:
            popt, _ = curve_fit(sine, xs, ys, p0=start, bounds=(lo, hi), xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=20000)
        except RuntimeError:
            continue
        cost = float(np.sum((sine(xs, *popt) - ys) ** 2))
        if cost < best_cost:
            best, best_cost = popt, cost
    if best is None:
        raise RuntimeError("sine fit did not converge")
    return _Fit(["amp", "f0", "phi", "ofs"], best, sine(xs, *best))


def fit_displacement(xs, ys):
    """Poisson model ``ofs + amp nbar^n exp(-nbar) / n!`` with ``nbar = (xscale x)^2``."""
    xs = np.asarray(xs, dtype=float)
    ys = np.asarray(ys, dtype=float)

    def displacement(x, xscale, amp, ofs, n):
        nbar = (x * xscale) ** 2
        return ofs + amp * nbar ** n / factorial(int(n)) * np.exp(-nbar)

    if xs[-1] > xs[0]:
        amp, ofs = ys[0] - ys[-1], np.min(ys)
    else:
        amp, ofs = ys[-1] - ys[0], np.max(ys)
    # ``n`` enters through int(n): it is not differentiable, and its initial value 0 is the
    # model the sweeps use (vacuum population), so it is held there.
    popt, _ = curve_fit(lambda x, xscale, a, o: displacement(x, xscale, a, o, 0), xs, ys, p0=[1.0, amp, ofs],
                        xtol=1e-15, ftol=1e-15, gtol=1e-15, maxfev=20000)
    vals = [popt[0], popt[1], popt[2], 0.0]
    return _Fit(["xscale", "amp", "ofs", "n"], vals, displacement(xs, *vals))


def _axes(plot, ax):
    if not plot:
        return None, None, None
    try:
        import matplotlib.pyplot as plt
    except Exception:
        return None, None, None
    if ax is None:
        fig, ax = plt.subplots(1, 1)
    else:
        fig = plt.gcf()
    return plt, fig, ax


def _setup(system, init_state, e_ops, mode_name, pulse_name):
    init_state = ket2dm(init_state)
    mode = system.get_mode(mode_name)
    pulse_name = pulse_name or mode.default_pulse
    pulse = getattr(mode, pulse_name)
    e_ops = ops2dms([init_state] if e_ops is None else e_ops)
    return init_state, mode, pulse_name, pulse, e_ops


def _use_device_synthesis(flag, pulse):
    """``device_synthesis`` argument of the sweep drivers: None = automatic."""
    if flag is False or not synth.supported(pulse):
        return False
    if flag:
        return True
    from .solver.engine import Engine
    try:
        return hasattr(Engine.get(), "synth_gaussian")      # (tests route Engine.get() to a CPU oracle stand-in without it)
    except Exception:
        return False


def tune_rabi(system, init_state, e_ops=None, mode_name="qubit", pulse_name=None, amp_range=(-2, 2, 51),
              progbar=True, plot=True, ax=None, ylabel=None, update=True, verify=True, device_synthesis=None):
    """Amplitude-Rabi calibration of a qubit pulse; returns ``((fig, ax), old_amp, new_amp)``.

    ``device_synthesis`` (an addition to the reference's signature): ONE template sequence is lowered and the
    coefficient tables of all sweep points are generated on the GPU (``solver/synth.py``)
    instead of compiling one sequence per point in Python.  Default (``None``): on whenever the pulse is one the
    device synthesis covers (noise-free Gaussian / DRAG) and the engine is the CUDA one; ``False`` forces the
    reference's per-point host compilation."""
    init_state, qubit, pulse_name, pulse, e_ops = _setup(system, init_state, e_ops, mode_name, pulse_name)
    amps = np.linspace(*amp_range)
    prog = tqdm if progbar else (lambda x, **kw: x)
    if _use_device_synthesis(device_synthesis, pulse):
        def template():
            with deferred_solves(execute=False) as d:
                seq = get_sequence(system)
                qubit.rotate_x(np.pi, pulse_name=pulse_name)
                seq.run(init_state, e_ops=e_ops, only_final_state=False)
            return d
        out = synth.qubit_rotation_sweep(system, qubit, pulse_name, [[(np.pi, 0.0, amp, pulse.drag)] for amp in amps],
                                         init_state, e_ops, template)
        e_pop = [float(v) for v in np.real(out.expect[:, 0, -1])]
    else:
        results = []
        with deferred_solves(), system.frozen():      # (the loop varies pulse parameters only: H0 / c_ops are built once)
            for amp in prog(amps):
                seq = get_sequence(system)
                with qubit.pulse_scale(amp, pulse_name=pulse_name):
                    qubit.rotate_x(np.pi, pulse_name=pulse_name)
                results.append(seq.run(init_state, e_ops=e_ops, only_final_state=False))
        e_pop = [r.expect[0][-1] for r in results]

    fit = fit_sine(amps, e_pop)
    old_amp = pulse.amp
    new_amp = old_amp / (2 * fit.params["f0"])

    plt, fig, ax = _axes(plot, ax)
    if plt is not None:
        ax.plot(amps, e_pop, "o")
        ax.plot(amps, fit.best_fit, "-")
        ax.set_xlabel("Pulse scale")
        if ylabel:
            ax.set_ylabel(ylabel)
        ax.grid(True)
        ax.set_title(f"{pulse.name} Rabi")
    if update:
        pulse.amp = new_amp
        print(f"Updating {qubit.name} unit amp from {old_amp:.2e} to {new_amp:.2e}.", flush=True)
    if verify:
        tune_rabi(system, init_state, e_ops=e_ops, mode_name=mode_name, pulse_name=pulse_name,
                  amp_range=amp_range, progbar=progbar, plot=plot, ax=ax, update=False, verify=False,
                  device_synthesis=device_synthesis)
    return (fig, ax), old_amp, new_amp


def _repeated(system, init_state, e_ops, mode_name, pulse_name, max_num_pulses, progbar, angle, title,
              plot, ax, ylabel, update, verify, me):
    init_state, qubit, pulse_name, pulse, e_ops = _setup(system, init_state, e_ops, mode_name, pulse_name)
    counts = np.arange(max_num_pulses + 1)
    prog = tqdm if progbar else (lambda x, **kw: x)
    state = init_state
    e_pop = []

    def play(state, n):
        seq = get_sequence(system)
        for _ in range(n):
            qubit.rotate_x(angle, pulse_name=pulse_name)
        result = seq.run(state, e_ops=e_ops, only_final_state=False)
        return result, result.states[-1]

    if angle == np.pi:
        for _ in prog(counts):
            result, state = play(state, 1)
            e_pop.append(result.expect[0][-1])
    else:
        result, state = play(state, 1)
        e_pop.append(result.expect[0][-1])
        for _ in prog(counts[:-1]):
            result, state = play(state, 2)
            e_pop.append(result.expect[0][-1])
    e_pop = np.array(e_pop)
    fit = fit_sine(counts, e_pop)
    old_amp = pulse.amp
    new_amp = old_amp * 2 * fit.params["f0"]
    plt, fig, ax = _axes(plot, ax)
    if plt is not None:
        ax.plot(counts, e_pop, "o")
        ax.plot(counts, fit.best_fit, "-")
        ax.set_xlabel("Number of pulses")
        if ylabel:
            ax.set_ylabel(ylabel)
        ax.grid(True)
        ax.set_title(f"{pulse.name} {title}")
    if update:
        pulse.amp = new_amp
        print(f"Updating {qubit.name} unit amp from {old_amp:.2e} to {new_amp:.2e}.", flush=True)
    if verify:
        me(system, init_state, e_ops=e_ops, mode_name=mode_name, pulse_name=pulse_name,
           max_num_pulses=max_num_pulses, progbar=progbar, plot=plot, ax=ax, update=False, verify=False)
    return (fig, ax), old_amp, new_amp


def tune_repeated_pi_pulses(system, init_state, e_ops=None, mode_name="qubit", pulse_name=None, max_num_pulses=100,
                            progbar=True, plot=True, ax=None, ylabel=None, update=True, verify=True):
    """Amplitude calibration from a train of pi pulses (state chained point to point: sequential)."""
    return _repeated(system, init_state, e_ops, mode_name, pulse_name, max_num_pulses, progbar, np.pi,
                     "Repeated pi pulses", plot, ax, ylabel, update, verify, tune_repeated_pi_pulses)


def tune_repeated_pio2_pulses(system, init_state, e_ops=None, mode_name="qubit", pulse_name=None, max_num_pulses=100,
                              progbar=True, plot=True, ax=None, ylabel=None, update=True, verify=True):
    """Amplitude check from a train of pi/2 pulses; like the reference (calibration.py:373) it
    never writes the new amplitude back."""
    return _repeated(system, init_state, e_ops, mode_name, pulse_name, max_num_pulses, progbar, np.pi / 2,
                     "Repeated pi/2 pulses", plot, ax, ylabel, False, verify, tune_repeated_pio2_pulses)


def tune_drag(system, init_state, e_ops=None, mode_name="qubit", pulse_name=None, drag_range=(-5, 5, 21),
              progbar=True, plot=True, ax=None, ylabel=None, update=True, device_synthesis=None):
    """DRAG calibration: ``Rx(pi);Ry(pi/2)`` vs ``Ry(pi);Rx(pi/2)`` over a DRAG sweep; the crossing
    of the two fitted lines is the optimum.  Returns ``((fig, ax), old_drag, new_drag)``.
    ``device_synthesis``: see ``tune_rabi``."""
    init_state, qubit, pulse_name, pulse, e_ops = _setup(system, init_state, e_ops, mode_name, pulse_name)
    old_drag = pulse.drag
    drags = np.linspace(*drag_range)
    prog = tqdm if progbar else (lambda x, **kw: x)
    if _use_device_synthesis(device_synthesis, pulse):
        def template():
            with deferred_solves(execute=False) as d:
                seq = get_sequence(system)
                qubit.rotate_x(np.pi, pulse_name=pulse_name)
                sync()
                qubit.rotate_y(np.pi / 2, pulse_name=pulse_name)
                seq.run(init_state, e_ops=e_ops, only_final_state=False)
            return d
        hp = np.pi / 2
        trajs = [[(np.pi, 0.0, 1.0, drag), (hp, hp, 1.0, drag)] for drag in drags] + \
                [[(np.pi, hp, 1.0, drag), (hp, 0.0, 1.0, drag)] for drag in drags]
        out = synth.qubit_rotation_sweep(system, qubit, pulse_name, trajs, init_state, e_ops, template)
        pops = np.real(out.expect[:, 0, -1])
        XpY9, YpX9 = [float(v) for v in pops[:len(drags)]], [float(v) for v in pops[len(drags):]]
    else:
        res_xy, res_yx = [], []
        with deferred_solves(), system.frozen():      # (the loop varies pulse parameters only: H0 / c_ops are built once)
            for drag in prog(drags):
                pulse.drag = drag
                seq = get_sequence(system)
                qubit.rotate_x(np.pi, pulse_name=pulse_name)
                sync()
                qubit.rotate_y(np.pi / 2, pulse_name=pulse_name)
                res_xy.append(seq.run(init_state, e_ops=e_ops, only_final_state=False))

                seq = get_sequence(system)
                qubit.rotate_y(np.pi, pulse_name=pulse_name)
                sync()
                qubit.rotate_x(np.pi / 2, pulse_name=pulse_name)
                res_yx.append(seq.run(init_state, e_ops=e_ops, only_final_state=False))
        XpY9 = [r.expect[0][-1] for r in res_xy]
        YpX9 = [r.expect[0][-1] for r in res_yx]

    r0, r1 = fit_line(drags, XpY9), fit_line(drags, YpX9)
    b0, b1 = r0.params["intercept"].value, r1.params["intercept"].value
    m0, m1 = r0.params["slope"].value, r1.params["slope"].value
    xopt = (b1 - b0) / (m0 - m1)

    plt, fig, ax = _axes(plot, ax)
    if plt is not None:
        ax.plot(drags, XpY9, "o", label="Rx(pi) - Ry(pi/2)")
        ax.plot(drags, r0.best_fit, "k-")
        ax.plot(drags, YpX9, "o", label="Ry(pi) - Rx(pi/2)")
        ax.plot(drags, r1.best_fit, "k-")
        ax.axvline(xopt, color="k", ls="--", label=f"DRAG: {xopt:.2e}")
        ax.legend(loc=0)
        ax.grid(True)
        ax.set_xlabel("DRAG coefficient")
        if ylabel:
            ax.set_ylabel("|e> population")
        ax.set_title(f"{pulse.name} DRAG")
    if update:
        pulse.drag = xopt
        print(f"Updating {pulse.name}.drag from {old_drag:.2e} to {xopt:.2e}.", flush=True)
    else:
        pulse.drag = old_drag
    return (fig, ax), old_drag, xopt


def tune_displacement(system, init_state, e_ops=None, mode_name="cavity", pulse_name=None, amp_range=(0.1, 3, 51),
                      progbar=True, plot=True, ax=None, ylabel=None, update=True, verify=True):
    """Displacement-amplitude calibration of a cavity pulse; returns ``((fig, ax), old_amp, new_amp)``."""
    init_state, cavity, pname, pulse, e_ops = _setup(system, init_state, e_ops, mode_name, pulse_name)
    amps = np.linspace(*amp_range)
    prog = tqdm if progbar else (lambda x, **kw: x)
    results = []
    with deferred_solves(), system.frozen():      # (the loop varies pulse parameters only: H0 / c_ops are built once)
        for amp in prog(amps):
            seq = get_sequence(system)
            with cavity.pulse_scale(amp):
                cavity.displace(1)
            results.append(seq.run(init_state, e_ops=e_ops, only_final_state=False))
    zero_pop = [r.expect[0][-1] for r in results]

    fit = fit_displacement(amps, zero_pop)
    old_amp = pulse.amp
    new_amp = old_amp / fit.params["xscale"].value

    plt, fig, ax = _axes(plot, ax)
    if plt is not None:
        ax.plot(amps, zero_pop, "o")
        ax.plot(amps, fit.best_fit, "-")
        ax.set_xlabel("Pulse scale")
        if ylabel:
            ax.set_ylabel(ylabel)
        ax.grid(True)
        ax.set_title(f"{pulse.name} displacement")
    if update:
        pulse.amp = new_amp
        print(f"Updating {cavity.name} unit amp from {old_amp:.2e} to {new_amp:.2e}.", flush=True)
    if verify:
        tune_displacement(system, init_state, e_ops=e_ops, mode_name=mode_name, pulse_name=pulse_name,
                          amp_range=amp_range, progbar=progbar, plot=plot, ax=ax, update=False, verify=False)
    return (fig, ax), old_amp, new_amp
