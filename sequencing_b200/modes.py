"""Oscillator modes (``Mode``/``PulseMode``/``Qubit``/``Transmon``/``Cavity``).

Host-side model objects that *produce the operators and waveforms* the engine consumes:
full-space ladder/number/quadrature operators (reference ``sequencing/modes.py:125-207``,
embedding by ``tensor_with_I`` ``:292-305``), collapse operators ``decay/excitation/dephasing``
``:194-207``, and the pulsed operations ``Qubit.rotate`` ``:719-753`` / ``Cavity.displace``
``:864-897`` which return ``Operation(duration, {channel: HTerm(op, samples)})``.  The public
surface (names, arguments, units: ns / GHz / rad) is the reference's; operator construction is
tiny and happens once per sweep, so it stays numpy on the host.

Ordering convention (reference ``sequencing/__init__.py:9-22``): modes are sorted descending by
``(name, levels)`` so the state is ``|m_n, ..., m_1, m_0>``.
"""
from contextlib import contextmanager

import attr
import numpy as np

from . import qobj as qt
from .parameters import (
    DictParameter,
    FloatParameter,
    GigahertzParameter,
    IntParameter,
    NanosecondParameter,
    Parameterized,
    StringParameter,
)
from .pulses import GaussianPulse, SmoothedConstantPulse, gaussian_chop_norm, pulse_factory
from .sequencing.ops import HTerm, Operation, capture_operation


def sort_modes(modes):
    """Descending ``(name, levels)`` order, e.g. ``[qubit1, qubit0, cavity1, cavity0]``."""
    return sorted(modes, key=lambda m: (m.name, m.levels), reverse=True)


_EMBED_CACHE = {}


@attr.s
class Mode(Parameterized):
    """One oscillator, embeddable in the product space ``self.space``."""

    levels = IntParameter(2)
    t1 = NanosecondParameter(np.inf, base=FloatParameter)
    t2 = NanosecondParameter(np.inf, base=FloatParameter)
    thermal_population = FloatParameter(0)
    df = GigahertzParameter(0)
    kerr = GigahertzParameter(0)

    order_modes = True

    OPERATORS = ("I", "a", "ad", "n", "x", "y", "sigmax", "sigmay", "sigmaz",
                 "Raxis", "Rx", "Ry", "Rz", "Rphi")

    def initialize(self):
        super().initialize()
        self._space = [self]
        self._logical_zero = None
        self._logical_one = None

    # -- Hilbert-space bookkeeping -------------------------------------------------------
    @property
    def space(self):
        return self._space

    @space.setter
    def space(self, modes):
        self._space = sort_modes(modes) if self.order_modes else modes

    @property
    def index(self):
        return self.space.index(self)

    @staticmethod
    def tensor(*args):
        return qt.tensor(*args)

    def tensor_with_I(self, operator):
        """``operator`` on this mode, identity on every other mode of ``self.space``."""
        factors = [qt.qeye(m.levels) for m in self.space]
        factors[self.index] = operator
        return self.tensor(*factors)

    def _embedded(self, name, build):
        """Kronecker-embedded elementary operator, memoised on (operator, levels, position, space
        dimensions).  The reference rebuilds these on every property access (modes.py:292-305); in a
        sweep that is the dominant host cost, and the result only depends on the dimensions."""
        key = (name, self.levels, self.index, tuple(m.levels for m in self.space))
        hit = _EMBED_CACHE.get(key)
        if hit is None:
            hit = self.tensor_with_I(build(self.levels))
            hit._a.flags.writeable = False
            if len(_EMBED_CACHE) > 512:
                _EMBED_CACHE.clear()
            _EMBED_CACHE[key] = hit
        return hit

    def tensor_with_zero(self, state):
        """``state`` on this mode, vacuum on every other mode of ``self.space``."""
        factors = [qt.basis(m.levels, 0) for m in self.space]
        factors[self.index] = state
        return self.tensor(*factors)

    @contextmanager
    def use_space(self, modes):
        if isinstance(modes, Mode):
            modes = [modes]
        previous = self.space
        try:
            self.space = modes
            yield
        finally:
            self.space = previous

    # -- operators -----------------------------------------------------------------------
    @property
    def I(self):  # noqa: E741, E743
        return self._embedded("I", qt.qeye)

    @property
    def a(self):
        return self._embedded("a", qt.destroy)

    @property
    def ad(self):
        return self._embedded("ad", qt.create)

    @property
    def n(self):
        # (the embedded number operator, memoised: a^dag a as a dense N x N product on every access - it sits under
        # detuning, dephasing and every "n" coupling term - was the largest single host cost of lowering a segment)
        return self._embedded("n", lambda levels: qt.create(levels) * qt.destroy(levels))      # (a^dag a as the reference forms it: sqrt(k)^2, not k)

    @property
    def _adada(self):
        """``a^dag a^dag a a`` (self-Kerr), memoised like the elementary operators."""
        return self._embedded("adadaa", lambda levels: qt.create(levels) * qt.create(levels) * qt.destroy(levels) * qt.destroy(levels))

    @property
    def x(self):
        a = self.a
        return a.dag() + a

    @property
    def y(self):
        a = self.a
        return -1j * (a - a.dag())

    @property
    def detuning(self):
        return 2 * np.pi * self.df * self.n

    @property
    def self_kerr(self):
        return np.pi * self.kerr * self._adada

    # -- decoherence ---------------------------------------------------------------------
    @property
    def tphi(self):
        if np.isinf(self.t2):
            return np.inf
        if self.t2 > 2 * self.t1:
            raise ValueError("Cannot have T2 > 2 * T1.")
        rate = 1 / self.t2 - 1 / (2 * self.t1)
        return np.inf if rate == 0 else 1 / rate

    @property
    def Gamma_up(self):
        return self.thermal_population / self.t1

    @property
    def Gamma_down(self):
        return 1 / self.t1 - self.Gamma_up

    @property
    def decay(self):
        return np.sqrt(self.Gamma_down) * self.a

    @property
    def excitation(self):
        return np.sqrt(self.Gamma_up) * self.ad

    @property
    def dephasing(self):
        return np.sqrt(2 / self.tphi) * self.n

    @contextmanager
    def no_loss(self):
        with self.temporarily_set(t1=np.inf, t2=np.inf, thermal_population=0):
            yield

    # -- states --------------------------------------------------------------------------
    def fock(self, n=0, full_space=True):
        ket = qt.basis(self.levels, n)
        return self.tensor_with_zero(ket) if full_space else ket

    basis = fock

    def fock_dm(self, n=0, full_space=True):
        return qt.ket2dm(self.fock(n=n, full_space=full_space))

    def set_logical_states(self, logical_zero=None, logical_one=None):
        dims = self.basis(0, full_space=False).dims
        for label, state in (("zero", logical_zero), ("one", logical_one)):
            if state is not None and state.dims != dims:
                raise ValueError(f"Logical {label} state must have the same dimension as the Mode.")
        self._logical_zero = logical_zero
        self._logical_one = logical_one

    def _logical(self, stored, default_level, full_space):
        state = stored if stored is not None else self.basis(default_level, full_space=False)
        return self.tensor_with_zero(state) if full_space else state

    def logical_zero(self, full_space=True):
        return self._logical(self._logical_zero, 0, full_space)

    def logical_one(self, full_space=True):
        return self._logical(self._logical_one, 1, full_space)

    def logical_states(self, full_space=True):
        return self.logical_zero(full_space=full_space), self.logical_one(full_space=full_space)

    # -- logical-subspace Paulis and rotations ---------------------------------------------
    def _pauli(self, builder, full_space):
        zero, one = self.logical_states(full_space=False)
        op = builder(zero, one)
        return self.tensor_with_I(op) if full_space else op

    def sigmax(self, full_space=True):
        return self._pauli(lambda z, o: z * o.dag() + o * z.dag(), full_space)

    def sigmay(self, full_space=True):
        return self._pauli(lambda z, o: -1j * (z * o.dag() - o * z.dag()), full_space)

    def sigmaz(self, full_space=True):
        return self._pauli(lambda z, o: z * z.dag() - o * o.dag(), full_space)

    def Raxis(self, theta, phi, full_space=True):
        axis = np.cos(phi) * self.sigmax(full_space=full_space) + np.sin(phi) * self.sigmay(full_space=full_space)
        return (-1j * theta / 2 * axis).expm()

    def Rx(self, theta, full_space=True):
        return (-1j * theta / 2 * self.sigmax(full_space=full_space)).expm()

    def Ry(self, theta, full_space=True):
        return (-1j * theta / 2 * self.sigmay(full_space=full_space)).expm()

    def Rz(self, theta, full_space=True):
        return (-1j * theta / 2 * self.sigmaz(full_space=full_space)).expm()

    def hadamard(self, full_space=True):
        return 1 / np.sqrt(2) * (self.sigmax(full_space=full_space) + self.sigmaz(full_space=full_space))

    def Rphi(self, phi, full_space=True):
        if full_space:
            number = self.n
        else:
            with self.use_space(self):
                number = self.n
        return (1j * phi * number).expm()

    def operator_expr(self, expr):
        """Evaluate an algebraic expression over this mode's operators, e.g. ``"a * ad"``.

        The reference delegates to ``asteval`` (modes.py:566-581); that package is absent here, so
        the expression is evaluated with Python's ``eval`` over a namespace containing only the
        operators in ``OPERATORS`` and numpy (no builtins).
        """
        # only the operators the expression names are built (each is an N x N Kronecker embedding; building all nine for
        # every coupling term of every System.H0 access was half of the host cost of lowering a Sequence segment)
        import re as _re
        wanted = set(_re.findall(r"[A-Za-z_][A-Za-z0-9_]*", expr))
        namespace = {name: getattr(self, name) for name in self.OPERATORS if name in wanted}
        namespace.update(np=np, pi=np.pi, sqrt=np.sqrt, exp=np.exp, sin=np.sin, cos=np.cos)
        return eval(expr, {"__builtins__": {}}, namespace)  # noqa: S307


@attr.s
class PulseMode(Mode):
    """A Mode that owns named ``Pulse`` objects."""

    pulses = DictParameter()
    default_pulse = StringParameter("gaussian_pulse")

    def initialize(self):
        super().initialize()
        self.add_pulse(cls=SmoothedConstantPulse)
        self.add_pulse(cls=GaussianPulse)
        self._dt = 1

    @property
    def dt(self):
        return self._dt

    @dt.setter
    def dt(self, dt):
        self._dt = dt
        for pulse in self.pulses.values():
            pulse.dt = dt

    def add_pulse(self, cls=GaussianPulse, name=None, error_if_exists=False, **kwargs):
        pulse = pulse_factory(cls, name=name, **kwargs)()
        if pulse.name in self.pulses and error_if_exists:
            raise ValueError(f"Pulse {pulse.name} already defined on {self.name}.")
        self.pulses[pulse.name] = pulse
        setattr(self, pulse.name, pulse)

    @contextmanager
    def amplitude(self, amp, pulse_name=None):
        pulse = self.pulses[pulse_name or self.default_pulse]
        saved = pulse.amp
        try:
            pulse.amp = amp
            yield
        finally:
            pulse.amp = saved

    @contextmanager
    def pulse_scale(self, scale, pulse_name=None):
        pulse = self.pulses[pulse_name or self.default_pulse]
        saved = pulse.amp
        try:
            pulse.amp = saved * scale
            yield
        finally:
            pulse.amp = saved


@attr.s
class Qubit(PulseMode):
    """Fixed-frequency qubit whose primitive pulsed operation is a Bloch-sphere rotation."""

    @property
    def anharmonicity(self):
        return self.kerr

    @anharmonicity.setter
    def anharmonicity(self, alpha):
        self.kerr = alpha

    @capture_operation
    def rotate(self, angle, phase, pulse_name=None, unitary=False, **kwargs):
        """Rotation by ``angle`` about the equatorial axis at ``phase``.

        The pulse with ``amp == 1`` is defined to be a pi rotation: the drive is
        ``angle * amp / gaussian_chop_norm`` times the unit waveform, its real part driving the
        ``x`` quadrature channel and its imaginary part the ``y`` channel.
        """
        if unitary:
            return self.Raxis(angle, phase, full_space=kwargs.get("full_space", True))
        pulse = getattr(self, pulse_name or self.default_pulse)
        scale = angle * pulse.amp / gaussian_chop_norm(pulse.sigma, pulse.chop)
        wave = pulse(amp=scale, phase=phase, **kwargs)
        return Operation(len(wave) * self.dt, {
            f"{self.name}.x": HTerm(self.x, wave.real),
            f"{self.name}.y": HTerm(self.y, wave.imag),
        })

    def rotate_x(self, angle, unitary=False, **kwargs):
        if unitary:
            return self.Rx(angle, full_space=kwargs.get("full_space", True))
        return self.rotate(angle, 0, **kwargs)

    def rotate_y(self, angle, unitary=False, **kwargs):
        if unitary:
            return self.Ry(angle, full_space=kwargs.get("full_space", True))
        return self.rotate(angle, np.pi / 2, **kwargs)


@attr.s
class Transmon(Qubit):
    """Fixed-frequency transmon (a ``Qubit`` with a negative self-Kerr)."""


@attr.s
class Cavity(PulseMode):
    """Weakly nonlinear cavity whose primitive pulsed operation is a displacement."""

    OPERATORS = Mode.OPERATORS + ("D",)

    def D(self, alpha, full_space=True):
        op = qt.displace(self.levels, alpha)
        return self.tensor_with_I(op) if full_space else op

    @capture_operation
    def displace(self, alpha, unitary=False, pulse_name=None, **kwargs):
        """Displacement by ``alpha``; the pulse with ``amp == 1`` is a unit displacement."""
        if unitary:
            return self.D(alpha)
        pulse = getattr(self, pulse_name or self.default_pulse)
        scale = 2 * alpha * pulse.amp / gaussian_chop_norm(pulse.sigma, pulse.chop)
        wave = pulse(amp=scale, **kwargs)
        return Operation(len(wave) * self.dt, {
            f"{self.name}.x": HTerm(self.x, -wave.imag),
            f"{self.name}.y": HTerm(self.y, wave.real),
        })
