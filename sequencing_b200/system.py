"""``System``: a set of coupled modes -> static drift ``H0`` and static collapse operators.

Mirrors the public surface of the reference's ``sequencing/system.py`` (``CouplingTerm :23-98``,
``System :102-512``).  On the hot path a ``System`` supplies ``sum(system.H0())`` (detunings +
self-Kerr + couplings, ``system.py:399-421``) and ``system.c_ops()`` (decay / excitation /
dephasing of every active mode, zero operators dropped, ``system.py:423-445``) to
``CompiledPulseSequence.run`` (``basic.py:486-492``); those become the engine's shared ``H0`` and
``L_j`` matrices.
"""
import json
import re
from collections import defaultdict
from contextlib import contextmanager
from functools import reduce

import attr
import numpy as np

from . import qobj as qt
from .modes import Mode, sort_modes
from .parameters import DictParameter, ListParameter, Parameterized


_EVAL_CACHE = None      # set inside ``System.frozen()``


class CouplingTerm(object):
    """``strength * prod_i mode_i.operator_expr(expr_i)`` (+ h.c. if ``add_hc``).

    ``strength`` is in units of 2 pi GHz (rad/ns).
    """

    def __init__(self, *terms, strength=1, add_hc=False):
        if len(terms) == 1:
            if not isinstance(terms[0], (list, tuple)):
                raise TypeError(f"Expected a list of (Mode, str), but got {type(terms)}.")
            terms = terms[0]
        elif len(terms) == 4 and all(isinstance(t, (Mode, str)) for t in terms):
            terms = [(terms[0], terms[1]), (terms[2], terms[3])]  # legacy two-mode syntax
        for mode, expr in terms:
            if not isinstance(mode, Mode):
                raise TypeError(f"Expected instance of Mode, but got {type(mode)}.")
            if not isinstance(expr, str):
                raise TypeError(f"Expected instance of str, but got {type(expr)}.")
        self.terms = list(terms)
        self.strength = float(strength)
        self.add_hc = bool(add_hc)

    @property
    def operators(self):
        return [mode.operator_expr(expr) for mode, expr in self.terms]

    def H(self, strength=None, add_hc=None):
        strength = self.strength if strength is None else strength
        add_hc = self.add_hc if add_hc is None else add_hc
        op = reduce(lambda left, right: left * right, self.operators)
        if add_hc:
            op = op + op.dag()
        return strength * op

    def __repr__(self):
        body = ", ".join(f"({mode.name}, '{expr}')" for mode, expr in self.terms)
        return f"{type(self).__name__}([{body}], strength={self.strength:.3e}, add_hc={self.add_hc})"


def _is_nn_pair(coupling, m1, m2):
    (a, ea), (b, eb) = coupling.terms[0], coupling.terms[1]
    return ea == "n" and eb == "n" and ((a is m1 and b is m2) or (a is m2 and b is m1))


@attr.s
class System(Parameterized):
    """A collection of modes with cross-Kerr and generic coupling terms."""

    modes = ListParameter()
    cross_kerrs = DictParameter()

    order_modes = True

    def initialize(self):
        super().initialize()
        if self.order_modes:
            self.modes = sort_modes(self.modes)
        self._dt = 1
        self.active_modes = self.modes
        self.coupling_terms = defaultdict(list)
        for key, chi in list(self.cross_kerrs.items()):
            # a System revived from a dict/JSON gets its cross-Kerr coupling terms back
            self.set_cross_kerr(*sorted(key), chi=chi)

    @property
    def dt(self):
        return self._dt

    @dt.setter
    def dt(self, dt):
        self._dt = dt
        for mode in self.modes:
            mode.dt = dt

    def __getattribute__(self, name):
        # ``system.qubit`` resolves to the mode named "qubit" (modes can change at any time)
        try:
            for mode in object.__getattribute__(self, "modes"):
                if mode.name == name:
                    return mode
        except AttributeError:
            pass
        return object.__getattribute__(self, name)

    def get_mode(self, mode):
        name = mode.name if isinstance(mode, Mode) else mode
        if name not in [m.name for m in self.modes]:
            raise ValueError(f"{name} is not a mode of {self.name}.")
        return getattr(self, name)

    @property
    def levels(self):
        return {mode.name: mode.levels for mode in self.modes}

    # -- active subspace -----------------------------------------------------------------
    @property
    def active_modes(self):
        for mode in self._active_modes:
            if mode not in self.modes:
                raise ValueError(
                    f"{mode.name} is not in {self.name}.modes; set {self.name}.active_modes "
                    f"to a subset of {self.name}.modes."
                )
        return self._active_modes

    @active_modes.setter
    def active_modes(self, modes):
        if isinstance(modes[0], str):
            modes = [getattr(self, m) for m in modes]
        if self.order_modes:
            modes = sort_modes(modes)
        if not all(mode in self.modes for mode in modes):
            raise ValueError("active_modes must be a subset of the system's modes.")
        self._active_modes = modes
        for mode in modes:
            mode.space = modes

    @contextmanager
    def use_modes(self, modes):
        if isinstance(modes, (str, Mode)):
            modes = [modes]
        previous = self.active_modes
        try:
            self.active_modes = modes
            yield
        finally:
            self.active_modes = previous

    # -- operators and states -------------------------------------------------------------
    @staticmethod
    def tensor(*args):
        return qt.tensor(*args)

    def I(self, modes=None):  # noqa: E741, E743
        if modes is None:
            modes = self.active_modes
        elif not all(mode in self.modes for mode in modes):
            raise ValueError("modes must be a subset of the system's modes.")
        return self.tensor(*[qt.qeye(mode.levels) for mode in modes])

    eye = I

    def _product_state(self, picker, args, kwargs):
        if args:
            if kwargs:
                raise ValueError("If positional arguments are provided, no keyword arguments are allowed.")
            if len(args) != len(self.active_modes):
                raise ValueError("The number of positional argument must match the number of active modes.")
            values = list(args)
        else:
            values = [kwargs.get(mode.name, 0) for mode in self.active_modes]
        return self.tensor(*[picker(mode, v) for mode, v in zip(self.active_modes, values)])

    def fock(self, *args, **kwargs):
        """Fock product state; per-mode occupations positionally or as ``mode_name=n``."""
        return self._product_state(lambda mode, n: qt.fock(mode.levels, n), args, kwargs)

    basis = fock

    def fock_dm(self, *args, **kwargs):
        return qt.ket2dm(self.fock(*args, **kwargs))

    def ground_state(self):
        return self.fock()

    def logical_basis(self, *args, **kwargs):
        return self._product_state(lambda mode, b: mode.logical_states(full_space=False)[b], args, kwargs)

    # -- couplings -----------------------------------------------------------------------
    def set_cross_kerr(self, mode1, mode2, chi=0):
        """Cross-Kerr ``2 pi chi n_1 n_2`` (chi in GHz); replaces an existing one for the pair."""
        mode1 = self.get_mode(mode1) if isinstance(mode1, str) else mode1
        mode2 = self.get_mode(mode2) if isinstance(mode2, str) else mode2
        if mode1 is mode2:
            raise ValueError("If mode1 is mode2, then it's not a cross-Kerr.")
        key = frozenset([mode1.name, mode2.name])
        if key in self.coupling_terms:
            self.coupling_terms[key] = [
                c for c in self.coupling_terms[key] if not _is_nn_pair(c, mode1, mode2)
            ]
        self.coupling_terms[key].append(
            CouplingTerm([(mode1, "n"), (mode2, "n")], strength=2 * np.pi * chi)
        )
        self.cross_kerrs[key] = chi

    # -- evaluation cache for loops over many sequences of one system state ------------------------
    @contextmanager
    def frozen(self):
        """Inside the block no parameter of any system or mode changes (``run_sequences`` lowers B sequences whose
        captured segments each carry a deep copy of the system, ``main.py:226-233``): ``H0()``, ``c_ops()`` and
        ``couplings()`` are evaluated once per system STATE - keyed by a fingerprint of every parameter, coupling term
        and tensor space, so that equal copies share one evaluation - and handed out again as the same ``Qobj`` objects."""
        global _EVAL_CACHE
        outer = _EVAL_CACHE
        _EVAL_CACHE = {"fp": {}, "vals": {}} if outer is None else outer
        try:
            yield self
        finally:
            _EVAL_CACHE = outer

    def _fingerprint(self):
        couplings = [(sorted(key), [([(m.name, e) for m, e in t.terms], t.strength.hex(), t.add_hc) for t in terms])
                     for key, terms in self.coupling_terms.items()]
        spaces = [(m.name, m.levels, [(x.name, x.levels) for x in m.space]) for m in self.modes]
        return repr((self.as_dict(), couplings, [m.name for m in self.active_modes], spaces, self._dt))

    def _cached(self, what, modes, clean, build):
        cache = _EVAL_CACHE
        if cache is None:
            return build()
        fp = cache["fp"].get(id(self))
        if fp is None:
            fp = (self, self._fingerprint())      # (the reference keeps the id alive)
            cache["fp"][id(self)] = fp
        key = (fp[1], what, tuple(mode.name for mode in modes), bool(clean))
        if key not in cache["vals"]:
            cache["vals"][key] = build()
        return list(cache["vals"][key])

    def couplings(self, modes=None, clean=True):
        modes = self.active_modes if modes is None else modes
        return self._cached("couplings", modes, clean, lambda: self._couplings(modes, clean))

    def _couplings(self, modes, clean):
        names = {mode.name for mode in modes}
        found = []
        with self.use_modes(modes):
            for key, terms in self.coupling_terms.items():
                if set(key) <= names:
                    found.extend(term.H() for term in terms)
        return [h for h in found if h.data.nnz] if clean else found

    def H0(self, modes=None, clean=True):
        """Static Hamiltonian as a list: detunings, self-Kerrs, then couplings."""
        modes = self.active_modes if modes is None else modes

        def build():
            parts = [m.detuning for m in modes] + [m.self_kerr for m in modes]
            parts += self.couplings(modes=modes, clean=clean)
            return [h for h in parts if h.data.nnz] if clean else parts
        return self._cached("H0", modes, clean, build)

    def H0_sum(self):
        """``sum(H0())`` as one operator (the static channel of a compiled sequence, ``basic.py:437-443``)."""
        def build():
            H0 = sum(self.H0())
            if isinstance(H0, int) and H0 == 0:
                H0 = 0 * self.I()
            return [H0]
        return self._cached("H0_sum", self.active_modes, True, build)[0]

    def c_ops(self, modes=None, clean=True):
        """Static collapse operators: all decays, then excitations, then dephasings."""
        modes = self.active_modes if modes is None else modes

        def build():
            ops = [m.decay for m in modes] + [m.excitation for m in modes] + [m.dephasing for m in modes]
            return [c for c in ops if c.data.nnz] if clean else ops
        return self._cached("c_ops", modes, clean, build)

    # -- (de)serialisation: cross_kerrs has frozenset keys ----------------------------------
    def as_dict(self, json_friendly=False):
        d = super().as_dict(json_friendly=json_friendly)
        kerrs = d.pop("cross_kerrs")
        if json_friendly:
            d["cross_kerrs"] = {"{" + ", ".join(key) + "}": val for key, val in kerrs.items()}
        else:
            d["cross_kerrs"] = dict(kerrs)
        return d

    @classmethod
    def from_json(cls, json_path=None, json_str=None):
        def decode(obj):
            for key in list(obj):
                if re.match(r"\{(.*?)\}", key):
                    obj[frozenset(key.strip("{}").split(", "))] = obj.pop(key)
            return obj

        if (json_path is None) == (json_str is None):
            raise ValueError("You must provide exactly one of json_path or json_str.")
        if json_str is not None:
            return cls.from_dict(json.loads(json_str, object_hook=decode))
        if not json_path.endswith(".json"):
            json_path += ".json"
        with open(json_path, "r") as f:
            return cls.from_dict(json.load(f, object_hook=decode))
