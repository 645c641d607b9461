"""``CompiledPulseSequence``: the solve boundary (L2 of SURVEY.md section 1).

``run()`` assembles the problem exactly as the reference does
(``sequencing/sequencing/basic.py:445-512``: default options ``max_step = dt``,
``store_states = True``; ``H0 = sum(system.H0())`` appended as the *last* channel; the caller's
``c_ops`` list extended in place with ``system.c_ops()`` and the collapse channels; ``e_ops``
mapped through ``ops2dms``; ``only_final_state`` -> outputs at ``[t_min, t_max]`` with the
coefficients still defined on the full grid) and then calls ``solver.mesolve`` - the batched
CUDA engine - where the reference calls ``qutip.mesolve``.
"""
import numpy as np

from .. import qobj as qt
from .. import solver
from .channels import HamiltonianChannels


class CompiledPulseSequence(object):
    """Time-dependent Hamiltonian channels built from a sequence of Operations."""

    def __init__(self, system=None, channels=None, t0=0):
        self.system = None
        self.modes = None
        self.hc = None
        self._t = None
        if system is not None:
            self.set_system(system, channels=channels, t0=t0)

    def set_system(self, system, channels=None, t0=0):
        self.system = system
        self.modes = system.modes
        self.hc = HamiltonianChannels(channels=channels, t0=t0, dt=system.dt)
        self._t = self.hc.tmax

    @property
    def dt(self):
        return self.hc.dt

    @property
    def times(self):
        return self.hc.times

    @property
    def channels(self):
        return self.hc.channels

    def add_channel(self, name, H=None, C_op=None, time_dependent=False, error_if_exists=True):
        self.hc.add_channel(name, H=H, C_op=C_op, time_dependent=time_dependent, error_if_exists=error_if_exists)

    def add_operation(self, operation):
        """Add every term of an ``Operation`` at the current sequence time."""
        from .ops import HTerm, Operation

        if not isinstance(operation, Operation):
            raise TypeError(f"Expected an Operation, but got {type(operation)}.")
        duration, terms = operation
        for channel_name, term in terms.items():
            op, coeffs, args, kwargs = term
            which = {"H": op} if isinstance(term, HTerm) else {"C_op": op}
            self.hc.add_operation(channel_name, t0=self._t, duration=duration, coeffs=coeffs,
                                  coeffs_args=args or [], coeffs_kwargs=kwargs or {}, **which)

    def delay(self, length, sync_before=True, sync_after=True):
        """Global delay: a zero operator with a unit table on the ``delay`` channel."""
        if sync_before:
            self.sync()
        if not length:
            return
        self.hc.add_operation("delay", H=0 * self.system.active_modes[0].I, t0=self._t, duration=length, coeffs=1)
        if sync_after:
            self.sync()

    def sync(self):
        """All later operations start one sample after the current end of every channel."""
        self._t = self.hc.tmax + self.hc.dt

    def build_hamiltonian(self):
        return self.hc.build_hamiltonian()

    def _assemble(self, c_ops):
        if "H0" not in self.hc.channels:
            self.hc.add_channel("H0", H=self.system.H0_sum(), time_dependent=False)
        c_ops.extend(self.system.c_ops())
        H, C_ops, times = self.build_hamiltonian()
        c_ops.extend(C_ops)
        return H, times

    def run(self, init_state, c_ops=None, e_ops=None, options=None, only_final_state=False, progress_bar=None):
        """Integrate the sequence.  Returns a ``solver.Result`` (``qutip.solver.Result`` attributes).

        Inside ``solver.deferred_solves()`` the returned Result is filled when the block exits
        and the whole sweep runs as one batched launch.
        """
        from .ops import ops2dms

        c_ops = [] if c_ops is None else c_ops
        e_ops = [] if e_ops is None else e_ops
        if options is None:
            options = solver.Options()
            options.max_step = self.hc.dt
            options.store_states = True
            options.rhs_reuse = True
            solver.rhs_clear()
        H, times = self._assemble(c_ops)
        e_ops = ops2dms(e_ops)
        if only_final_state:
            tlist = [times.min(), times.max()]
            H = solver.QobjEvo(H, tlist=times, state0=init_state, e_ops=e_ops)
        else:
            tlist = times
        return solver.mesolve(H, init_state, tlist, c_ops, e_ops, options=options, progress_bar=progress_bar)

    def propagator(self, c_ops=None, options=None, unitary_mode="batch", parallel=False):
        """Propagators ``U(t)`` at every sample (a batched solve over basis states)."""
        c_ops = [] if c_ops is None else c_ops
        if options is None:
            options = solver.Options()
            options.max_step = self.hc.dt
            options.rhs_reuse = True
            solver.rhs_clear()
        H, times = self._assemble(c_ops)
        props = solver.propagator(H, times, c_op_list=c_ops, unitary_mode=unitary_mode, parallel=parallel, options=options)
        if isinstance(props, qt.Qobj):
            props = [props]
        return list(props)

    def plot_coefficients(self, subplots=True, plot_imag=False, step=False):
        return self.hc.plot_coefficients(subplots=subplots, plot_imag=plot_imag, step=step)
