"""Sequence containers: ``PulseSequence``, ``Sequence`` and ``SequenceResult``.

Callers of the hot path (reference ``sequencing/sequencing/main.py``): ``PulseSequence.run``
compiles and runs one pulse block (``:67-136``); ``Sequence.run`` walks a list of pulse blocks,
bare Operations and instantaneous unitaries, handing the final state of each segment to the
next (``:250-340``) - segments are sequentially dependent, so batching is across independent
sequences, not within one.
"""
import copy

import numpy as np

from .. import qobj as qt
from .compiled import CompiledPulseSequence
from .ops import (
    DelayChannelsOperation,
    DelayOperation,
    Operation,
    SyncOperation,
    ValidatedList,
    delay_channels,
    get_sequence,
)

try:  # progress bars are cosmetic
    from tqdm import tqdm
except Exception:  # pragma: no cover
    def tqdm(iterable, **kwargs):
        return iterable


class PulseSequence(ValidatedList):
    """List of Operation / Sync / Delay / DelayChannels items, compiled lazily."""

    VALID_TYPES = (Operation, SyncOperation, DelayOperation, DelayChannelsOperation)

    def __init__(self, system=None, t0=0, items=None):
        self.system = None
        self.t0 = None
        super().__init__(items)
        self.set_system(system=system, t0=t0)

    def set_system(self, system=None, t0=0):
        self.system = system
        self.t0 = t0
        self.clear()

    def compile(self):
        """Replay the items into a fresh ``CompiledPulseSequence`` starting at ``self.t0``."""
        seq = CompiledPulseSequence(system=self.system, t0=self.t0)
        for item in self:
            item = self._validate(item)
            if isinstance(item, Operation):
                seq.add_operation(item)
            elif isinstance(item, SyncOperation):
                seq.sync()
            elif isinstance(item, DelayOperation):
                seq.delay(item.length, sync_before=item.sync_before, sync_after=item.sync_after)
            else:
                delay_channels(item.channels, item.length, seq=seq)
        return seq

    @property
    def times(self):
        return self.compile().times

    @property
    def channels(self):
        return self.compile().channels

    def run(self, init_state, c_ops=None, e_ops=None, options=None, only_final_state=False, progress_bar=None):
        return self.compile().run(init_state, c_ops=c_ops, e_ops=e_ops, options=options,
                                  only_final_state=only_final_state, progress_bar=progress_bar)

    def propagator(self, c_ops=None, options=None, unitary_mode="batch", parallel=False, progress_bar=None):
        if all(isinstance(op, SyncOperation) for op in self):
            return [self.system.I()]
        return self.compile().propagator(c_ops=c_ops, options=options, unitary_mode=unitary_mode, parallel=parallel)

    def plot_coefficients(self, subplots=True, plot_imag=False, step=False):
        return self.compile().plot_coefficients(subplots=subplots, plot_imag=plot_imag, step=step)


class Sequence(ValidatedList):
    """PulseSequences, Operations and instantaneous unitaries applied in series."""

    VALID_TYPES = (qt.Qobj, Operation, PulseSequence)

    def __init__(self, system, operations=None):
        self.system = system
        self.pulse_sequence = get_sequence(system)
        self._t = 0
        super().__init__(operations)

    def _validate(self, item):
        item = super()._validate(item)
        if item is self.pulse_sequence:
            # appending the global sequence: keep a snapshot and start a new one
            item = copy.deepcopy(item)
            self.reset_pulse_sequence()
        elif len(self.pulse_sequence):
            self.capture()
        return item

    def reset_pulse_sequence(self):
        self.pulse_sequence = get_sequence(self.system)

    def capture(self):
        """Append the global pulse sequence if anything was captured into it."""
        if len(self.pulse_sequence):
            self.append(self.pulse_sequence)

    def _segment(self, item):
        """Compile one pulsed item (PulseSequence or Operation) at the current sequence time."""
        if isinstance(item, PulseSequence):
            if item.system != self.system:
                raise ValueError("All operations must have the same system.")
            item.t0 = self._t
            seq = item.compile()
        else:
            seq = CompiledPulseSequence(self.system, t0=self._t)
            seq.add_operation(item)
        seq.sync()
        return seq

    def run(self, init_state, e_ops=None, options=None, full_evolution=True, progress_bar=False):
        self.capture()
        wrap = tqdm if progress_bar else (lambda it, **kw: it)
        e_ops = e_ops or []
        self._t = 0
        times = [self._t]
        states = [init_state]
        for item in wrap(self):
            item = self._validate(item)
            if isinstance(item, qt.Qobj):
                state = states[-1]
                state = item * state if state.isket else item * state * item.dag()
                new_states, new_times = [state], [times[-1]]
            else:
                seq = self._segment(item)
                result = seq.run(states[-1], options=options, only_final_state=(not full_evolution))
                keep = slice(None) if full_evolution else slice(-1, None)
                new_states, new_times = result.states[keep], list(result.times[keep])
                self._t = seq._t
            states.extend(new_states)
            times.extend(new_times)
        times = np.array(times)
        order = np.argsort(times, kind="stable")
        times = times[order]
        states = [states[i] for i in order]
        expect = [qt.expect(op, states) for op in e_ops]
        return SequenceResult(times=times, states=states, expect=expect,
                              num_collapse=len(self.system.c_ops(clean=True)))

    def run_batch(self, init_states, e_ops=None, options=None):
        """``run(init_state, e_ops, full_evolution=False)`` for a whole batch of initial states with the states
        kept on the GPU between segments (replaces the per-state loop over ``main.py:281-335``).

        Pulsed segments are lowered once and integrated for all B states by one engine launch whose initial
        states are the previous segment's device-resident final states; instantaneous unitaries
        (``main.py:316-322``) are applied by ``seq_engine_apply_unitary`` and the expectation values at the
        segment boundaries (``main.py:333-335``) come from ``seq_engine_expect`` - nothing but the boundary
        states' expectation values and the final states crosses PCIe.  Returns one ``SequenceResult`` per state
        (times / states / expect at the segment boundaries, exactly what ``full_evolution=False`` keeps)."""
        init_states = list(init_states)
        return run_sequences([self] * len(init_states), init_states, e_ops=e_ops, options=options)

    def propagator(self, c_ops=None, options=None, unitary_mode="batch", parallel=False, progress_bar=False):
        self.capture()
        wrap = tqdm if progress_bar else (lambda it, **kw: it)
        c_ops = c_ops or []
        self._t = 0
        props = [self.system.I()]
        for item in wrap(self):
            item = self._validate(item)
            if isinstance(item, qt.Qobj):
                props.append(item * props[-1])
                continue
            seq = self._segment(item)
            step = seq.propagator(c_ops=c_ops, options=options, unitary_mode=unitary_mode, parallel=parallel)[-1]
            self._t = seq._t
            props.append(step * props[-1])
        return props

    def plot_coefficients(self, subplots=True, sharex=True, sharey=True):
        import matplotlib.pyplot as plt

        self._t = 0
        traces = {}
        marks = []
        for item in self:
            item = self._validate(item)
            if isinstance(item, qt.Qobj):
                marks.append(self._t)
                continue
            seq = self._segment(item)
            seg_times = seq.times
            self._t = seg_times.max()
            for name, info in seq.channels.items():
                if "coeffs" not in info:
                    continue
                ts, cs = traces.get(name, (np.zeros(0), np.zeros(0)))
                traces[name] = (np.concatenate([ts, seg_times]), np.concatenate([cs, info["coeffs"]]))
        names = [n for n in traces if n != "delay"]
        if not names:
            raise ValueError("There are no channels with coefficients to plot.")
        if subplots:
            fig, axes = plt.subplots(len(names), sharex=sharex, sharey=sharey)
            axes = np.atleast_1d(axes)
        else:
            fig, ax = plt.subplots(1, 1)
            axes = [ax] * len(names)
        for name, ax in zip(names, axes):
            ax.plot(*traces[name], label=name)
            ax.legend(loc=0)
            ax.grid(True)
        for ax in (axes if subplots else axes[:1]):
            for t in marks:
                ax.axvline(t, ls="--", lw=1.5, color="k", alpha=0.25)
        fig.suptitle("Hamiltonian coefficients")
        axes[-1].set_xlabel("Time")
        return (fig, axes) if subplots else (fig, axes[0])


def run_sequences(sequences, init_states, e_ops=None, options=None):
    """B ``Sequence`` objects of identical structure (the same kinds of items in the same order, pulsed segments
    of the same length - e.g. one qasm-built sequence under B amplitude / phase jitters, BASELINE config 4) run
    as ONE batch: trajectory b is ``sequences[b].run(init_states[b], e_ops, full_evolution=False)``
    (``main.py:250-340``).  Segment k of every sequence is lowered on the host (per-trajectory coefficient
    tables) and integrated for all B trajectories by one engine launch that starts from the previous segment's
    device-resident final states; instantaneous unitaries (``main.py:316-322``; one shared, or one per
    trajectory) are applied by ``seq_engine_apply_unitary``; boundary expectation values (``main.py:333-335``)
    by ``seq_engine_expect``.  States that start as kets turn into density matrices at the first segment that
    carries collapse terms (as ``qutip.mesolve`` would return them) and stay density matrices."""
    from .. import solver
    from ..solver.engine import Engine
    from .ops import ops2dms

    sequences = list(sequences)
    init_states = list(init_states)
    given = list(init_states)        # states[0] of every result is the initial state as it was passed in
    B = len(init_states)
    if B == 0:
        return []
    if len(sequences) != B:
        raise ValueError("one Sequence per initial state")
    distinct = list({id(sq): sq for sq in sequences}.values())
    if len(distinct) == 1:
        distinct[0].capture()       # (several sequences share the process-global PulseSequence: capture them as they are built)
    system = sequences[0].system
    e_ops = ops2dms(e_ops or [])
    as_ket = all(st.isket for st in init_states) and len(system.c_ops(clean=True)) == 0
    if not as_ket:
        init_states = [qt.ket2dm(st) if st.isket else st for st in init_states]
    eng = Engine.get()
    torch = eng.torch
    ket_dims, dm_dims = (init_states[0].dims if as_ket else None), (None if as_ket else init_states[0].dims)
    cur = torch.from_numpy(np.ascontiguousarray(np.stack([np.asarray(st.full()).reshape(-1) if as_ket else np.asarray(st.full())
                                                          for st in init_states]), dtype=np.complex128)).to(eng.device)
    N = cur.shape[1]
    n_items = len(sequences[0])
    if any(len(sq) != n_items for sq in distinct):
        raise ValueError("sequences of a batch must have the same number of items")
    for sq in distinct:
        sq._t = 0
    times = [0]
    snaps = [(cur.clone(), as_ket)]
    with system.frozen():      # one system state for the whole batch: static operators are built (and hashed) once
        cur, as_ket, ket_dims, dm_dims = _run_items(sequences, distinct, n_items, system, eng, cur, as_ket, ket_dims, dm_dims, N, B, options, times, snaps)
    n_snap = len(snaps)
    ex = np.zeros((n_snap, B, len(e_ops)), dtype=complex)
    if e_ops:
        te = torch.from_numpy(np.ascontiguousarray(np.stack([op.full() for op in e_ops]), dtype=np.complex128)).to(eng.device)
        for q, (st, k_) in enumerate(snaps):
            exq = torch.empty((B, len(e_ops)), dtype=torch.complex128, device=eng.device)
            eng._check(eng.lib.seq_engine_expect(eng.handle, te.data_ptr(), st.data_ptr(), exq.data_ptr(), N, B, len(e_ops), int(k_)))
            ex[q] = exq.cpu().numpy()
    host = [(st.cpu().numpy(), k_) for st, k_ in snaps]
    times = np.array(times)
    num_collapse = len(system.c_ops(clean=True))
    results = []
    for b in range(B):
        states = [given[b]]
        for q in range(1, n_snap):
            arr, k_ = host[q]
            states.append(qt.Qobj(arr[b].reshape(-1, 1), dims=ket_dims) if k_ else qt.Qobj(arr[b], dims=dm_dims))
        expect = []
        for m, op in enumerate(e_ops):
            row = ex[:, b, m]
            expect.append(np.array(row.real if op.isherm else row))
        results.append(SequenceResult(times=times, states=states, expect=expect, num_collapse=num_collapse))
    return results


def _run_items(sequences, distinct, n_items, system, eng, cur, as_ket, ket_dims, dm_dims, N, B, options, times, snaps):
    """Item k of every sequence of the batch, k = 0, 1, ...: see ``run_sequences``."""
    from .. import solver
    torch = eng.torch
    for k in range(n_items):
        items = {id(sq): sq._validate(sq[k]) for sq in distinct}
        first = items[id(sequences[0])]
        if isinstance(first, qt.Qobj):
            if any(not isinstance(it, qt.Qobj) for it in items.values()):
                raise ValueError(f"item {k}: unitary in one sequence, pulsed segment in another")
            if len(distinct) == 1:
                tu = torch.from_numpy(np.ascontiguousarray(first.full(), dtype=np.complex128)).to(eng.device)
                eng._check(eng.lib.seq_engine_apply_unitary(eng.handle, tu.data_ptr(), cur.data_ptr(), N, B, int(as_ket)))
            else:       # one unitary per trajectory
                for b, sq in enumerate(sequences):
                    tu = torch.from_numpy(np.ascontiguousarray(items[id(sq)].full(), dtype=np.complex128)).to(eng.device)
                    eng._check(eng.lib.seq_engine_apply_unitary(eng.handle, tu.data_ptr(), cur[b:b + 1].data_ptr(), N, 1, int(as_ket)))
            times.append(times[-1])
        else:
            # any state of the right kind: only the operators, tables and options of the lowered problem are used
            probe = qt.Qobj(np.eye(N)[:, [0]], dims=ket_dims) if as_ket else qt.Qobj(np.eye(N) / N, dims=dm_dims)
            with solver.deferred_solves(execute=False) as d:
                for sq in distinct:
                    seg = sq._segment(items[id(sq)])
                    seg.run(probe, options=options, only_final_state=True)
                    sq._t = seg._t
                    sq._tmax = float(seg.times.max())
            bps = d.batch_problems()
            if len(bps) != 1:
                raise ValueError(f"item {k}: the sequences' segments differ in operators, length or options")
            bp = bps[0]
            order = {id(sq): q for q, sq in enumerate(distinct)}
            pick = np.array([order[id(sq)] for sq in sequences])
            if bp.coeff is not None and bp.coeff.shape[0] == len(distinct) and len(distinct) > 1:
                bp.coeff = bp.coeff[pick]
            if bp.rate is not None and bp.rate.shape[0] == len(distinct) and len(distinct) > 1:
                bp.rate = bp.rate[pick]
            if as_ket and not bp.is_ket:
                # the segment carries collapse terms: qutip.mesolve evolves |psi><psi| and returns density matrices
                cur = cur.unsqueeze(2) * cur.conj().unsqueeze(1)
                as_ket = False
                dm_dims = [ket_dims[0], ket_dims[0]]
            elif not as_ket and bp.is_ket:
                raise RuntimeError("internal: density-matrix states reached a ket segment")
            bp.state0 = np.broadcast_to(bp.state0[:1], (B,) + bp.state0.shape[1:])
            tens = eng.to_device(bp)
            tens["state0"] = cur.contiguous()
            out = eng.solve_device(tens, bp)["out"]
            if bool((out["status"] != 0).any()):
                raise Exception(solver.ODE_ERROR + " (batched Sequence segment)")
            cur = out["final_state"]
            times.append(sequences[0]._tmax)
        snaps.append((cur.clone(), as_ket))
    return cur, as_ket, ket_dims, dm_dims


class SequenceResult(object):
    """``qutip.solver.Result``-like record for a ``Sequence``."""

    def __init__(self, times=None, states=None, expect=None, num_collapse=0):
        self.times = np.array([]) if times is None else times
        self.expect = [] if expect is None else expect
        self.states = [] if states is None else states
        self.num_collapse = num_collapse
        self.solver = "sequencing.Sequence"

    @property
    def num_expect(self):
        return len(self.expect)

    def __str__(self):
        lines = ["SequenceResult", "-" * 14]
        if self.times is not None and len(self.times) > 0:
            lines.append(f"Number of times: {len(self.times)}")
        if self.states is not None and len(self.states) > 0:
            lines.append("states = True")
        if self.expect is not None and len(self.expect) > 0:
            lines.append(f"expect = True, num_expect = {self.num_expect}")
        lines.append(f"num_collapse = {self.num_collapse}")
        return "\n".join(lines)

    __repr__ = __str__


_global_pulse_sequence = PulseSequence()
