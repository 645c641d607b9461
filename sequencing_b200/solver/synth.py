"""On-device waveform synthesis for sweeps.

The reference builds every sweep point's coefficient tables sample by sample in Python
(``Qubit.rotate`` -> ``Pulse.__call__`` -> ``array_pulse``, ``modes.py:719-753`` / ``pulses.py:31-117,262-293``)
and the channel compiler pads and sums them (``basic.py:87-207``).  For a sweep over pulse parameters the
tables of all trajectories differ only in ``amp / drag / detune / phase``, so here ONE template sequence is
lowered through the ordinary host path (operators, channel order, time grid) and the ``B x n_ch x n_t`` tables
are generated on the GPU by ``seq_engine_synth_gaussian`` from one small descriptor per pulse: they never exist
on the host and nothing but the descriptors crosses PCIe.
"""
import copy

import numpy as np

from ..pulses import GaussianPulse, gaussian_chop_norm
from .engine import Engine


def _channel_index(bp, op):
    """Index of the Hamiltonian channel whose operator equals ``op`` (full-space matrix)."""
    m = np.asarray(op.full() if hasattr(op, "full") else op)
    for k in range(bp.Hk.shape[0]):
        if bp.Hk[k].shape == m.shape and np.array_equal(bp.Hk[k], m):
            return k
    raise LookupError("the template sequence has no channel for this operator")


def supported(pulse):
    """Device synthesis covers noise-free chopped Gaussian(-DRAG) pulses."""
    return isinstance(pulse, GaussianPulse) and not pulse.noise_sigma


def qubit_rotation_sweep(system, qubit, pulse_name, trajectories, init_state, e_ops, run_template):
    """Integrate ``len(trajectories)`` pulse sequences on one qubit mode in one engine launch.

    ``trajectories[b]`` is a list of rotations ``(angle, phase, pulse_scale, drag)`` played back to back (a
    ``sync()`` between pulses on the same mode is what the reference's sweeps do).  ``run_template()`` must build
    and run ONE sequence of the same shape inside ``deferred_solves(execute=False)`` and return the deferral; it
    provides operators, channel order, grid and solver options.  Returns the engine's ``BatchOutput``.
    """
    pulse = getattr(qubit, pulse_name or qubit.default_pulse)
    d = run_template()
    (bp,) = d.batch_problems()
    ch_x, ch_y = _channel_index(bp, qubit.x), _channel_index(bp, qubit.y)
    n = int(pulse.chop * pulse.sigma / pulse.dt)
    n_t = int(bp.coeff.shape[-1])
    norm = gaussian_chop_norm(pulse.sigma, pulse.chop)
    descs = []
    for b, rotations in enumerate(trajectories):
        if len(rotations) * n > n_t:
            raise ValueError("template sequence is shorter than the trajectory's pulses")
        for q, (angle, phase, scale, drag) in enumerate(rotations):
            descs.append(dict(traj=b, ch_re=ch_x, ch_im=ch_y, start=q * n, sign_re=1.0, sign_im=1.0,
                              sigma=float(pulse.sigma), chop=float(pulse.chop), dt=float(pulse.dt),
                              amp=float(angle * (pulse.amp * scale) / norm), drag=float(drag),
                              detune=float(pulse.detune), phase=float(phase)))
    B = len(trajectories)
    eng = Engine.get()
    coeff = eng.synth_gaussian(descs, B, int(bp.Hk.shape[0]), n_t)
    batch = copy.copy(bp)
    batch.state0 = np.ascontiguousarray(np.broadcast_to(bp.state0[:1], (B,) + bp.state0.shape[1:]))
    batch.coeff_scale = None
    return eng.solve(batch, coeff_device=coeff)
