"""sequencing_b200: B200-native time-evolution engine behind seQuencing's pulse-sequence API.

Drop-in surface of ``sequencing`` (reference ``sequencing/__init__.py``) for the hot path
``CompiledPulseSequence.run() -> qutip.mesolve``: same ``System`` / ``Mode`` / ``PulseSequence``
/ ``Sequence`` names, with the integration done by hand-written sm_100a CUDA kernels through the
C ABI in ``include/seq_engine.h``.  Mode ordering convention: ``|m_n, ..., m_1, m_0>``, larger
Hilbert spaces to the right (reference ``sequencing/__init__.py:9-22``).
"""
from . import qobj
from .modes import Cavity, Mode, PulseMode, Qubit, Transmon, sort_modes
from .system import CouplingTerm, System
from .sequencing import (
    CTerm,
    HTerm,
    Operation,
    PulseSequence,
    Sequence,
    SequenceResult,
    capture_operation,
    delay,
    delay_channels,
    get_sequence,
    ket2dm,
    ops2dms,
    sync,
    tqdm,
)
from .solver import Options, deferred_solves

__version__ = "0.1.0"
