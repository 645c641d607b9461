"""Pulse-sequence construction and the solve boundary (reference: ``sequencing/sequencing/``)."""
from .channels import HamiltonianChannels
from .compiled import CompiledPulseSequence
from .containers import PulseSequence, Sequence, SequenceResult, run_sequences, tqdm
from .ops import (
    CTerm,
    DelayChannelsOperation,
    DelayOperation,
    HTerm,
    Operation,
    SyncOperation,
    ValidatedList,
    capture_operation,
    delay,
    delay_channels,
    get_sequence,
    ket2dm,
    ops2dms,
    sync,
)
