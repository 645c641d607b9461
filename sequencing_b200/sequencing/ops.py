"""Sequence vocabulary: Hamiltonian/collapse terms, operations, and the global capture hook.

Public surface of the reference's ``sequencing/sequencing/common.py``: ``HTerm/CTerm/Operation``
(``:109-158``), ``Sync/Delay/DelayChannels`` markers (``:161-209``), the process-global
``PulseSequence`` with ``@capture_operation`` (``:212-264``), and ``ket2dm/ops2dms``
(``:16-31``; ``ops2dms`` is applied to ``e_ops`` on the hot path, ``basic.py:496``).
Pure bookkeeping - nothing here touches the device.
"""
import inspect
from collections import namedtuple
from functools import wraps

from .. import qobj as qt


def ket2dm(obj):
    """Ket -> projector; density matrices pass through."""
    return qt.ket2dm(obj) if obj.isket else obj


def ops2dms(ops):
    """One Qobj or a sequence of them -> list of density matrices / operators."""
    if isinstance(ops, qt.Qobj):
        ops = [ops]
    return [ket2dm(op) for op in ops]


class ValidatedList(object):
    """List that only accepts instances of ``VALID_TYPES`` (``None`` = anything)."""

    VALID_TYPES = None

    def __init__(self, iterable=None):
        self._items = []
        if iterable is not None:
            self.extend(iterable)

    def _validate(self, item):
        allowed = self.VALID_TYPES
        if allowed is not None and not isinstance(item, allowed):
            names = ", ".join(t.__name__ for t in allowed)
            raise TypeError(f"{type(self).__name__} expected instance of [{names}] , but got {type(item)}.")
        return item

    def __repr__(self):
        return f"{type(self).__name__}([{','.join(map(repr, self))}])"

    def __len__(self):
        return len(self._items)

    def __iter__(self):
        return iter(list(self._items))

    def __getitem__(self, index):
        return self._items[index]

    def append(self, item):
        self._items.append(self._validate(item))

    def extend(self, iterable):
        self._items.extend([self._validate(item) for item in iterable])

    def insert(self, index, item):
        self._items.insert(index, self._validate(item))

    def pop(self, index=-1):
        return self._items.pop(index)

    def clear(self):
        self._items.clear()


HTerm = namedtuple("HTerm", ["H", "coeffs", "args", "kwargs"], defaults=[1, None, None])
HTerm.__doc__ = """Time-dependent Hamiltonian term ``coeffs(t) * H`` (coeffs: number, array or callable)."""

CTerm = namedtuple("CTerm", ["op", "coeffs", "args", "kwargs"], defaults=[1, None, None])
CTerm.__doc__ = """Time-dependent collapse operator ``coeffs(t) * op``."""

Operation = namedtuple("Operation", ["duration", "terms"])
Operation.__doc__ = """Terms applied simultaneously for ``duration`` samples: ``{channel_name: HTerm | CTerm}``."""


class SyncOperation(object):
    """Marker: everything after it starts once every channel before it has finished."""


class DelayOperation(object):
    """Marker: a global delay of ``length`` on all channels."""

    def __init__(self, length, sync_before=True, sync_after=True):
        self.length = length
        self.sync_before = sync_before
        self.sync_after = sync_after


class DelayChannelsOperation(object):
    """Marker: delay only ``channels`` (name, list of names, or ``{name: H}`` / ``{name: (H, C)}``)."""

    def __init__(self, channels, length):
        self.channels = channels
        self.length = length


def get_sequence(system=None):
    """The process-global ``PulseSequence``; passing a ``System`` resets it at ``t0 = 0``."""
    from .containers import _global_pulse_sequence

    if system is not None:
        _global_pulse_sequence.set_system(system, t0=0)
    return _global_pulse_sequence


def capture_operation(func):
    """Decorator: an ``Operation`` returned by ``func`` is appended to the global sequence.

    ``capture=False`` returns the ``Operation`` instead.  A ``t0=None`` parameter of ``func`` is
    filled with the global sequence's start time.
    """

    t0_param = inspect.signature(func).parameters.get("t0")
    fill_t0 = t0_param is not None and t0_param.default is None

    @wraps(func)
    def wrapper(*args, **kwargs):
        sequence = get_sequence()
        if sequence.system is not None:
            if fill_t0 and kwargs.get("t0") is None:
                kwargs["t0"] = sequence.t0
        capture = kwargs.pop("capture", True)
        produced = func(*args, **kwargs)
        if not (capture and isinstance(produced, Operation)):
            return produced
        if sequence.system is None:
            raise Exception("The global PulseSequence is not currently associated with a System.")
        sequence.append(produced)
        return None

    return wrapper


def sync(seq=None):
    """Align all channels here (on ``seq``, or on the global sequence when ``seq`` is None)."""
    if seq is None:
        get_sequence().append(SyncOperation())
    elif seq.system is not None:
        seq.sync()


def delay(length, sync_before=True, sync_after=True, seq=None):
    """Global delay of ``length`` samples."""
    if seq is None:
        get_sequence().append(DelayOperation(length))
    elif seq.system is not None:
        seq.delay(length, sync_before=sync_before, sync_after=sync_after)


def delay_channels(channels, length, seq=None):
    """Delay only the named channels by ``length``."""
    if seq is None:
        get_sequence().append(DelayChannelsOperation(channels, length))
        return
    if seq.system is None:
        return
    if isinstance(channels, str):
        channels = [channels]
    if isinstance(channels, dict):
        for name, spec in channels.items():
            H, C = spec if isinstance(spec, (list, tuple)) else (spec, None)
            if name not in seq.channels:
                seq.add_channel(name, H=H, C_op=C, time_dependent=True)
    elif not isinstance(channels, (list, tuple)):
        raise TypeError("Channels must be either a sequence of channel names or a dict like {name: operator}.")
    seq.hc.delay_channels(list(channels), length)
