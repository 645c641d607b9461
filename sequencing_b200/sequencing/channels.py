"""Channel compiler: per-channel coefficient tables on one uniform sample grid.

Produces the inputs the engine streams - ``H = [[H_k, c_k[n_t]], ..., H0]``, collapse terms and
``times = arange(tmin, tmax + dt, dt)`` - with the reference's semantics
(``sequencing/sequencing/basic.py:13-248``): every channel is zero-padded left/right so all
tables share the grid (integer ``//`` arithmetic on sample offsets), operations on the same
channel that overlap in time are summed, per-channel ``delay`` offsets shift later operations,
and a collapse channel whose table is all zero is emitted as a constant operator.
"""
import numpy as np

from .. import qobj as qt


def _check_operator(op, what):
    # ``sum([])`` gives the integer 0, which is accepted as "no operator"
    if not isinstance(op, qt.Qobj) and not (isinstance(op, (int, float)) and op == 0):
        raise TypeError(f"Excpected instance of Qobj for {what}, but got {type(op)}.")


def _pad(values, left, right):
    return np.concatenate([np.zeros(left), values, np.zeros(right)])


class HamiltonianChannels(object):
    """Named Hamiltonian / collapse channels with time-dependent coefficients.

    ``channels`` maps name -> ``{"H", "time_dependent", "delay"[, "coeffs"]}`` and
    ``collapse_channels`` maps name -> ``{"op", "time_dependent", "delay"[, "coeffs"]}``.
    """

    def __init__(self, channels=None, collapse_channels=None, t0=0, dt=1):
        self.channels = {}
        self.collapse_channels = {}
        for name, info in (channels or {}).items():
            self.add_channel(name, H=info["H"], time_dependent=info.get("time_dependent", False))
        for name, info in (collapse_channels or {}).items():
            self.add_channel(name, C_op=info["op"], time_dependent=info.get("time_dependent", False))
        self.tmin = t0
        self.tmax = t0
        self.dt = dt

    @property
    def times(self):
        return np.arange(self.tmin, self.tmax + self.dt, self.dt).astype(float)

    def _all(self):
        yield from self.channels.items()
        yield from self.collapse_channels.items()

    def add_channel(self, name, H=None, C_op=None, time_dependent=False, error_if_exists=True):
        if name in self.channels and error_if_exists:
            raise ValueError(f"Channel {name} already exists!")
        if H is not None:
            _check_operator(H, "H")
            self.channels[name] = {"H": H, "time_dependent": time_dependent, "delay": 0}
        elif C_op is not None:
            _check_operator(C_op, "C_op")
            self.collapse_channels[name] = {"op": C_op, "time_dependent": time_dependent, "delay": 0}
        else:
            raise ValueError("Expected H or C_op.")

    def add_operation(self, channel_name, t0=None, duration=None, times=None, H=None, C_op=None,
                      coeffs=1, coeffs_args=None, coeffs_kwargs=None):
        """Place ``coeffs`` on ``channel_name`` over ``times`` (or ``t0 .. t0+duration``), growing
        and re-padding every other channel so that all tables stay aligned.

        ``coeffs`` may be a number, an array, or a callable ``f(times, *args, **kwargs)``; with
        ``reset_t0=True`` in ``coeffs_kwargs`` the callable receives ``times - times[0]``.
        """
        if H is not None and C_op is not None:
            raise ValueError("Expected only H or C_op.")
        if channel_name not in self.channels and channel_name not in self.collapse_channels:
            self.add_channel(channel_name, H=H, C_op=C_op, time_dependent=True)
        owner = self.channels if channel_name in self.channels else self.collapse_channels
        target = owner[channel_name]
        if not target["time_dependent"]:
            raise ValueError(
                f"Channel {channel_name} is time-independent, so you cannot add time-dependent coefficients."
            )
        if times is None:
            if t0 is None or duration is None:
                raise ValueError("You must either specify an array of times or t0 and duration.")
            start = t0 + int(target["delay"])
            times = np.arange(start, start + duration, self.dt)
        if isinstance(coeffs, (int, float, complex)):
            coeffs = coeffs * np.ones_like(times)
        elif callable(coeffs):
            kw = dict(coeffs_kwargs or {})
            shift = times[0] if kw.pop("reset_t0", False) else 0
            coeffs = coeffs(times - shift, *(coeffs_args or []), **kw)

        first, last = times.min(), times.max()
        # the new table, extended to the grid as it was before this operation
        new = _pad(coeffs,
                   int(max((first - self.tmin) // self.dt, 0)),
                   int(max((self.tmax - last) // self.dt, 0)))
        # how far the grid itself grows on either side
        grow_l = int(max((self.tmin - first) // self.dt, 0))
        grow_r = int(max((last - self.tmax) // self.dt, 0))
        self.tmin = min(first, self.tmin)
        self.tmax = max(last, self.tmax)
        for name, channel in self._all():
            if "coeffs" in channel:
                old = _pad(channel["coeffs"], grow_l, grow_r)
                # signals on one channel that are not separated by a sync() add up
                channel["coeffs"] = new + old if name == channel_name else old
            elif name == channel_name:
                channel["coeffs"] = new
        n_t = len(self.times)
        for name, channel in self._all():
            if "coeffs" in channel and len(channel["coeffs"]) != n_t:
                raise ValueError(
                    f"Channel {name}: Number of time points {n_t} does not match "
                    f'number of coefficients {len(channel["coeffs"])}.'
                )

    def delay_channels(self, channel_names, duration):
        if duration == 0:
            return
        if duration < 0:
            raise ValueError("Delay cannot be less than 0.")
        if isinstance(channel_names, str):
            channel_names = [channel_names]
        for name in channel_names:
            if name not in self.channels and name not in self.collapse_channels:
                raise ValueError(f"Channel {name} does not exist.")
        for name in channel_names:
            self.channels[name]["delay"] += duration

    def build_hamiltonian(self):
        """``(H, C_ops, times)`` in list format: ``[op, table]`` for time-dependent entries."""
        H = []
        for channel in self.channels.values():
            if "coeffs" in channel:
                H.append([channel["H"], channel["coeffs"]])
            elif isinstance(channel["H"], (list, tuple)):
                H.extend(channel["H"])
            else:
                H.append(channel["H"])
        C_ops = []
        for channel in self.collapse_channels.values():
            if "coeffs" in channel and np.any(channel["coeffs"]):
                C_ops.append([channel["op"], channel["coeffs"]])
            elif isinstance(channel["op"], (list, tuple)):
                C_ops.extend(channel["op"])
            else:
                C_ops.append(channel["op"])
        return H, C_ops, self.times

    def plot_coefficients(self, subplots=True, plot_imag=False, step=False):
        import matplotlib.pyplot as plt

        names = [n for n, ch in self.channels.items() if n != "delay" and "coeffs" in ch]
        xs = self.times
        if subplots:
            fig, axes = plt.subplots(len(names), 1, sharex=True)
            axes = np.atleast_1d(axes)
        else:
            fig, ax = plt.subplots(1, 1)
            axes = [ax] * len(names)
        for i, name in enumerate(names):
            ax = axes[i]
            ys = self.channels[name]["coeffs"]
            draw = ax.step if step else ax.plot
            draw(xs, np.real(ys), "-", label=name, color=f"C{i}")
            if plot_imag:
                draw(xs, np.imag(ys), "--", color=f"C{i}")
            ax.grid(True)
            ax.legend(loc=0)
        if subplots:
            axes[0].set_title("Hamiltonian coefficients")
            axes[-1].set_xlabel("Time")
            return fig, axes
        axes[0].set_xlabel("Time")
        axes[0].set_ylabel("Hamiltonian coefficient")
        return fig, axes[0]
