"""Device staging + calls into the C ABI.  PyTorch is plumbing only (HBM buffers, streams)."""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _lib
from ._lib import EngineUnavailable, SolveDesc


@dataclass
class BatchProblem:
    """B trajectories sharing operators (host numpy arrays, complex128 / float64).

    One trajectory = one (sequence x initial state x sweep value): what one ``seq.run()`` of the
    reference's sweep loops integrates (``calibration.py:120-125``).
    """

    H0: np.ndarray                      # [N,N] or [B,N,N]
    Hk: np.ndarray                      # [n_ch,N,N]
    coeff: np.ndarray                   # [B,n_ch,n_t] or [1,n_ch,n_t] (shared)
    state0: np.ndarray                  # [B,N,N] or [B,N] (kets)
    t0: float
    dt: float
    out_idx: np.ndarray                 # [n_out] int32
    Lc: Optional[np.ndarray] = None     # [n_c,N,N]
    Ltd: Optional[np.ndarray] = None    # [n_ctd,N,N]
    rate: Optional[np.ndarray] = None   # [B,n_ctd,n_t] or [1,...]
    E: Optional[np.ndarray] = None      # [n_e,N,N]
    coeff_scale: Optional[np.ndarray] = None  # [B,n_ch]
    is_ket: bool = False
    normalize: bool = True
    store_states: bool = False
    expect_complex: bool = False
    atol: float = 1e-8
    rtol: float = 1e-6
    max_step: float = 0.0
    first_step: float = 0.0
    min_step: float = 0.0
    nsteps: int = 1000
    method: int = _lib.METHOD_AUTO
    order_switch: bool = True           # let the controller drop to the 4-stage RK4(3) pair on capped steps

    @property
    def N(self):
        return int(self.Hk.shape[-1]) if self.Hk is not None and self.Hk.size else int(self.H0.shape[-1])

    @property
    def B(self):
        return int(self.state0.shape[0])


@dataclass
class BatchOutput:
    expect: np.ndarray          # [B,n_e,n_out] float64 or complex128
    final_state: np.ndarray     # [B,N,N] / [B,N]
    states: Optional[np.ndarray]
    status: np.ndarray          # [B] int32
    stats: np.ndarray           # [B,4] int32
    kernel_ms: float = -1.0
    launches: int = 0
    method: str = "auto"
    variant: str = ""


def _with_batched_coeff(bp):
    """A shallow copy of ``bp`` whose (host) coefficient array has the per-trajectory shape - only the shape is
    used once the tables themselves come from a device tensor."""
    import copy

    out = copy.copy(bp)
    out.coeff = np.broadcast_to(bp.coeff[:1], (bp.B,) + bp.coeff.shape[1:])
    return out


def _c128(a, shape_tail):
    a = np.ascontiguousarray(a, dtype=np.complex128)
    return a.reshape((-1,) + tuple(shape_tail)) if a.size else a.reshape((0,) + tuple(shape_tail))


class Engine:
    """One C-ABI engine handle bound to one CUDA device (and torch's current stream on it)."""

    _instances = {}

    def __init__(self, device_index: int = 0):
        import torch

        if not torch.cuda.is_available():
            raise EngineUnavailable("no CUDA device: sequencing_b200 integrates on B200 only (no CPU fallback)")
        self.lib = _lib.load()
        self.torch = torch
        self.device = torch.device("cuda", device_index)
        with torch.cuda.device(self.device):
            self.stream = torch.cuda.current_stream(self.device)
            handle = C.c_void_p()
            rc = self.lib.seq_engine_create(device_index, C.c_void_p(self.stream.cuda_stream), C.byref(handle))
        if rc != 0:
            raise EngineUnavailable(f"seq_engine_create failed with code {rc}")
        self.handle = handle
        self.last_kernel_ms = -1.0
        self.last_launches = 0
        self.total_launches = 0
        self._pinned_cache = {}

    @classmethod
    def get(cls, device_index: Optional[int] = None) -> "Engine":
        import torch

        if not torch.cuda.is_available():
            raise EngineUnavailable("no CUDA device: sequencing_b200 integrates on B200 only (no CPU fallback)")
        if device_index is None:
            device_index = torch.cuda.current_device()
        if device_index not in cls._instances:
            cls._instances[device_index] = cls(device_index)
        return cls._instances[device_index]

    def _check(self, rc):
        if rc != 0:
            msg = self.lib.seq_engine_last_error(self.handle)
            raise RuntimeError(f"seq_engine error {rc}: {msg.decode() if msg else ''}")

    # -- descriptor assembly ---------------------------------------------------------------
    def _desc(self, bp: BatchProblem, ptr, host: bool) -> SolveDesc:
        N, B = bp.N, bp.B
        d = SolveDesc()
        d.size = C.sizeof(SolveDesc)
        flags = 0
        if bp.is_ket:
            flags |= _lib.FLAG_KET
            if bp.normalize:
                flags |= _lib.FLAG_NORMALIZE
        if bp.store_states:
            flags |= _lib.FLAG_STORE_STATES
        if bp.expect_complex:
            flags |= _lib.FLAG_EXPECT_COMPLEX
        if not bp.order_switch:
            flags |= _lib.FLAG_NO_ORDER_SWITCH
        if host:
            flags |= _lib.FLAG_HOST_POINTERS
        d.flags = flags
        d.N, d.B = N, B
        d.n_t = int(bp.coeff.shape[-1]) if bp.coeff is not None and bp.coeff.size else int(self._n_t(bp))
        d.n_ch = int(bp.Hk.shape[0]) if bp.Hk is not None else 0
        d.n_c = int(bp.Lc.shape[0]) if bp.Lc is not None else 0
        d.n_ctd = int(bp.Ltd.shape[0]) if bp.Ltd is not None else 0
        d.n_e = int(bp.E.shape[0]) if bp.E is not None else 0
        d.n_out = int(len(bp.out_idx))
        d.nsteps = int(bp.nsteps)
        d.method = int(bp.method)
        d.t0, d.dt = float(bp.t0), float(bp.dt)
        d.atol, d.rtol = float(bp.atol), float(bp.rtol)
        d.max_step, d.first_step, d.min_step = float(bp.max_step), float(bp.first_step), float(bp.min_step)
        d.H0 = ptr["H0"]
        d.H0_batch_stride = N * N if (bp.H0.ndim == 3 and bp.H0.shape[0] == B and B > 1) else 0
        d.Hk = ptr.get("Hk")
        d.coeff = ptr.get("coeff")
        d.coeff_batch_stride = d.n_ch * d.n_t if (d.n_ch and bp.coeff.shape[0] == B and B > 1) else 0
        d.coeff_scale = ptr.get("coeff_scale")
        d.Lc = ptr.get("Lc")
        d.Ltd = ptr.get("Ltd")
        d.rate = ptr.get("rate")
        d.rate_batch_stride = d.n_ctd * d.n_t if (d.n_ctd and bp.rate.shape[0] == B and B > 1) else 0
        d.state0 = ptr["state0"]
        d.E = ptr.get("E")
        d.out_idx = ptr["out_idx"]
        d.expect = ptr.get("expect")
        d.final_state = ptr["final_state"]
        d.states = ptr.get("states")
        d.status = ptr["status"]
        d.stats = ptr["stats"]
        return d

    @staticmethod
    def _n_t(bp):
        if bp.rate is not None and bp.rate.size:
            return bp.rate.shape[-1]
        return int(np.max(bp.out_idx)) + 1

    def _host_arrays(self, bp: BatchProblem):
        N, B = bp.N, bp.B
        arrs = {"H0": _c128(bp.H0, (N, N))}
        if bp.Hk is not None and bp.Hk.shape[0]:
            arrs["Hk"] = _c128(bp.Hk, (N, N))
            arrs["coeff"] = np.ascontiguousarray(bp.coeff, dtype=np.float64)
            if bp.coeff_scale is not None:
                arrs["coeff_scale"] = np.ascontiguousarray(bp.coeff_scale, dtype=np.float64)
        if bp.Lc is not None and bp.Lc.shape[0]:
            arrs["Lc"] = _c128(bp.Lc, (N, N))
        if bp.Ltd is not None and bp.Ltd.shape[0]:
            arrs["Ltd"] = _c128(bp.Ltd, (N, N))
            arrs["rate"] = np.ascontiguousarray(bp.rate, dtype=np.float64)
        arrs["state0"] = np.ascontiguousarray(bp.state0, dtype=np.complex128)
        if bp.E is not None and bp.E.shape[0]:
            arrs["E"] = _c128(bp.E, (N, N))
        arrs["out_idx"] = np.ascontiguousarray(bp.out_idx, dtype=np.int32)
        return arrs

    def _out_shapes(self, bp: BatchProblem):
        N, B = bp.N, bp.B
        n_e = int(bp.E.shape[0]) if bp.E is not None else 0
        n_out = len(bp.out_idx)
        st = (N,) if bp.is_ket else (N, N)
        shapes = {
            "expect": ((B, n_e, n_out), np.complex128 if bp.expect_complex else np.float64),
            "final_state": ((B,) + st, np.complex128),
            "status": ((B,), np.int32),
            "stats": ((B, 4), np.int32),
        }
        if bp.store_states:
            shapes["states"] = ((B, n_out) + st, np.complex128)
        return shapes

    # -- solves ----------------------------------------------------------------------------
    def solve(self, bp: BatchProblem, coeff_device=None) -> BatchOutput:
        """Public host-buffer solve: numpy in, numpy out; H2D/D2H staging via pinned memory.
        ``coeff_device``: a device tensor f64 [B, n_ch, n_t] (e.g. from ``synth_gaussian``) used instead of
        ``bp.coeff`` - the tables then never exist on the host."""
        tens = self.to_device(bp)
        if coeff_device is not None:
            assert tuple(coeff_device.shape) == (bp.B, bp.Hk.shape[0], bp.coeff.shape[-1]), coeff_device.shape
            tens["coeff"] = coeff_device
            bp = _with_batched_coeff(bp)
        dev = self.solve_device(tens, bp)
        out = {k: v.cpu().numpy() for k, v in dev["out"].items()}
        return self._finish(bp, out)

    def to_device(self, bp: BatchProblem):
        torch = self.torch
        tens = {}
        for name, arr in self._host_arrays(bp).items():
            t = torch.from_numpy(arr)
            tens[name] = t.pin_memory().to(self.device, non_blocking=True) if t.numel() else t.to(self.device)
        return tens

    def solve_device(self, tens, bp: BatchProblem):
        """Inputs already resident in HBM (torch tensors); outputs stay on the device."""
        torch = self.torch
        with torch.cuda.device(self.device):
            outs = {}
            for name, (shape, dt) in self._out_shapes(bp).items():
                tdt = {np.float64: torch.float64, np.complex128: torch.complex128, np.int32: torch.int32}[dt]
                outs[name] = torch.empty(shape, dtype=tdt, device=self.device)
            ptr = {k: (v.data_ptr() if v.numel() else None) for k, v in tens.items()}
            ptr.update({k: (v.data_ptr() if v.numel() else None) for k, v in outs.items()})
            desc = self._desc(bp, ptr, host=False)
            self._check(self.lib.seq_engine_solve(self.handle, C.byref(desc)))
            self.last_launches = self.lib.seq_engine_last_launch_count(self.handle)
            self.total_launches += self.last_launches
            self._last_method = self.lib.seq_engine_last_method(self.handle)
        return {"out": outs, "in": tens}

    def _pinned_out(self, name, shape, dt, pinned):
        """Output buffer of a host-ABI solve.  ``pinned``: a page-locked buffer owned by the engine and re-used by
        the next call of the same shape (page-locking 66 MB per call would cost more than the copy it speeds up;
        D2H into pageable memory is staged by the driver and cannot overlap the kernels of the next chunk)."""
        if not pinned or int(np.prod(shape)) == 0:
            return np.empty(shape, dtype=dt)
        key = (name, tuple(shape), np.dtype(dt).str)
        buf = self._pinned_cache.get(key)
        if buf is None:
            tdt = {np.float64: self.torch.float64, np.complex128: self.torch.complex128, np.int32: self.torch.int32}[dt]
            buf = self.torch.empty(tuple(shape), dtype=tdt, pin_memory=True)
            if len(self._pinned_cache) > 16:
                self._pinned_cache.clear()
            self._pinned_cache[key] = buf
        return buf.numpy()

    def solve_host_abi(self, bp: BatchProblem, pinned_out: bool = False) -> BatchOutput:
        """The reference-facing C-ABI call on HOST buffers (SEQ_FLAG_HOST_POINTERS): the engine
        itself stages host->device, integrates, and copies results back before returning.  Large batches
        are pipelined in chunks (copies of one chunk behind the kernels of another).  ``pinned_out=True``
        returns views of page-locked buffers the engine re-uses on its next call - copy what must outlive it."""
        arrs = self._host_arrays(bp)
        outs = {name: self._pinned_out(name, shape, dt, pinned_out) for name, (shape, dt) in self._out_shapes(bp).items()}
        ptr = {k: (v.ctypes.data if v.size else None) for k, v in arrs.items()}
        ptr.update({k: (v.ctypes.data if v.size else None) for k, v in outs.items()})
        desc = self._desc(bp, ptr, host=True)
        self._check(self.lib.seq_engine_solve(self.handle, C.byref(desc)))
        self.last_launches = self.lib.seq_engine_last_launch_count(self.handle)
        self.total_launches += self.last_launches
        self._last_method = self.lib.seq_engine_last_method(self.handle)
        return self._finish(bp, outs)

    def init_native_comm(self, group=None):
        """NCCL communicator INSIDE the engine (``seq_engine_nccl_init``: libnccl opened at run time, collectives issued
        from C on the engine's stream - no Python callback per all-gather).  Rank 0's id is broadcast over the
        ``torch.distributed`` group.  Afterwards ``solve_rows_sharded(bp, None)`` shards over it."""
        import torch.distributed as dist

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        buf = (C.c_char * 128)()
        if rank == 0:
            self._check(self.lib.seq_engine_nccl_unique_id(buf))
        box = [bytes(buf)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        idb = (C.c_char * 128).from_buffer_copy(box[0])
        with self.torch.cuda.device(self.device):
            self._check(self.lib.seq_engine_nccl_init(self.handle, idb, rank, world))
        self.native_world = world
        return world

    def solve_rows_sharded(self, bp: BatchProblem, comm: "_lib.Comm", tens=None, to_host=True):
        """ONE large system with rho row-sharded across the ranks of ``comm`` (C5): every rank calls
        this with the same ``bp`` and gets the full result.  ``comm`` carries the all-gather /
        all-reduce hooks (``sequencing_b200.distributed.make_comm``).  ``tens``: operands already staged
        by ``to_device``; ``to_host=False`` returns the device output tensors instead of a BatchOutput."""
        if tens is None:
            tens = self.to_device(bp)
        torch = self.torch
        with torch.cuda.device(self.device):
            outs = {}
            for name, (shape, dt) in self._out_shapes(bp).items():
                tdt = {np.float64: torch.float64, np.complex128: torch.complex128, np.int32: torch.int32}[dt]
                outs[name] = torch.zeros(shape, dtype=tdt, device=self.device)
            ptr = {k: (v.data_ptr() if v.numel() else None) for k, v in tens.items()}
            ptr.update({k: (v.data_ptr() if v.numel() else None) for k, v in outs.items()})
            desc = self._desc(bp, ptr, host=False)
            self._check(self.lib.seq_engine_solve_sharded(self.handle, C.byref(desc), C.byref(comm) if comm is not None else None))
            self.last_launches = self.lib.seq_engine_last_launch_count(self.handle)
            self.total_launches += self.last_launches
            self._last_method = self.lib.seq_engine_last_method(self.handle)
            torch.cuda.synchronize(self.device)
        if not to_host:
            return {"out": outs, "in": tens}
        return self._finish(bp, {k: v.cpu().numpy() for k, v in outs.items()})

    def _finish(self, bp, out) -> BatchOutput:
        self.last_kernel_ms = float(self.lib.seq_engine_last_kernel_ms(self.handle))
        return BatchOutput(
            expect=out["expect"], final_state=out["final_state"], states=out.get("states"),
            status=out["status"], stats=out["stats"], kernel_ms=self.last_kernel_ms,
            launches=self.last_launches, method=_lib.METHOD_NAMES.get(getattr(self, "_last_method", 0), "?"),
            variant=(self.lib.seq_engine_last_variant(self.handle) or b"").decode(),
        )

    def kernel_ms(self) -> float:
        return float(self.lib.seq_engine_last_kernel_ms(self.handle))

    # -- on-device waveform synthesis ------------------------------------
    def synth_gaussian(self, descs, B: int, n_ch: int, n_t: int, out=None):
        """Coefficient tables f64 [B, n_ch, n_t] generated on the device from pulse descriptors: ``descs`` is a
        list of dicts / tuples with the fields of ``_lib.PulseDesc`` (traj, ch_re, ch_im, start, sign_re, sign_im,
        sigma, chop, dt, amp, drag, detune, phase).  ``out``: accumulate into an existing device tensor."""
        torch = self.torch
        arr = (_lib.PulseDesc * max(len(descs), 1))()
        names = [f[0] for f in _lib.PulseDesc._fields_]
        for q, d in enumerate(descs):
            vals = d if isinstance(d, dict) else dict(zip(names, d))
            for k in names:
                setattr(arr[q], k, vals[k])
        with torch.cuda.device(self.device):
            coeff = out if out is not None else torch.empty((B, n_ch, n_t), dtype=torch.float64, device=self.device)
            self._check(self.lib.seq_engine_synth_gaussian(self.handle, arr, len(descs), coeff.data_ptr(), B, n_ch, n_t, 0 if out is None else 1))
            self.total_launches += self.lib.seq_engine_last_launch_count(self.handle)
        return coeff

    # -- auxiliary entry points --------------------------------------------------------------
    def spline_prep(self, y: np.ndarray, dt: float) -> np.ndarray:
        torch = self.torch
        y = np.ascontiguousarray(y, dtype=np.float64)
        n_t = y.shape[-1]
        ty = torch.from_numpy(y.reshape(-1, n_t)).to(self.device)
        tm = torch.empty_like(ty)
        self._check(self.lib.seq_engine_spline_prep(self.handle, ty.data_ptr(), tm.data_ptr(), ty.shape[0], n_t, float(dt)))
        return tm.cpu().numpy().reshape(y.shape)

    def apply_unitary(self, U: np.ndarray, states: np.ndarray, is_ket: bool) -> np.ndarray:
        torch = self.torch
        N = U.shape[0]
        tu = torch.from_numpy(np.ascontiguousarray(U, dtype=np.complex128)).to(self.device)
        ts = torch.from_numpy(np.ascontiguousarray(states, dtype=np.complex128)).to(self.device)
        B = ts.shape[0]
        self._check(self.lib.seq_engine_apply_unitary(self.handle, tu.data_ptr(), ts.data_ptr(), N, B, int(is_ket)))
        return ts.cpu().numpy()

    def expect(self, E: np.ndarray, states: np.ndarray, is_ket: bool) -> np.ndarray:
        torch = self.torch
        N = E.shape[-1]
        te = torch.from_numpy(_c128(E, (N, N))).to(self.device)
        ts = torch.from_numpy(np.ascontiguousarray(states, dtype=np.complex128)).to(self.device)
        B, n_e = ts.shape[0], te.shape[0]
        out = torch.empty((B, n_e), dtype=torch.complex128, device=self.device)
        self._check(self.lib.seq_engine_expect(self.handle, te.data_ptr(), ts.data_ptr(), out.data_ptr(), N, B, n_e, int(is_ket)))
        return out.cpu().numpy()
