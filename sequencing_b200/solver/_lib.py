"""ctypes binding of ``include/seq_engine.h`` (the C-ABI drop-in boundary).

The product path has no CPU fallback: if ``_seq_engine.so`` is missing or a symbol is absent the
import of the library raises, and any solve without a CUDA device raises ``EngineUnavailable``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "_seq_engine.so")

# mirrors of the #defines in include/seq_engine.h
ABI_VERSION = 3
SEQ_OK = 0
STATUS_OK, STATUS_MAX_STEPS, STATUS_NONFINITE, STATUS_STEP_UNDERFLOW = 0, 1, 2, 3
FLAG_KET, FLAG_NORMALIZE, FLAG_STORE_STATES, FLAG_EXPECT_COMPLEX, FLAG_HOST_POINTERS = 0x1, 0x2, 0x4, 0x8, 0x10
FLAG_NO_ORDER_SWITCH = 0x20
METHOD_AUTO, METHOD_GENERIC, METHOD_DMMA, METHOD_SMALL, METHOD_DMMA_STREAM, METHOD_LARGE, METHOD_SPARSE, METHOD_DIAG = 0, 1, 2, 3, 4, 5, 6, 7
METHOD_NAMES = {0: "auto", 1: "generic", 2: "dmma", 3: "small", 4: "dmma_stream", 5: "large", 6: "sparse", 7: "diag"}

EXPORTS = (
    "seq_engine_abi_version", "seq_engine_build_info", "seq_engine_create", "seq_engine_destroy",
    "seq_engine_last_error", "seq_engine_workspace_bytes", "seq_engine_solve",
    "seq_engine_select_method", "seq_engine_last_launch_count", "seq_engine_last_kernel_ms",
    "seq_engine_spline_prep", "seq_engine_apply_unitary", "seq_engine_expect", "seq_engine_solve_sharded",
    "seq_engine_nccl_unique_id", "seq_engine_nccl_init", "seq_engine_nccl_destroy",
    "seq_engine_last_method", "seq_engine_last_variant", "seq_engine_last_fma_per_rhs", "seq_engine_last_step_extra_ops", "seq_engine_synth_gaussian",
)


class EngineUnavailable(RuntimeError):
    """The CUDA extension or a CUDA device is missing; there is deliberately no CPU fallback."""


class SolveDesc(C.Structure):
    """``struct seq_solve_desc`` - field order and types must match include/seq_engine.h."""

    _fields_ = [
        ("size", C.c_uint32), ("flags", C.c_uint32),
        ("N", C.c_int32), ("B", C.c_int32), ("n_t", C.c_int32), ("n_ch", C.c_int32),
        ("n_c", C.c_int32), ("n_ctd", C.c_int32), ("n_e", C.c_int32), ("n_out", C.c_int32),
        ("nsteps", C.c_int32), ("method", C.c_int32),
        ("t0", C.c_double), ("dt", C.c_double), ("atol", C.c_double), ("rtol", C.c_double),
        ("max_step", C.c_double), ("first_step", C.c_double), ("min_step", C.c_double),
        ("H0", C.c_void_p), ("H0_batch_stride", C.c_int64),
        ("Hk", C.c_void_p),
        ("coeff", C.c_void_p), ("coeff_batch_stride", C.c_int64), ("coeff_scale", C.c_void_p),
        ("Lc", C.c_void_p), ("Ltd", C.c_void_p),
        ("rate", C.c_void_p), ("rate_batch_stride", C.c_int64),
        ("state0", C.c_void_p), ("E", C.c_void_p), ("out_idx", C.c_void_p),
        ("expect", C.c_void_p), ("final_state", C.c_void_p), ("states", C.c_void_p),
        ("status", C.c_void_p), ("stats", C.c_void_p),
    ]


class PulseDesc(C.Structure):
    """``struct seq_pulse_desc`` - one chopped Gaussian(-DRAG) pulse of one trajectory (on-device synthesis)."""

    _fields_ = [("traj", C.c_int32), ("ch_re", C.c_int32), ("ch_im", C.c_int32), ("start", C.c_int32),
                ("sign_re", C.c_double), ("sign_im", C.c_double),
                ("sigma", C.c_double), ("chop", C.c_double), ("dt", C.c_double),
                ("amp", C.c_double), ("drag", C.c_double), ("detune", C.c_double), ("phase", C.c_double)]


ALL_GATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64)
ALL_REDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64)


class Comm(C.Structure):
    """``struct seq_comm`` - collective hooks of the row-sharded large-system solve."""

    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("ctx", C.c_void_p),
                ("all_gather", ALL_GATHER_FN), ("all_reduce_sum", ALL_REDUCE_FN)]


_lib = None


def load():
    """Load the shared library once and declare every prototype; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineUnavailable(
            f"{LIB_PATH} not found - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). sequencing_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    missing = [name for name in EXPORTS if not hasattr(lib, name)]
    if missing:
        raise EngineUnavailable(f"{LIB_PATH} lacks symbols {missing}")
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.seq_engine_abi_version.restype = C.c_int
    lib.seq_engine_build_info.restype = C.c_char_p
    lib.seq_engine_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    lib.seq_engine_destroy.argtypes = [vp]
    lib.seq_engine_last_error.argtypes = [vp]
    lib.seq_engine_last_error.restype = C.c_char_p
    lib.seq_engine_workspace_bytes.argtypes = [vp, C.POINTER(SolveDesc)]
    lib.seq_engine_workspace_bytes.restype = i64
    lib.seq_engine_solve.argtypes = [vp, C.POINTER(SolveDesc)]
    lib.seq_engine_select_method.argtypes = [vp, C.POINTER(SolveDesc)]
    lib.seq_engine_solve_sharded.argtypes = [vp, C.POINTER(SolveDesc), C.POINTER(Comm)]
    lib.seq_engine_last_launch_count.argtypes = [vp]
    lib.seq_engine_last_method.argtypes = [vp]
    lib.seq_engine_last_variant.argtypes = [vp]
    lib.seq_engine_last_variant.restype = C.c_char_p
    lib.seq_engine_nccl_unique_id.argtypes = [vp]
    lib.seq_engine_nccl_init.argtypes = [vp, vp, C.c_int32, C.c_int32]
    lib.seq_engine_nccl_destroy.argtypes = [vp]
    lib.seq_engine_last_fma_per_rhs.argtypes = [vp]
    lib.seq_engine_last_fma_per_rhs.restype = C.c_int64
    lib.seq_engine_last_step_extra_ops.argtypes = [vp]
    lib.seq_engine_last_step_extra_ops.restype = C.c_int64
    lib.seq_engine_last_kernel_ms.argtypes = [vp]
    lib.seq_engine_last_kernel_ms.restype = C.c_float
    lib.seq_engine_spline_prep.argtypes = [vp, vp, vp, i64, i32, C.c_double]
    lib.seq_engine_apply_unitary.argtypes = [vp, vp, vp, i32, i32, C.c_int]
    lib.seq_engine_synth_gaussian.argtypes = [vp, C.POINTER(PulseDesc), i64, vp, i32, i32, i32, C.c_int]
    lib.seq_engine_expect.argtypes = [vp, vp, vp, vp, i32, i32, i32, C.c_int]
    if lib.seq_engine_abi_version() != ABI_VERSION:
        raise EngineUnavailable("seq_engine ABI version mismatch")
    if C.sizeof(SolveDesc) <= 0:
        raise EngineUnavailable("bad SolveDesc layout")
    _lib = lib
    return lib
