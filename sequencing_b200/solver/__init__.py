"""Batched time-evolution engine front end (the L1 integrator of SURVEY.md section 1)."""
from ._lib import EngineUnavailable
from .api import (
    ODE_ERROR,
    Options,
    QobjEvo,
    Result,
    deferred_solves,
    mesolve,
    propagator,
    rhs_clear,
    sesolve,
    solve_batch,
)
from .engine import BatchOutput, BatchProblem, Engine
