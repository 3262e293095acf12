"""gpupsat_b200 — B200-native drop-in for the data-parallel hot path of nvzoll/gpupsat.

The product is gpupsat_b200/libgpsat.so (hand-written sm_100a CUDA behind the C ABI of include/gpsat.h) and the
`gpupsat` command-line front end.  This package is the thin ctypes binding used by tests and bench.py."""
from .binding import (SAT, UNSAT, UNDEF, DECIDE_REFERENCE, DECIDE_VSIDS, STRATEGY_DISTRIBUTED, STRATEGY_UNIFORM,
                      Cnf, Solver, MultiSolver, GpsatError, solve_cnf, default_opts, lib, RECORD_DTYPE, mesh_attach_local)
from . import instances

__all__ = ["SAT", "UNSAT", "UNDEF", "DECIDE_REFERENCE", "DECIDE_VSIDS", "STRATEGY_DISTRIBUTED", "STRATEGY_UNIFORM",
           "Cnf", "Solver", "MultiSolver", "mesh_attach_local", "GpsatError", "solve_cnf", "default_opts", "lib", "instances", "RECORD_DTYPE"]
