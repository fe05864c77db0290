"""Host emulation of the product's device code -- TEST INFRASTRUCTURE ONLY (see emu_driver.cpp)."""
from __future__ import annotations

import os
import subprocess

from oracle import pyref

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
EMU_SO = os.path.join(HERE, "liblcqp_emu.so")
SOURCES = [os.path.join(HERE, "emu_driver.cpp"), os.path.join(ROOT, "lcqpow_b200", "csrc", "lcqp_device.cuh"),
           os.path.join(ROOT, "lcqpow_b200", "csrc", "lcqp_pas.cuh"), os.path.join(ROOT, "lcqpow_b200", "csrc", "lcqp_osqp.cuh"), os.path.join(ROOT, "lcqpow_b200", "csrc", "lcqp_osqp_impl.inc"),
           os.path.join(ROOT, "lcqpow_b200", "csrc", "lcqp_sparse_host.hpp"),
           os.path.join(ROOT, "include", "lcqp_cuda.h")]


def build(force: bool = False) -> str:
    if not force and os.path.exists(EMU_SO) and all(os.path.getmtime(s) <= os.path.getmtime(EMU_SO) for s in SOURCES):
        return EMU_SO
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DLCQP_HOST_EMU", "-DLCQP_COUNT_WORK", "-ffp-contract=off",
                    "-Wno-unused-function", "-o", EMU_SO, SOURCES[0]], check=True)
    return EMU_SO


import ctypes as C


class EmuOptions(C.Structure):
    """lcqp_cuda_options (include/lcqp_cuda.h, ABI 2): the oracle's option block followed by the OSQPSettings fields."""
    _fields_ = [(("osqp_admm" if n == "reserved0" else n), t) for n, t in pyref.OracleOptions._fields_] + [
        ("osqp_rho", C.c_double), ("osqp_sigma", C.c_double), ("osqp_alpha", C.c_double), ("osqp_delta", C.c_double),
        ("osqp_eps_abs", C.c_double), ("osqp_eps_rel", C.c_double), ("osqp_eps_prim_inf", C.c_double), ("osqp_eps_dual_inf", C.c_double),
        ("osqp_adaptive_rho_tolerance", C.c_double),
        ("osqp_max_iter", C.c_int), ("osqp_check_termination", C.c_int), ("osqp_scaling", C.c_int), ("osqp_adaptive_rho", C.c_int),
        ("osqp_adaptive_rho_interval", C.c_int), ("osqp_polish", C.c_int), ("osqp_polish_refine_iter", C.c_int), ("osqp_reserved", C.c_int),
        ("qpoases_terminationTolerance", C.c_double), ("qpoases_boundTolerance", C.c_double)]


class EmuLib(pyref._Lib):
    """Same calling convention as oracle.pyref.OracleLib (lcqp_cuda_options is layout-identical to
    lcqp_oracle_options; lcqp_cuda_stats to lcqp_oracle_result)."""
    so_path = EMU_SO
    prefix = "lcqp_emu"
    kind = "emu"
    options_cls = EmuOptions

    def __init__(self):
        build()
        super().__init__()
