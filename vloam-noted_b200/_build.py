"""Build recipes: the CUDA C-ABI library (nvcc, sm_100a) and the host-only synthetic generator."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvloam_b200.so")
SYNTH_LIB = os.path.join(HERE, "libvloam_synth.so")
CU_SOURCES = ["scan_registration.cu", "voxel_grid.cu", "laser_odometry.cu", "laser_mapping.cu", "lm_solver.cu", "capi.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # bit-exact parity with the reference's x86-64 arithmetic: no FMA contraction,
    # IEEE division and square root
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_cuda(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in CU_SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh", ".hpp"))]
    deps.append(os.path.join(ROOT, "include", "vloam_b200.h"))
    if force or _stale(LIB, deps):
        cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs
        subprocess.run(cmd, check=True)
    return LIB


def build_synth(force=False):
    src = os.path.join(HERE, "synth", "synth.cpp")
    if force or _stale(SYNTH_LIB, [src]):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SYNTH_LIB, src], check=True)
    return SYNTH_LIB
