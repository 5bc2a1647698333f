"""The C-ABI library loads and exports every symbol include/vloam_b200.h declares (no compute)."""
import ctypes
import os
import re


def test_exports_match_header(pkg):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "vloam_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(vloam_b200_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 20
    lib_path = pkg._build.build_cuda()
    lib = ctypes.CDLL(lib_path)
    for sym in declared:
        assert hasattr(lib, sym), "missing export " + sym
    assert sorted(pkg.EXPORTS) == declared


def test_default_params(pkg):
    L = pkg.load_lib(build=True)
    p = pkg.Params()
    L.vloam_b200_default_params(ctypes.byref(p))
    # shipped KITTI values, loam_velodyne_HDL_64_kitti.launch:3-16
    assert (p.n_scans, p.mapping_skip_frame) == (64, 1)
    assert abs(p.minimum_range - 5.0) < 1e-7 and abs(p.line_res - 0.4) < 1e-7 and abs(p.plane_res - 0.8) < 1e-7


def test_no_oracle_in_product(pkg):
    """The product path must not reference the oracle."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pdir = os.path.join(root, "vloam-noted_b200")
    for dp, _, files in os.walk(pdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle_py" not in txt and "liboracle" not in txt and "vloam_oracle" not in txt, f
