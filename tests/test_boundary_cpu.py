"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/b200md.h
declares, fails loudly without a GPU (no CPU fallback), and the LAMMPS-facing plugins register the
reference's style names and mirror its input-error behaviour.  No compute calls are made here."""
import ctypes
import os
import re

import pytest

import lammps_plugins_b200 as b2
import support as S

HEADER = os.path.join(S.REPO, "include", "b200md.h")


def declared_functions():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200md_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    names = declared_functions()
    assert len(names) >= 30
    L = ctypes.CDLL(b2.library_path())
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    bound = {n for n, _, _ in b2.SYMBOLS}
    assert set(names) <= bound, sorted(set(names) - bound)      # the Python binding covers the whole header
    assert b2.lib().b200md_version() == 100


def test_header_cites_reference_for_each_entry_point():
    txt = open(HEADER).read()
    for cite in ("pair_rebomos.cpp:102-111", "pair_rebomos.cpp:281-352", "pair_aeam.cpp:110-479", "pair_aeam.cpp:752-942"):
        assert cite in txt


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(has_gpu(), reason="a GPU is present; the no-device error path cannot be provoked")
def test_header_documents_every_option_and_counter():
    """every name b200md_set_option / b200md_get_counter accepts (csrc/ctx.cu) appears, quoted, in include/b200md.h"""
    src = open(os.path.join(S.REPO, "lammps_plugins_b200", "csrc", "ctx.cu")).read()
    a, b = src.index('extern "C" int b200md_set_option'), src.index("unknown option")
    names = set(re.findall(r'n == "([a-z0-9_]+)"', src[a:]))
    assert len(names) > 30 and "deterministic" in names and "kernel_launches" in names
    hdr = open(HEADER).read()
    missing = sorted(n for n in names if '"%s"' % n not in hdr)
    assert not missing, missing
    assert b > a


def test_no_gpu_means_loud_failure_not_fallback():
    with pytest.raises(b2.B200MDError, match="no CPU fallback"):
        b2.Context(0)


def test_plugins_export_lammpsplugin_init_and_register_reference_names(oracle_built):
    for so, style in ((S.B200_REBOMOS_SO, "rebomos"), (S.B200_AEAM_SO, "aeam")):
        assert os.path.exists(so), so + " not built"
        lmp = S.MiniLmp()
        lmp.command("plugin load " + so)               # dlopen + lammpsplugin_init + register
        if style == "rebomos":
            lmp.commands(S.rebomos_bulk_commands()[:-2])
        else:
            lmp.commands(S.aeam_commands()[:8])
        lmp.command("pair_style " + style)              # the name the reference registers
        with pytest.raises(S.LammpsError, match="Illegal pair_style command"):
            lmp.command("pair_style %s 1.0" % style)
        lmp.close()


def test_b200_plugin_mirrors_reference_input_errors(oracle_built):
    pot = os.path.join(S.potential_dir(), "MoS.REBO.set5b")
    lmp = S.MiniLmp()
    lmp.command("plugin load " + S.B200_REBOMOS_SO)
    lmp.commands(S.rebomos_bulk_commands()[:-2])
    lmp.command("pair_style rebomos")
    for bad in ("pair_coeff * * %s M" % pot, "pair_coeff 1 2 %s M S" % pot, "pair_coeff * * %s Mo Q" % pot,
                "pair_coeff * * %s NULL NULL" % pot):
        with pytest.raises(S.LammpsError, match="Incorrect args for pair coefficients"):
            lmp.command(bad)
    lmp.command("pair_coeff * * %s Mo S" % pot)         # both spellings of molybdenum are accepted
    lmp.command("pair_coeff * * %s M S" % pot)
    lmp.close()
    apot = os.path.join(S.potential_dir(), "AlSi.aeam")
    lmp = S.MiniLmp()
    lmp.command("plugin load " + S.B200_AEAM_SO)
    lmp.commands(S.aeam_commands()[:8])
    lmp.command("pair_style aeam")
    with pytest.raises(S.LammpsError, match="no matching atom order of input file and potential file"):
        lmp.command("pair_coeff * * %s Si Al" % apot)
    with pytest.raises(S.LammpsError, match="No matching element in AEAM potential file"):
        lmp.command("pair_coeff * * %s Al Cu" % apot)
    with pytest.raises(S.LammpsError, match="Cannot open AEAM potential file"):
        lmp.command("pair_coeff * * /nonexistent.aeam Al Si")
    lmp.command("pair_coeff * * %s Al Si" % apot)
    assert lmp.mass()[1] == 27.0 and lmp.mass()[2] == 28.0    # masses come from the file (pair_aeam.cpp:588)
    lmp.close()


@pytest.mark.skipif(has_gpu(), reason="needs a box without GPU")
def test_b200_plugin_fails_loudly_without_gpu(oracle_built):
    lmp = S.make_rebomos_system(S.B200_REBOMOS_SO)
    with pytest.raises(S.LammpsError, match="Cannot open the B200 device"):
        lmp.commands(["fix 1 all nve", "run 0"])
    lmp.close()


def test_cmake_configures_in_shim_mode(tmp_path):
    """the CMake glue mirrors the reference's plugin build (LAMMPS_SOURCE_DIR +
    LAMMPSInterfacePlugin); without a LAMMPS tree the B200MD_USE_SHIM option must configure, and the LAMMPS branch
    must refuse to configure without LAMMPS_SOURCE_DIR, with the reference's message."""
    import shutil
    import subprocess
    cmake = shutil.which("cmake")
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not cmake or not os.path.exists(nvcc):
        pytest.skip("cmake or nvcc not available")
    src = os.path.join(S.REPO, "lammps_plugins_b200")
    r = subprocess.run([cmake, "-S", src, "-B", str(tmp_path / "shim"), "-DB200MD_USE_SHIM=ON",
                        "-DCMAKE_CUDA_COMPILER=" + nvcc], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    r = subprocess.run([cmake, "-S", src, "-B", str(tmp_path / "real"), "-DCMAKE_CUDA_COMPILER=" + nvcc],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "Must set LAMMPS_SOURCE_DIR" in r.stderr
