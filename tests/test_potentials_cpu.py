"""The package's own potential-file readers (lammps_plugins_b200/potentials.py) and the files it ships: same numbers as
the committed golden fixtures, and -- where /root/reference is present -- as the reference's files."""
import os

import numpy as np
import pytest

import support as S
from lammps_plugins_b200 import potentials as P


def test_packaged_rebomos_file_equals_fixture():
    _, params = S.load_rebomos_fixture()
    vals = P.rebomos_values(P.default_path("MoS.REBO.set5b"))
    assert vals == [v for _, v in params]
    a, b = P.read_rebomos(), S.rebomos_params_struct()
    for name, _ in a._fields_:
        x, y = getattr(a, name), getattr(b, name)
        assert np.array_equal(np.ctypeslib.as_array(x), np.ctypeslib.as_array(y)), name


def test_packaged_aeam_file_equals_fixture():
    t, g = P.read_aeam(), S.load_aeam_fixture()
    assert (t["nelements"], t["nnonangular"], t["nangular"], t["names"]) == (g["nelements"], g["nnonangular"], g["nangular"], g["names"])
    assert t["nrho"] == g["nrho"] and t["drho"] == g["drho"] and t["mass"] == g["mass"]
    for k in ("nr", "dr", "cut"):
        assert np.array_equal(t[k], g[k])
    nel = t["nelements"]
    for i in range(nel):
        assert np.array_equal(t["frho"][i], g["frho"][i])
        for j in range(nel):
            assert np.array_equal(t["rhor"][i][j], g["rhor"][i][j])
            if j <= i:
                assert np.array_equal(t["z2r"][i][j], g["z2r"][i][j])


@pytest.mark.skipif(not S.have_reference_tree(), reason="reference tree not present on this box")
def test_packaged_files_equal_reference_files():
    ref = P.rebomos_values(os.path.join(S.REFERENCE, "USER-REBOMOS", "MoS.REBO.set5b"))
    assert ref == P.rebomos_values(P.default_path("MoS.REBO.set5b"))
    r, t = P.read_aeam(os.path.join(S.REFERENCE, "USER-AEAM", "AlSi.aeam")), P.read_aeam()
    assert r["nrho"] == t["nrho"] and r["drho"] == t["drho"] and np.array_equal(r["nr"], t["nr"])
    assert np.array_equal(r["dr"], t["dr"]) and np.array_equal(r["cut"], t["cut"])
    for i in range(r["nelements"]):
        assert np.array_equal(r["frho"][i], t["frho"][i])
        for j in range(r["nelements"]):
            assert np.array_equal(r["rhor"][i][j], t["rhor"][i][j])
            if j <= i:
                assert np.array_equal(r["z2r"][i][j], t["z2r"][i][j])


def test_reader_errors():
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".set5b", delete=False) as fh:
        fh.write("# DATE: x UNITS: metal\n1.0 a\n2.0 b\n")
    with pytest.raises(ValueError):
        P.read_rebomos(fh.name)
    os.unlink(fh.name)
