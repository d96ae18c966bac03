"""Host-side logic of the product (readers, inverse-model setup, synthetic generator) against the oracle's
restatement of the reference readers; no GPU needed."""
import os

import numpy as np
import pytest

from hmcmt2d_b200 import api, fileio, synthetic
from oracle import fileio as ofio
from oracle import sampler as osamp
from tests.helpers import GOLDEN


@pytest.mark.parametrize("name,modfile,datfile", [("dprism3d", "dprism2d_G96x49.mod", "dprism2dobs.dat"),
                                                 ("coprod2", "coprod2.mod", "coprod2data.dat")])
def test_readers_match_oracle(name, modfile, datfile):
    d = os.path.join(GOLDEN, name)
    pm, om = fileio.readEMModel2D(os.path.join(d, modfile)), ofio.readEMModel2D(os.path.join(d, modfile))
    assert tuple(pm.gridSize) == tuple(om.gridSize)
    for a in ("yLen", "zLen", "airLayer", "origin", "sigma"):
        assert np.array_equal(getattr(pm, a), getattr(om, a)), a
    (pd, pobs, perr), (od, oobs, oerr) = fileio.readMT2DData(os.path.join(d, datfile)), ofio.readMT2DData(os.path.join(d, datfile))
    for a in ("rxLoc", "freqs", "rxID", "freqID", "dtID", "dataID"):
        assert np.array_equal(getattr(pd, a), getattr(od, a)), a
    assert pd.dataComp == od.dataComp == ["ZXY", "ZYX"] and pd.compTE and pd.compTM
    assert np.array_equal(pobs, oobs) and np.array_equal(perr, oerr)


def test_example_sizes():
    """dprism 96x56, N=5225, 11 freqs, 41 rx, 902 data, 4704 params; coprod2 76x52, 470 data, 3420."""
    mesh, data, inv, prior = api.readstartupFile(os.path.join(GOLDEN, "dprism3d", "startupfile"))
    assert tuple(mesh.gridSize) == (96, 56) and len(data.freqs) == 11 and data.rxLoc.shape == (41, 2)
    assert len(inv.obsData) == 902 and len(inv.strModel) == 4704
    assert prior.dt == 0.03 and prior.timestep == [6, 10] and prior.totalsamples == 10000 and prior.burninsamples == 100
    assert np.allclose(prior.sigBounds, [1e-4, 1.0]) and prior.regParam == 1.0
    mesh, data, inv, prior = api.readstartupFile(os.path.join(GOLDEN, "coprod2", "startupfile"))
    assert tuple(mesh.gridSize) == (76, 52) and len(inv.obsData) == 470 and len(inv.strModel) == 3420 and prior.dt == 0.015


def test_inverse_model_setup_matches_oracle():
    d = os.path.join(GOLDEN, "coprod2")
    mesh, data, inv, prior = api.readstartupFile(os.path.join(d, "startupfile"))
    omesh, odata, oinv, oprior = osamp.readstartupFile(os.path.join(d, "startupfile"), d)
    assert (inv.activeCell != oinv.activeCell).nnz == 0
    assert np.array_equal(inv.bgModel, oinv.bgModel) and np.array_equal(inv.strModel, oinv.strModel)
    assert abs(inv.Wm - oinv.Wm).max() == 0
    assert np.array_equal(inv.dataW, oinv.dataW)
    # Wm = (G P)^T (G P): differences across the air/earth interface are one-sided
    ny = mesh.gridSize[0]
    diag = inv.Wm.diagonal()
    assert diag[0] == 3.0 and diag[1] == 4.0 and diag[ny + 1] == 4.0          # corner / top-row / interior earth cells


def test_model_and_data_writers_roundtrip(tmp_path):
    d = os.path.join(GOLDEN, "dprism3d")
    mesh = fileio.readEMModel2D(os.path.join(d, "dprism2d_G96x49.mod"))
    fileio.writeEMModel2D(str(tmp_path / "m.mod"), mesh, stamp="t")
    ofio.writeEMModel2D(str(tmp_path / "o.mod"), ofio.readEMModel2D(os.path.join(d, "dprism2d_G96x49.mod")), stamp="t")
    assert (tmp_path / "m.mod").read_text() == (tmp_path / "o.mod").read_text()
    back = fileio.readEMModel2D(str(tmp_path / "m.mod"))
    assert np.array_equal(back.yLen, mesh.yLen) and np.array_equal(back.zLen, mesh.zLen)
    assert np.allclose(back.sigma, mesh.sigma, rtol=5e-3) and np.allclose(back.origin, mesh.origin)
    data, obs, err = fileio.readMT2DData(os.path.join(d, "dprism2dobs.dat"))
    fileio.writeMT2DData(str(tmp_path / "d.dat"), data, obs, err, stamp="t")
    od, oobs, oerr = ofio.readMT2DData(os.path.join(d, "dprism2dobs.dat"))
    ofio.writeMT2DData(str(tmp_path / "o.dat"), od, oobs, oerr, stamp="t")
    assert (tmp_path / "d.dat").read_text() == (tmp_path / "o.dat").read_text()
    d2, obs2, err2 = fileio.readMT2DData(str(tmp_path / "d.dat"))
    assert np.allclose(obs2, obs, rtol=1e-6) and np.array_equal(d2.freqID, data.freqID)


def test_startup_parser_quirks(tmp_path):
    p = tmp_path / "startupfile"
    p.write_text("# comment\ndatafile: a.dat\nmodelfile: b.mod\n\nburninsamples: 7\ntotalsamples: 33\n"
                 "resistivity: 0.5 2e3 0.05\ntimeinterval: 0.02\ntimestep: 3 9\nlinearsolver: mumps\nsmoothparameter: 2.5\n")
    for parse in (fileio.parseStartup, ofio.parse_startup):
        datafile, modelfile, smin, smax, sigfix, prior = parse(str(p))
        assert (datafile, modelfile) == ("a.dat", "b.mod") and sigfix == [1e-8]
        assert prior.burninsamples == 7 and prior.totalsamples == 33 and prior.dt == 0.02 and prior.timestep == [3, 9]
        assert np.isclose(smin, 1 / 2e3) and np.isclose(smax, 2.0) and prior.regParam == 2.5 and prior.linearSolver == "mumps"


def test_synthetic_generator_sizes():
    """cfg2 of SURVEY.md section 8: ny=200, nz=100, 30 freqs -> N=19701, nCell=20000, nAC=18600, 40 rx."""
    mesh, data, inv, prior = synthetic.make_problem(200, 100, 30)
    ny, nz = mesh.gridSize
    assert (ny - 1) * (nz - 1) == 19701 and ny * nz == 20000 and len(inv.strModel) == 18600
    assert len(data.freqs) == 30 and data.rxLoc.shape == (40, 2) and len(inv.obsData) == 30 * 40 * 2
    assert np.isclose(data.freqs[0], 100.0) and np.isclose(data.freqs[-1], 1e-3)
    znode = np.concatenate([[0], np.cumsum(mesh.zLen)]) - mesh.origin[1]
    assert np.abs(znode[7]) < 1e-9                               # receivers sit on the air/earth interface
    assert np.all(np.diff(np.lexsort((data.dtID, data.rxID, data.freqID))) == 1)      # rows sorted (freq, rx, comp)
    assert mesh.yLen[0] == 200 * 2 ** 8 and mesh.zLen[-1] == 100 * 2 ** 8


def test_output_files_format(tmp_path):
    """outputHMCSamples (HMCSampler.jl:785-828): file names and line formats."""
    nparam, ns, nd = 5, 3, 4
    rng = np.random.default_rng(0)
    model = rng.standard_normal((nparam, ns))
    data = rng.standard_normal((nd, ns + 1)) + 1j * rng.standard_normal((nd, ns + 1))
    st = api.HMCStatus(2, 1, np.array([True, False, True]), np.abs(rng.standard_normal((4, ns + 1))))
    api.outputHMCSamples(model, st, data, ichain=2, cputime=1.5, outdir=str(tmp_path))
    lines = (tmp_path / "hmcsamples_id2.model").read_text().splitlines()
    assert len(lines) == ns and lines[0].split()[0] == "%8.4e" % model[0, 0]
    lines = (tmp_path / "hmcstatistics_id2.log").read_text().splitlines()
    assert lines[0] == "Total elapsed time (s):     1.50" and lines[1].startswith("Totalsamples:      3, nAccept:      2")
    assert len(lines) == 4 + ns and lines[4].split()[-1] == "1" and lines[5].split()[-1] == "0"
    assert len((tmp_path / "hmcsamples_id2.data").read_text().splitlines()) == ns + 1
