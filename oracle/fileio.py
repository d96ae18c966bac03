"""Model / data / startup file formats — restates `HMCMT/src/HMCFileIO/*.jl` and
`HMCMT/src/HMCSampler/readstartupFile.jl` (test infrastructure; the product has its own
readers in `hmcmt2d_b200/fileio.py`, which the tests compare against these).
"""
from __future__ import annotations

from dataclasses import dataclass, field
import numpy as np

from . import operators as ops


@dataclass
class TensorMesh2D:
    """`TensorMesh2D` HMCFileIO.jl:46-60.  zLen includes the air layers (air first)."""
    yLen: np.ndarray
    zLen: np.ndarray
    airLayer: np.ndarray
    gridSize: tuple
    origin: np.ndarray
    sigma: np.ndarray
    Face: object = None
    Grad: object = None
    AveCN: object = None
    AveCF: object = None
    setup: bool = False


@dataclass
class MTData:
    """`MTData` HMCFileIO.jl:26-41.  rxID/freqID/dtID are kept 1-based as in the file."""
    rxLoc: np.ndarray
    freqs: np.ndarray
    dataType: str
    dataComp: list
    rxID: np.ndarray
    freqID: np.ndarray
    dtID: np.ndarray
    dataID: np.ndarray
    compTE: bool
    compTM: bool


def setupTensorMesh2D(mesh: TensorMesh2D) -> TensorMesh2D:
    """`setupTensorMesh2D!` MT2DOperators.jl:16-27."""
    mesh.Face = ops.meshGeoFace2D(mesh.yLen, mesh.zLen)
    mesh.Grad = ops.getNodalGradient2D(mesh.yLen, mesh.zLen)
    mesh.AveCN = ops.aveCell2Node2D(mesh.gridSize)
    mesh.AveCF = ops.aveCell2Face2D(mesh.gridSize)
    mesh.setup = True
    return mesh


def _lines(path):
    """Non-empty, non-comment stripped lines (readEMModel2D.jl:31-37 skip rule)."""
    out = []
    with open(path, "r") as f:
        for raw in f:
            s = raw.strip()
            if not s or s[0] == "#":
                continue
            out.append(s)
    return out


def _read_floats(lines, pos, n):
    vals = []
    while len(vals) < n:
        vals.extend(float(t) for t in lines[pos].split())
        pos += 1
    return np.array(vals[:n], dtype=np.float64), pos


def readEMModel2D(modelfile: str) -> TensorMesh2D:
    """`readEMModel2D` readEMModel2D.jl:11-154.

    Keywords are matched by substring (NY, NZ, NAIR, 'Resistivity Type', 'Model Type',
    'Origin'); air layers are listed bottom-up, reversed and prepended (:135-144).
    """
    lines = _lines(modelfile)
    ny = nz = nAir = 0
    yLen = zLen = None
    airLayer = np.zeros(0)
    sigma = None
    origin = np.zeros(2)
    resType = ""
    pos = 0
    while pos < len(lines):
        cl = lines[pos]
        pos += 1
        if "NY" in cl:
            ny = int(cl.split()[-1])
            yLen, pos = _read_floats(lines, pos, ny)
        elif "NZ" in cl:
            nz = int(cl.split()[-1])
            zLen, pos = _read_floats(lines, pos, nz)
        elif "NAIR" in cl:
            nAir = int(cl.split()[-1])
            airLayer, pos = _read_floats(lines, pos, nAir)
        elif "Resistivity Type" in cl:
            resType = cl.split()[-1]
        elif "Model Type" in cl:
            modType = cl.split()[-1]
            sigma, pos = _read_floats(lines, pos, ny * nz)
            if resType == "Resistivity":
                sigma = 1.0 / sigma
            if modType == "log":
                # readEMModel2D.jl:117-119 calls exp(::Vector) which throws in Julia 1.x
                raise ValueError("Model Type 'log' is not supported by the reference (exp(::Vector))")
        elif "Origin" in cl:
            t = cl.split()
            origin = np.array([float(t[-2]), float(t[-1])])
    if airLayer.size:
        zLen = np.concatenate([airLayer[::-1], zLen])
        origin = origin.copy()
        origin[1] += airLayer.sum()
        sigma = np.concatenate([np.full(ny * nAir, 1e-8), sigma])
    nzt = len(zLen)
    return TensorMesh2D(yLen, zLen, airLayer, (ny, nzt), origin, sigma)


def readMT2DData(datafile: str):
    """`readMT2DData` readMT2DData.jl:14-179 -> (MTData, obsData, dataErr)."""
    lines = _lines(datafile)
    pos = 0
    rxLoc = freqs = None
    dataType = ""
    dataComp = []
    isComplex = False
    rxID = freqID = dtID = obs = err = None
    while pos < len(lines):
        cl = lines[pos]
        pos += 1
        if "Format" in cl:
            pass
        elif "Receiver Location" in cl:
            nr = int(cl.split()[-1])
            rxLoc = np.zeros((nr, 2))
            for i in range(nr):
                t = lines[pos].split()
                pos += 1
                rxLoc[i, 0], rxLoc[i, 1] = float(t[0]), float(t[1])
        elif "Frequencies" in cl:
            nf = int(cl.split()[-1])
            freqs = np.zeros(nf)
            for i in range(nf):
                freqs[i] = float(lines[pos])
                pos += 1
        elif "DataType" in cl:
            dataType = cl.split()[-1]
            if dataType not in ("Impedance", "Rho_Pha"):
                raise ValueError(f"{dataType} is not supported.")
            isComplex = dataType == "Impedance"
        elif "DataComp" in cl:
            nDt = int(cl.split()[-1])
            dataComp = []
            for i in range(nDt):
                dataComp.append(lines[pos].strip())
                pos += 1
        elif "Data Block" in cl:
            nData = int(cl.split()[-1])
            rxID = np.zeros(nData, dtype=np.int64)
            dtID = np.zeros(nData, dtype=np.int64)
            freqID = np.zeros(nData, dtype=np.int64)
            obs = np.zeros(nData, dtype=np.complex128 if isComplex else np.float64)
            err = np.zeros(nData)
            for i in range(nData):
                t = lines[pos].split()
                pos += 1
                freqID[i], rxID[i], dtID[i] = int(t[0]), int(t[1]), int(t[2])
                if isComplex:
                    obs[i] = float(t[3]) + 1j * float(t[4])
                    err[i] = float(t[5])
                else:
                    obs[i] = float(t[3])
                    err[i] = float(t[4])
    compTE = any("XY" in c for c in dataComp)
    compTM = any("YX" in c for c in dataComp)
    nr, nf, nDt = rxLoc.shape[0], len(freqs), len(dataComp)
    # dataID = vec(Bool[nDt, nr, nf]) column-major: comp fastest, then rx, then freq (:165-172)
    dataID = np.zeros((nf, nr, nDt), dtype=bool)
    dataID[freqID - 1, rxID - 1, dtID - 1] = True
    dataID = dataID.reshape(-1)
    return MTData(rxLoc, freqs, dataType, dataComp, rxID, freqID, dtID, dataID, compTE, compTM), obs, err


def writeEMModel2D(modelfile: str, mesh: TensorMesh2D, stamp: str = "") -> None:
    """`writeEMModel2D` writeEMModel2D.jl:11-82 (time stamp replaced by `stamp`)."""
    ny, nz = len(mesh.yLen), len(mesh.zLen)
    with open(modelfile, "w") as f:
        f.write("%-18s %s\n" % ("#Format:", "EMModel2DFile"))
        f.write("%-18s %s\n" % ("#Description:", "file generated in " + stamp))
        f.write("%-6s %4d\n" % ("NY:", ny))
        for i in range(1, ny + 1):
            f.write("%10.2f" % mesh.yLen[i - 1])
            if i % 8 == 0:
                f.write("\n")
        if ny % 8 != 0:
            f.write("\n")
        nAir = len(mesh.airLayer)
        if nAir:
            f.write("%-6s %4d\n" % ("NAIR:", nAir))
            for i in range(1, nAir + 1):
                f.write("%12.2f" % mesh.airLayer[i - 1])
                if i % 8 == 0:
                    f.write("\n")
            if nAir % 8 != 0:
                f.write("\n")
        f.write("%-6s %4d\n" % ("NZ:", nz - nAir))
        for i in range(nAir + 1, nz + 1):
            f.write("%10.2f" % mesh.zLen[i - 1])
            if (i - nAir) % 8 == 0:
                f.write("\n")
        if (nz - nAir) % 8 != 0:
            f.write("\n")
        sig = np.asarray(mesh.sigma)[ny * nAir:].reshape(nz - nAir, ny)
        f.write("%-18s %s\n" % ("Resistivity Type:", "Conductivity"))
        f.write("%-18s %s\n" % ("Model Type:", "Linear"))
        for k in range(nz - nAir):
            for j in range(ny):
                f.write("%4.2e " % sig[k, j])
            f.write("\n")
        origin = np.array(mesh.origin, dtype=float)
        if nAir:
            origin[1] -= np.sum(mesh.airLayer)
        f.write("%-15s %4.2e %4.2e" % ("Origin (m):", origin[0], origin[1]))


def writeMT2DData(datafile: str, d: MTData, predData, dataErr=None, stamp: str = "") -> None:
    """`writeMT2DData` writeMT2DData.jl:12-86."""
    predData = np.asarray(predData)
    if dataErr is None or len(dataErr) == 0:
        dataErr = np.abs(predData) * 0.03
    elif len(dataErr) == 1:
        dataErr = np.abs(predData) * dataErr[0]
    with open(datafile, "w") as f:
        f.write("%-20s%s\n" % ("Format:", "MT2DData_1.0"))
        f.write("# %s\n" % ("file generated in " + stamp))
        nr = d.rxLoc.shape[0]
        f.write("%-25s %4d\n" % ("Receiver Location (m):", nr))
        f.write("# %5s %5s\n" % ("Y", "Z"))
        for i in range(nr):
            f.write("%12.2f %12.2f\n" % (d.rxLoc[i, 0], d.rxLoc[i, 1]))
        f.write("%-20s%3d\n" % ("Frequencies (Hz):", len(d.freqs)))
        for fr in d.freqs:
            f.write("%8.4e\n" % fr)
        f.write("%-12s %12s\n" % ("DataType:", d.dataType))
        f.write("%-15s %d\n" % ("DataComp:", len(d.dataComp)))
        for c in d.dataComp:
            f.write("%4s\n" % c)
        f.write("%-15s %d\n" % ("Data Block:", len(predData)))
        if np.iscomplexobj(predData):
            f.write("# %6s %6s %10s %10s %15s %12s\n" % ("FreqNo.", "RxNo.", "dataComp", "RealValue", "ImagValue", "Error"))
            for i in range(len(predData)):
                f.write("%5d %6d %8d %15.6e %15.6e %15.6e\n" % (d.freqID[i], d.rxID[i], d.dtID[i],
                                                             predData[i].real, predData[i].imag, dataErr[i]))
        else:
            f.write("# %6s %6s %10s %10s %12s\n" % ("FreqNo.", "RxNo.", "dataComp", "RealValue", "Error"))
            for i in range(len(predData)):
                f.write("%5d %6d %8d %15.6e %15.6e\n" % (d.freqID[i], d.rxID[i], d.dtID[i], predData[i], dataErr[i]))


@dataclass
class HMCPrior:
    """`HMCPrior` HMCStruct.jl:18-38; defaults `initHMCPrior` :129-140."""
    burninsamples: int = 100
    totalsamples: int = 500
    sigBounds: list = field(default_factory=lambda: [0.01, 10.0])
    sigmastd: float = 0.05
    dt: float = 0.01
    timestep: list = field(default_factory=lambda: [10, 15])
    linearSolver: str = ""
    massType: str = "diagonal"
    regParam: float = 1.0
    nfevals: int = 0


def parse_startup(startupfile: str):
    """Key/value part of `readstartupFile` readstartupFile.jl:4-83.

    Substring matching in the reference's branch order: a `fixedresistivity:` line also
    contains `resistivity:` and is therefore consumed by the earlier `resistivity:` branch
    (readstartupFile.jl:46-60) — restated as is (it raises in the reference too unless the
    line happens to carry three numbers).
    Returns (datafile, modelfile, sigmin, sigmax, sigfix, HMCPrior).
    """
    datafile = modelfile = None
    sigmin = sigmax = 0.0
    sigfix = [1e-8]
    prior = HMCPrior()
    for cl in _lines(startupfile):
        t = cl.split()
        if "datafile:" in cl:
            datafile = t[-1]
        elif "modelfile:" in cl:
            modelfile = t[-1]
        elif "burninsamples:" in cl:
            prior.burninsamples = int(t[-1])
        elif "totalsamples:" in cl:
            prior.totalsamples = int(t[-1])
        elif "resistivity:" in cl:
            rhomin, rhomax, _rhostd = float(t[-3]), float(t[-2]), float(t[-1])
            sigmin, sigmax = 1.0 / rhomax, 1.0 / rhomin
            prior.sigBounds = [sigmin, sigmax]
            prior.sigmastd = (np.log(sigmax) - np.log(sigmin)) * 0.05
        elif "fixedresistivity:" in cl:      # unreachable, kept for fidelity
            sigfix.append(float(t[-1]))
        elif "timeinterval:" in cl:
            prior.dt = float(t[-1])
        elif "timestep:" in cl:
            prior.timestep = [int(t[-2]), int(t[-1])]
        elif "linearsolver:" in cl:
            prior.linearSolver = t[-1]
        elif "masstype:" in cl:
            prior.massType = t[-1]
        elif "smoothparameter:" in cl:
            prior.regParam = float(t[-1])
    return datafile, modelfile, sigmin, sigmax, sigfix, prior
