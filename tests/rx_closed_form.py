"""numpy statement of the closed-form reverse mode of the receiver functional — the same
recipe `k_rx_adjoint` (hmcmt2d_b200/csrc/mt_kernels.cuh) implements; checked against the oracle's sparse
L^T d, Q^T d in test_oracle_known_answers.py."""
import numpy as np

from oracle.sensitivity import linearInterp

MU0 = 4e-7 * np.pi


def rx_adjoint_closed_form(mode, omega, yLen, h, sig1, F01, rxY, yNode, d):
    ny = len(yLen)
    F0, F1 = F01[:, 0], F01[:, 1]
    iw, iwmu = 1j * omega, 1j * omega * MU0
    yb = 0.5 * yLen[:-1] + 0.5 * yLen[1:]                     # interior nodes 1..ny-1
    if mode == 0:
        Qc = (0.75 * np.diff(F0) / yLen / iw + 0.25 * np.diff(F1) / yLen / iw) / MU0
        HyH = -(F1[1:-1] - F0[1:-1]) / h / iwmu
        ExQ = 0.75 * F0[1:-1] + 0.25 * F1[1:-1]
        sv = (0.5 * sig1[:-1] * yLen[:-1] + 0.5 * sig1[1:] * yLen[1:]) / yb
        G0 = np.zeros(ny + 1, complex)
        G0[1:-1] = HyH - (np.diff(Qc) / yb - sv * ExQ) * (0.5 * h)
    else:
        Qc = (0.75 * (-np.diff(F0) / yLen) + 0.25 * (-np.diff(F1) / yLen)) / sig1
        JyH = (F1[1:-1] - F0[1:-1]) / h
        rv = (0.5 * yLen[:-1] / sig1[:-1] + 0.5 * yLen[1:] / sig1[1:]) / yb
        HxQ = 0.75 * F0[1:-1] + 0.25 * F1[1:-1]
        G0 = np.zeros(ny + 1, complex)
        G0[1:-1] = JyH * rv - (np.diff(Qc) / yb + iwmu * HxQ) * (0.5 * h)
    G0[0], G0[-1] = G0[1], G0[-2]
    aF0 = np.zeros(ny + 1, complex); aF1 = np.zeros(ny + 1, complex); aG0 = np.zeros(ny + 1, complex)
    for r, y in enumerate(rxY):
        iL, iR, wL, wR = linearInterp(y, yNode)
        if mode == 0:
            num, den = wL * F0[iL] + wR * F0[iR], wL * G0[iL] + wR * G0[iR]
        else:
            num, den = wL * G0[iL] + wR * G0[iR], wL * F0[iL] + wR * F0[iR]
        nbar, dbar = d[r] / den, -d[r] * num / den ** 2
        tn, td = (aF0, aG0) if mode == 0 else (aG0, aF0)
        tn[iL] += wL * nbar; tn[iR] += wR * nbar
        td[iL] += wL * dbar; td[iR] += wR * dbar
    aG0[1] += aG0[0]; aG0[-2] += aG0[-1]
    u = aG0[1:-1]
    q = np.zeros(ny, complex)
    if mode == 0:
        a = u / h / iwmu
        aF1[1:-1] -= a; aF0[1:-1] += a
        e = sv * (0.5 * h) * u
        aF0[1:-1] += 0.75 * e; aF1[1:-1] += 0.25 * e
        svbar = ExQ * (0.5 * h) * u
        q[:-1] += svbar * 0.5 * yLen[:-1] / yb
        q[1:] += svbar * 0.5 * yLen[1:] / yb
    else:
        a = rv * u / h
        aF1[1:-1] += a; aF0[1:-1] -= a
        e = -(iwmu * 0.5 * h) * u
        aF0[1:-1] += 0.75 * e; aF1[1:-1] += 0.25 * e
        rvbar = JyH * u
        q[:-1] += rvbar * (-0.5 * yLen[:-1] / (sig1[:-1] ** 2 * yb))
        q[1:] += rvbar * (-0.5 * yLen[1:] / (sig1[1:] ** 2 * yb))
    gq = -(0.5 * h) / yb * u                                  # adjoint of dQ/dy
    aQ = np.zeros(ny, complex)
    aQ[1:] += gq
    aQ[:-1] -= gq
    if mode == 0:
        wq = aQ / MU0 / yLen / iw
    else:
        wq = -(aQ / sig1 / yLen)
        q += aQ * (-(Qc * sig1) / sig1 ** 2)
    aF0[1:] += 0.75 * wq; aF0[:-1] -= 0.75 * wq
    aF1[1:] += 0.25 * wq; aF1[:-1] -= 0.25 * wq
    return aF0, aF1, q
