"""TEST INFRASTRUCTURE -- an independent known answer for the synthetic systems: full configuration interaction in numpy,
straight from the FCIDUMP integrals (second quantisation over alpha / beta occupation strings, dense Hamiltonian, Sz = 0
sector).  It shares NOTHING with the product or the oracle (no MPO builder, no symmetry containers, no C++): a DMRG sweep
that keeps the full bond dimension must reproduce its ground-state energy, which pins term generation + MPO construction +
contraction + solver for the systems the reference holds no printed energy for (the 6e/6o and 8e/8o = BASELINE config 1
synthetic integrals).  H = sum_pq h_pq a+_p a_q + 1/2 sum_pqrs (pq|rs) a+_p a+_r a_s a_q + E_core (chemists' notation, the
FCIDUMP convention the reference reads in dmrg/models/chem/parse_integrals.h)."""
import itertools
import numpy as np


def read_fcidump(path):
    txt = open(path).read()
    head, body = txt.split("&END")
    norb = int(head.split("NORB=")[1].split(",")[0])
    nelec = int(head.split("NELEC=")[1].split(",")[0])
    h = np.zeros((norb, norb)); eri = np.zeros((norb,) * 4); core = 0.0
    for line in body.strip().splitlines():
        v, i, j, k, l = line.split()
        v = float(v); i, j, k, l = int(i), int(j), int(k), int(l)
        if i == 0:
            core = v
        elif k == 0:
            h[i - 1, j - 1] = h[j - 1, i - 1] = v
        else:
            i, j, k, l = i - 1, j - 1, k - 1, l - 1
            for a, b, c, d in ((i, j, k, l), (j, i, k, l), (i, j, l, k), (j, i, l, k), (k, l, i, j), (l, k, i, j), (k, l, j, i), (l, k, j, i)):
                eri[a, b, c, d] = v
    return norb, nelec, h, eri, core


def _apply(op_list, det):
    """apply a product of (creator?, spin-orbital) operators, rightmost first, to an occupation bit string; returns (sign, det) or None"""
    sign = 1
    for create, p in reversed(op_list):
        occ = (det >> p) & 1
        if create == occ:
            return None
        sign *= -1 if bin(det & ((1 << p) - 1)).count("1") & 1 else 1
        det ^= 1 << p
    return sign, det


def fci_ground_state_energy(path, nup=None, ndown=None, total_spin=None, n_states=1):
    """lowest energy in the (nup, ndown) sector (n_states > 1: the n_states lowest, as a list); total_spin = 2S: lowest state(s) of that
    total spin (what an SU2 run with spin = 2S targets).  H = sum_pq [h_pq - 1/2 sum_r (pr|rq)] E_pq + 1/2 sum_pqrs (pq|rs) E_pq E_rs with the spin-summed excitation
    operators E_pq = sum_sigma a+_{p sigma} a_{q sigma}, built as sparse matrices over the determinants."""
    import scipy.sparse as sp
    norb, nelec, h, eri, core = read_fcidump(path)
    nup = nelec // 2 if nup is None else nup
    ndown = nelec - nup if ndown is None else ndown
    dets = []       # spin orbital index: 2 * orbital + spin
    for a in itertools.combinations(range(norb), nup):
        for b in itertools.combinations(range(norb), ndown):
            d = 0
            for p in a: d |= 1 << (2 * p)
            for p in b: d |= 1 << (2 * p + 1)
            dets.append(d)
    index = {d: i for i, d in enumerate(dets)}
    n = len(dets)
    E = {}
    for p in range(norb):
        for q in range(norb):
            rows, cols, vals = [], [], []
            for col, d in enumerate(dets):
                for s in (0, 1):
                    r = _apply([(1, 2 * p + s), (0, 2 * q + s)], d)
                    if r: rows.append(index[r[1]]); cols.append(col); vals.append(float(r[0]))
            E[p, q] = sp.csr_matrix((vals, (rows, cols)), shape=(n, n))
    heff = h - 0.5 * np.einsum("prrq->pq", eri)
    H = sp.csr_matrix((n, n))
    for p in range(norb):
        for q in range(norb):
            G = sp.csr_matrix((n, n))
            for r in range(norb):
                for s_ in range(norb):
                    if eri[p, q, r, s_] != 0.0: G = G + eri[p, q, r, s_] * E[r, s_]
            H = H + heff[p, q] * E[p, q] + 0.5 * (E[p, q] @ G)
    H = H.toarray()
    assert np.abs(H - H.T).max() < 1e-10
    if total_spin is None:
        w = np.linalg.eigvalsh(H)
        return float(w[0]) + core if n_states == 1 else [float(x) + core for x in w[:n_states]]
    # lowest state of total spin S = total_spin / 2.  S^2 = S- S+ + Sz (Sz + 1) in the same basis; H commutes with it, a small
    # multiple of S^2 added to H lifts accidental degeneracies between multiplets, so the eigenvectors are spin pure
    S2 = np.zeros((n, n))
    sz = 0.5 * (nup - ndown)
    for col, d in enumerate(dets):
        S2[col, col] += sz * (sz + 1)
        for p_ in range(norb):
            for q_ in range(norb):
                r = _apply([(1, 2 * p_ + 1), (0, 2 * p_), (1, 2 * q_), (0, 2 * q_ + 1)], d)      # S-_p S+_q
                if r: S2[index[r[1]], col] += r[0]
    assert np.abs(H @ S2 - S2 @ H).max() < 1e-9
    w, v = np.linalg.eigh(H + 1e-3 * S2)
    target = 0.5 * total_spin * (0.5 * total_spin + 1)
    found = []
    for k in range(n):
        s2 = float(v[:, k] @ S2 @ v[:, k])
        if abs(s2 - target) < 1e-6:
            found.append(float(v[:, k] @ H @ v[:, k]) + core)
            if len(found) == n_states:
                return found[0] if n_states == 1 else sorted(found)
    raise RuntimeError("not enough states of the requested spin")
