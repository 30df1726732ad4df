#!/usr/bin/env python
"""Regenerates the golden fixtures of tests/golden/ from the reference tree (run in the build container, where
/root/reference is mounted; the GPU box only sees the committed outputs).

What is taken from the reference are INPUT DATA and KNOWN ANSWERS of its own tests and shipped example runs --
integral tables, printed energies, printed MPO bond dimensions, the Wigner 9j table -- never source code:

  h2_2o.fcidump       integrals of dmrg/tests/test1.cpp:36-47                    E = -0.980724992658492 (test1.cpp:93)
  lih_4o.fcidump      integrals of dmrg/tests/Fixtures/LiHFixture.h:28-96         E = -7.90435750473166  (LiHFixture.h:112)
                      (the same table is test_mps_mpo_ops/test_siteproblem_fixture.h:21-88)
  h2_4o.fcidump       examples/iTD-DMRG/FCIDUMP_H2_TI                             E = -1.151682732118105 (H2_2e4o.TI.SS.out:70)
  benzene_6o.fcidump  examples/TD-DMRG/FCIDUMP_Benzene_631Gs_CAS6-6 (real parts)
  reference_values.json   energies, "MPO Bond p: dim/pairs" lines, "The hamiltonian will contain N terms" lines,
                      the 24 nine-j values of dmrg/tests/test_wigner.cpp:21-46
"""
import json, os, re, sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def integral_map(path, first_line=0, last_line=10 ** 9):
    pat = re.compile(r"\{\s*\{\s*(\d+)\s*,\s*(\d+)\s*,\s*(\d+)\s*,\s*(\d+)\s*\}\s*,\s*([-+0-9.eE]+)\s*\}")
    out = []
    for n, line in enumerate(open(path), 1):
        if n < first_line or n > last_line or line.lstrip().startswith("//"):
            continue
        m = pat.search(line)
        if m:
            out.append((m.group(5), [int(m.group(i)) for i in range(1, 5)]))
    return out


def write_fcidump(path, norb, nelec, entries):
    with open(path, "w") as f:
        f.write(" &FCI NORB=%d,NELEC=%d,MS2=0,\n  ORBSYM=%s\n  ISYM=1\n &END\n" % (norb, nelec, ",".join(["1"] * norb)))
        for val, idx in entries:
            f.write(" %s %d %d %d %d\n" % (val, idx[0], idx[1], idx[2], idx[3]))


def mpo_bonds(path):
    dims, pairs, terms = [], [], None
    for line in open(path):
        m = re.match(r"MPO Bond (\d+): (\d+)/(\d+)", line)
        if m:
            dims.append(int(m.group(2))); pairs.append(int(m.group(3)))
        m = re.match(r"The hamiltonian will contain (\d+) terms", line)
        if m:
            terms = int(m.group(1))
    return {"dims": dims, "pairs": pairs, "terms": terms}


def energies(path, n):
    e = []
    for line in open(path):
        m = re.match(r"\s*Energy = ([-0-9.eE]+)", line)
        if m:
            e.append(float(m.group(1)))
    return e[:n]


def main():
    if not os.path.isdir(REF):
        sys.exit("the reference tree is not mounted at " + REF)
    write_fcidump(os.path.join(HERE, "h2_2o.fcidump"), 2, 2, integral_map(REF + "/dmrg/tests/test1.cpp", 36, 48))
    write_fcidump(os.path.join(HERE, "lih_4o.fcidump"), 4, 2, integral_map(REF + "/dmrg/tests/Fixtures/LiHFixture.h"))
    # FCIDUMP_H2_TI is already in the 4-header-line form the reference parser expects (parse_integrals.h:100-102)
    lines = open(REF + "/examples/iTD-DMRG/FCIDUMP_H2_TI").read().splitlines()
    ent = []
    for l in lines[4:]:
        t = l.split()
        if len(t) == 5:
            ent.append((t[0], [int(x) for x in t[1:]]))
    write_fcidump(os.path.join(HERE, "h2_4o.fcidump"), 4, 2, ent)
    # the benzene dump carries (real, imaginary) value columns; the imaginary parts are all zero
    ent = []
    for l in open(REF + "/examples/TD-DMRG/FCIDUMP_Benzene_631Gs_CAS6-6").read().splitlines()[4:]:
        t = l.split()
        if len(t) == 6:
            assert float(t[1]) == 0.0
            ent.append((t[0], [int(x) for x in t[2:]]))
    write_fcidump(os.path.join(HERE, "benzene_6o.fcidump"), 6, 6, ent)

    nine_j = []
    pat = re.compile(r"\{\s*\{([0-9,\s]+)\}\s*,\s*([-0-9.]+)\s*\}")
    for line in open(REF + "/dmrg/tests/test_wigner.cpp"):
        m = pat.search(line)
        if m:
            nine_j.append([[int(x) for x in m.group(1).split(",")], float(m.group(2))])
    assert len(nine_j) == 24
    vals = {
        "energies": {
            "h2_2o": {"value": -0.980724992658492, "source": "dmrg/tests/test1.cpp:93", "symmetries": ["su2u1pg", "su2u1", "2u1pg", "2u1"]},
            "lih_4o": {"value": -7.90435750473166, "source": "dmrg/tests/Fixtures/LiHFixture.h:112"},
            "h2_4o": {"value": -1.151682732118105, "source": "examples/iTD-DMRG/H2_2e4o.TI.SS.out:70"},
        },
        "h2_4o_microiteration_energies_2u1pg_singlesite_const_init": energies(REF + "/examples/iTD-DMRG/H2_2e4o.TI.SS.out", 8),
        "mpo": {
            "h2_4o/2u1pg": dict(mpo_bonds(REF + "/examples/iTD-DMRG/H2_2e4o.TI.SS.out"), source="examples/iTD-DMRG/H2_2e4o.TI.SS.out:15-19"),
            "h2_4o/su2u1pg": dict(mpo_bonds(REF + "/examples/DMRG-FEAST/H2_2e4o.FEAST.TS.out"), source="examples/DMRG-FEAST/H2_2e4o.FEAST.TS.out:14-18"),
            "benzene_6o/su2u1pg": dict(mpo_bonds(REF + "/examples/TD-DMRG/Benzene_6e6o.TD.TS.out"), source="examples/TD-DMRG/Benzene_6e6o.TD.TS.out:15-21"),
        },
        "wigner_9j": {"source": "dmrg/tests/test_wigner.cpp:21-46 (arguments are 2*j, tolerance 1e-6 relative)", "cases": nine_j},
    }
    json.dump(vals, open(os.path.join(HERE, "reference_values.json"), "w"), indent=1)
    print("wrote fixtures to", HERE)


if __name__ == "__main__":
    main()
