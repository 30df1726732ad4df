"""Regenerates tests/golden/fci_values.json: full-CI energies of the synthetic systems by tests/fci_numpy.py (an implementation
that shares nothing with the product or the oracle): the lowest state of the Sz = 0 sector (what a 2u1 run with nup = ndown
targets) and the lowest singlet (what an su2u1 run with spin = 0 targets).  The 8e/8o system is BASELINE config 1's."""
import json, os, sys, tempfile, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from fci_numpy import fci_ground_state_energy
from qcmaquis_b200.fcidump import make_fcidump

out = {}
for name in ("synth_4o4e", "synth_6o6e"):
    f = os.path.join(HERE, name + ".fcidump")
    out[name] = {"sz0_ground_state": fci_ground_state_energy(f), "singlet_ground_state": fci_ground_state_energy(f, total_spin=0)}
path = os.path.join(tempfile.mkdtemp(), "synth_8o8e.fcidump")
make_fcidump(path, 8, 8)
t = time.time()
out["synth_8o8e"] = {"sz0_ground_state": fci_ground_state_energy(path), "singlet_ground_state": fci_ground_state_energy(path, total_spin=0),
                     "source": "qcmaquis_b200.fcidump.make_fcidump(path, 8, 8), the conftest fixture fcidump_8o8e"}
json.dump(out, open(os.path.join(HERE, "fci_values.json"), "w"), indent=1)
print(out, time.time() - t)
