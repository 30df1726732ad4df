import ctypes, sys, os
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(root, "tests", "harness", "libqcm_harness.so"))
def parity(f, symm, L, ne, M, seed=42, eng=1, world=1, budget=1<<28):
    out = (ctypes.c_double*16)(); err = ctypes.create_string_buffer(512)
    r = lib.qcmt_chain_parity(os.path.join(root, f).encode(), symm.encode(), L, ne, M, seed, eng, world, ctypes.c_longlong(budget), out, 16, err, 512)
    return r, list(out), err.value.decode()
print("gpus", lib.qcmt_gpu_available())
for f, L, ne in [("tests/golden/synth_4o4e.fcidump", 4, 4), ("tests/golden/synth_6o6e.fcidump", 6, 6)]:
    for symm in ["2u1", "su2u1", "2u1pg", "su2u1pg"]:
        for budget in [1<<28, 300]:
            r, o, e = parity(f, symm, L, ne, 20, budget=budget)
            print(L, symm, budget, "rc", r, e, ["%.3g" % x for x in o[:14]], flush=True)
