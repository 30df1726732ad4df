"""Measure the FP64 roofline denominators on the box (DFMA, DMMA, HBM copy) through the C ABI."""
import ctypes, json, os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(root, "qcmaquis_b200", "lib", "libqcm_b200.so"))
lib.qcm_last_error.restype = ctypes.c_char_p
assert lib.qcm_init(0) == 0, lib.qcm_last_error()
res = {}
for name in ["qcm_measure_fp64_fma_peak", "qcm_measure_fp64_dmma_peak", "qcm_measure_hbm_copy"]:
    v = ctypes.c_double()
    best = 0.0
    for _ in range(3):
        rc = getattr(lib, name)(ctypes.byref(v))
        assert rc == 0, lib.qcm_last_error()
        best = max(best, v.value)
    res[name.replace("qcm_measure_", "")] = best
buf = ctypes.create_string_buffer(128); lib.qcm_device_name(buf, 128); res["device"] = buf.value.decode()
print(json.dumps(res))
