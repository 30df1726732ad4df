"""The descriptor entry points of include/qcm_b200.h (qcm_mpo_upload, qcm_plan_sigma, qcm_plan_left_step, qcm_plan_right_step,
qcm_plan_out_size / qcm_plan_out_blocks): the problem crosses the C ABI as plain arrays only -- MPO tensor in CSC form, operator
table as sparse entries, bond spins, Hermitian maps, block structures -- is planned inside libqcm_b200.so and executed on the
GPU; the result (values AND block structure, read back through qcm_plan_out_blocks) must equal the CPU oracle's.  The
flattening code the test uses (tests/harness/flatten_desc.hpp) is the binding INTEGRATION.md shows for the QCMaquis side."""
import ctypes
import pytest
from conftest import golden

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.mark.parametrize("symm", ["su2u1", "2u1", "su2u1pg", "2u1pg"])
@pytest.mark.parametrize("site,twosite", [(2, True), (3, False)])
def test_plans_built_from_descriptors_match_the_oracle(harness_gpu, symm, site, twosite):
    out = (ctypes.c_double * 8)(); err = ctypes.create_string_buffer(1024)
    rc = harness_gpu.lib.qcmt_desc_parity(golden("synth_6o6e.fcidump"), symm.encode(), 6, 6, site, int(twosite), 40, 3, out, err, 1024)
    assert rc == 0, err.value.decode()
    assert out[0] == 1 and out[1] < TOL, list(out)
    if not twosite:
        assert out[2] == 1 and out[3] < TOL and out[4] == 1 and out[5] < TOL, list(out)
        assert out[6] == 2 and out[7] < TOL, list(out)        # qcm_plan_noise_left / qcm_plan_noise_right


def test_descriptor_plan_at_config1_size(harness_gpu, fcidump_8o8e):
    out = (ctypes.c_double * 8)(); err = ctypes.create_string_buffer(1024)
    rc = harness_gpu.lib.qcmt_desc_parity(fcidump_8o8e, b"su2u1", 8, 8, 3, 1, 256, 1, out, err, 1024)
    assert rc == 0, err.value.decode()
    assert out[0] == 1 and out[1] < TOL, list(out)
