"""N>1 on real devices: two ranks, one process per GPU, NCCL communicator inside libqcm_b200.so.  Each rank builds ITS share
of the sigma plan (edges of the MPO bond graph sharded by their step-1 index), the library sums the partial sigma vectors
with ncclAllReduce; rank 0 compares the result with the CPU oracle on the same instance.  Skips when fewer than two devices
are visible (the driver's GPU test box has one; the 2/4/8-GPU runs exercise the same path through bench.py, which reports
parity_rel_err_vs_oracle at every N)."""
import ctypes, os, socket, sys, tempfile
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, libs, path, symm, norb, nelec, site, M, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cu = ctypes.CDLL(libs["cuda"], mode=ctypes.RTLD_GLOBAL)
    cu.qcm_last_error.restype = ctypes.c_char_p
    host = ctypes.CDLL(libs["host"]); host.qcmd_create.restype = ctypes.c_void_p
    assert cu.qcm_init(rank) == 0, cu.qcm_last_error()
    idbuf = ctypes.create_string_buffer(128)
    if rank == 0:
        assert cu.qcm_comm_unique_id(idbuf) == 0, cu.qcm_last_error()
    t = torch.tensor(list(idbuf.raw), dtype=torch.uint8)
    dist.broadcast(t, 0)
    idbuf = ctypes.create_string_buffer(bytes(t.tolist()), 128)
    assert cu.qcm_comm_init(rank, world, idbuf) == 0, cu.qcm_last_error()
    err = ctypes.create_string_buffer(1024)
    h = ctypes.c_void_p(host.qcmd_create(path.encode(), symm.encode(), norb, nelec, err, 1024)); assert h.value, err.value
    info = (ctypes.c_double * 32)()
    assert host.qcmd_setup_site(h, site, 1, M, 1, rank, rank, world, info, err, 1024) == 0, err.value
    n_psi, n_sig = int(info[5]), int(info[6])
    psi = torch.empty(n_psi, dtype=torch.float64); sig = torch.empty(n_sig, dtype=torch.float64)
    host.qcmd_get_psi(h, ctypes.c_void_p(psi.data_ptr()))
    assert host.qcmd_sigma_host(h, ctypes.c_void_p(psi.data_ptr()), ctypes.c_void_p(sig.data_ptr()), err, 1024) == 0, err.value
    if rank == 0:
        olib = ctypes.CDLL(libs["oracle"]); olib.orc_create.restype = ctypes.c_void_p
        oh = ctypes.c_void_p(olib.orc_create(path.encode(), symm.encode(), norb, nelec, err, 1024)); assert oh.value, err.value
        pe, sec, ov, se = ctypes.c_double(), ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        assert olib.orc_setup_site(oh, site, 1, M, 1, ctypes.byref(pe), err, 1024) == 0, err.value
        assert olib.orc_sigma(oh, 1, ctypes.byref(sec), ctypes.byref(ov), ctypes.byref(se), err, 1024) == 0, err.value
        ref = torch.empty(int(se.value), dtype=torch.float64)
        olib.orc_get_sigma(oh, ctypes.c_void_p(ref.data_ptr()))
        q.put((int(se.value) == n_sig, float((sig - ref).norm() / ref.norm()) if int(se.value) == n_sig else 1.0, info[0]))
    dist.barrier()
    host.qcmd_destroy(h)
    cu.qcm_comm_destroy()
    dist.destroy_process_group()


@pytest.mark.parametrize("symm", ["su2u1", "2u1"])
def test_two_rank_nccl_sigma_matches_the_oracle(built, symm):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    from qcmaquis_b200.fcidump import make_fcidump
    path = os.path.join(tempfile.mkdtemp(prefix="qcm_test_"), "s.fcidump")
    make_fcidump(path, 12, 12)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    libs = {k: built[k] for k in ("cuda", "host", "oracle")}
    mp.spawn(_worker, args=(2, port, libs, path, symm, 12, 12, 5, 300, q), nprocs=2, join=True)
    same, rel, flops = q.get()
    assert same, "sigma block structure differs from the oracle"
    assert rel < 1e-10
