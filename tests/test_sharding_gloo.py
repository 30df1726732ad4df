"""N>1 path on CPU: two processes (gloo), each builds ITS share of the sigma plan exactly as one GPU rank does
(the edges of the MPO bond graph sharded by their step-1 index, plan.hpp shard_sources), executes it with the plan interpreter, and the
partial sigma vectors are summed with torch.distributed.all_reduce -- the CPU stand-in for the NCCL allreduce in
qcm_site_hamil2; the exchange of partial W sums inside the plan (reduce-scatter of the exchange region) is played by gloo too.  The sum must equal the oracle's sigma; each rank's share alone must not."""
import ctypes, os, socket, sys
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, libpath, symm, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = ctypes.CDLL(libpath)
    f = os.path.join(ROOT, "tests", "golden", "synth_6o6e.fcidump").encode()
    err = ctypes.create_string_buffer(1024); n = ctypes.c_longlong()
    args = (f, symm.encode(), 6, 6, 2, 1, 40, 1)
    ns = ctypes.c_longlong()
    # phase 0: this rank's plan up to the W pass of the exchange wave (partial sums of the panels fed by both ranks)
    assert lib.qcmt_rank_sigma_phase(0, *args, rank, world, None, ctypes.byref(n), ctypes.byref(ns), err, 1024) == 0, err.value
    xe = n.value
    if xe > 0:
        # the reduce-scatter of qcm_site_hamil2's exchange wave, played by gloo: sum the regions, every rank keeps its chunk
        x = torch.zeros(xe, dtype=torch.float64)
        assert lib.qcmt_rank_sigma_phase(1, *args, rank, world, ctypes.c_void_p(x.data_ptr()), ctypes.byref(n), ctypes.byref(ns), err, 1024) == 0, err.value
        dist.all_reduce(x)
        assert lib.qcmt_rank_sigma_phase(2, *args, rank, world, ctypes.c_void_p(x.data_ptr()), ctypes.byref(n), ctypes.byref(ns), err, 1024) == 0, err.value
    part = torch.zeros(ns.value, dtype=torch.float64)
    assert lib.qcmt_rank_sigma_phase(3, *args, rank, world, ctypes.c_void_p(part.data_ptr()), ctypes.byref(n), ctypes.byref(ns), err, 1024) == 0, err.value
    n = ns
    total = part.clone()
    dist.all_reduce(total)
    if rank == 0:
        ref = torch.zeros(n.value, dtype=torch.float64); m = ctypes.c_longlong()
        assert lib.qcmt_rank_sigma(*args, 0, 1, 2, ctypes.c_void_p(ref.data_ptr()), ctypes.byref(m), err, 1024) == 0, err.value
        q.put((m.value == n.value, float((total - ref).norm() / ref.norm()), float((part - ref).norm() / ref.norm()), xe))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("symm", ["su2u1", "2u1"])
def test_two_rank_sharded_sigma(harness_cpu_path, symm):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    mp.spawn(_worker, args=(2, port, harness_cpu_path, symm, q), nprocs=2, join=True)
    same_size, rel_sum, rel_part, xe = q.get()
    assert same_size
    assert xe > 0                   # the exchange wave is exercised
    assert rel_sum < 1e-12          # allreduced sigma == oracle sigma
    assert rel_part > 1e-3          # a single rank really holds only part of the sum
