"""N>1 path on CPU: two processes (gloo), each builds ITS share of the sigma plan exactly as one GPU rank does
(the edges of the MPO bond graph sharded by their step-1 index, plan.hpp shard_sources), executes it with the plan interpreter, and the
partial sigma vectors are summed with torch.distributed.all_reduce -- the CPU stand-in for the NCCL allreduce in
qcm_site_hamil2.  The sum must equal the oracle's sigma; each rank's share alone must not."""
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
    assert lib.qcmt_rank_sigma(*args, rank, world, 0, None, ctypes.byref(n), err, 1024) == 0, err.value
    part = torch.zeros(n.value, dtype=torch.float64)
    assert lib.qcmt_rank_sigma(*args, rank, world, 1, ctypes.c_void_p(part.data_ptr()), ctypes.byref(n), err, 1024) == 0, err.value
    total = part.clone()
    dist.all_reduce(total)
    if rank == 0:
        ref = torch.zeros(n.value, dtype=torch.float64); m = ctypes.c_longlong()
        assert lib.qcmt_rank_sigma(*args, 0, 1, 2, ctypes.c_void_p(ref.data_ptr()), ctypes.byref(m), err, 1024) == 0, err.value
        q.put((m.value == n.value, float((total - ref).norm() / ref.norm()), float((part - ref).norm() / ref.norm())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("symm", ["su2u1", "2u1"])
def test_two_rank_sharded_sigma(harness_cpu_path, symm):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    mp.spawn(_worker, args=(2, port, harness_cpu_path, symm, q), nprocs=2, join=True)
    same_size, rel_sum, rel_part = q.get()
    assert same_size
    assert rel_sum < 1e-12          # allreduced sigma == oracle sigma
    assert rel_part > 1e-3          # a single rank really holds only part of the sum
