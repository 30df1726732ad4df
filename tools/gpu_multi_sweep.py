"""Two-site sweeps with the engine sharded over N GPUs (torchrun): every rank drives the same host loop, sigma / boundaries are
combined by the library's NCCL allreduce; rank 0 compares the energies with a single-GPU run of the same sweeps.
usage: python -m torch.distributed.run --nproc-per-node N tools/gpu_multi_sweep.py [norb] [M]"""
import ctypes, json, os, sys, time
sys.path.insert(0, os.getcwd())
os.environ["OMP_NUM_THREADS"] = str(max(1, len(os.sched_getaffinity(0)) // int(os.environ.get("WORLD_SIZE", "1"))))
import torch, torch.distributed as dist
import bench
from qcmaquis_b200 import build
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
norb = int(sys.argv[1]) if len(sys.argv) > 1 else 12
M = int(sys.argv[2]) if len(sys.argv) > 2 else 300
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cu = ctypes.CDLL(build.build_cuda(), mode=ctypes.RTLD_GLOBAL); host = ctypes.CDLL(build.build_host()); host.qcmd_create.restype = ctypes.c_void_p
cu.qcm_last_error.restype = ctypes.c_char_p
assert cu.qcm_init(local) == 0, cu.qcm_last_error()
path = bench.make_fcidump(norb, norb) if rank == 0 else None
obj = [path]; dist.broadcast_object_list(obj, 0); path = obj[0]
e = bench.errbuf()
def sweeps(r, w):
    h = ctypes.c_void_p(host.qcmd_create(path.encode(), b"su2u1", norb, norb, e, 1024))
    en = (ctypes.c_double * 4096)(); n = ctypes.c_int(); info = (ctypes.c_double * 8)()
    t = time.time()
    assert host.qcmd_ts_sweeps_ranked(h, 60, M, 1, 42, local, r, w, en, 4096, ctypes.byref(n), info, e, 1024) == 0, e.value
    return list(en[:n.value]), time.time() - t, info[1]
single = sweeps(0, 1) if rank == 0 else None          # before the communicator exists: plain single-GPU engine
idbuf = ctypes.create_string_buffer(128)
if rank == 0: assert cu.qcm_comm_unique_id(idbuf) == 0
t = torch.tensor(list(idbuf.raw), dtype=torch.uint8, device="cuda"); dist.broadcast(t, 0)
idbuf = ctypes.create_string_buffer(bytes(t.cpu().tolist()), 128)
assert cu.qcm_comm_init(rank, world, idbuf) == 0, cu.qcm_last_error()
dist.barrier()
multi = sweeps(rank, world)
dist.barrier()
if rank == 0:
    print(json.dumps({"workload": "%de%do su2u1 two-site sweep, M=%d" % (norb, norb, M), "n_gpus": world, "sweep_seconds_1gpu": single[2], "sweep_seconds_ngpu": multi[2],
                      "micro_iterations": len(multi[0]), "max_abs_energy_diff_1_vs_n": max(abs(a - b) for a, b in zip(single[0], multi[0])), "final_energy": multi[0][-1]}))
cu.qcm_comm_destroy(); dist.destroy_process_group()
