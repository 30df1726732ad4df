"""In-tree build of the native libraries (nvcc / g++ directly; no JIT cache, the .so files travel with the tree).

  qcmaquis_b200/lib/libqcm_b200.so   CUDA kernels + the C ABI of include/qcm_b200.h        (product)
  qcmaquis_b200/lib/libqcm_host.so   C++ host driver above the C ABI                       (product)
  oracle/_build/liboracle.so         CPU oracle entry points                               (checker only)
  tests/harness/libqcm_harness*.so   parity harness (oracle + engine under test)           (tests only)
"""
import glob, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "qcmaquis_b200", "csrc")
LIB = os.path.join(ROOT, "qcmaquis_b200", "lib")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]
CXX_FLAGS = ["-O2", "-std=c++17", "-fopenmp", "-fPIC", "-shared"]


def scipy_openblas():
    import scipy
    d = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
    libs = sorted(glob.glob(os.path.join(d, "libscipy_openblas*.so")))
    if not libs:
        raise RuntimeError("scipy's OpenBLAS not found under " + d)
    return d, libs[0]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))


def _headers():
    return glob.glob(os.path.join(CSRC, "qcm", "*.hpp")) + glob.glob(os.path.join(ROOT, "include", "*.h"))


def build_cuda(force=False):
    """Every .cu under csrc/ is one translation unit (compiled side by side), linked into libqcm_b200.so."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(LIB, exist_ok=True)
    odir = os.path.join(ROOT, "build", "obj")
    os.makedirs(odir, exist_ok=True)
    out = os.path.join(LIB, "libqcm_b200.so")
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = _headers() + glob.glob(os.path.join(CSRC, "*.cuh"))
    objs, jobs = [], []
    for src in srcs:
        obj = os.path.join(odir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + deps):
            jobs.append(["nvcc"] + [f for f in NVCC_FLAGS if f != "-shared"] + ["-c", src, "-o", obj])
    # host-side translation units of the same library (the schedule builder behind the descriptor entry points): g++, OpenMP
    for src in sorted(glob.glob(os.path.join(CSRC, "plan_capi.cpp"))):
        obj = os.path.join(odir, os.path.basename(src)[:-4] + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + deps):
            jobs.append(["g++", "-O2", "-std=c++17", "-fopenmp", "-fPIC", "-I", CSRC, "-I", "/usr/local/cuda/include", "-c", src, "-o", obj])
    if jobs:
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            list(ex.map(_run, jobs))
    if jobs or force or _newer(out, objs):
        _run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", out, "-ldl", "-lgomp"])
    return out


def build_host(force=False):
    build_cuda(force)
    out = os.path.join(LIB, "libqcm_host.so")
    src = os.path.join(CSRC, "driver_capi.cpp")
    if force or _newer(out, [src] + _headers()):
        # host-side BLAS/LAPACK (site-operator algebra at MPO construction) comes from scipy's OpenBLAS, as the
        # reference host links a BLAS; nothing on the sigma / boundary path calls it
        d, blas = scipy_openblas()
        _run(["g++"] + CXX_FLAGS + ["-I", CSRC, src, "-o", out, "-L", LIB, "-lqcm_b200", blas, "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + d])
    return out


def build_oracle(force=False):
    d, blas = scipy_openblas()
    odir = os.path.join(ROOT, "oracle", "_build")
    os.makedirs(odir, exist_ok=True)
    out = os.path.join(odir, "liboracle.so")
    srcs = [os.path.join(ROOT, "oracle", "oracle_capi.cpp"), os.path.join(ROOT, "oracle", "oracle_engine.hpp")]
    if force or _newer(out, srcs + _headers()):
        _run(["g++"] + CXX_FLAGS + ["-I", CSRC, "-I", os.path.join(ROOT, "oracle"), srcs[0], "-o", out, blas, "-Wl,-rpath," + d])
    return out


def build_harness(force=False, gpu=True):
    d, blas = scipy_openblas()
    hdir = os.path.join(ROOT, "tests", "harness")
    src = os.path.join(hdir, "harness.cpp")
    deps = [src, os.path.join(hdir, "plan_interp.hpp"), os.path.join(ROOT, "oracle", "oracle_engine.hpp")] + _headers()
    out_cpu = os.path.join(hdir, "libqcm_harness_cpu.so")
    base = ["g++"] + CXX_FLAGS + ["-I", CSRC, "-I", os.path.join(ROOT, "oracle"), "-I", hdir, src]
    if force or _newer(out_cpu, deps):
        _run(base + ["-o", out_cpu, blas, "-Wl,-rpath," + d])
    out_gpu = os.path.join(hdir, "libqcm_harness.so")
    if gpu:
        build_cuda(force)
        if force or _newer(out_gpu, deps + [os.path.join(LIB, "libqcm_b200.so")]):
            _run(base + ["-DQCMT_WITH_GPU", "-o", out_gpu, "-L", LIB, "-lqcm_b200", blas, "-Wl,-rpath," + d,
                         "-Wl,-rpath,$ORIGIN/../../qcmaquis_b200/lib"])
    return out_cpu, out_gpu


def build_all(force=False):
    return {"cuda": build_cuda(force), "host": build_host(force), "oracle": build_oracle(force), "harness": build_harness(force)}


if __name__ == "__main__":
    for k, v in build_all("--force" in sys.argv).items():
        print(k, v)
