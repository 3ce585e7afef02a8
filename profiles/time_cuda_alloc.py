"""How long do cudaMalloc / cudaFree take on this box?  (scmc_run_metropolis / scmc_anneal allocate their scratch per call.)
python profiles/time_cuda_alloc.py"""
import ctypes as C, time, sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
torch.cuda.init(); torch.zeros(1, device="cuda"); torch.cuda.synchronize()
rt = None
for name in ("libcudart.so.12", "libcudart.so"):
    try: rt = C.CDLL(name); break
    except OSError: pass
if rt is None:
    import glob
    rt = C.CDLL(sorted(glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*")) + glob.glob("/usr/local/cuda/lib64/libcudart.so*"))[0])
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]; rt.cudaFree.argtypes = [C.c_void_p]
for size in (8 * 1024, 512 * 1024, 17 * 2**20, 512 * 2**20):
    tm, tf = [], []
    for i in range(60):
        p = C.c_void_p()
        t = time.perf_counter(); rc = rt.cudaMalloc(C.byref(p), size); t1 = time.perf_counter()
        assert rc == 0
        rc = rt.cudaFree(p); t2 = time.perf_counter()
        tm.append(1e3 * (t1 - t)); tf.append(1e3 * (t2 - t1))
    tm, tf = np.array(tm), np.array(tf)
    print(f"size {size/2**20:8.2f} MiB: cudaMalloc ms first {tm[0]:8.3f} median {np.median(tm):7.3f} max {tm.max():8.3f} | cudaFree ms first {tf[0]:8.3f} median {np.median(tf):7.3f} max {tf.max():8.3f}", flush=True)
# with work in flight on the device (cudaFree synchronises)
x = torch.randn(8192, 8192, device="cuda")
for i in range(5):
    y = x @ x
    p = C.c_void_p(); rt.cudaMalloc(C.byref(p), 17 * 2**20)
    t = time.perf_counter(); rt.cudaFree(p); dt = time.perf_counter() - t
    print(f"cudaFree with a matmul in flight: {1e3*dt:.2f} ms")
