"""Where the host binary's start-up time goes on a box: CUDA context creation vs pinned / device
allocations vs the first kernel launch (module load). python scripts/cuda_startup.py"""
import ctypes as C
import time

t0 = time.perf_counter()
rt = C.CDLL("libcudart.so.12")
t1 = time.perf_counter()
n = C.c_int()
rt.cudaGetDeviceCount(C.byref(n))
t2 = time.perf_counter()
rt.cudaSetDevice(0)
rt.cudaFree(None)
t3 = time.perf_counter()
p = C.c_void_p()
rt.cudaHostAlloc(C.byref(p), C.c_size_t(64 << 20), 0)
t4 = time.perf_counter()
q = C.c_void_p()
rt.cudaMalloc(C.byref(q), C.c_size_t(64 << 20))
t5 = time.perf_counter()
s = C.c_void_p()
rt.cudaStreamCreate(C.byref(s))
rt.cudaMemcpyAsync(q, p, C.c_size_t(64 << 20), 1, s)
rt.cudaStreamSynchronize(s)
t6 = time.perf_counter()
print(f"dlopen {t1 - t0:.3f} s, cudaGetDeviceCount {t2 - t1:.3f} s ({n.value} devices), context {t3 - t2:.3f} s, "
      f"cudaHostAlloc 64 MiB {t4 - t3:.3f} s, cudaMalloc 64 MiB {t5 - t4:.3f} s, first 64 MiB H2D {t6 - t5:.3f} s")
