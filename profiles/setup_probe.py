import ctypes, time, sys, os
sys.path.insert(0, "/root/repo/hm-16.2_b200")
t0 = time.perf_counter()
rt = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else ctypes.CDLL("libcudart.so")
t1 = time.perf_counter()
rt.cudaFree(ctypes.c_void_p(0))
t2 = time.perf_counter()
import hmgpu
t3 = time.perf_counter()
c = hmgpu.Context(416, 240, 8, 16)
t4 = time.perf_counter()
c2 = hmgpu.Context(416, 240, 8, 16)
t5 = time.perf_counter()
print("load cudart %.3f  cudaFree(0)=context %.3f  import hmgpu(+lib load) %.3f  first Context %.3f  second Context %.3f" % (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4))
