#!/usr/bin/env python
"""Pinned host -> device copy rate of one 88.1 MB batch: default pinned memory (what torch.pin_memory gives) against
write-combined pinned memory (cudaHostAllocWriteCombined), one cudaMemcpyAsync each, CUDA events."""
import ctypes as C, glob, json, os, sys
import numpy as np
import torch

torch.cuda.init()
cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + \
        glob.glob("/usr/local/cuda/lib64/libcudart.so*") + \
        glob.glob(os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
rt = C.CDLL(cands[0])
n = 256 * 21 * 64 * 64 * 4
dst = torch.empty(n // 4, dtype=torch.float32, device="cuda")
out = {"cudart": cands[0]}
for name, flags in (("default", 0), ("write_combined", 4), ("portable_mapped", 1 | 2)):
    p = C.c_void_p()
    rc = rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(flags))
    if rc != 0:
        out[name] = f"cudaHostAlloc rc={rc}"
        continue
    buf = (C.c_char * n).from_address(p.value)
    np.frombuffer(buf, dtype=np.float32)[:] = 1.0
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        rt.cudaMemcpyAsync(C.c_void_p(dst.data_ptr()), p, C.c_size_t(n), C.c_int(1), C.c_void_p(stream))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        rt.cudaMemcpyAsync(C.c_void_p(dst.data_ptr()), p, C.c_size_t(n), C.c_int(1), C.c_void_p(stream))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    out[name] = {"ms": ms, "GBps": n / (ms * 1e-3) / 1e9}
    rt.cudaFreeHost(p)
print(json.dumps(out, indent=1))
