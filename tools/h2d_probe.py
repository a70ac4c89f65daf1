#!/usr/bin/env python
"""H2D bandwidth from pinned memory: contiguous vs the strided band copy the engine issues (cudaMemcpy2DAsync)."""
import torch, time, ctypes
cudart = ctypes.CDLL("libcudart.so.12")
G, n, band = 20000, 1000, 8405
h = torch.empty(G * n, dtype=torch.float64).pin_memory()
d = torch.empty(band * n, dtype=torch.float64, device="cuda")
s = torch.cuda.current_stream().cuda_stream
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
bytes_ = band * n * 8
c1 = lambda: cudart.cudaMemcpyAsync(ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(h.data_ptr()), ctypes.c_size_t(bytes_), 1, ctypes.c_void_p(s))
c2 = lambda: cudart.cudaMemcpy2DAsync(ctypes.c_void_p(d.data_ptr()), ctypes.c_size_t(band * 8), ctypes.c_void_p(h.data_ptr()), ctypes.c_size_t(G * 8), ctypes.c_size_t(band * 8), ctypes.c_size_t(n), 1, ctypes.c_void_p(s))
for name, fn in (("1D contiguous", c1), ("2D band", c2)):
    dt = t(fn)
    print(f"{name}: {dt*1e3:.3f} ms  {bytes_/dt/1e9:.1f} GB/s")
