"""fp32 FFMA peak vs legacy mma.sync tf32 peak on this GPU (njode_diag.cu microbenchmarks)"""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from njode_b200 import _ext
lib = _ext.cuda_lib().dll
for fn in (lib.njode_fma_peak_launch, lib.njode_mma_tf32_peak_launch):
    fn.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_void_p]
scratch = torch.empty(148 * 8 * 256 * 2, dtype=torch.float32, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for name, fn, scale in (("fp32 FFMA", lib.njode_fma_peak_launch, 1.0), ("mma.sync m16n8k8 tf32", lib.njode_mma_tf32_peak_launch, 1.0)):
    best = 0.0
    n = C.c_double()
    for it in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(C.c_void_p(scratch.data_ptr()), 4096, C.byref(n), st); b.record(); torch.cuda.synchronize()
        best = max(best, n.value / (a.elapsed_time(b) * 1e-3))
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    print("%-24s %8.2f T MAC/s = %7.1f MAC/clk/SM at 1.965 GHz (x2 = %.1f TFLOP/s)" % (name, best / 1e12, best / sms / 1.965e9, 2 * best / 1e12))
