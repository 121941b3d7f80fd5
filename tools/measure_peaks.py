"""Measure the roofline denominators MEASURED_PEAKS.json does not carry (same method as the driver:
torch.matmul 8192^3, best of 10 = burst, back-to-back for 3 s = sustained)."""
import json, os, sys, time
import torch
dev = torch.device("cuda:0")
out = {}
def gemm_peak(dtype, tf32):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    n = 8192
    a = torch.randn(n, n, device=dev, dtype=dtype); b = torch.randn(n, n, device=dev, dtype=dtype)
    for _ in range(3): a @ b
    best = 1e9
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    burst = 2 * n ** 3 / best / 1e9
    t0 = time.time(); cnt = 0
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(10): a @ b
        cnt += 10; torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    sustained = 2 * n ** 3 * cnt / e0.elapsed_time(e1) / 1e9
    return burst, sustained
out["tf32_tflops"], out["tf32_tflops_sustained"] = gemm_peak(torch.float32, True)
out["fp32_simt_tflops"], out["fp32_simt_tflops_sustained"] = gemm_peak(torch.float32, False)
out["bf16_tflops"], out["bf16_tflops_sustained"] = gemm_peak(torch.bfloat16, False)
# HBM copy
x = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev); y = torch.empty_like(x)
best = 1e9
for _ in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); y.copy_(x); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
out["hbm_gbs"] = 2 * x.numel() * 2 / best / 1e6
out["how"] = "torch.matmul 8192^3 (allow_tf32 on/off, bf16), best of 10 and 3 s back-to-back; copy_ of 1 Gi bf16"
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/measured_peaks_extra.json", "w"), indent=1)
# issue-rate peaks of the SIMT pipes (the denominators of the L1-evaluation and log-domain Sinkhorn rooflines)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
from gnn_mtl_b200 import _lib
scratch = torch.zeros(4, device=dev)
for kind, name in ((0, "fp32_fma_lane_ops_per_s"), (1, "mufu_ex2_lane_ops_per_s"), (2, "fp64_add_lane_ops_per_s"),
                   (3, "fp32_absdiff_add_lane_ops_per_s")):
    v = C.c_double(0.0)
    _lib.check(_lib.lib.eg_issue_peak(kind, 2000, C.byref(v), C.c_void_p(scratch.data_ptr()), None), "eg_issue_peak")
    out[name] = v.value
out["how_issue"] = "eg_issue_peak: 148 x 8 CTAs x 256 threads, 8 independent chains per thread, best of 3 after a warm-up"
print(json.dumps(out))
json.dump(out, open("gpurun_out/measured_peaks_extra.json", "w"), indent=1)
