"""Does sustained throughput drift with time under load (thermal / power-averaging)?  Cycles the four kernel
families, 1 s each, for ~70 s and prints TF/s + NVML clock/power/temperature."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, pynvml
from hint_sweep import FL, fam_fn  # noqa: E402

pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
t_start = time.time()
order = sys.argv[1:] or ["FWD", "DZ", "DH", "DW"]
for rep in range(int(os.environ.get("REPS", 16))):
    for fam in order:
        fn = fam_fn(fam)
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 0; t0 = time.time(); e0.record()
        while time.time() - t0 < 1.0:
            for _ in range(4):
                fn(); n += 1
            torch.cuda.synchronize()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        clk = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000
        tmp = pynvml.nvmlDeviceGetTemperature(h, pynvml.NVML_TEMPERATURE_GPU)
        print(f"t={time.time()-t_start:5.1f}s {fam:3s} {ms:6.2f} ms {FL/ms/1e9:5.0f} TF/s  clk {clk} MHz {pw:.0f} W {tmp} C", flush=True)
