"""Summarise an .ncu-rep (raw page) into the handful of metrics DESIGN.md / profiles/ quote."""
import csv, subprocess, sys
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr,units,data=rows[0],rows[1],rows[2:]
want=[('Kernel Name','kernel'),('gpu__time_duration.sum','dur'),('sm__cycles_elapsed.avg.per_second','clk'),
('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','tensor%act'),
('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','tensor%el'),
('dram__bytes_read.sum','dram_rd'),('dram__bytes_write.sum','dram_wr'),('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','dram%'),
('lts__t_sector_hit_rate.pct','l2hit%'),('lts__throughput.avg.pct_of_peak_sustained_elapsed','lts%'),
('l1tex__throughput.avg.pct_of_peak_sustained_elapsed','l1tex%'),('launch__registers_per_thread','regs'),('launch__grid_size','grid'),
('launch__cluster_size','cluster'),('sm__warps_active.avg.pct_of_peak_sustained_active','warps%')]
idx=[(hdr.index(k),n) for k,n in want if k in hdr]
print(" | ".join(f"{n}[{units[i]}]" for i,n in idx))
for r in data:
    print(" | ".join((r[i][-28:] if n=='kernel' else r[i][:10]) for i,n in idx))
