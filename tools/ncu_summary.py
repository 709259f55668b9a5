import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
r=list(csv.reader(out.splitlines()))
hdr=r[0]
want=["Kernel Name","gpu__time_duration.sum","launch__grid_size","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","sm__throughput.avg.pct_of_peak_sustained_elapsed","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","dram__bytes_read.sum","dram__bytes_write.sum","lts__t_bytes.sum","lts__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__throughput.avg.pct_of_peak_sustained_elapsed","sm__issue_active.avg.pct_of_peak_sustained_elapsed","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","sm__cycles_elapsed.max","smsp__inst_executed.sum","sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active","lts__t_sector_hit_rate.pct","sm__cycles_active.avg"]
idx={h:i for i,h in enumerate(hdr)}
for row in r[2:]:
    print("----")
    for w in want:
        if w in idx: print(f"  {w:70s} {r[1][idx[w]]:12s} {row[idx[w]]}")
