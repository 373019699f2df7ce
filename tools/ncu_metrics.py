"""Key metrics of ncu --set full captures.  usage: python tools/ncu_metrics.py rep..."""
import csv, subprocess, sys
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','smsp__inst_executed.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','lts__t_bytes.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__waves_per_multiprocessor','smsp__thread_inst_executed_per_inst_executed.ratio','local_load','lts__t_sectors_op_red.sum','lts__t_sectors_srcunit_tex_op_read.sum','lts__t_sectors_srcunit_tex_op_write.sum']
for rep in sys.argv[1:]:
    txt=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
    rows=list(csv.reader(txt.splitlines())); hdr,units,vals=rows[0],rows[1],rows[2]
    print("==",rep, vals[hdr.index("Kernel Name")][:60])
    for w in want:
        for i,h in enumerate(hdr):
            if h==w or (w=='local_load' and 'local' in h and 'sum' in h and ('ld' in h or 'st' in h) and 'inst_executed' in h): print(f'  {h:78s} {vals[i]:>16s} {units[i]}')
