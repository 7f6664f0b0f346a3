"""Selected raw metrics of one profiled launch -> the text summary committed under profiles/.
Usage: python tools/ncu_summary.py gpurun_out/prof_k_step_fused.ncu-rep [kernel-regex] > profiles/r02_k_step_fused_ncu_summary.txt
(runs `ncu -i <rep> --page raw --csv`; works on the CPU box, no GPU needed)"""
import csv, io, re, subprocess, sys

KEEP = re.compile(r"^(dram__bytes_(read|write)\.sum|dram__cycles_active|gpu__dram_throughput|gpu__time_duration\.sum|"
                  r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|"
                  r"launch__(block_size|grid_size|registers_per_thread|shared_mem_per_block_dynamic)|"
                  r"sass__inst_executed_local_(loads|stores)|sm__cycles_elapsed\.max|sm__inst_executed_pipe_(alu|fma|fmaheavy|lsu|xu|uniform)\.avg\.pct|"
                  r"sm__inst_executed\.avg\.per_cycle_elapsed|sm__instruction_throughput|sm__pipe_tensor_cycles_active|sm__pipe_tensor_op|sm__warps_active|"
                  r"smsp__average_warp|smsp__inst_executed\.sum|smsp__issue_active\.avg\.pct|lts__t_bytes\.sum|lts__t_sector_hit_rate|"
                  r"l1tex__t_bytes_pipe_lsu_mem_global_op_ld\.sum|smsp__cycles_active\.avg)")

rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if pat and not pat.search(name):
        continue
    print(f"# kernel: {name}   grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
    for h, u, v in zip(hdr, units, r):
        if KEEP.match(h):
            print(f"{h} [{u}] = {v}")
    break
