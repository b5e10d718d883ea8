#!/usr/bin/env python3
"""Per-kernel summary of an ncu --set full capture of ONE CHAIN12 pass (tools/r02_profiles.sh) + the pass's DRAM traffic.
usage: tools/ncu_pass.py gpurun_out/TAG_chain12.ncu-rep profiles/TAG_chain12_pass.txt [profiles/kchain_traffic.json]"""
import csv, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]
names = ["k_chain #1 (noise_gate, compressor)", "k_os (excess, fuzz, overdrive, distortion; x4)", "k_chain #2 (tone_stack, chorus, delay, reverb, power_amp, cabinet)"]
lines = ["# ncu --set full --clock-control none: the three segment kernels of ONE CHAIN12 pass (1024 channels x 8192 samples, 192 kHz), GDG_WAYS=1",
         "# %s (tools/r02_profiles.sh); units as printed by ncu" % rep, ""]
tot_r = tot_w = 0.0
def mb(v, u): return float(v.replace(",", "")) * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u]
for i, r in enumerate(rows[2:5]):
    d = dict(zip(hdr, r))
    lines.append("== %s   [%s, grid %s, block %s]" % (names[i] if i < 3 else "?", d["Kernel Name"].split("(")[0], d["Grid Size"], d["Block Size"]))
    for k in KEYS:
        if k in d: lines.append("  %-78s %-16s %s" % (k, units[hdr.index(k)], d[k]))
    st = sorted([(float(d[h]), h) for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and d[h]], reverse=True)[:5]
    lines.append("  top stalls per issue: " + ", ".join("%s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for v, h in st))
    lines.append("")
    tot_r += mb(d["dram__bytes_read.sum"], units[hdr.index("dram__bytes_read.sum")]); tot_w += mb(d["dram__bytes_write.sum"], units[hdr.index("dram__bytes_write.sum")])
lines.append("pass total: dram read %.1f MB + write %.1f MB = %.1f MB (algorithmic 134.2 MB; unfused byte model 3 x 134.2 MB)" % (tot_r, tot_w, tot_r + tot_w))
open(out, "w").write("\n".join(lines) + "\n")
if len(sys.argv) > 3:
    json.dump({"source": rep, "dram_bytes_read": tot_r * 1e6, "dram_bytes_write": tot_w * 1e6, "dram_bytes_per_pass": (tot_r + tot_w) * 1e6,
               "kernels": 3, "channels": 1024, "block": 8192}, open(sys.argv[3], "w"), indent=1)
print("\n".join(lines))
