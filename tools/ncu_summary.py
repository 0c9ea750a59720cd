"""Summarise an `ncu --set full` report into profiles/<name>.json + .txt: per kernel (median over its launches) duration,
DRAM bytes, issue-slot utilisation, registers, top stall reasons.  bench.py reads `traffic` for its roofline line from
the JSON.   python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_ncu_summary"""
import collections, csv, io, json, statistics, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
def num(r, k):
    try:
        return float(r[col[k]].replace(",", ""))
    except Exception:
        return None
def scaled(r, k):          # to base units (bytes, seconds)
    v = num(r, k)
    if v is None:
        return None
    u = units[col[k]]
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1,
                "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1}.get(u, 1)
stall_cols = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and "_per_issue_active" in h]
groups = collections.OrderedDict()
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("biear::", "")
    groups.setdefault(name, []).append(r)
summary = {}
lines = []
for name, rs in groups.items():
    def med(f):
        xs = [x for x in (f(r) for r in rs) if x is not None]
        return statistics.median(xs) if xs else None
    d = {
        "launches_captured": len(rs),
        "duration_us": med(lambda r: scaled(r, "gpu__time_duration.sum")) * 1e6,
        "dram_read_bytes": med(lambda r: scaled(r, "dram__bytes_read.sum")),
        "dram_write_bytes": med(lambda r: scaled(r, "dram__bytes_write.sum")),
        "issue_active_pct": med(lambda r: num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")),
        "warp_instructions": med(lambda r: num(r, "smsp__inst_executed.sum")),
        "registers_per_thread": med(lambda r: num(r, "launch__registers_per_thread")),
        "grid": rs[0][col["launch__grid_size"]], "block": rs[0][col["launch__block_size"]],
        "smem_bank_conflicts": med(lambda r: num(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")),
        # pipe utilisation (per cent of the peak sustained rate while the SM is active)
        "pipe_fma_inst_pct": med(lambda r: num(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active")),
        "pipe_fma_cycles_pct": med(lambda r: num(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active")),
        "pipe_alu_pct": med(lambda r: num(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active")),
        "pipe_xu_pct": med(lambda r: num(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active")),
        "pipe_lsu_pct": med(lambda r: num(r, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active")),
        "smem_wavefronts_pct": med(lambda r: num(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed")),
        "sm_cycles_active_avg": med(lambda r: num(r, "sm__cycles_active.avg")),
        "sm_cycles_elapsed_max": med(lambda r: num(r, "sm__cycles_elapsed.max")),
    }
    d["dram_bytes"] = d["dram_read_bytes"] + d["dram_write_bytes"]
    d["dram_gbs"] = d["dram_bytes"] / (d["duration_us"] * 1e-6) / 1e9
    st = sorted(((statistics.median([num(r, h) or 0 for r in rs]), h) for h in stall_cols), reverse=True)[:5]
    d["top_stalls_per_issue"] = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(v, 3)
                                 for v, h in st}
    summary[name] = d
    lines.append(f"{name:28s} n={d['launches_captured']:2d} {d['duration_us']:9.1f} us  grid {d['grid']:>5s} x {d['block']:>4s}  "
                 f"regs {int(d['registers_per_thread']):3d}  DRAM {d['dram_bytes'] / 1e6:8.2f} MB ({d['dram_gbs']:7.1f} GB/s)  "
                 f"issue {d['issue_active_pct']:5.1f}%  fma-pipe {d['pipe_fma_cycles_pct'] or 0:5.1f}%  xu {d['pipe_xu_pct'] or 0:4.1f}%  "
                 f"smem-wavefronts {d['smem_wavefronts_pct'] or 0:5.1f}%  stalls " + ", ".join(f"{k} {v}" for k, v in d["top_stalls_per_issue"].items()))
json.dump({"source": rep.split("/")[-1], "note": "ncu --set full --clock-control none; per-launch medians; cold-cache, serialised",
           "kernels": summary}, open(out + ".json", "w"), indent=1)
open(out + ".txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
