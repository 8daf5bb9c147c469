#!/usr/bin/env python
"""Turn what a gpurun call brought back in gpurun_out/ into the tracked summaries under profiles/:

  final_launches.csv   (ncu --metrics gpu__time_duration.sum --clock-control none ... bench.py)   -> profiles/<tag>_launches_cfg2.csv,
                                                                                                   profiles/kernel_times.json
  final_prof.ncu-rep   (ncu --set full --clock-control none --import-source on ... bench.py)      -> profiles/<tag>_ncu_full_cfg2.txt,
                                                                                                   profiles/traffic.json
  final_bench_*.json, final_pytest_gpu.log, final_parity.txt                                        -> profiles/<tag>_*

  python tools/summarise_profiles.py [tag]        (default tag r1_final; needs `ncu` on PATH to read the .ncu-rep)"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2_final"
# `--selected`: the files written by tools/_gpu_job_r2b.sh (A/B selection, then tests / bench / ncu under the selected switches)
SELECTED = "--selected" in sys.argv
SRC = {"bench": "bench_{w}_selected.json", "launches": "launches_selected.csv", "rep": "prof_selected.ncu-rep", "pytest": "pytest_gpu_selected.log"} if SELECTED else \
      {"bench": "final_bench_{w}.json", "launches": "final_launches.csv", "rep": "final_prof.ncu-rep", "pytest": "final_pytest_gpu.log"}

WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def copy(src, dst):
    if os.path.exists(os.path.join(OUT, src)):
        shutil.copy(os.path.join(OUT, src), os.path.join(PROF, dst))


for w in ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5", "n2", "n4", "n8"):
    copy(SRC["bench"].format(w=w), f"{tag}_bench_{w}.json")
copy("final_parity_tri.txt", f"{tag}_parity_vs_cv2_fft_rows.txt")
copy("final_bench_ref.json", f"{tag}_bench_cfg2_reference.json")
copy(SRC["pytest"], f"{tag}_pytest_gpu.log")
if SELECTED:
    copy("ab.log", f"{tag}_ab_variants.log")
    copy(os.path.join("ab", "results.json"), f"{tag}_ab_variants.json")
    copy(os.path.join("ab", "selected.json"), f"{tag}_ab_selected.json")
copy("final_parity.txt", f"{tag}_parity_vs_cv2.txt")
copy(SRC["launches"], f"{tag}_launches_cfg2.csv")

launches = os.path.join(PROF, f"{tag}_launches_cfg2.csv")
if os.path.exists(launches):
    agg = collections.defaultdict(list)
    for row in csv.DictReader(l for l in open(launches) if not l.startswith("==")):
        agg[row["Kernel Name"]].append(float(row["Metric Value"].replace(",", "")) / 1e6)  # ns -> ms

    def full_size(sub):  # the device-resident step launches the full ROI; the banded host path launches smaller pieces
        v = [max(x) for k, x in agg.items() if sub in k]
        return max(v) if v else None

    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge

    stamp = ge.library_hash(ge.LIB)
    kt = {"library_stamp": stamp,
          "cfg2": {"rhs": full_size("rhs_fold") or full_size("rhs_kernel"), "rows_fwd": full_size("rows_fwd"), "rows_inv": full_size("rows_inv"),
                   "cols": full_size("tri_solve"), "i8_gemm_fwd": full_size("kernel<2, 4"), "i8_gemm_inv": full_size("kernel<4, 3"),
                   "_note": f"gpu__time_duration.sum (ms) of the full-size launch, profiles/{tag}_launches_cfg2.csv (ncu --metrics gpu__time_duration.sum --clock-control none: "
                            "cold cache, serialised); cols = the tridiagonal solve kernel alone (tri_solve_smem_kernel / tri_solve_kernel); valid for the library whose source hash is library_stamp (bench.py drops them otherwise)"}}
    json.dump(kt, open(os.path.join(PROF, "kernel_times.json"), "w"), indent=1)
    total = sum(sum(v) for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{100 * sum(v) / total:6.2f} %  n={len(v):3d}  mean {1e3 * sum(v) / len(v):8.2f} us  max {1e3 * max(v):8.2f} us  {k[:70]}")

rep = os.path.join(OUT, SRC["rep"])
if os.path.exists(rep) and shutil.which("ncu"):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge

    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = [f"# ncu --set full --clock-control none, bench.py cfg2 (ROI 1810x1339), default engine (INT8 tensor-core DST along x + tridiagonal solve along y); "
           f"report gpurun_out/{SRC['rep']} (not committed)"]
    seen, traffic = set(), {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if name in seen:
            continue
        seen.add(name)
        out.append("\n" + name)
        for w in WANT:
            if w in hdr:
                out.append(f"  {w:70s} {r[hdr.index(w)]} {units[hdr.index(w)]}")
        stalls = [(num(r[i]), h) for i, h in enumerate(hdr) if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio") and num(r[i]) is not None]
        for v, h in sorted(stalls, reverse=True)[:5]:
            out.append(f"  stall {v:7.3f} {h.split('stalled_')[1].split('_per_issue')[0]}")
        i, j = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        traffic[name.split("(")[0]] = num(r[i]) * SCALE[units[i]] + num(r[j]) * SCALE[units[j]]  # (template arguments kept: i8_gemm_pkernel<2, 4, ..> vs <4, 3, ..>)
    open(os.path.join(PROF, f"{tag}_ncu_full_cfg2.txt"), "w").write("\n".join(out) + "\n")
    pick = lambda sub: next((v for k, v in traffic.items() if sub in k), None)
    t = {"library_stamp": ge.library_hash(ge.LIB),
         "cfg2": {"rows_fwd": pick("rows_fwd"), "rows_inv": pick("rows_inv"), "cols": pick("tri_solve"), "rhs": pick("rhs_fold") or pick("rhs_kernel"),
                  "i8_gemm_fwd": pick("kernel<2, 4"), "i8_gemm_inv": pick("kernel<4, 3"),
                  "_note": f"dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full of bench.py cfg2, profiles/{tag}_ncu_full_cfg2.txt "
                           "(writes mostly stay in the 126 MB L2 during ncu's kernel replay); cols = the tridiagonal solve kernel"}}
    json.dump(t, open(os.path.join(PROF, "traffic.json"), "w"), indent=1)

for w in ("cfg2", "cfg1", "cfg5", "cfg4", "cfg3"):
    f = os.path.join(PROF, f"{tag}_bench_{w}.json")
    if os.path.exists(f):
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])  # (NCCL prints its version line to stdout first)
        print(w, "Mpix/s", round(d["value"]), "ms", round(d["ms_per_step"], 4), "| e2e Mpix/s", round(d["e2e"]["value"]), "ms", round(d["e2e"]["ms_per_step"], 4),
              "| cpu", round((d.get("cpu_baseline") or {}).get("value") or 0, 2), "| roofline", (d.get("roofline") or {}).get("frac"), (d.get("roofline_stencil") or {}).get("frac_ncu"))
