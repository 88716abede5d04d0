#!/bin/bash
# ncu --set full captures of the kernels other than K1 (one launch each), summarised into gpurun_out/kernel_summaries_<tag>.csv
tag=${1:-rX}
out=gpurun_out; mkdir -p $out
cap() { # name regex skip script args...
  local name=$1 regex=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none -k regex:$regex -s $skip -c 1 -o $out/k_${name}_$tag -f "$@" > $out/k_${name}_$tag.log 2>&1
  ncu -i $out/k_${name}_$tag.ncu-rep --page details --csv > $out/k_${name}_details_$tag.csv 2>/dev/null
}
cap vm_mark        '^vm_mark_kernel'        0 python tools/profile_step.py headline 1 0
cap vm_accumulate  '^vm_accumulate_kernel'  0 python tools/profile_step.py headline 1 0
cap vm_finalize    '^vm_finalize_kernel'    0 python tools/profile_step.py headline 1 0
cap nn1            'nn1_kernel'             0 python tools/profile_step.py headline 1 0
cap nn_scatter     'nn_scatter_kernel'      0 python tools/profile_step.py headline 1 0
cap vg_accumulate  'vg_accumulate_kernel'   0 python tools/profile_step.py headline 1 0
cap unpack_bounds  'unpack_points_bounds'   0 python tools/profile_step.py headline 1 0
cap gicp_cov       'gicp_cov_kernel'        1 python tools/profile_gicp.py 1
cap gicp_inner     'gicp_inner_kernel'      1 python tools/profile_gicp.py 1
cap gicp_corr      'gicp_corr_kernel'       1 python tools/profile_gicp.py 1
cap nn1_far_c4     'nn1_far_kernel'         1 python tools/profile_c4.py
cap nn1_c4         'nn1_kernel'             1 python tools/profile_c4.py
python - <<PY
import csv, glob, os
want = ["Duration", "DRAM Throughput", "Memory Throughput", "L1/TEX Hit Rate", "L2 Hit Rate", "Compute (SM) Throughput", "Achieved Occupancy",
        "Theoretical Occupancy", "Registers Per Thread", "Issue Slots Busy", "No Eligible", "Executed Ipc Active", "Block Size", "Grid Size"]
rows_out = [["kernel", "capture"] + want]
for f in sorted(glob.glob("$out/k_*_details_$tag.csv")):
    rows = list(csv.reader(open(f)))
    if len(rows) < 2: continue
    hdr = rows[0]
    mi, vi, ui, ki = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Kernel Name")
    vals = {}
    for r in rows[1:]:
        if r[mi] in want and r[mi] not in vals: vals[r[mi]] = (r[vi] + " " + r[ui]).strip()
    rows_out.append([rows[1][ki].split("(")[0].replace("void ", "").replace("b200::", "").replace("<unnamed>::", ""), os.path.basename(f)] + [vals.get(w, "") for w in want])
csv.writer(open("$out/kernel_summaries_$tag.csv", "w")).writerows(rows_out)
for r in rows_out: print(" | ".join(r[:1] + r[2:9]))
PY
