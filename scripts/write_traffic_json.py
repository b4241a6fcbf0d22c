"""Writes profiles/<round>_traffic.json from the ncu --set full exports of scripts/gpu_final_profiles.sh:
dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernels, with the commit they were taken at
(bench.py reads `roofline.traffic` from this file instead of carrying literals)."""
import csv
import json
import subprocess
import sys

rnd = sys.argv[1] if len(sys.argv) > 1 else "r02"


def metrics(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def to_bytes(v, u):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
    return float(v) * mult


def entry(path):
    m = metrics(path)
    rd, wr = to_bytes(*m["dram__bytes_read.sum"]), to_bytes(*m["dram__bytes_write.sum"])
    dur = m["gpu__time_duration.sum"]
    return dict(kernel=m["Kernel Name"][0][:120], dram_bytes=rd + wr, dram_read_bytes=rd, dram_write_bytes=wr,
                duration=f"{dur[0]} {dur[1]}", source=path)


out = dict(commit=subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip(),
           how="ncu --set full --clock-control none, one launch each, bench.py workloads at scale 1.0 on one B200")
out["tc_eval_kernel_main"] = entry(f"gpurun_out/prof_{rnd}_tc_eval_main_raw.csv")
out["rescore_lists_kernel"] = entry(f"gpurun_out/prof_{rnd}_rescore_lists_raw.csv")
rows_k, hub_k = entry(f"gpurun_out/prof_{rnd}_spmm_rows_raw.csv"), entry(f"gpurun_out/prof_{rnd}_spmm_hub_chunk_raw.csv")
out["spmm_rows_kernel"], out["spmm_hub_chunk_kernel"] = rows_k, hub_k
out["spmm_layer"] = dict(dram_bytes=rows_k["dram_bytes"] + hub_k["dram_bytes"], note="rows kernel + hub-chunk kernel of one layer "
                         "(the hub-finish kernel moves < 0.3 GB and was not captured)")
json.dump(out, open(f"profiles/{rnd}_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
