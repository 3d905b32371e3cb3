"""Summarise an ncu --csv launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch)
into the per-kernel table the bench's `roofline.traffic` cites.

    python tools/summarise_launches.py gpurun_out/fin1_launches.csv 10000000 "<command>" > profiles/r02_n_moment_traffic.json
"""
import collections
import csv
import json
import re
import sys


def main():
    path, samples, command = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    table_bytes = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0      # bytes per sample pcq_mom_table_kernel writes
    hdr, per = None, collections.OrderedDict()
    for r in csv.reader(open(path, errors="replace")):
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "").strip()
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        k = per.setdefault((d["ID"], name), {})
        if d["Metric Name"] == "gpu__time_duration.sum":
            k["ms"] = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[u]
        else:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            k["rd" if "read" in d["Metric Name"] else "wr"] = v * scale
    agg = collections.OrderedDict()
    for (_, name), k in per.items():
        a = agg.setdefault(name, {"n": 0, "ms": 0.0, "rd": 0.0, "wr": 0.0})
        a["n"] += 1
        for f in ("ms", "rd", "wr"):
            a[f] += k.get(f, 0.0)
    main_k = {n: a for n, a in agg.items() if n.startswith("pcq_mom_kernel") or n.startswith("pcq_mom_ws_kernel") or n.startswith("pcq_syrk_kernel")}
    main_bytes = sum(a["rd"] + a["wr"] for a in main_k.values())
    all_bytes = sum(a["rd"] + a["wr"] for a in agg.values())
    total_ms = sum(a["ms"] for a in agg.values())
    processed = None
    if table_bytes and "pcq_mom_table_kernel" in agg:
        processed = agg["pcq_mom_table_kernel"]["wr"] / table_bytes      # samples all captured calls went over
    out = {
        "command": command,
        "note": "one ncu pass over every pcq_* launch of the command (the three warm-up passes, the timed pass and the end-to-end calls); "
                "bytes and times are sums over ALL those launches, per-sample figures divide by the samples they processed",
        "samples_per_step": samples,
        "launches": agg,
        "total_ms_serialised": total_ms,
        "share_of_time": {n: a["ms"] / total_ms for n, a in agg.items()},
        "dram_bytes_main_kernel_all_calls": main_bytes,
        "dram_bytes_all_kernels_all_calls": all_bytes,
    }
    if processed:
        out["samples_processed_by_the_captured_calls"] = processed
        out["samples_note"] = "DRAM bytes written by pcq_mom_table_kernel / its %g B per sample" % table_bytes
        out["dram_bytes_per_sample_main_kernel"] = main_bytes / processed
        out["dram_bytes_per_sample_all_kernels"] = all_bytes / processed
        out["dram_bytes_per_step_main_kernel"] = main_bytes / processed * samples
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
