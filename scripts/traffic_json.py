"""profiles/<tag>_traffic.csv (ncu --csv metric log written by scripts/gpu_traffic.sh) -> profiles/r2_traffic.json, the static
`roofline.traffic` source bench.py reads.  usage: python scripts/traffic_json.py gpurun_out/traffic_r2_final.csv profiles/r2_traffic.json"""
import csv, json, sys
src, dst = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src)) if len(r) >= 15 and r[0].isdigit()]
m = {}
for r in rows:
    kern = "blind_rotate" if "blind_rotate" in r[4] else "keyswitch"
    m.setdefault(kern, {"name": r[4].split("(")[0].replace("void ", "")})[r[12]] = float(r[14].replace(",", ""))
br, ks = m["blind_rotate"], m.get("keyswitch", {})
out = {
    "gates_per_launch": 65536,
    "captured": "round 2, scripts/gpu_traffic.sh (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum, one launch each), converted by scripts/traffic_json.py",
    "blind_rotate_kernel": br["name"],
    "blind_rotate_dram_bytes_per_launch": br["dram__bytes_read.sum"] + br["dram__bytes_write.sum"],
    "blind_rotate_dram_read": br["dram__bytes_read.sum"],
    "blind_rotate_dram_write": br["dram__bytes_write.sum"],
    "blind_rotate_l2_bytes": br["lts__t_bytes.sum"],
    "blind_rotate_ns_under_ncu": br["gpu__time_duration.sum"],
}
if ks:
    out.update({
        "keyswitch_kernel": ks["name"],
        "keyswitch_dram_bytes_per_launch": ks["dram__bytes_read.sum"] + ks["dram__bytes_write.sum"],
        "keyswitch_l2_bytes": ks["lts__t_bytes.sum"],
        "keyswitch_ns_under_ncu": ks["gpu__time_duration.sum"],
    })
out["note"] = ("static figure from one ncu capture (bench.py does not run under a profiler); the Fourier key (114.7 MB) is re-streamed from DRAM "
               "%.1f times per wave of CTAs in this capture" % (br["dram__bytes_read.sum"] / 114.7e6 / (65536 / 592.0)))
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
