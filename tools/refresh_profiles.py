"""Copy the evidence of the last `tools/gpu_profile.sh` run from gpurun_out/ (scratch) into profiles/ (tracked):
bench line, launch list of the timed region, `ncu --page raw` CSVs of the full captures, and the DRAM traffic of the dominant kernel
that bench.py reports as roofline.traffic.   usage: python tools/refresh_profiles.py [round-tag, default r1]"""
import csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"


def raw_csv(rep, dst):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    lines = [l for l in out.splitlines() if l.startswith('"')]
    with open(dst, "w") as f:
        f.write("\n".join(lines) + "\n")
    rows = list(csv.reader(lines))
    return dict(zip(rows[0], zip(rows[1], rows[2])))


def to_bytes(unit, val):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return int(round(float(val) * scale))


with open(os.path.join(G, "bench.log")) as f:
    line = next(l for l in f if l.startswith("{"))
with open(os.path.join(P, f"{tag}_bench_1gpu.json"), "w") as f:
    f.write(line)
shutil.copy(os.path.join(G, "launches.csv"), os.path.join(P, f"{tag}_launches_timed_region.csv"))
traffic_path = os.path.join(P, "dominant_kernel_traffic.json")
traffic = json.load(open(traffic_path)) if os.path.isfile(traffic_path) else {}
for rep, name, key in (("fused_full.ncu-rep", f"{tag}_fused_ncu_full_raw.csv", "fused_pack_resize_kernel<float>"),
                       ("k2_full.ncu-rep", f"{tag}_k2_ncu_full_raw.csv", None)):
    src = os.path.join(G, rep)
    if not os.path.isfile(src):
        continue
    d = raw_csv(src, os.path.join(P, name))
    if key:
        rd, wr = to_bytes(*d["dram__bytes_read.sum"]), to_bytes(*d["dram__bytes_write.sum"])
        old = traffic.get(key, {})
        traffic[key] = {"source": f"profiles/{name} (ncu --set full --clock-control none, one launch, bench config 2)",
                        "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
                        "algorithmic_bytes_per_launch": old.get("algorithmic_bytes_per_launch"),
                        "gpu_time_ms": float(d["gpu__time_duration.sum"][1]) * ({"us": 1e-3, "ms": 1.0}[d["gpu__time_duration.sum"][0]])}
with open(traffic_path, "w") as f:
    json.dump(traffic, f, indent=1)
print("profiles refreshed:", tag)
