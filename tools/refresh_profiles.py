"""Copy the evidence of the last `tools/gpu_profile_r2.sh` run from gpurun_out/<tag>/ (scratch) into profiles/ (tracked):
the launch list of the bench's timed region, the `ncu --page raw` CSV of every full capture, a compact summary of all captures
(`<tag>_ncu_summary.json`), and `dominant_kernel_traffic.json` — the DRAM traffic bench.py reports as roofline.traffic, stamped with the
kernel's register count so that a stale entry is detectable (tests/test_profiles_consistency.py compares it with the built library, and
this script FAILS if a capture's register count differs from the library's).
usage: python tools/refresh_profiles.py [round-tag, default r2] [source directory under gpurun_out/, default = the tag]"""
import csv, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
G, P = os.path.join(ROOT, "gpurun_out", sys.argv[2] if len(sys.argv) > 2 else tag), os.path.join(ROOT, "profiles")

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "launch__registers_per_thread": "registers", "launch__grid_size": "grid", "launch__block_size": "block",
    "launch__occupancy_limit_registers": "occ_limit_regs", "launch__occupancy_limit_shared_mem": "occ_limit_smem",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefronts_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_instruction",
}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def _norm(name):
    """c++filt prints bool template arguments as words, ncu as 0 / 1."""
    return re.sub(r"\bfalse\b", "0", re.sub(r"\btrue\b", "1", name))


def library_registers():
    from sola_b200 import _build
    res = subprocess.run(["cuobjdump", "-res-usage", _build.build()], capture_output=True, text=True).stdout
    out = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+)", res):
        dn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        out[_norm(re.sub(r"^void\s+", "", dn).split("(")[0])] = int(m.group(2))
    return out


def summarise(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    lines = [l for l in out.splitlines() if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, zip(units, vals)))
    s = {"kernel": d["Kernel Name"][1]}
    for k, name in KEYS.items():
        if k in d:
            u, v = d[k]
            try:
                x = float(v.replace(",", ""))
            except ValueError:
                continue
            s[name] = x * SCALE[u] if u in SCALE else x
    if "duration" in s:
        s["duration_ms"] = s.pop("duration")
    stalls = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(v[1])
              for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")}
    s["stalls_per_issue"] = {k: round(v, 3) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:6] if k != "selected"}
    return lines, s


def main():
    os.makedirs(P, exist_ok=True)
    regs = library_registers()
    summary, bad = {}, []
    for f in sorted(os.listdir(G)):
        if not f.endswith(".ncu-rep"):
            continue
        name = f[:-len(".ncu-rep")]
        lines, s = summarise(os.path.join(G, f))
        with open(os.path.join(P, f"{tag}_ncu_{name}_raw.csv"), "w") as out:
            out.write("\n".join(lines) + "\n")
        base = _norm(re.sub(r"^void\s+", "", s["kernel"]).split("(")[0])
        lib_regs = next((r for k, r in regs.items() if k.replace("sola::", "") == base.replace("sola::", "")), None)
        s["library_registers"] = lib_regs
        if lib_regs is not None and "registers" in s and int(s["registers"]) != lib_regs:
            bad.append(f"{name}: captured with {int(s['registers'])} registers, the built library has {lib_regs} ({base})")
        summary[name] = s
    with open(os.path.join(P, f"{tag}_ncu_summary.json"), "w") as f:
        json.dump(summary, f, indent=1)
    if os.path.isfile(os.path.join(G, "launches.csv")):
        shutil.copy(os.path.join(G, "launches.csv"), os.path.join(P, f"{tag}_launches_timed_region.csv"))
    traffic = {}
    for key, cap in (("fused_pack_resize_kernel<float>", "fused_f32_720p"), ("jf_fused_kernel", "jf_region"), ("jf_fused_kernel[boundary]", "jf_boundary")):
        if cap in summary:
            s = summary[cap]
            traffic[key] = {"source": f"profiles/{tag}_ncu_{cap}_raw.csv (ncu --set full --clock-control none, one launch)", "kernel": s["kernel"].split("(")[0],
                            "registers": int(s.get("registers", 0)), "dram_bytes_read": int(s.get("dram_read", 0)), "dram_bytes_write": int(s.get("dram_write", 0)),
                            "dram_bytes_per_launch": int(s.get("dram_read", 0) + s.get("dram_write", 0)), "gpu_time_ms": s.get("duration_ms")}
    if traffic:
        with open(os.path.join(P, "dominant_kernel_traffic.json"), "w") as f:
            json.dump(traffic, f, indent=1)
    for name, s in summary.items():
        print(f"{name:22s} {s['kernel'].split('(')[0][:48]:48s} {s.get('duration_ms', 0) * 1e3:9.1f} us  regs {int(s.get('registers', 0)):3d}  "
              f"issue {s.get('issue_active_pct', 0):5.1f}%  alu {s.get('pipe_alu_pct', 0):5.1f}%  xu {s.get('pipe_xu_pct', 0):5.1f}%  "
              f"dram {s.get('dram_throughput_pct', 0):5.1f}%  smem {s.get('smem_wavefronts_pct', 0):5.1f}%")
    if bad:
        raise SystemExit("STALE CAPTURES (rebuild or re-capture):\n  " + "\n  ".join(bad))
    print("profiles refreshed:", tag)


if __name__ == "__main__":
    main()
