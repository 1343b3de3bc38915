"""Disassemble the shipped library (cuobjdump -sass) and write, per kernel of interest, the instructions that prove how it moves data:
TMA (UTMALDG / UBLKCP), mbarrier (SYNCS.*), streaming 128-bit loads (LDG.E...128), LOP3 / POPC / SHF counts, registers.
usage: python tools/sass_excerpts.py [round-tag]   -> profiles/<tag>_sass_excerpts.txt + profiles/<tag>_sass_summary.json (no GPU needed)"""
import json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sola_b200 import _build
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
lib = _build.build()
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
regs = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+)", res):
    regs[m.group(1)] = int(m.group(2))
blocks = re.split(r"\n\s*Function : ", sass)[1:]
KEEP = re.compile(r"UTMALDG|UBLKCP|SYNCS\.|LDG\.E\.[A-Z0-9.]*128|UTMAPF|CCTL|BAR\.SYNC|REDG|ATOMG|LDGSTS|LDS\.128|STS\.128")
want = ("pair_iou_st_ring_kernel", "pair_iou_st_ring_peer_kernel", "fused_pack_resize_kernel", "jf_fused_kernel", "pack_flat_kernel", "raw_counts_vec_kernel",
        "band_pack_generic_kernel", "pair_iou_st_kernel")
summary, sections = {}, []
for b in blocks:
    name = b.split("\n", 1)[0].strip()
    dn = demangle(name)
    short = next((w for w in want if w + "<" in dn or w + "(" in dn), None)
    if short is None:
        continue
    ins = [l.strip() for l in b.splitlines() if "/*" in l and ";" in l]
    ops = [re.sub(r"/\*[0-9a-f]+\*/\s*", "", l).split(";")[0].strip() for l in ins]
    ops = [re.sub(r"^@!?U?P\d+\s+", "", o) for o in ops]
    count = lambda rx: sum(1 for o in ops if re.match(rx, o))
    key = re.sub(r"[^A-Za-z0-9_<>,]", "", re.sub(r"^void\s+", "", dn.split("(")[0]).replace("sola::", "").replace(" ", ""))
    summary[key] = {"mangled": name, "registers": regs.get(name), "instructions": len(ops), "UTMALDG": count(r"UTMALDG"), "UBLKCP": count(r"UBLKCP"),
                    "SYNCS": count(r"SYNCS"), "LDG_128": count(r"LDG\.E\.[A-Z0-9.]*128"), "LDS": count(r"LDS"), "LOP3": count(r"LOP3"),
                    "POPC": count(r"POPC"), "SHF": count(r"SHF"), "BAR": count(r"BAR")}
    keep = [o for o in ops if KEEP.search(o)]
    sections.append(f"## {dn}\n## {len(ops)} instructions, {regs.get(name)} registers; data-movement / synchronisation instructions only:\n"
                    + "\n".join(keep) + "\n")
with open(os.path.join(ROOT, "profiles", f"{tag}_sass_excerpts.txt"), "w") as f:
    f.write(f"# cuobjdump -sass of sola_b200/lib/libsola_maskpath.so, built from csrc digest {_build.source_digest()[:16]} (tools/sass_excerpts.py)\n\n")
    f.write("\n".join(sections))
with open(os.path.join(ROOT, "profiles", f"{tag}_sass_summary.json"), "w") as f:
    json.dump({"library_digest": _build.source_digest(), "kernels": summary}, f, indent=1)
for k, v in summary.items():
    print(k, {a: b for a, b in v.items() if a != "mangled"})
