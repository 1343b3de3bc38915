"""Per-source-line summary of an ncu report's source page: python tools/ncu_lines.py report.ncu-rep [top_n]
(lines ranked by warp-level instructions executed; also stall samples and average active threads)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file, hdr, out, total = None, None, [], 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit():
        if len(r) != len(hdr):                      # a source line with unescaped quotes (inline asm) splits into extra fields: the
            k = len(r) - len(hdr)                   # numeric columns are still the LAST ones, so realign from the end
            r = [r[0], ",".join(r[1:2 + k])] + r[2 + k:]
        d = dict(zip(hdr, r))
        ie = int(float((d.get("Instructions Executed") or "0").replace("-", "0") or 0))
        if ie:
            out.append((ie, cur_file, int(r[0]), r[1].strip()[:110], int(float((d.get("# Samples") or "0").replace("-", "0") or 0)), d.get("Avg. Threads Executed")))
            total += ie
out.sort(reverse=True)
print(f"total warp instructions attributed: {total}")
for ie, f, ln, src, smp, thr in out[:top]:
    print(f"{ie:>12} {100 * ie / total:5.1f}%  smp {smp:>6}  thr {thr:>5}  {f}:{ln}  {src}")
