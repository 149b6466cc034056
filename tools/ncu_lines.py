#!/usr/bin/env python
"""Per-source-line hot spots from an ncu report:  python tools/ncu_lines.py rep.ncu-rep [topN]
Uses `ncu --page source --csv --print-source cuda,sass` (needs -lineinfo at compile time)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; data = []; fpath = ""
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fpath = r[1].split("/")[-1]
    if len(r) > 10 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-": continue
    ix = {h: i for i, h in enumerate(hdr)}
    def g(name):
        try: return int(r[ix[name]])
        except Exception: return 0
    data.append(dict(file=fpath, line=int(r[0]), src=r[1].strip()[:100], samp=g("# Samples"), inst=g("Instructions Executed"),
                     wf=g("L1 Wavefronts Shared"), wfx=g("L1 Wavefronts Shared Excessive"),
                     bar=g("stall_barrier"), mio=g("stall_mio"), lsb=g("stall_long_sb"), ssb=g("stall_short_sb"),
                     math=g("stall_math"), wait=g("stall_wait")))
ts = sum(d["samp"] for d in data) or 1; ti = sum(d["inst"] for d in data) or 1; tw = sum(d["wf"] for d in data) or 1
print(f"total samples {ts}  warp-instructions {ti}  shared wavefronts {tw} (excess {sum(d['wfx'] for d in data)})")
print("line  samp%  inst%  wf%  wfx%  | bar mio lsb ssb math | source")
for d in sorted(data, key=lambda x: -x["samp"])[:top]:
    print(f"{d['file'][:14]}:{d['line']:4d} {100*d['samp']/ts:5.1f} {100*d['inst']/ti:5.1f} {100*d['wf']/tw:5.1f} {100*d['wfx']/tw:5.1f} |"
          f" {d['bar']:5d} {d['mio']:5d} {d['lsb']:5d} {d['ssb']:5d} {d['math']:5d} | {d['src']}")
