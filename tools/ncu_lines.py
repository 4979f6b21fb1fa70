"""Aggregate warp-stall samples and executed warp instructions per CUDA source line from an .ncu-rep
(needs -lineinfo + --import-source on).  usage: ncu_lines.py <rep> [top] [samples|inst]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
key = sys.argv[3] if len(sys.argv) > 3 else "samples"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg = collections.defaultdict(float); inst = collections.defaultdict(float); src = {}; cur = None; curfile = None; kern = None
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == 'File Path': curfile = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': kern = r[1]; continue
    if r[0] == 'Line No': continue
    if r[0].isdigit(): cur = (kern, curfile, int(r[0])); src[cur] = r[1]; continue
    if r[0] == '' and len(r) > 7 and r[2].startswith('0x'):
        try: agg[cur] += float(r[4]); inst[cur] += float(r[7])
        except ValueError: pass
bykern = collections.defaultdict(float); ibykern = collections.defaultdict(float)
for k, v in agg.items(): bykern[k[0]] += v; ibykern[k[0]] += inst[k]
sel = agg if key == "samples" else inst
for kn, tot in bykern.items():
    print("==", kn[:60], "samples", tot, "warp instructions", ibykern[kn])
    t = tot if key == "samples" else ibykern[kn]
    for k, v in sorted(((k, v) for k, v in sel.items() if k[0] == kn), key=lambda kv: -kv[1])[:top]:
        print("%10.0f %5.1f%% (samples %5.1f%% inst %5.1f%%) %s:%d  %s" % (v, 100 * v / t, 100 * agg[k] / tot, 100 * inst[k] / ibykern[kn], k[1], k[2], src[k].strip()[:100]))
