"""Per CUDA source line: shared-memory wavefronts (total / excessive = bank conflicts), global L1 tag requests,
executed warp instructions and stall samples, from an .ncu-rep captured with --set full --import-source on.
usage: ncu_smem_lines.py <rep> [top]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
hdr = None; cur = None; kern = None; curfile = None
agg = collections.defaultdict(lambda: [0.0] * 5); src = {}
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == 'File Path': curfile = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': kern = r[1].split('(')[0]; continue
    if r[0] == 'Line No': hdr = {h: i for i, h in enumerate(r)}; continue
    if r[0].isdigit(): cur = (kern, curfile, int(r[0])); src[cur] = r[1]; continue
    if r[0] == '' and len(r) > 20 and r[2].startswith('0x'):
        def f(name):
            try: return float(r[hdr[name]])
            except (ValueError, KeyError): return 0.0
        a = agg[cur]
        a[0] += f("L1 Wavefronts Shared"); a[1] += f("L1 Wavefronts Shared Excessive"); a[2] += f("L1 Tag Requests Global")
        a[3] += f("Instructions Executed"); a[4] += f("# Samples")
kerns = sorted(set(k[0] for k in agg))
for kn in kerns:
    rows = [(k, v) for k, v in agg.items() if k[0] == kn]
    tot = [sum(v[i] for _, v in rows) for i in range(5)]
    print("== %s: shared wavefronts %.3g (excessive %.3g), global tag requests %.3g, warp inst %.3g, samples %.0f" % (kn, *tot))
    for k, v in sorted(rows, key=lambda kv: -(kv[1][0] + kv[1][2]))[:top]:
        print("%10.3g sh-wf (%9.3g exc) %10.3g gl-req %10.3g inst %6.1f%% smp  %s:%d  %s" % (v[0], v[1], v[2], v[3], 100 * v[4] / max(tot[4], 1), k[1], k[2], src[k].strip()[:90]))
