"""Dynamic instruction counts per source line of K3 (encode_kernel<4>): joins the line table of the built object
(nvdisasm -g) with the per-instruction 'Instructions Executed' of an ncu report taken from the same build.
usage: python scripts/k3_line_profile.py gpurun_out/x.ncu-rep [frames] [K]"""
import collections, csv, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3330
kinst = sys.argv[3] if len(sys.argv) > 3 else "3"      # encode_kernel<K>: events per lane of the profiled launch
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "calcium_b200", "_obj", "encode.o")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout.splitlines()
line_of, cur, infun = {}, None, False
for l in dis:
    if l.startswith(".text."):
        infun = f"encode_kernelILi{kinst}E" in l
        continue
    if not infun: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m: line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ia = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
per_line, samp_line, base, total = collections.Counter(), collections.Counter(), None, 0
mism = 0
for r in rows[2:]:
    try: n = int(r[ia])
    except (ValueError, IndexError): continue
    a = int(r[0], 16)
    if base is None: base = a
    ent = line_of.get(a - base)
    if ent is None: mism += 1; continue
    if ent[1].split()[0 if not ent[1].startswith("@") else 1].split(".")[0] != r[1].strip().split()[0 if not r[1].strip().startswith("@") else 1].split(".")[0]: mism += 1
    per_line[ent[0]] += n; samp_line[ent[0]] += int(r[isamp] or 0); total += n
src = {}
print(f"total warp instructions {total}, per frame-row {total / frames / 512:.1f}; opcode mismatches {mism}")
tots = sum(samp_line.values())
for (f, ln), n in sorted(per_line.items(), key=lambda kv: -kv[1])[:70]:
    if f not in src:
        path = os.path.join(ROOT, "calcium_b200", "csrc", f)
        src[f] = open(path).read().splitlines() if os.path.exists(path) else None
    text = src[f][ln - 1].strip()[:110] if src[f] else "(toolkit header)"
    print(f"{n / frames / 512:7.1f} {100 * samp_line[(f, ln)] / tots:5.1f}%  {f}:{ln:<5d} {text}")
