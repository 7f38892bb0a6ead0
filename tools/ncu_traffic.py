"""profiles/traffic.json from the round's `ncu --set full` captures (the reduced per-launch CSVs tools/ncu_reduce.py writes):
    python tools/ncu_traffic.py profiles/<trunk>.csv profiles/<lbs>.csv
trunk: dram read + write bytes summed over every trunk launch of ONE 128-image forward (the CSV must hold exactly one forward);
lbs: dram read + write bytes of the vertex kernel's launch at B = 8192.  bench.py reads the file for `roofline.traffic` /
`roofline_lbs.traffic`; without it both are null (never a stale constant)."""
import csv, json, os, sys

def col(hdr, name):
    return next(i for i, h in enumerate(hdr) if h.startswith(name))

def unit_scale(h):
    u = h[h.index("[") + 1:h.index("]")].lower()
    return {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]

def load(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    ir, iw, it, ik = col(hdr, "dram__bytes_read.sum"), col(hdr, "dram__bytes_write.sum"), col(hdr, "gpu__time_duration.sum"), col(hdr, "Kernel Name")
    out = []
    for r in rows[1:]:
        out.append((r[ik], float(r[it]), float(r[ir]) * unit_scale(hdr[ir]) + float(r[iw]) * unit_scale(hdr[iw])))
    return out

trunk, lbs = load(sys.argv[1]), load(sys.argv[2])
tr = [r for r in trunk if "avgpool" not in r[0]]
vk = max((r for r in lbs if "vertex" in r[0]), key=lambda r: r[1])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {"trunk": {"dram_bytes_per_128_images": sum(r[2] for r in tr), "launches": len(tr), "serialized_us": sum(r[1] for r in tr),
                 "source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum over the %d trunk launches of one 128-image forward (%s; cold-cache, serialised)" % (len(tr), os.path.relpath(sys.argv[1], root))},
       "lbs": {"dram_bytes_per_launch": vk[2], "kernel_us": vk[1],
               "source": "ncu --set full, dram read + write of %s at B=8192 (%s)" % (vk[0].split("(")[0][-40:], os.path.relpath(sys.argv[2], root))}}
json.dump(out, open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
