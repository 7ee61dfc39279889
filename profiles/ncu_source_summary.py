"""Summarises `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda` per CUDA source line."""
import csv
import sys


def f(x):
    try:
        return float(x.replace(',', ''))
    except ValueError:
        return 0.0


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hdr, cur_file, data = None, None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[0] != "":
            d = {h: v for h, v in zip(hdr, r) if h not in ("Address",)}
            d["line"], d["src"], d["file"] = r[0], r[1], cur_file
            d["inst"] = f(r[hdr.index("Instructions Executed")])
            d["tinst"] = f(r[hdr.index("Thread Instructions Executed")])
            d["samples"] = f(r[hdr.index("# Samples")])
            d["noinst"] = f(r[hdr.index("stall_no_inst")])
            d["barrier"] = f(r[hdr.index("stall_barrier")])
            d["conf"] = f(r[hdr.index("L1 Wavefronts Shared Excessive")])
            data.append(d)
    tot = sum(d["inst"] for d in data) or 1
    ts = sum(d["samples"] for d in data) or 1
    print(f"total warp instructions {tot:.3e}, samples {ts:.0f}")
    for d in sorted(data, key=lambda d: -d["inst"])[:top]:
        eff = d["tinst"] / d["inst"] / 32 if d["inst"] else 0
        print(f"{(d['file'] or '').split('/')[-1][:20]:20s} L{d['line']:>4s} inst {100 * d['inst'] / tot:5.1f}%  lanes {100 * eff:4.0f}%  "
              f"samples {100 * d['samples'] / ts:5.1f}%  noinst {d['noinst']:6.0f} bar {d['barrier']:6.0f} smemX {d['conf']:9.0f} | {d['src'].strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
