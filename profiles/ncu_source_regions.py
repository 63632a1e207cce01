"""Per-instruction warp-state samples of the first kernel in an .ncu-rep (`--set full --import-source on`), grouped by how
often an instruction executed (= which loop nest it sits in), plus the instructions with the most samples.  This is the view
that showed a third of custom::Correlation's warp time sitting in its per-tile epilogue (round 2).

    python profiles/ncu_source_regions.py gpurun_out/x.ncu-rep [top_n [kernel_index]]
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    path = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    print("##", path)
    print("###", rows[0][1][:160])
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    # the page lists the kernels one after the other; a new kernel starts with a "Kernel Name" row
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kernels.append(cur)
        elif cur is not None and len(r) >= len(hdr) and r[0].startswith("0x"):
            cur["rows"].append(r)
    first = kernels[which]["rows"]
    print("### kernel", which, "of", len(kernels), ":", kernels[which]["name"][:150])
    S = lambda r: int(r[idx["# Samples"]])
    E = lambda r: int(r[idx["Instructions Executed"]])
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(S(r) for r in first)
    print(f"instructions {len(first)}, warp samples {tot}, warp instructions executed {sum(E(r) for r in first)}")
    agg = collections.Counter()
    for r in first:
        for s in stalls:
            agg[s] += int(r[idx[s]])
    print("all:", ", ".join(f"{k[6:]} {v}" for k, v in agg.most_common(8)))
    reg = {}
    for r in first:
        d = reg.setdefault(E(r), {"n": 0, "samples": 0, "st": collections.Counter()})
        d["n"] += 1
        d["samples"] += S(r)
        for s in stalls:
            d["st"][s] += int(r[idx[s]])
    print("by execution count (instructions that ran equally often = one loop level):")
    for e, d in sorted(reg.items(), key=lambda x: -x[1]["samples"])[:8]:
        print(f"  executed {e:>9d} x  {d['n']:5d} instr  {d['samples']:6d} samples ({100.0 * d['samples'] / max(tot, 1):4.1f} %)  "
              + ", ".join(f"{k[6:]} {v}" for k, v in d["st"].most_common(5)))
    base = int(first[0][0], 16)
    print("instructions with the most samples:")
    for r in sorted(first, key=S, reverse=True)[:top_n]:
        st = collections.Counter({s: int(r[idx[s]]) for s in stalls if int(r[idx[s]]) > 0})
        print(f"  {int(r[0], 16) - base:#7x}  {r[idx['Source']].strip()[:58]:58s} {S(r):5d} {E(r):9d}  "
              + ", ".join(f"{k[6:]} {v}" for k, v in st.most_common(3)))


if __name__ == "__main__":
    main()
