#!/usr/bin/env python
"""Summarise an .ncu-rep (from `ncu --set full`) into profiles/<name>.md: key raw metrics per kernel
launch and the top stall instructions from the source page.  Run here (no GPU needed)."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "launch__cluster_max_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct"]
STALL = "smsp__average_warps_issue_stalled_"


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {}
        for i, h in enumerate(hdr):
            if h in KEYS or h == "Kernel Name" or (h.startswith(STALL) and h.endswith("_per_issue_active.ratio")):
                d[h] = (r[i], units[i])
        res.append(d)
    return res


def source(rep, top=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": [], "hdr": None}
            blocks.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    res = []
    for b in blocks[:1]:
        h = b["hdr"]
        iS, iSrc, iEx = h.index("Warp Stall Sampling (All Samples)"), h.index("Source"), h.index("Instructions Executed")
        data = [(int(r[iS] or 0), r[iSrc].strip(), r[iEx]) for r in b["rows"]]
        tot = sum(d[0] for d in data) or 1
        topi = sorted(range(len(data)), key=lambda k: -data[k][0])[:top]
        res.append((b["name"], tot, [(k,) + data[k] for k in sorted(topi)]))
    return res


def main():
    rep, outp = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    with open(outp, "w") as f:
        f.write("# %s\n\nSource: `%s` (`ncu --set full --clock-control none --import-source on`). "
                "Times under ncu are serialised and cold-cache; never bench values.\n\n" % (title, rep))
        for d in raw(rep):
            f.write("## %s\n\n| metric | value | unit |\n|---|---|---|\n" % d.get("Kernel Name", ("?",))[0][:80])
            for k in KEYS:
                if k in d:
                    f.write("| %s | %s | %s |\n" % (k, d[k][0], d[k][1]))
            f.write("\nWarp stall reasons (per issue-active):\n\n")
            st = sorted(((float(v[0].replace(",", "") or 0), k) for k, v in d.items() if k.startswith(STALL)), reverse=True)
            for v, k in st[:8]:
                f.write("- %s: %.2f\n" % (k[len(STALL):-len("_per_issue_active.ratio")], v))
            f.write("\n")
        for name, tot, rows in source(rep):
            f.write("## Top stall instructions (%d samples) -- %s\n\n| idx | samples | %% | executed | SASS |\n|---|---|---|---|---|\n" % (tot, name[:60]))
            for k, s, src, ex in rows:
                f.write("| %d | %d | %.1f | %s | `%s` |\n" % (k, s, 100.0 * s / tot, ex, src[:100]))


if __name__ == "__main__":
    main()
