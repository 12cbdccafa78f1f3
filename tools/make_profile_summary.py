"""Round summary from the ncu outputs of tools/gpu_round.sh:
  python tools/make_profile_summary.py <run dir> <round tag>      e.g.  gpurun_out/final r01
writes profiles/<tag>_summary.md (per-kernel table of the last bench step + the --set full figures of the hot kernels)
and profiles/<tag>_traffic.json (dram bytes per launch of the kernels bench.py may name in its roofline line)."""
import collections
import csv
import json
import os
import re
import sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except (ValueError, AttributeError):
        return float("nan")


def main():
    run, tag = sys.argv[1], sys.argv[2]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out_md = os.path.join(root, "profiles", tag + "_summary.md")
    rows = [r for r in csv.reader(open(os.path.join(run, "launches.csv"))) if len(r) > 14 and r[0].isdigit()]
    ends = [i for i, r in enumerate(rows) if "adam_kernel" in r[4]]
    step = rows[ends[-2] + 1:ends[-1] + 1]
    # bench.py's L2 flush between steps (a 256 MiB uint8 fill) is harness, not the step
    step = [r for r in step if "FillFunctor<unsigned char>" not in r[4] and "adam_advance" not in r[4]] + \
           [r for r in step if "adam_advance" in r[4]][-1:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in step:
        n = re.sub(r"^(pn2::)?(<?unnamed>::)?", "", r[4].replace("void ", "")).split("(")[0][:80]
        agg[n][0] += 1
        agg[n][1] += float(r[14]) / 1e3
    tot = sum(v[1] for v in agg.values())
    lines = ["# %s: kernel breakdown of one bench step (B=32, N=4096; ncu gpu__time_duration, serialised, cold cache)" % tag, "",
             "%d launches, %.3f ms summed kernel time" % (len(step), tot / 1e3), "",
             "| kernel | launches | us/step | share |", "|---|---|---|---|"]
    for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        lines.append("| `%s` | %d | %.1f | %.1f %% |" % (n, v[0], v[1], 100 * v[1] / tot))
    full = os.path.join(run, "full_raw.csv")
    traffic = {}
    if os.path.exists(full):
        raw = list(csv.reader(open(full)))
        hdr = raw[0]
        col = {k: hdr.index(k) for k in ("Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
                                         "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                                         "smsp__issue_active.avg.pct_of_peak_sustained_active") if k in hdr}
        units = raw[1]

        def to_bytes(v, u):
            return num(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

        def to_us(v, u):
            return num(v) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)

        lines += ["", "## `ncu --set full` of the hot kernels (one eager step; per launch)", "",
                  "| kernel | grid | us | DRAM read MB | DRAM write MB | DRAM GB/s | SM thr % | issue active % |", "|---|---|---|---|---|---|---|---|"]
        seen = collections.OrderedDict()
        for r in raw[2:]:
            name = re.sub(r"^(pn2::)?(<?unnamed>::)?", "", r[col["Kernel Name"]].replace("void ", "")).split("(")[0][:60]
            key = (name, r[col["Grid Size"]])
            us = to_us(r[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]])
            rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
            wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
            if key not in seen or us > seen[key][0]:
                seen[key] = (us, rd, wr, num(r[col["sm__throughput.avg.pct_of_peak_sustained_elapsed"]]),
                             num(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]]))
        for (name, grid), (us, rd, wr, smt, iss) in sorted(seen.items(), key=lambda kv: -kv[1][0])[:40]:
            lines.append("| `%s` | %s | %.1f | %.1f | %.1f | %.0f | %.0f | %.0f |" % (name, grid, us, rd / 1e6, wr / 1e6,
                                                                                   (rd + wr) / us / 1e3, smt, iss))
        # kernels bench.py can name as dominant: (ncu name fragment, grid) -> bench key
        wanted = {
            ("wgrad_kernel<1>", "(6, 1, 49)"): "pn2_mlp_gemm_wgrad:131072,384,128,128,384,384",
            ("wgrad_tc_kernel<1>", "(49, 3, 1)"): "pn2_mlp_gemm_wgrad:131072,384,128,128,384,384",
            ("wgrad_tc_kernel<(bool)1>", "(49, 3, 1)"): "pn2_mlp_gemm_wgrad:131072,384,128,128,384,384",
            ("cm_to_rows_bwd_kernel", "(32, 32, 1)"): "pn2_pool_bwd:32,4096,1,384,0,0",
            ("cm_to_rows_bwd_kernel", "(296, 1, 1)"): "pn2_pool_bwd:32,4096,1,384,0,0",  # persistent grid (the largest launch wins below)
        }
        for (name, grid), v in seen.items():
            for (frag, g), key in wanted.items():
                if frag in name and grid == g and (key not in traffic or v[0] > traffic[key]["us_under_ncu"]):
                    traffic[key] = {"dram_bytes": int(v[1] + v[2]), "us_under_ncu": round(v[0], 1),
                                    "source": "profiles/%s_summary.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)" % tag}
    open(out_md, "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open(os.path.join(root, "profiles", tag + "_traffic.json"), "w"), indent=1)
    print(out_md, len(traffic), "traffic entries")


if __name__ == "__main__":
    main()
