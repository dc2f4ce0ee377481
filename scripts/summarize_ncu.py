"""Turn the raw ncu outputs that come back in gpurun_out/ into the tracked summaries under profiles/.
  python scripts/summarize_ncu.py launches gpurun_out/launches_r01b.csv profiles/launches_r01_summary.csv "<command>"
  python scripts/summarize_ncu.py report gpurun_out/prof_k1_r01b.ncu-rep profiles/k1_fm_eval_r01.md "<title>"
The report mode shells out to `ncu -i <rep> --page raw --csv` (ncu is in this image; no GPU needed to read a report)."""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "smsp__inst_executed_op_tma_ld.sum", "launch__occupancy_limit_registers",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum"]


def short(name):
    name = re.sub(r"\(.*\)$", "", name)
    name = name.replace("pxr::", "").replace("pxr_chol2::", "").replace("pxr_chol::", "")
    return name.strip()


def launches(src, dst, command):
    rows = []
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)
        rows.append((short(r["Kernel Name"]), ms))
    agg = OrderedDict()
    for k, ms in rows:
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ms
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list — `%s`\n" % command)
        f.write("# B200 (sm_100a). Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n")
        f.write("# %d launches captured, %.3f ms total (includes one-off setup kernels: synthetic patch generator, reference extraction)\n" % (len(rows), total))
        f.write("kernel,launches,total_ms,avg_ms,share_pct\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%s,%d,%.4f,%.5f,%.2f\n" % (k, n, ms, ms / n, 100 * ms / total))
    print("wrote", dst, len(rows), "launches")


def report(src, dst, title):
    # src: a .ncu-rep, or the raw-page CSV made from one on the GPU box (`ncu -i x.ncu-rep --page raw --csv`)
    if src.endswith(".csv"):
        out = open(src).read()
    else:
        out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    header, units = rd[0], rd[1]
    with open(dst, "w") as f:
        f.write("# %s\n\n" % title)
        for li, row in enumerate(rd[2:]):
            d = dict(zip(header, row)); u = dict(zip(header, units))
            f.write("## launch %d: `%s`  grid %s x block %s\n\n| metric | value | unit |\n|---|---|---|\n" %
                    (li, short(d.get("Kernel Name", "?")), d.get("launch__grid_size", "?"), d.get("launch__block_size", "?")))
            for k in KEYS:
                if k in d and d[k] != "":
                    f.write("| %s | %s | %s |\n" % (k, d[k], u.get(k, "")))
            stalls = sorted(((float(d[k].replace(",", "")), k) for k in d if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and d[k] not in ("", "n/a")), reverse=True)[:6]
            if stalls:
                f.write("\ntop warp-stall reasons (per issue-active): " + ", ".join("%s %.2f" % (k.split("stalled_")[1].split("_per_issue")[0], v) for v, k in stalls) + "\n")
            ops = sorted(((float(d[k].replace(",", "")), k) for k in d if k.startswith("smsp__sass_inst_executed_op_") or k.startswith("sm__sass_inst_executed_op_")
                          if d[k] not in ("", "n/a")), reverse=True)[:8]
            if ops:
                f.write("\n" + ", ".join("%s %.3g" % (k.split("op_")[1], v) for v, k in ops) + "\n")
            f.write("\n")
    print("wrote", dst)


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        report(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else sys.argv[2])
