"""Turn the scratch output of scripts/gpu_round.sh (gpurun_out/) into the tracked evidence under profiles/.

usage: python scripts/make_profiles.py [round_tag]     (default r01)
Needs `ncu` on PATH to read the .ncu-rep captures (works without a GPU)."""
import csv, io, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
PEAK = 6549.8

COLS = [("grid", "launch__grid_size"), ("block", "launch__block_size"), ("regs", "launch__registers_per_thread"),
        ("time_us", "gpu__time_duration.sum"), ("dram_read_MB", "dram__bytes_read.sum"), ("dram_write_MB", "dram__bytes_write.sum"),
        ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("l1tex_pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"), ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("stall_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        ("stall_short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
        ("stall_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
        ("stall_mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
        ("stall_math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
        ("stall_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio")]


def to_unit(val, unit, want):
    v = float(val.replace(",", ""))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
             "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}
    if want == "MB":
        return v * scale.get(unit, 1.0) / 1e6
    if want == "us":
        return v * scale.get(unit, 1.0)
    return v


def ncu_summary(rep, dst, traffic):
    """One row per distinct kernel (its LONGEST captured launch; all are warm), selected columns of the raw page."""
    if not os.path.exists(rep):
        return []
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        return []
    head, units, body = rows[0], rows[1], rows[2:]
    ci = {n: i for i, n in enumerate(head)}
    last = {}
    for r in body:
        name = r[ci["Kernel Name"]].split("(")[0].replace("void ", "").replace("bdet::", "")
        try:
            if name in last and float(r[ci["gpu__time_duration.sum"]].replace(",", "")) <= float(last[name][ci["gpu__time_duration.sum"]].replace(",", "")):
                continue
        except ValueError:
            pass
        last[name] = r
    out = []
    for name, r in last.items():
        row = {"kernel": name}
        for col, metric in COLS:
            if metric not in ci:
                row[col] = ""
                continue
            want = "MB" if col.endswith("_MB") else ("us" if col == "time_us" else "")
            try:
                row[col] = round(to_unit(r[ci[metric]], units[ci[metric]], want), 3)
            except ValueError:
                row[col] = ""
        out.append(row)
        base = name.split("<")[0]
        if row["dram_read_MB"] != "" and row["dram_write_MB"] != "":
            traffic[base] = int((row["dram_read_MB"] + row["dram_write_MB"]) * 1e6)
    with open(dst, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=["kernel"] + [c for c, _ in COLS])
        w.writeheader()
        w.writerows(out)
    return out


def copy(src, dst):
    s = os.path.join(OUT, src)
    if os.path.exists(s):
        shutil.copyfile(s, os.path.join(PROF, "%s_%s" % (TAG, dst)))
        return True
    return False


def last_json_line(path):
    if not os.path.exists(path):
        return None
    for line in reversed(open(path).read().strip().splitlines()):
        line = line.strip()
        if line.startswith("{"):
            try:
                return json.loads(line)
            except ValueError:
                pass
    return None


def main():
    os.makedirs(PROF, exist_ok=True)
    copy("bench_fused.json", "bench_fused.json")
    copy("bench_mat.json", "bench_materialised.json")
    copy("bench_reference.json", "bench_reference.json")
    copy("launches_bench_fused.csv", "launches_bench_fused.csv")
    copy("launches_bench_materialised.csv", "launches_bench_materialised.csv")
    copy("pytest_gpu.log", "pytest_gpu.log")
    copy("perf_all.json", "perf_all.json")
    bw = last_json_line(os.path.join(OUT, "bw_probe.log"))
    if bw:
        json.dump(bw, open(os.path.join(PROF, "%s_bw_probe.json" % TAG), "w"), indent=1)
    traffic = {}
    tpath = os.path.join(PROF, "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    tg = ncu_summary(os.path.join(OUT, "prof_targets.ncu-rep"), os.path.join(PROF, "%s_ncu_target_assignment.csv" % TAG), traffic)
    po = ncu_summary(os.path.join(OUT, "prof_post.ncu-rep"), os.path.join(PROF, "%s_ncu_postprocess_roi.csv" % TAG), traffic)
    json.dump(traffic, open(tpath, "w"), indent=1)

    fused = last_json_line(os.path.join(OUT, "bench_fused.json"))
    mat = last_json_line(os.path.join(OUT, "bench_mat.json"))
    ref = last_json_line(os.path.join(OUT, "bench_reference.json"))
    perf = json.load(open(os.path.join(OUT, "perf_all.json"))) if os.path.exists(os.path.join(OUT, "perf_all.json")) else {}
    L = []
    A = L.append
    A("# profiles/ — round %s evidence (B200, sm_100a)\n" % TAG[1:].lstrip("0"))
    A("Generated by `scripts/make_profiles.py %s` from one `scripts/gpu_round.sh` visit to a B200 through `gpurun`." % TAG)
    A("Timings quoted as *bench* / *events* are CUDA-event measurements outside any profiler; ncu captures are used for DRAM bytes,")
    A("issue activity and stall attribution only (their per-launch times are cold-cache and serialised).\n")
    A("| file | what |\n|---|---|")
    A("| `%s_bench_fused.json` | `python bench.py --steps 2000 --warmup 20` (headline line: fused target assignment) |" % TAG)
    A("| `%s_bench_materialised.json` | `python bench.py --path materialised` (drop-in op sequence: IoU matrix -> Matcher -> encode) |" % TAG)
    A("| `%s_bench_reference.json` | `python bench.py --impl reference` (C restatement of the reference path on the box's host cores) |" % TAG)
    A("| `%s_launches_bench_*.csv` | `ncu --metrics gpu__time_duration.sum --clock-control none` launch list of the same bench command |" % TAG)
    A("| `%s_ncu_target_assignment.csv` | `ncu --set full` summary of assign_main / assign_lq / pairwise / match_colmax / match_lq at config 2 (B=16) |" % TAG)
    A("| `%s_ncu_postprocess_roi.csv` | `ncu --set full` summary of score_filter / select_sort / select_decode / nms_* / roi_align_* (B=2 drivers) |" % TAG)
    A("| `traffic.json` | `dram__bytes_read.sum + dram__bytes_write.sum` per launch from those captures (bench.py copies it into `roofline.traffic`) |")
    A("| `%s_bw_probe.json` | write-only / read-only / copy HBM ceilings of the device (`scripts/bw_probe.py`) |" % TAG)
    A("| `%s_perf_all.json` | per-kernel event timings at every BASELINE config size (`scripts/perf_all.py`) |" % TAG)
    A("| `%s_pytest_gpu.log` | `pytest -m gpu` on the same box |\n" % TAG)
    if fused:
        A("## Headline (config 2: RetinaNet target assignment, 16 images/GPU, A=120 087, G=100)\n")
        A("| arm | images/s | ms/step | note |\n|---|---|---|---|")
        A("| fused `bdet_assign_targets` (value) | %.0f | %.4f | device-resident inputs, per-step CUDA events, L2 flushed between steps |"
          % (fused["value"], fused["ms_per_step"]))
        A("| fused, end to end | %.0f | — | pinned host GT -> H2D -> step -> label census -> D2H |" % fused["e2e"]["value"])
        if mat:
            A("| drop-in op sequence (materialised) | %.0f | %.4f | `Boxes.iou` + `Matcher` + `BoxCoder.encode` kernels |" % (mat["value"], mat["ms_per_step"]))
        if ref and "value" in ref:
            A("| CPU, C restatement, %s threads | %.1f | %.2f | reference arm (`--impl reference`) |"
              % (ref.get("cpu_baseline", {}).get("cores", "?"), ref["value"], ref.get("ms_per_step", 0)))
        A("")
        A("roofline (dominant kernel): `%s`\n" % json.dumps(fused.get("roofline")))
        A("clocks during the timed region: `%s`\n" % json.dumps(fused.get("clocks")))
    if bw:
        A("## HBM ceilings of this device (%s_bw_probe.json)\n\n| probe | GB/s |\n|---|---|" % TAG)
        for k, v in bw.items():
            if isinstance(v, (int, float)):
                A("| %s | %.0f |" % (k, v))
        A("")
    if perf:
        A("## Per-kernel timings (events, `%s_perf_all.json`; peak = %.1f GB/s measured copy)\n" % (TAG, PEAK))
        A("| config | total ms | kernel | avg us | launches/iter | algorithmic GB/s | of measured peak |\n|---|---|---|---|---|---|---|")
        for cfg, d in perf.items():
            for k, v in d["kernels"].items():
                gb = "%.0f" % v["GBps"] if "GBps" in v else ""
                fr = "%.1f %%" % (100 * v["frac_of_measured_peak"]) if "GBps" in v else ""
                A("| %s | %.3f | %s | %.1f | %g | %s | %s |" % (cfg, d["total_ms_median"], k, v["avg_us"], v["launches_per_iter"], gb, fr))
        A("")
    for title, rows in (("target assignment", tg), ("post-processing / ROIAlign", po)):
        if rows:
            A("## ncu --set full, %s (longest warm launch per kernel)\n" % title)
            A("| kernel | time us | DRAM rd MB | DRAM wr MB | dram % | sm % | issue active % | warps active % | top stalls |\n|---|---|---|---|---|---|---|---|---|")
            for r in rows:
                st = sorted(((r[c], c[6:]) for c in r if c.startswith("stall_") and r[c] != ""), reverse=True)[:2]
                A("| %s | %s | %s | %s | %s | %s | %s | %s | %s |" % (r["kernel"], r["time_us"], r["dram_read_MB"], r["dram_write_MB"], r["dram_pct"],
                                                                     r["sm_pct"], r["issue_active_pct"], r["warps_active_pct"],
                                                                     ", ".join("%s %.1f" % (n, v) for v, n in st)))
            A("")
    notes = os.path.join(PROF, "NOTES.md")
    if os.path.exists(notes):
        A(open(notes).read())
    open(os.path.join(PROF, "README.md"), "w").write("\n".join(L) + "\n")
    print("profiles/ refreshed:", sorted(os.listdir(PROF)))


if __name__ == "__main__":
    main()
