"""Turn one `ncu --set full` report of a bench step into the artefacts kept under profiles/:
   <out>_summary.txt  one block per distinct kernel (duration, DRAM traffic, occupancy, top stalls)
   <out>_kernels.json bench-kernel-name -> {dram_bytes_per_launch, duration_us, launches, fp64_pipe_pct, occupancy_pct, registers}
                      (read by bench.py for roofline.traffic and the FP64-pipe figure of the predicate kernels; <out> must be
                      profiles/r02_<workload id> for bench.py to find it)
usage: python tools/ncu_to_profiles.py report.ncu-rep profiles/r02_c2"""
import csv, json, re, subprocess, sys, collections

rep, out = sys.argv[1], sys.argv[2]
exclude = sys.argv[3].split(",") if len(sys.argv) > 3 else []  # kernel-name prefixes kept out of the share column (other legs of bench.py)
# (a `--page raw --csv` export works as input too: the GPU box exports it, the report itself is too big to bring back)
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def num(r, h):
    try:
        return float(r[ix[h]].replace(",", ""))
    except Exception:
        return float("nan")


def scale(h, v):
    u = units[ix[h]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)


def bench_name(n):
    n = n.replace("<unnamed>::", "").replace("rsort::", "").replace("void ", "")
    m = re.match(r"k_onesweep_pass<unsigned int, unsigned int", n)
    if m:
        return "onesweep_pass_u32_kv"
    m = re.match(r"k_onesweep_pass<unsigned long long, unsigned int, (\(bool\))?([01])", n)
    if m:
        return "onesweep_pass_u64_kv" if m.group(2) == "1" else "onesweep_pass_u64_k"
    if n.startswith("k_histogram<unsigned long long"):
        return "sort_histogram_u64"
    if n.startswith("k_histogram<unsigned int"):
        return "sort_histogram_u32"
    m = re.match(r"k_tests<(\(bool\))?([01]), (\(bool\))?([01])>", n)
    if m:
        return "k_tests_%s_%s" % ("exact" if m.group(4) == "1" else "filter", "tri" if m.group(2) == "1" else "poly")
    m = re.match(r"k_(face_bbox|face_codes|planes|tree)<(\(bool\))?([01])>", n)
    if m:
        return "k_%s<%s>" % (m.group(1), "true" if m.group(3) == "1" else "false")
    return re.sub(r"[(<].*", "", n)


agg = collections.OrderedDict()
for r in rows[2:]:
    name = bench_name(r[ix["Kernel Name"]])
    dur = scale("gpu__time_duration.sum", num(r, "gpu__time_duration.sum"))
    rd = scale("dram__bytes_read.sum", num(r, "dram__bytes_read.sum"))
    wr = scale("dram__bytes_write.sum", num(r, "dram__bytes_write.sum"))
    stalls = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): num(r, h) for h in hdr
              if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
    if not stalls:
        stalls = {h.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "")
                   .replace("_per_issue_active.ratio", ""): num(r, h) for h in hdr
                  if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_issue_active.ratio")}
    a = agg.setdefault(name, {"n": 0, "dur": 0.0, "rd": 0.0, "wr": 0.0, "occ": 0.0, "regs": r[ix["launch__registers_per_thread"]],
                              "grid": r[ix["launch__grid_size"]], "inst": 0.0, "fp64": 0.0, "stalls": collections.Counter()})
    a["n"] += 1
    a["dur"] += dur
    a["rd"] += rd
    a["wr"] += wr
    a["occ"] += num(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
    a["inst"] += num(r, "smsp__inst_executed.sum")
    f64 = num(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")
    a["fp64"] += f64 if f64 == f64 else 0.0
    for k, v in stalls.items():
        if v == v:
            a["stalls"][k] += v

lines = ["ncu --set full --clock-control none, one bench step (cold caches, kernels serialised: use SHARES, not absolutes)", ""]
tot = sum(a["dur"] for k, a in agg.items() if not any(k.startswith(x) for x in exclude))
traffic = {}
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["dur"]):
    n = a["n"]
    st = sum(a["stalls"].values()) or 1
    top = ", ".join(f"{k}={v / st:.0%}" for k, v in a["stalls"].most_common(4))
    share = "  n/a " if any(name.startswith(x) for x in exclude) else f"{a['dur'] / tot:5.1%}"
    lines.append(f"{name:24s} x{n:<2d} {a['dur'] / n:7.1f} us/launch  share {share}  dram rd {a['rd'] / n / 1e6:6.1f} MB wr {a['wr'] / n / 1e6:6.1f} MB"
                 f" ({(a['rd'] + a['wr']) / a['dur'] / 1e3:5.0f} GB/s)  occ {a['occ'] / n:4.1f}%  fp64 pipe {a['fp64'] / n:4.1f}%  regs {a['regs']}  grid {a['grid']}  inst {a['inst'] / n / 1e6:5.1f}M")
    lines.append(f"{'':24s} stalls: {top}")
    traffic[name] = {"dram_bytes_per_launch": (a["rd"] + a["wr"]) / n, "duration_us": a["dur"] / n, "launches": n,
                     "fp64_pipe_pct": a["fp64"] / n, "occupancy_pct": a["occ"] / n, "registers": a["regs"]}
lines.append("")
lines.append(f"sum of kernel durations of the resident step (excluding {exclude}): {tot:.1f} us")
open(out + "_summary.txt", "w").write("\n".join(lines) + "\n")
json.dump(traffic, open(out + "_kernels.json", "w"), indent=1, sort_keys=True)
print("\n".join(lines))
