"""Turn one measurement directory under gpurun_out/ into the tracked artefacts under profiles/.
Usage: python tools/make_profiles.py gpurun_out/r1g r1     (reads bench_*.json, bench_launches.csv, iter_v.ncu-rep)"""
import collections, csv, json, os, re, shutil, subprocess, sys

src, tag = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def last_json(path):
    for l in reversed(open(path).read().splitlines()):
        if l.startswith("{"):
            return json.loads(l)
    return None


names = {"n1": "bench_n1", "seq": "bench_sequence_n1", "4k": "bench_4k_n1", "vga": "bench_vga_n1", "lk": "bench_lk_n1",
         "ref": "bench_reference_arm", "n2": "bench_n2", "n4": "bench_n4", "n8": "bench_n8"}
for k, v in names.items():
    f = os.path.join(src, "bench_%s.json" % k)
    if os.path.exists(f) and last_json(f):
        shutil.copy(f, os.path.join(P, "%s_%s.json" % (tag, v)))

# ---- launch list
f = os.path.join(src, "bench_launches.csv")
if os.path.exists(f):
    shutil.copy(f, os.path.join(P, "%s_bench_launches.csv" % tag))
    rows = list(csv.reader(open(f)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("ofb::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    b = last_json(os.path.join(src, "bench_n1.json"))
    with open(os.path.join(P, "%s_bench_launches.md" % tag), "w") as o:
        o.write("# %s — ncu launch list of the benchmark command\n\n" % tag)
        o.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file %s_bench_launches.csv "
                "python bench.py --steps 2 --warmup 3 --no-cpu-baseline`\n(first 600 launches: warm-up and timed device steps of "
                "18 pairs, then the host-buffer steps, which run in chunks).  Times are cold-cache and serialised: compare "
                "SHARES with bench.py's stage timers.\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n" % tag)
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
            o.write("| `%s` | %d | %.1f | %.1f %% |\n" % (k, v[0], v[1] / 1000.0, 100 * v[1] / tot))
        it = sum(v[1] for k, v in agg.items() if k.startswith("k_iter_v"))
        if b:
            o.write("\nShare of the fused iteration kernel: %.1f %% here vs `roofline.share_of_step` = %.3f from the CUDA-event "
                    "stage timers inside bench.py's timed region (profiles/%s_bench_n1.json).\n"
                    % (100 * it / tot, b["roofline"]["share_of_step"], tag))

# ---- ncu capture of the dominant kernel
rep = os.path.join(src, "iter_v.ncu-rep")
if os.path.exists(rep):
    raw = os.path.join(P, "%s_iter_v_ncu_raw.csv" % tag)
    srcp = os.path.join(src, "iter_v_src.csv")
    subprocess.run("ncu -i %s --page raw --csv > %s 2>/dev/null" % (rep, raw), shell=True, check=True)
    subprocess.run("ncu -i %s --page source --csv > %s 2>/dev/null" % (rep, srcp), shell=True, check=True)
    summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), raw, srcp, "12"],
                          capture_output=True, text=True).stdout
    rows = list(csv.reader(open(raw)))
    hdr, r = rows[0], rows[2]
    g = lambda k: float(r[hdr.index(k)].replace(",", ""))
    units = rows[1]
    def in_bytes(k):
        v, u = g(k), units[hdr.index(k)]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    dram = in_bytes("dram__bytes_read.sum") + in_bytes("dram__bytes_write.sum")
    pairs = int(sys.argv[3]) if len(sys.argv) > 3 else 18
    json.dump({"kernel": r[hdr.index("Kernel Name")][:60], "pairs_per_launch": pairs, "dram_bytes_per_launch": dram,
               "gpu_time_us": g("gpu__time_duration.sum"), "source": "%s_iter_v_ncu_raw.csv" % tag},
              open(os.path.join(P, "%s_iter_v_ncu.json" % tag), "w"))
    open(os.path.join(P, "%s_iter_v_ncu_summary.txt" % tag), "w").write(summ)
    print(summ)
    print("dram bytes per launch", dram, "algorithmic", 56 * 2073600 * pairs)
