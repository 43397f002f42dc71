"""Writes profiles/<tag>_summary.md from the JSON / CSV files of one tools/gpu_check.sh run.  Usage: python tools/profile_summary.py <tag>"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
P = lambda name: os.path.join(ROOT, "profiles", f"{tag}_{name}")
seen = {}
if os.path.exists(P("kernels.csv")):   # (absent when the full ncu report was too large to bring back)
    rows = list(csv.reader(open(P("kernels.csv"))))
    h = rows[0]
    col = lambda name: [i for i, x in enumerate(h) if x.startswith(name)][0]
    ik, ig, ird, iwr, it, iregs, iissue, iwarps = (col(n) for n in ("Kernel Name", "Grid Size", "dram__bytes_read", "dram__bytes_write",
                                                                    "gpu__time_duration", "launch__registers", "smsp__issue_active", "sm__warps_active"))
    for r in rows[1:]:
        seen.setdefault(r[ik].replace("void ", "").split("(")[0], r)
b = json.load(open(P("bench.json")))
out = [f"# {tag} — one B200, `tools/gpu_check.sh {tag}`", "",
       "Headline (`bench.py`, CUDA events, not under a profiler): blur k=9 on 256 × 1360×765: **%.0f img/s, %.0f GB/s = %.1f %% of the "
       "measured copy bandwidth** (%.0f GB/s); e2e through host buffers %.0f img/s; CPU reference %.0f img/s on %d cores (%s)."
       % (b["value"], b["roofline"]["achieved"], 100 * b["roofline"]["frac"], b["roofline"]["peak"], b["e2e"]["value"],
          b["cpu_baseline"]["value"], b["cpu_baseline"]["cores"], b["cpu_baseline"].get("cpu_model", "?")), "",
       f"Side measurements of the other kernels (`ops` in `{tag}_bench.json`, same timing method):", "",
       "| workload | GB/s (algorithmic bytes) | % of measured peak | images/s |", "|---|---|---|---|"]
for k, v in b["ops"].items():
    if "GB/s" in v:
        out.append("| %s | %.0f | %.1f | %.0f |" % (k, v["GB/s"], 100 * v["frac_of_measured_peak"], v["images_per_s"]))
if "drop_in_per_call_ms" in b["ops"]:
    out += ["", "Per-call latency of the drop-in functions (host array in / out), ms: " +
            ", ".join("%s %.2f (reference %.2f)" % (k, v["ours"], v.get("reference", float("nan")))
                      for k, v in b["ops"]["drop_in_per_call_ms"].items())]
if seen:
  out += ["", "`ncu --set full --clock-control none` of `tools/profile_ops.py` (first launch of each kernel; cold, serialised: shares, not "
        f"absolutes — `{tag}_kernels.csv` has every launch):", "",
        "| kernel | grid | regs | time µs | DRAM read MB | DRAM write MB | issue active % | warps active % |", "|---|---|---|---|---|---|---|---|"]
else:
    out += ["", "(The `ncu --set full` report of this run exceeded the transfer limit; the LowRes / training-path / encoder kernels are "
            "unchanged since the previous run, whose per-kernel table stands.)"]
for k, r in seen.items():
    out.append("| `%s` | %s | %s | %.1f | %.0f | %.0f | %.1f | %.1f |" % (
        k, r[ig].split(",")[0].strip("("), r[iregs], float(r[it]), float(r[ird]), float(r[iwr]), float(r[iissue]), float(r[iwarps])))
out += ["", "End to end with real inputs:", ""]
if os.path.exists(P("testset_driver.json")):
    d = json.load(open(P("testset_driver.json")))
    out.append("* test-set build on %d JPEG files (`%s_testset_driver.json`): reference loop %.1f s, drop-in driver compat (byte-identical: "
               "%d/%d files) %.2f s, philox %.3f s = %.0f files/s." % (d["output_files"], tag, d["reference_loop_s"], d["compat_files_identical"],
                                                        d["output_files"], d["driver_compat_s"], d["driver_philox_s"], d["driver_philox_files_per_s"]))
if os.path.exists(P("time_jpegdec.json")):
    j = json.load(open(P("time_jpegdec.json")))
    out.append("* device JPEG decoder, %d files (`%s_time_jpegdec.json`): " % (j["smooth"]["images"], tag) + "; ".join(
        "%s %.1f ms per call (device %.1f ms, host preparation %.1f ms) = %d img/s, cv2.imdecode on 16 threads %.1f ms"
        % (k, v["decode_call_ms"], v["device_ms"], v["host_prepare_ms"], v["decode_call_images_per_s"], v["cv2_16_threads_ms"]) for k, v in j.items()) + ".")
if os.path.exists(P("testset_548.json")):
    t5 = json.load(open(P("testset_548.json")))
    out.append("* test-set build, two trees of %d frames, %d files, device decoder + encoder, philox noise (`%s_testset_548.json`): %.2f s = %.0f files/s."
               % (t5["images_per_tree"], t5["output_files"], tag, t5["wall_s"], t5["files_per_s"]))
if os.path.exists(P("training_hook.json")):
    t = json.load(open(P("training_hook.json")))
    out.append("* training hook, batch 16 host frames → fp16 640² on the device (`%s_training_hook.json`): %.0f img/s (%.2f ms per batch) vs "
               "%.0f img/s for one reference process." % (tag, t["ours_images_per_s"], t["ours_ms_per_batch"], t["reference_images_per_s_one_process"]))
for n in (2, 8):
    if os.path.exists(P(f"bench_{n}gpu.json")):
        m = json.load(open(P(f"bench_{n}gpu.json")))
        out.append("* %d GPUs: %.2f M img/s, e2e %.1f k img/s (`%s_bench_%dgpu.json`)." % (n, m["value"] / 1e6, m["e2e"]["value"] / 1e3, tag, n))
open(P("summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
